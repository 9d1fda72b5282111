"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:8.1f}  {k[:100]}")
