"""Print selected metrics per kernel launch from an .ncu-rep (reads `ncu -i rep --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
metrics = sys.argv[2].split(",") if len(sys.argv) > 2 else [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [m for m in metrics if m in idx]
print("kernel | " + " | ".join(f"{m} [{units[idx[m]]}]" for m in cols))
for r in rows[2:]:
    print(r[idx["Kernel Name"]][:48] + " | " + " | ".join(r[idx[m]] for m in cols))
