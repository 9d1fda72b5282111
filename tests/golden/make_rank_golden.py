"""Freeze outputs of the REAL reference IntensityExtractor (rank_model/model.py) on seeded inputs.
Run in the build container (needs /root/reference):  python tests/golden/make_rank_golden.py
Weights come from tests/golden/rank_weights.py (rebuilt from seeds, not stored)."""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rank_weights as RW  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_rank_model", "/root/reference/emo_rank_tts/rank_model/model.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

model = ref.IntensityExtractor(**RW.CFG).eval()
missing = model.load_state_dict(RW.state_dict(), strict=True)        # same keys and shapes as the reference: strict
g = torch.Generator().manual_seed(1234)
B, T = 3, 45
length = torch.tensor([45, 31, 12])
x = torch.randn(B, T, 82, generator=g)
for b in range(B):
    x[b, length[b]:] = 0.0                                   # collate zero-pads (fastspeech2/dataset.py:94-95)
emotions = torch.tensor([0, 3, 4])
with torch.no_grad():
    out = model(x, length, emotions)
torch.save({"x": x, "length": length, "emotions": emotions, "out": out}, os.path.join(HERE, "rank_extractor.pt"))
print("saved", tuple(out.shape), float(out.abs().max()), missing)
