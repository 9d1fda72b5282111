"""SURVEY 8f rows 1-2 measured: the frozen IntensityExtractor forward + duration-segment mean that the reference runs in
front of every FastSpeech2 training step (train.py:16-51, 69), on the bench's own batches (batch 32, length-bucketed).
Prints one JSON line: frames/s of get_intensity_representation and its share next to the training step."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "fine-grained-emotional-control-of-tts_b200"


def main():
    pkg = importlib.import_module(PKG)
    data = importlib.import_module(PKG + ".data")
    dev = torch.device("cuda")
    torch.manual_seed(0)
    ext = pkg.IntensityExtractor(**pkg.DEFAULT_RANK_MODEL_CONFIG).to(dev).eval()
    batches = [tuple(t.to(dev) if torch.is_tensor(t) else t for t in b) for b, _ in data.synthetic_batches(32, 4, seed=1234, rank=0)]
    frames = [int(b[7].sum()) for b in batches]

    def step(i):
        return pkg.get_intensity_representation(ext, batches[i % 4], dev)

    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    tot = 0
    for i in range(n):
        step(i)
        tot += frames[i % 4]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # algorithmic FLOPs of the padded rectangles: per frame and layer 2*(3D*D + D*D + 2*k*D*4D) + attention
    D, k, nl = 384, 9, 6
    fl = 0.0
    for b in batches:
        B, T = b[10].shape[0], b[10].shape[2]
        fl += nl * (B * T * 2 * (3 * D * D + D * D + 2 * k * D * 4 * D) + B * 2 * (2 * T * T * (D // 2)) * 2) + B * T * 2 * 82 * D
    fl *= n / 4
    print(json.dumps({"metric": "intensity_representation_mel_frames_per_sec", "value": tot / (ms * 1e-3), "unit": "mel_frames/s",
                      "ms_per_call": ms / n, "tflops": fl / (ms * 1e-3) / 1e12, "steps": n, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "get_intensity_representation = IntensityExtractor forward (6 FFT blocks, conv k=9 both "
                                             "ways, GELU) + duration-segment mean, batch 32, the bench's 4 length-bucketed batches",
                                 "padded_shapes_Tp_Tm": [(b[0].shape[1], b[10].shape[2]) for b in batches]}}))


if __name__ == "__main__":
    main()
