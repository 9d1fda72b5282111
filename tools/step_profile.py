"""Phase breakdown of the training step on one GPU (CUDA events between phases) and host-side timing of the
e2e loop.  Run on the GPU box:  python tools/step_profile.py [--no-graphs]"""
import argparse
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "fine-grained-emotional-control-of-tts_b200"
pkg = importlib.import_module(PKG)
data = importlib.import_module(PKG + ".data")
L = importlib.import_module(PKG + "._lib")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16").to(dev).train()
    model.use_cuda_graphs = not args.no_graphs
    model.async_mel_lens = True
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    opt = pkg.FusedAdamW(model, lr=1e-4)
    host = [([t.pin_memory() for t in b[:8]], i.pin_memory()) for b, i in data.synthetic_batches(32, 4, seed=1234, rank=0)]
    res = [([t.to(dev) for t in b], i.to(dev)) for b, i in host]

    def step(b, it, evs=None, stamps=None):
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = b
        mark = (lambda: evs.append(_rec())) if evs is not None else (lambda: None)
        hs = (lambda: stamps.append(time.perf_counter())) if stamps is not None else (lambda: None)
        mark(); hs()
        opt.zero_grad()
        mark(); hs()
        preds = model(tokens, speakers, dur, pitch, energy, intensity=it)
        mark(); hs()
        losses = crit(preds, (mel, dur, pitch, energy, out_lens, in_lens), 0)
        mark(); hs()
        losses["total_loss"].backward()
        mark(); hs()
        opt.step()
        mark(); hs()
        return losses

    def _rec():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for i in range(10):
        step(*res[i % 4])
    torch.cuda.synchronize()
    names = ["zero_grad", "forward", "loss", "backward", "adamw"]
    for j in range(4):
        tot = [0.0] * 5
        host_t = [0.0] * 5
        for _ in range(args.steps):
            evs, stamps = [], []
            torch.cuda.synchronize()
            step(*res[j], evs=evs, stamps=stamps)
            torch.cuda.synchronize()
            for k in range(5):
                tot[k] += evs[k].elapsed_time(evs[k + 1])
                host_t[k] += 1e3 * (stamps[k + 1] - stamps[k])
        shape = (res[j][0][0].shape[1], res[j][0][3].shape[1])
        print(f"batch {j} (Tp,Tm)={shape}: " + "  ".join(f"{n} {t / args.steps:.3f}" for n, t in zip(names, tot)) +
              f"  | sum {sum(tot) / args.steps:.3f} ms (device, one step in flight)")
        print(f"      host ms: " + "  ".join(f"{n} {t / args.steps:.3f}" for n, t in zip(names, host_t)) +
              f"  | sum {sum(host_t) / args.steps:.3f}")
    # e2e loop: host stamps
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg = [0.0, 0.0, 0.0]
    n = 16
    for i in range(n):
        a = time.perf_counter()
        b = [t.to(dev, non_blocking=True) for t in host[i % 4][0]]
        it = host[i % 4][1].to(dev, non_blocking=True)
        c = time.perf_counter()
        losses = step(b, it)
        d = time.perf_counter()
        float(losses["total_loss"])
        e = time.perf_counter()
        seg[0] += c - a
        seg[1] += d - c
        seg[2] += e - d
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"e2e loop: {1e3 * dt / n:.3f} ms/step; host: h2d-enqueue {1e3 * seg[0] / n:.3f}  step-enqueue {1e3 * seg[1] / n:.3f}  "
          f"loss.item wait {1e3 * seg[2] / n:.3f}")
    # same loop without the item() sync
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        b = [t.to(dev, non_blocking=True) for t in host[i % 4][0]]
        it = host[i % 4][1].to(dev, non_blocking=True)
        step(b, it)
    torch.cuda.synchronize()
    print(f"e2e loop without loss.item(): {1e3 * (time.perf_counter() - t0) / n:.3f} ms/step")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        step(*res[i % 4])
    torch.cuda.synchronize()
    print(f"resident loop: {1e3 * (time.perf_counter() - t0) / n:.3f} ms/step")


if __name__ == "__main__":
    main()
