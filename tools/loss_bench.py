"""Event-timed Loss.forward (+ gradients) and its backward at the configs[2] size and at the bench's own batch shapes
(tools/hbm_bench.py methodology: ring of operand sets larger than L2, CUDA events, no profiler).
Usage: python tools/loss_bench.py  -> one JSON line per shape."""
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
    dev = "cuda"
    peak = 6455.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    for B, Tp, Tm, ragged in ((32, 128, 800, False), (32, 117, 632, True), (32, 74, 376, True)):
        g = torch.Generator().manual_seed(B + Tm)
        mel_len = torch.randint(int(0.85 * Tm), Tm + 1, (B,), generator=g) if ragged else torch.full((B,), Tm)
        mel_len[0] = Tm
        ph_len = torch.randint(int(0.8 * Tp), Tp + 1, (B,), generator=g) if ragged else torch.full((B,), Tp)
        R = 8
        mk = lambda *s: torch.randn(*s, device=dev)
        sets = []
        for _ in range(R):
            tgt = torch.rand(B, Tm, 80, device=dev) * 13.5 - 11.5
            mo = (tgt + 0.7 * mk(B, Tm, 80)).requires_grad_()
            po = (tgt + 0.5 * mk(B, Tm, 80)).requires_grad_()
            preds = (mo, po, mk(B, Tp).requires_grad_(), mk(B, Tp, 1).requires_grad_(), mk(B, Tp, 1), mk(B, Tp, 1).requires_grad_(),
                     mk(B, Tp, 1), mel_len)
            tg = (tgt, torch.randint(0, 12, (B, Tp), device=dev), None, None, mel_len.to(dev), ph_len.to(dev))
            sets.append((preds, tg))

        def fwd(i):
            return crit(sets[i % R][0], sets[i % R][1], 0)

        def fwd_bwd(i):
            fwd(i)["total_loss"].backward()

        res = {}
        for name, fn in (("forward_with_gradients", fwd), ("forward_plus_backward", fwd_bwd)):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 40
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) * 1e3 / n
        alg = 5 * B * Tm * 80 * 4         # 3 reads (mel_out, postnet_out, target) + 2 gradient writes
        us = res["forward_with_gradients"]
        print(json.dumps({"kernel": "Loss.forward (5xMSE + SSIM, values + gradients): fs2_loss_fused, 3 launches",
                          "B": B, "Tp": Tp, "Tm": Tm, "ragged": ragged, "alg_bytes": alg, "us": round(us, 2),
                          "us_forward_plus_backward": round(res["forward_plus_backward"], 2),
                          "GB/s": round(alg / us * 1e-3, 1), "frac_of_hbm_peak": round(alg / us * 1e-3 / peak, 3), "peak_GB/s": peak}),
              flush=True)


if __name__ == "__main__":
    main()
