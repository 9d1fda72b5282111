"""Stage-by-stage parity report of the CUDA path against the oracle (run on the GPU box).
Usage: python tests/parity_report.py [--B 4] [--tp 24] [--frames 150] [--precision fp32 bf16]"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200")
data = importlib.import_module("fine-grained-emotional-control-of-tts_b200.data")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.shape != b.shape:
        return float("nan"), f"shape {tuple(a.shape)} vs {tuple(b.shape)}"
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d / (s + 1e-12), f"abs {d:.3e} scale {s:.3e}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--tp", type=int, default=24)
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--precision", nargs="+", default=["fp32", "bf16"])
    ap.add_argument("--dtype", default="float64")
    args = ap.parse_args()
    (batch, intensity), = data.synthetic_batches(args.B, 1, seed=11, min_tp=8, max_tp=args.tp, max_frames=args.frames, pool_factor=2)
    tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
    odt = getattr(torch, args.dtype)
    oracle = O.build(seed=0, dtype=odt).eval()
    oracle.trace = {}
    crit_o = O.Loss(**O.DEFAULT_LOSS_CONFIG)
    preds_o = oracle(tokens, speakers, dur, pitch.to(odt), energy.to(odt), intensity=intensity.to(odt))
    loss_o = crit_o(preds_o, (mel.to(odt), dur, pitch.to(odt), energy.to(odt), out_lens, in_lens), 0)
    loss_o["total_loss"].backward()
    grads_o = {k: p.grad for k, p in oracle.named_parameters()}
    sd = {k: v.float() for k, v in oracle.state_dict().items()}
    worst = {}
    for prec in args.precision:
        print(f"\n================ precision={prec} ================")
        model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision=prec)
        model.load_state_dict(sd)
        model = model.cuda().eval()
        model.trace = {}
        crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
        c = lambda t: t.cuda()
        preds = model(c(tokens), c(speakers), c(dur), c(pitch), c(energy), intensity=c(intensity))
        torch.cuda.synchronize()
        for k in ["enc_in", "encoder.layer0", "enc_out", "cond", "after_pitch", "after_energy", "dec_in", "decoder.layer0", "dec_out"]:
            if k in oracle.trace and k in model.trace:
                r, info = rel(model.trace[k], oracle.trace[k])
                print(f"  trace {k:14s} rel {r:.3e}  {info}")
        names = ["mel_post", "postnet_output", "predict_durations", "predict_pitch", "avg_pitch", "predict_energy", "avg_energy"]
        for n, a, b in zip(names, preds[:7], preds_o[:7]):
            r, info = rel(a, b)
            worst[(prec, n)] = r
            print(f"  out   {n:18s} rel {r:.3e}  {info}")
        print("  mel_lens equal:", torch.equal(preds[7], preds_o[7]), preds[7].tolist())
        loss = crit(preds, (c(mel), c(dur), c(pitch), c(energy), c(out_lens), c(in_lens)), 0)
        for k in loss:
            r, info = rel(loss[k], loss_o[k])
            worst[(prec, k)] = r
            print(f"  loss  {k:18s} {loss[k].item():.6f} vs {loss_o[k].item():.6f} rel {r:.3e}")
        model.zero_grad()
        loss["total_loss"].backward()
        torch.cuda.synchronize()
        bad = []
        for k, p in model.named_parameters():
            r, info = rel(p.grad, grads_o[k])
            bad.append((r, k, info))
        bad.sort(reverse=True)
        print("  worst parameter gradients:")
        for r, k, info in bad[:25]:
            print(f"    {k:55s} rel {r:.3e} {info}")
        import statistics
        print("  median grad rel err:", statistics.median([b[0] for b in bad]))
        worst[(prec, "grad_max")] = bad[0][0]
        print("  tc error flag:", importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib").gemm_tc_error_flag())


if __name__ == "__main__":
    main()
