"""Micro-benchmark of the memory-bound kernels of the step against the measured HBM peak (BASELINE
configs[1]: LengthRegulator + VarianceAdaptor pieces standalone at batch 64, plus the loss / LayerNorm / AdamW
passes at the configs[2] size).  Every kernel is called through the C ABI, timed with CUDA events over a ring
of buffers whose total size exceeds the 126 MB L2 (so every launch reads HBM, not L2), and reported as
algorithmic bytes / time / peak.  Prints one JSON object per kernel; `python tools/hbm_bench.py > profiles/...`.
"""
import importlib
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "fine-grained-emotional-control-of-tts_b200"
PAD = 4


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbs_burst", "hbm_copy_gbs_burst", "hbm_gbs", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured"
        for v in d.values():
            if isinstance(v, dict):
                for k in ("burst", "gbs_burst", "gbs"):
                    if k in v and "hbm" in json.dumps(v).lower():
                        return float(v[k]), "measured"
    except Exception:
        pass
    return 6455.6, "fallback (B200_PROFILING.md)"


def timeit(fn, ring, iters=40, warm=5):
    iters = int(os.environ.get("HBM_ITERS", iters))
    warm = int(os.environ.get("HBM_WARM", warm))
    for i in range(warm):
        fn(i % ring)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ring)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters       # us


def main():
    L = importlib.import_module(PKG + "._lib")
    peak, src = peak_gbs()
    dev = "cuda"
    g = torch.Generator().manual_seed(7)
    out = []

    def report(name, alg_bytes, us, note=""):
        gbs = alg_bytes / us * 1e-3
        r = {"kernel": name, "alg_bytes": int(alg_bytes), "us": round(us, 2), "GB/s": round(gbs, 1),
             "frac_of_hbm_peak": round(gbs / peak, 3), "peak_GB/s": peak, "peak_source": src, "note": note}
        out.append(r)
        print(json.dumps(r), flush=True)

    # ------------------------------------------------- reference point: a plain device copy of the same byte count
    # (the measured HBM peak comes from a 2 GiB copy; a 91 MB launch pays its ramp and tail inside the timed window)
    n_ref = 45_629_440 // 4
    src_c = [torch.randn(n_ref, device=dev) for _ in range(6)]
    dst_c = [torch.empty(n_ref, device=dev) for _ in range(6)]
    us = timeit(lambda i: dst_c[i].copy_(src_c[i]), 6)
    report("torch copy_ of 45.6 MB (91.3 MB read+write): same bytes as lr_expand", 2 * n_ref * 4, us,
           "what a small launch can reach at all; compare the LengthRegulator rows against this, not only against the 2 GiB peak")
    del src_c, dst_c

    # ------------------------------------------------------------- cfg-2: LengthRegulator, B=64, Tp=128, D=384
    B, Tp, D = 64, 128, 384
    dur = torch.exp(torch.randn(B, Tp, generator=g) * 0.6 + 1.6).round().clamp(0, 40).long()
    scale = 800.0 / dur.sum(1, keepdim=True).clamp(min=1).float()
    dur = (dur.float() * scale.clamp(max=1.0)).floor().long()
    # top up so that every item has exactly 800 frames (full rectangle: bytes are then exactly B*Tm*D*e)
    dur[:, 0] += 800 - dur.sum(1)
    Tm = 800
    dur_c = dur.to(dev)
    ends = torch.zeros(B, Tp, dtype=torch.int32, device=dev)
    mel_lens = torch.zeros(B, dtype=torch.int32, device=dev)
    L.call("fs2_lr_prepare", dur_c, None, 1.0, B, Tp, ends, mel_lens)
    assert mel_lens.cpu().tolist() == [Tm] * B
    R = 6                                            # ring: 6 x (12.6 + 78.6) MB >> L2
    feats = [torch.randn(B, Tp, D, device=dev) for _ in range(R)]
    outs = [torch.empty(B, Tm, D, device=dev) for _ in range(R)]
    us = timeit(lambda i: L.call("fs2_lr_prepare", dur_c, None, 1.0, B, Tp, ends, mel_lens), 1)
    report("lr_prepare (duration scan, B=64,Tp=128)", B * Tp * (8 + 4) + B * 4, us, "latency-bound: 98 KB")
    pe = torch.randn(Tm, D, device=dev)
    of = [torch.empty(B * (Tm + 8), D, device=dev) for _ in range(R)]
    oa = [torch.empty(B * (Tm + 8), D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
    fin = [torch.randn(B * (Tp + 8), D, device=dev) for _ in range(R)]
    df = [torch.randn(B * (Tm + 8), D, device=dev) for _ in range(R)]
    dph = [torch.zeros(B * (Tp + 8), D, device=dev) for _ in range(R)]
    lib = L.load()
    lib.fs2_lr_tune.argtypes = [L.C.c_int]
    lib.fs2_lr_bulk_rows.argtypes = [L.C.c_int]
    lib.fs2_lr_bulk_rows(0)                          # SIMT kernels in this loop; the bulk-copy form is swept below
    for rb in [int(x) for x in os.environ.get("LR_RB", "4").split(",")]:
        lib.fs2_lr_tune(rb)
        tag = f" [rows in flight {rb}]"
        us = timeit(lambda i: L.call("fs2_lr_expand", feats[i], Tp, 0, ends, mel_lens, None, B, Tp, Tm, D, outs[i], None, 0,
                                     Tm, 0, None), R)
        report("lr_expand fp32 (B=64,Tp=128,Tm=800,D=384)" + tag, B * Tp * D * 4 + B * Tm * D * 4 + B * Tp * 4, us,
               "SURVEY 8d K10: read B*Tp*D*4 + write B*Tm*D*4")
        # the fused form the model uses: padded rows, + pos-enc, fp32 residual stream + bf16 operand copy
        us = timeit(lambda i: L.call("fs2_lr_expand", fin[i], Tp + 8, PAD, ends, mel_lens, pe, B, Tp, Tm, D, of[i], oa[i], 1,
                                     Tm + 8, PAD, None), R)
        report("lr_expand fused (+pos-enc, fp32 + bf16 outputs, padded rows)" + tag,
               B * Tp * D * 4 + B * (Tm + 8) * D * 6 + Tm * D * 4, us,
               "model form: writes the fp32 stream and the bf16 GEMM operand in one pass")
        us = timeit(lambda i: L.call("fs2_lr_bwd", df[i], None, Tm + 8, PAD, ends, mel_lens, B, Tp, Tm, D, dph[i], Tp + 8, PAD), R)
        report("lr_bwd (frame-parallel segment sums: memset + vector atomics, B=64)" + tag, B * Tm * D * 4 + B * Tp * D * 4, us,
               "read B*Tm*D*4, write B*Tp*D*4 (+ the 13 MB memset of dphon inside the call)")
        if hasattr(lib, "fs2_lr_tune_bwd"):
            lib.fs2_lr_tune_bwd.argtypes = [L.C.c_int]
            lib.fs2_lr_tune_bwd(1)
            us = timeit(lambda i: L.call("fs2_lr_bwd", df[i], None, Tm + 8, PAD, ends, mel_lens, B, Tp, Tm, D, dph[i], Tp + 8, PAD), R)
            report("lr_bwd (reproducible form: one CTA per 8 phonemes, plain stores, B=64)" + tag, B * Tm * D * 4 + B * Tp * D * 4, us,
                   "opt-in (fs2_lr_tune_bwd(1)): no memset, no atomics, fixed summation order")
            lib.fs2_lr_tune_bwd(0)
    lib.fs2_lr_tune(4)
    # plain fp32 expansion on the bulk-copy engine (cp.async.bulk in and out of shared memory), rows per CTA swept;
    # 0 = the SIMT kernel measured above
    for rows in [int(x) for x in os.environ.get("LR_BULK", "0,8,16,32,64,128").split(",")]:
        lib.fs2_lr_bulk_rows(rows)
        us = timeit(lambda i: L.call("fs2_lr_expand", feats[i], Tp, 0, ends, mel_lens, None, B, Tp, Tm, D, outs[i], None, 0,
                                     Tm, 0, None), R)
        report("lr_expand fp32 (B=64,Tp=128,Tm=800,D=384)" + (f" [bulk-copy engine, {rows} rows per CTA]" if rows else " [SIMT kernel]"),
               B * Tp * D * 4 + B * Tm * D * 4 + B * Tp * 4, us, "SURVEY 8d K10: read B*Tp*D*4 + write B*Tm*D*4")
    lib.fs2_lr_bulk_rows(8)
    del feats, outs, of, oa, fin, df, dph

    # ------------------------------------------------------------------------ average_over_durations (K8)
    pitch = [torch.randn(B, Tm, device=dev) for _ in range(4)]
    avg = torch.empty(B, Tp, device=dev)
    try:
        us = timeit(lambda i: L.call("fs2_avg_over_durations", pitch[i], dur_c, B, Tp, Tm, avg, None, None, None), 4)
        report("avg_over_durations (B=64,Tp=128,Tm=800)", B * Tm * 4 + B * Tp * (8 + 4), us, "latency-bound: 0.3 MB")
    except Exception as e:                                          # signature drift: report, don't die
        print(json.dumps({"kernel": "avg_over_durations", "error": str(e)[:200]}), flush=True)

    # ---------------------------------------------------------------- masked MSE losses (K13), cfg-3 size
    pkg = importlib.import_module(PKG)
    B3, Tm3, Tp3 = 32, 800, 128
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    mk = lambda *s: torch.randn(*s, device=dev)
    mel_len = torch.full((B3,), Tm3, dtype=torch.int64, device=dev)
    ph_len = torch.full((B3,), Tp3, dtype=torch.int64, device=dev)
    durs = torch.full((B3, Tp3), Tm3 // Tp3, dtype=torch.int64, device=dev)
    R3 = 8
    preds = [(mk(B3, Tm3, 80).requires_grad_(), mk(B3, Tm3, 80).requires_grad_(), mk(B3, Tp3).requires_grad_(),
              mk(B3, Tp3, 1).requires_grad_(), mk(B3, Tp3, 1), mk(B3, Tp3, 1).requires_grad_(), mk(B3, Tp3, 1),
              mel_len.cpu()) for _ in range(R3)]
    tgts = [(mk(B3, Tm3, 80), durs, mk(B3, Tm3), mk(B3, Tm3), mel_len, ph_len) for _ in range(R3)]
    with torch.no_grad():
        us = timeit(lambda i: crit(preds[i], tgts[i], 0), R3, iters=20)
    report("Loss.forward (5xMSE + SSIM, values + gradients, B=32,Tm=800)", 3 * B3 * Tm3 * 80 * 4 + 2 * B3 * Tm3 * 80 * 4, us,
           "whole Loss call (mse + ssim kernels, several launches); bytes = 3 reads + 2 gradient writes of (B,Tm,80) fp32")

    # ------------------------------------------------------------------------------------ AdamW (K16)
    n = 85295299
    p_, g_, m_, v_ = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v_.abs_()
    us = timeit(lambda i: L.call("fs2_adamw", p_, g_, m_, v_, n, 1e-4, 0.9, 0.999, 1e-8, 0.01, i + 1, 1.0), 1, iters=10)
    report("adamw (85.3 M params, fp32 state)", n * 7 * 4, us, "4 reads + 3 writes of 341 MB")
    del p_, g_, m_, v_

    # -------------------------------------------------------- LayerNorm fwd (residual + LN + bf16 copy), cfg-3
    model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16").to(dev)
    model._ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    rows = B3 * (Tm3 + 8)
    C = 384
    R4 = 6
    xs = [torch.randn(rows, C, device=dev) for _ in range(R4)]
    brs = [torch.randn(rows, C, device=dev) for _ in range(R4)]
    o32 = [torch.empty(rows, C, device=dev) for _ in range(R4)]
    o16 = [torch.empty(rows, C, device=dev, dtype=torch.bfloat16) for _ in range(R4)]
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    raw = L.load()
    raw.fs2_ln_tune.argtypes = [L.C.c_int]
    PFS = [int(x) for x in os.environ.get("LN_PF", "0,1").split(",")]   # L2 prefetch distance in rows of a warp (default 1)
    for pf in PFS:
        raw.fs2_ln_tune(pf)
        us = timeit(lambda i: model._ln_fwd(B3, Tm3, C, xs[i], gam, bet, 1e-6, branch=brs[i], drop_b=(0.1, 11), out_f32=o32[i],
                                            out_act=o16[i], halo=4, mean=mean, rstd=rstd), R4)
        report("ln_fwd (residual + dropout + LN -> fp32 + bf16, B=32,Tm=800,C=384)" + (f" [L2 prefetch distance {pf}]" if pf else " [no L2 prefetch]"),
               rows * C * (4 + 4 + 4 + 2) + rows * 8, us,
               "read x fp32 + branch fp32, write fp32 stream + bf16 operand + mean/rstd")
    dys = [torch.randn(rows, C, device=dev) for _ in range(R4)]
    dx = [torch.empty(rows, C, device=dev) for _ in range(R4)]
    da = [torch.empty(rows, C, device=dev, dtype=torch.bfloat16) for _ in range(R4)]
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    for pf in PFS:
        raw.fs2_ln_tune(pf)
        us = timeit(lambda i: model._ln_bwd(B3, Tm3, C, xs[i], gam, bet, 1e-6, mean, rstd, dy=dys[i], branch=brs[i],
                                            drop_b=(0.1, 11), dx_f32=dx[i], dact=da[i], dgamma=dg, dbeta=db), R4)
        report("ln_bwd (B=32,Tm=800,C=384)" + (f" [L2 prefetch distance {pf}]" if pf else " [no L2 prefetch]"), rows * C * (4 + 4 + 4 + 4 + 2) + rows * 8, us,
               "read dy fp32 + x fp32 + branch fp32, write dx fp32 + dbranch bf16")
    # the form the FFT blocks use since fs2_gemm_ln_tc: x_hat from the saved LayerNorm output (no x, no branch)
    raw.fs2_ln_tune(PFS[-1])
    us = timeit(lambda i: model._ln_bwd(B3, Tm3, C, None, gam, bet, 1e-6, None, rstd, dy=dys[i], y=o32[i],
                                        drop_b=(0.1, 11), dx_f32=dx[i], dact=da[i], dgamma=dg, dbeta=db), R4)
    report("ln_bwd, x_hat from the forward output (B=32,Tm=800,C=384)", rows * C * (4 + 4 + 4 + 2) + rows * 4, us,
           "read dy fp32 + y fp32, write dx fp32 + dbranch bf16")
    cs = torch.zeros(1536, device=dev)
    big = [torch.randn(rows, 1536, device=dev).bfloat16() for _ in range(3)]
    us = timeit(lambda i: L.call("fs2_colsum", big[i], 1, rows, 1536, 1536, cs), 3)
    report("colsum bf16 (bias gradient, rows=25856, C=1536)", rows * 1536 * 2, us, "one read")
    us = timeit(lambda i: L.call("fs2_colsum", da[i], 1, rows, C, C, dg), R4)
    report("colsum bf16 (bias gradient, rows=25856, C=384)", rows * C * 2, us, "one read")
    return out


if __name__ == "__main__":
    main()
