"""Event-timed comparison of fs2_gemm_ln_tc (one launch) with the two launches it replaces (fs2_gemm_tc with an fp32
branch in HBM + fs2_ln_fwd) at the step's shapes: out-projection (K = 384) and FFN conv 2 (K = 1536), batch 32, the
bench's frame counts.  A ring of buffers larger than the 126 MB L2 keeps every launch on HBM.
`python tools/gemm_ln_bench.py [T ...]` prints one JSON line per (K, T)."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "fine-grained-emotional-control-of-tts_b200"
PAD, C = 4, 384
PDROP = float(os.environ.get("GEMM_LN_P", "0.1"))      # branch dropout (0 selects the kernel instantiation without it)


def timeit(fn, ring, iters=30, warm=4):
    iters, warm = int(os.environ.get("GEMM_LN_ITERS", iters)), int(os.environ.get("GEMM_LN_WARM", warm))   # 1 / 1 under ncu
    for i in range(warm):
        fn(i % ring)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % ring)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    L = importlib.import_module(PKG + "._lib")
    Ts = [int(a) for a in sys.argv[1:] if a.isdigit()] or [int(a) for a in os.environ.get("GEMM_LN_T", "800,632,488,376,128").split(",")]
    B = 32
    out = []
    dbg = None
    if os.environ.get("GEMM_LN_DBG"):          # probe build: FS2_B200_LIB=<pkg>/libfs2_b200_probe.so
        dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
        raw = L.load()
        raw.fs2_gemm_ln_set_debug.argtypes = [L.C.c_void_p]
        raw.fs2_gemm_ln_set_debug(L.C.c_void_p(dbg.data_ptr()))
    for K in (384, 1536):
        for T in Ts:
            rows = B * (T + 2 * PAD)
            per = rows * (K * 2 + C * 4 * 3 + C * 2)
            ring = max(2, min(8, int(300e6 // per) + 1))
            A = [torch.randn(rows, K, device="cuda").to(torch.bfloat16) for _ in range(ring)]
            x = [torch.randn(rows, C, device="cuda") for _ in range(ring)]
            proj = [torch.empty(rows, C, device="cuda") for _ in range(ring)]
            of = [torch.empty(rows, C, device="cuda") for _ in range(ring)]
            oa = [torch.empty(rows, C, device="cuda", dtype=torch.bfloat16) for _ in range(ring)]
            W = (torch.randn(C, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
            bias, gamma, beta = torch.randn(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
            mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
            ctr = torch.zeros(1, dtype=torch.int64, device="cuda")

            def two(i):
                L.gemm(mode=0, M=rows, N=C, K=K, A=A[i], lda=K, a_rows=rows, a_inner=K, B=W, ldb=K, b_rows=C, b_inner=K,
                       Cout=proj[i], ldc=C, c_bf16=False, ab_bf16=True, bias=bias, rs_T=T, rs_Tp=T + 2 * PAD)
                p = L.Fs2LnFwd()
                p.B, p.T, p.C = B, T, C
                p.x, p.branch = x[i].data_ptr(), proj[i].data_ptr()
                p.drop_b_p, p.drop_b_seed = PDROP, 1234
                p.gamma, p.beta, p.eps = gamma.data_ptr(), beta.data_ptr(), 1e-6
                p.out_f32, p.out_act, p.act_bf16, p.halo = of[i].data_ptr(), oa[i].data_ptr(), 1, 4
                p.mean, p.rstd, p.seed_dev = mean.data_ptr(), rstd.data_ptr(), ctr.data_ptr()
                L.call("fs2_ln_fwd", L.C.addressof(p))

            def one(i):
                q = L.Fs2GemmLn()
                q.B, q.T, q.K, q.lda, q.ldw = B, T, K, K, K
                q.A, q.W, q.bias, q.x = A[i].data_ptr(), W.data_ptr(), bias.data_ptr(), x[i].data_ptr()
                q.drop_p, q.drop_seed, q.seed_dev = PDROP, 1234, ctr.data_ptr()
                q.gamma, q.beta, q.eps = gamma.data_ptr(), beta.data_ptr(), 1e-6
                q.out_f32, q.out_act, q.halo = of[i].data_ptr(), oa[i].data_ptr(), 4
                q.mean, q.rstd = mean.data_ptr(), rstd.data_ptr()
                L.call("fs2_gemm_ln_tc", L.C.addressof(q))

            t2, t1 = timeit(two, ring), timeit(one, ring)
            alg = rows * (K * 2 + C * (4 + 4 + 2))            # A + residual in, fp32 + bf16 out
            out.append({"K": K, "T": T, "rows": rows, "two_launches_us": round(t2, 1), "fused_us": round(t1, 1),
                        "alg_bytes": alg, "fused_alg_GB/s": round(alg / t1 * 1e-3, 1),
                        "fused_TFLOP/s": round(2.0 * rows * C * K / t1 * 1e-6, 1)})
            print(json.dumps(out[-1]), flush=True)
            if dbg is not None:
                torch.cuda.synchronize()
                print("   epilogue of CTA 0, first tile [cycles]: setup %d, wait for MMAs %d, pass 1 %d, pass 2 %d, pass 3 %d"
                      % tuple(dbg[:5].tolist()), flush=True)
                d = dbg.tolist()
                print("   pass 1, third step: loads + dropout words issued %d, wait TMEM %d, math + TMEM store %d, residual of the next step arrived %d"
                      % (d[9] - d[8], d[10] - d[9], d[11] - d[10], d[13] - d[11]), flush=True)
                print("   pass 1, cycles per step: %s, loop end -> tensor-memory stores complete %d"
                      % (" ".join(str(d[23 + k] - d[22 + k]) for k in range(5)) + " " + str(d[28] - d[27]), d[29] - d[28]), flush=True)
                print("   pass 3, third step: wait TMEM (+ gamma / beta) %d, staging tile free %d, math + st.shared %d, proxy fence %d, 3 bulk stores issued %d"
                      % (d[17] - d[16], d[18] - d[17], d[19] - d[18], d[20] - d[19], d[21] - d[20]), flush=True)
            del A, x, proj, of, oa
    assert L.gemm_tc_error_flag() == 0
    return out


if __name__ == "__main__":
    main()
