"""Fused attention kernels vs the unfused GEMM + softmax + GEMM sequence, B=32, H=2, head_dim 192 (GPU time via CUDA graph).
ATTN_DBG=1 prints the cycle breakdown of CTA 0; it needs a library built with `make EXTRA=-DFS2_TC_PROBE`."""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_sweep import timeit  # noqa: E402

PAD, HD = 4, 192


def main():
    B, H = int(os.environ.get("B", 32)), 2
    D = H * HD
    for T in [int(t) for t in (sys.argv[1:] or ["800", "488", "128"])]:
        TP, ldk, ld = T + 2 * PAD, (T + 7) // 8 * 8, 3 * D
        R = 3
        qkv = [(torch.randn(B * TP, ld, device="cuda") * 0.7).to(torch.bfloat16) for _ in range(R)]
        dO = [(torch.randn(B * TP, D, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(R)]
        lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
        P = [torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16) for _ in range(R)]
        Pd = [torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16) for _ in range(R)]
        dS = [torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
        O = [torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16) for _ in range(R)]
        dqkv = [torch.zeros(B * TP, ld, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
        S = [torch.zeros(B * H, T, ldk, device="cuda") for _ in range(2)]
        sc = 1.0 / math.sqrt(HD)
        flops = 2.0 * B * H * T * T * HD
        for p_drop in (0.0, 0.1):
            f = timeit(lambda i: L.call("fs2_attn_fwd", qkv[i % R], lens, B, H, T, D, ldk, sc, p_drop, 5, None, P[i % R],
                                        Pd[i % R] if p_drop > 0 else None, O[i % R]))
            b = timeit(lambda i: L.call("fs2_attn_bwd", dO[i % R], O[i % R], qkv[i % R], P[i % R], lens, B, H, T, D, ldk, sc, p_drop,
                                        5, None, dS[i % 2], dqkv[i % 2]))
            print(f"T={T} p={p_drop}: fused fwd {f:8.1f} us ({2 * flops / f / 1e6:5.0f} TF/s useful)   fused bwd(dS,dQ) {b:8.1f} us", flush=True)

        def unfused_fwd(i):
            q = qkv[i % R]
            L.gemm(mode=0, M=T, N=T, K=HD, A=q, A_off=PAD * ld, lda=ld, a_rows=T, a_inner=HD, a_s1=HD, a_s2=TP * ld,
                   B=q, B_off=PAD * ld + D, ldb=ld, b_rows=T, b_inner=HD, b_s1=HD, b_s2=TP * ld, batch1=H, batch2=B,
                   Cout=S[i % 2], ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=True)
            L.call("fs2_softmax_fwd", S[i % 2], lens, B, H, T, ldk, sc, 0.1, 5, None, P[i % R], Pd[i % R], 1)
            L.gemm(mode=1, M=T, N=HD, K=T, A=Pd[i % R], lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
                   B=q, B_off=PAD * ld + 2 * D, ldb=ld, b_rows=T, b_inner=HD, b_s1=HD, b_s2=TP * ld, batch1=H, batch2=B,
                   Cout=O[i % R], C_off=PAD * D, ldc=D, c_s1=HD, c_s2=TP * D, c_bf16=True, ab_bf16=True)

        if os.environ.get("ATTN_DBG"):
            lib = L.load()
            dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
            lib.fs2_attn_set_debug.argtypes = [L.C.c_void_p]
            lib.fs2_attn_set_debug(L.C.c_void_p(dbg.data_ptr()))
            for name, fn in (("fwd", lambda: L.call("fs2_attn_fwd", qkv[0], lens, B, H, T, D, ldk, sc, 0.1, 5, None, P[0], Pd[0], O[0])),
                             ("bwd", lambda: L.call("fs2_attn_bwd", dO[0], O[0], qkv[0], P[0], lens, B, H, T, D, ldk, sc, 0.1, 5, None, dS[0], dqkv[0]))):
                fn(); fn()
                torch.cuda.synchronize()
                d = dbg.tolist()
                print(f"  {name} CTA0: MMA thread total {d[0]} cyc over {d[5]} jobs; waits K {d[1]} S-free {d[2]} P-ready {d[3]} V {d[4]} | "
                      f"softmax thread total {d[8]}: wait-S {d[9]} ld {d[10]} math+gst {d[11]} wait-Pbuf {d[12]} sts+fence+arrive {d[13]}")
            lib.fs2_attn_set_debug(None)
        print(f"T={T}: unfused fwd (QK^T + softmax/dropout + PV) {timeit(unfused_fwd):8.1f} us", flush=True)
    print("flag", L.gemm_tc_error_flag())


if __name__ == "__main__":
    main()
