"""Flash-style attention kernels (fs2_flash_attn_fwd / _bwd), B=32, H=2, head_dim 192: GPU time via CUDA graph replay,
useful TFLOP/s (forward 4*T^2*192 per (b,h); backward 10*T^2*192), probabilities through tensor memory vs shared memory."""
import ctypes
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
    lib = L.load()
    for T in [int(t) for t in (sys.argv[1:] or ["800", "632", "488", "376", "128"])]:
        TP, ld = T + 2 * PAD, 3 * D
        R = 3
        Tl = int(lib.fs2_flash_attn_lse_len(ctypes.c_int(T)))
        qkv = [(torch.randn(B * TP, ld, device="cuda") * 0.7).to(torch.bfloat16) for _ in range(R)]
        dO = [(torch.randn(B * TP, D, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(R)]
        lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
        O = [torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16) for _ in range(R)]
        lse = [torch.zeros(B * H, Tl, device="cuda") for _ in range(R)]
        dvec = torch.zeros(B * H, Tl, device="cuda")
        dqkv = [torch.zeros(B * TP, ld, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
        sc = 1.0 / math.sqrt(HD)
        flops = 2.0 * B * H * T * T * HD
        if os.environ.get("FLASH_ONCE"):           # for `ncu --set full -k regex:fa_ -c 6`: forward, dQ, dK/dV twice each
            for i in range(2):
                L.call("fs2_flash_attn_fwd", qkv[i], lens, B, H, T, D, sc, 0.1, 5, None, lse[i], O[i], 0)
            for i in range(2):
                L.call("fs2_flash_attn_bwd", dO[i], O[i], qkv[i], lse[i], lens, B, H, T, D, sc, 0.1, 5, None, dvec, dqkv[i], 0)
            torch.cuda.synchronize()
            continue
        for p_drop in (0.0, 0.1):
            res = {}
            for pt in (1, 0):
                lib.fs2_flash_attn_tune(ctypes.c_int(pt))
                res[pt] = timeit(lambda i: L.call("fs2_flash_attn_fwd", qkv[i % R], lens, B, H, T, D, sc, p_drop, 5, None,
                                                  lse[i % R], O[i % R], 0))
            lib.fs2_flash_attn_tune(ctypes.c_int(1))
            for i in range(R):
                L.call("fs2_flash_attn_fwd", qkv[i], lens, B, H, T, D, sc, p_drop, 5, None, lse[i], O[i], 0)
            b = timeit(lambda i: L.call("fs2_flash_attn_bwd", dO[i % R], O[i % R], qkv[i % R], lse[i % R], lens, B, H, T, D, sc,
                                        p_drop, 5, None, dvec, dqkv[i % 2], 0))
            print(f"T={T} p={p_drop}: fwd P-in-TMEM {res[1]:7.1f} us ({2 * flops / res[1] / 1e6:5.0f} TF/s useful)  P-in-smem "
                  f"{res[0]:7.1f} us   bwd (dQ + dK/dV kernels) {b:7.1f} us ({5 * flops / b / 1e6:5.0f} TF/s useful)", flush=True)
        if os.environ.get("FLASH_DBG"):
            dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
            lib.fs2_flash_attn_set_debug.argtypes = [L.C.c_void_p]
            lib.fs2_flash_attn_set_debug(L.C.c_void_p(dbg.data_ptr()))
            for _ in range(2):
                L.call("fs2_flash_attn_fwd", qkv[0], lens, B, H, T, D, sc, 0.1, 5, None, lse[0], O[0], 0)
            torch.cuda.synchronize()
            d = dbg.tolist()
            print(f"  fwd CTA0 softmax thread total {d[0]} cyc | pass1: wait-S {d[1]} ld {d[2]} max {d[3]} | pass2: wait-S {d[4]} ld {d[5]} "
                  f"math {d[6]} wait-Pbuf {d[7]} st+fence+arrive {d[8]} || MMA thread total {d[16]} over {d[21]} jobs: waits K {d[17]} "
                  f"S-free {d[18]} P-ready {d[19]} V {d[20]}")
            dbg.zero_()
            for _ in range(2):
                L.call("fs2_flash_attn_bwd", dO[0], O[0], qkv[0], lse[0], lens, B, H, T, D, sc, 0.1, 5, None, dvec, dqkv[0], 0)
            torch.cuda.synchronize()
            d = dbg.tolist()
            print(f"  dQ kernel CTA0 softmax thread total {d[0]}: wait-S {d[1]} ld {d[2]} math {d[3]} wait-dSbuf {d[4]} st {d[5]} || MMA "
                  f"total {d[16]} over {d[21]} jobs: waits K {d[17]} V {d[20]} S-free {d[18]} dS-ready {d[19]}")
            print(f"  dK/dV kernel CTA0 softmax thread (group 0) total {d[32]}: wait-S {d[33]} ld {d[34]} math {d[35]} st {d[36]} || MMA "
                  f"total {d[48]} over {d[52]} jobs: waits Q/dO {d[49]} buffer-free {d[50]} P-ready {d[51]}")
            lib.fs2_flash_attn_set_debug(None)
    print("flag", L.gemm_tc_error_flag())


if __name__ == "__main__":
    main()
