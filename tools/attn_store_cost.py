"""How much of the fused attention forward is the P / Pd global stores?  Same launch with and without the two outputs."""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
L = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")
from gemm_sweep import timeit  # noqa: E402

B, H, HD, PAD = 32, 2, 192, 4
D = H * HD
for T in (800, 488):
    TP, ldk, ld = T + 2 * PAD, (T + 7) // 8 * 8, 3 * D
    qkv = [(torch.randn(B * TP, ld, device="cuda") * 0.7).to(torch.bfloat16) for _ in range(3)]
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    P = [torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    Pd = [torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    O = [torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    sc = 1.0 / math.sqrt(HD)
    full = timeit(lambda i: L.call("fs2_attn_fwd", qkv[i % 3], lens, B, H, T, D, ldk, sc, 0.1, 5, None, P[i % 3], Pd[i % 3], O[i % 3]))
    only_p = timeit(lambda i: L.call("fs2_attn_fwd", qkv[i % 3], lens, B, H, T, D, ldk, sc, 0.1, 5, None, P[i % 3], None, O[i % 3]))
    none = timeit(lambda i: L.call("fs2_attn_fwd", qkv[i % 3], lens, B, H, T, D, ldk, sc, 0.1, 5, None, None, None, O[i % 3]))
    print(f"T={T}: forward with P+Pd stores {full:.1f} us, P only {only_p:.1f} us, no stores (dropout math kept) {none:.1f} us", flush=True)
