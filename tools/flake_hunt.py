"""Repeat one GEMM test case many times and characterise any mismatch (which elements, how large)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gemm_ref import ref_gemm  # noqa: E402
import test_gemm_gpu as T  # noqa: E402

lib = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["wgrad_split"]
paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["simt_bf16"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
for case in T.CASES:
    name, mode, M, N, K, taps, ar, ai, br, bi, extra = case
    if name not in names:
        continue
    for path in paths:
        dt = torch.float32 if path == "simt_f32" else torch.bfloat16
        A = T._rand((ar, ai), 1, dt)
        B = T._rand((br, bi), 2, dt)
        kw = {k: v for k, v in extra.items() if k in ("a_row_off", "a_tap_step", "b_row_off", "b_tap_step")}
        ref = ref_gemm(mode, M, N, K, taps, A.float(), B.float(), **kw)
        bad = 0
        for it in range(reps):
            # churn the allocator / leave other work in flight like a test session does
            junk = torch.randn(1 << (10 + it % 12), device="cuda")
            C = T._run(lib, path == "tc", dt, mode, M, N, K, taps, A, B, **extra)
            d = (C.double().view(ref.shape) - ref).abs()
            err = d.max().item()
            if err > 2e-4 * max(ref.abs().max().item(), 1.0):
                bad += 1
                idx = (d > 1e-3).nonzero()
                print(f"{name}/{path} it={it}: err {err:.4f}, {idx.shape[0]} bad elements, first {idx[:6].tolist()}, "
                      f"nan={torch.isnan(C).sum().item()}", flush=True)
            del junk
        print(f"{name}/{path}: {bad}/{reps} bad", flush=True)
