"""Where does a short-K GEMM launch spend its time?  Times the same (N, K) GEMM at M = r * 148 tiles-worth of rows
for r = 1..8 rounds of the persistent tile loop: the slope is the per-tile-round cost, the intercept the per-launch
cost (launch, TMEM alloc, pipeline fill, tail).
GEMM_DBG=3|4 (no epilogue stores / staging without global stores) and GEMM_PROBE=1 (cycle probe of CTA 0's TMA thread, MMA
thread and first epilogue warp) need a library built with the measurement code: `make -C .../csrc clean all EXTRA=-DFS2_TC_PROBE`
(the product build compiles it out -- the GEMM kernels are sensitive to their instruction footprint)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")


def timeit(fn, iters=30, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def run(name, N, K, c_bf16, bn, graph):
    bf = torch.bfloat16
    w = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
    out = []
    for r in (1, 2, 3, 4, 6, 8, 16):
        tiles_n = (N + bn - 1) // bn
        M = 128 * (148 * r // tiles_n)
        R = 4
        xs = [torch.randn(M, K, device="cuda").to(bf) for _ in range(R)]
        os_ = [torch.empty(M, N, device="cuda", dtype=bf if c_bf16 else torch.float32) for _ in range(R)]

        def one(i):
            L.gemm(mode=0, M=M, N=N, K=K, A=xs[i % R], lda=K, a_rows=M, a_inner=K, B=w, ldb=K, b_rows=N, b_inner=K,
                   Cout=os_[i % R], ldc=N, c_bf16=c_bf16, ab_bf16=True)
        if graph:
            one(0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(20):
                    one(i)
            us = timeit(lambda i: g.replay(), iters=10, warm=2) / 20
        else:
            us = timeit(one)
        out.append((r, M, us))
    (r0, _, t0), (r1, _, t1) = out[0], out[-1]
    slope = (t1 - t0) / (r1 - r0)
    print(f"{name:34s} " + "  ".join(f"r={r}: {us:6.1f}" for r, _, us in out) + f"   | per round {slope:5.2f} us, intercept {t0 - slope * r0:5.2f} us",
          flush=True)


if os.environ.get("GEMM_DBG"):
    L.gemm_tc_tune(pair=2 | (int(os.environ["GEMM_DBG"]) << 4), cfg=-1)
    print("dbg_mode", os.environ["GEMM_DBG"], "(results are garbage, timings only)")
for graph in (True,):
    print("CUDA-graph replay of 20 launches" if graph else "eager launches from Python")
    run("out_proj 384->384 fp32 out", 384, 384, False, 192, graph)
    run("qkv 384->1152 bf16 out", 1152, 384, True, 192, graph)
    run("conv1 1536->384 fp32 out", 384, 1536, False, 192, graph)
    run("ffn-in 384->1536 bf16 out (k=1)", 1536, 384, True, 256, graph)
print("flag", L.gemm_tc_error_flag())

# ---- role probe of CTA 0 (single-CTA kernel): who waits on whom (needs a library built with `make EXTRA=-DFS2_TC_PROBE`)
if os.environ.get("GEMM_PROBE"):
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.gemm_tc_set_debug(dbg)
    bf = torch.bfloat16
    for name, N, K, c_bf16, bn in (("out_proj 384->384 fp32", 384, 384, False, 192), ("qkv 384->1152 bf16", 1152, 384, True, 192),
                                   ("conv1 1536->384 fp32", 384, 1536, False, 192)):
        r = 8
        M = 128 * (148 * r // ((N + bn - 1) // bn))
        w = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
        x = torch.randn(M, K, device="cuda").to(bf)
        o = torch.empty(M, N, device="cuda", dtype=bf if c_bf16 else torch.float32)
        for _ in range(3):
            L.gemm(mode=0, M=M, N=N, K=K, A=x, lda=K, a_rows=M, a_inner=K, B=w, ldb=K, b_rows=N, b_inner=K, Cout=o, ldc=N,
                   c_bf16=c_bf16, ab_bf16=True)
        torch.cuda.synchronize()
        d = dbg.tolist()
        print(f"{name}: tiles/CTA {d[7]} | TMA thread total {d[2]} cyc, waiting for a free stage {d[3]} | MMA thread total {d[4]}, "
              f"waiting for accumulator {d[5]}, for operands {d[6]} | epilogue warp total {d[8]}, waiting for MMA {d[9]}, "
              f"in chunk loop {d[10]} (of which tcgen05.ld + wait {d[11]})")
    L.gemm_tc_set_debug(None)
