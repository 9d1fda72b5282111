"""Sweep every tensor-core kernel / tile configuration over the GEMM shapes of the FastSpeech2 step
(B=32; T from the command line) and print microseconds per launch.  Run on the GPU box."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200")
L = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")
PAD = 4


def timeit(fn, iters=8, warm=2):
    """GPU time per launch: the launches are captured into one CUDA graph so that the host's launch rate
    (ctypes + tensor-map lookup, ~10 us) cannot bound the small GEMMs."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    B = int(os.environ.get("B", 32))
    Ts = [int(t) for t in (sys.argv[1:] or ["800"])]
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4).cuda()
    m.store.pack(True)
    m.store.ensure_grads()
    bf = torch.bfloat16
    R = 3
    dbg = torch.zeros(2, dtype=torch.int64, device="cuda")
    L.gemm_tc_set_debug(dbg)
    for T in Ts:
        rows = B * (T + 2 * PAD)
        buf = {}

        def X(c, dt=bf, n=R):
            key = (c, dt)
            if key not in buf:
                buf[key] = [torch.randn(rows, c, device="cuda").to(dt) for _ in range(n)]
            return buf[key]

        lay = lambda i: f"decoder.layers.{i % 6}"
        pn = lambda i: f"postnet.convs_intermedite.{i % 3}.conv"
        cases = [
            ("fwd conv9 384->1536", 2.0 * rows * 1536 * 384 * 9, lambda i: m._conv(X(384)[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", X(1536)[i % R], c_bf16=True, bias=m._P(lay(i) + ".pos_ffn.0.conv.bias"), relu=1)),
            ("fwd conv1 1536->384", 2.0 * rows * 1536 * 384, lambda i: m._conv(X(1536)[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", X(384, torch.float32)[i % R], c_bf16=False, bias=m._P(lay(i) + ".pos_ffn.2.conv.bias"))),
            ("fwd qkv 384->1152", 2.0 * rows * 1152 * 384, lambda i: m._conv(X(384)[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", X(1152)[i % R], c_bf16=True)),
            ("fwd out_proj 384->384", 2.0 * rows * 384 * 384, lambda i: m._conv(X(384)[i % R], B, T, lay(i) + ".self_att.att.out_proj.weight", X(384, torch.float32)[i % R], c_bf16=False)),
            ("fwd pred conv3 384->384", 2.0 * rows * 384 * 384 * 3, lambda i: m._conv(X(384)[i % R], B, T, "durPred.conv1.conv.weight", X(384, torch.float32)[i % R], c_bf16=False)),
            ("fwd postnet conv5 512->512", 2.0 * rows * 512 * 512 * 5, lambda i: m._conv(X(512)[i % R], B, T, pn(i) + ".weight", X(512)[(i + 1) % R], c_bf16=True, halo=2)),
            ("fwd mel linear 384->80", 2.0 * rows * 384 * 80, lambda i: m._conv(X(384)[i % R], B, T, "linear.w.weight", X(80, torch.float32)[i % R], c_bf16=False)),
            ("fwd postnet conv5 80->512", 2.0 * rows * 80 * 512 * 5, lambda i: m._conv(X(80)[i % R], B, T, "postnet.conv_pre.conv.weight", X(512, torch.float32)[i % R], c_bf16=False)),
            ("dgrad conv9 1536->384", 2.0 * rows * 1536 * 384 * 9, lambda i: m._conv_dgrad(X(1536)[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", X(384, torch.float32)[i % R])),
            ("dgrad conv1 384->1536", 2.0 * rows * 1536 * 384, lambda i: m._conv_dgrad(X(384)[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", X(1536)[i % R], c_bf16=True, relu_aux=X(1536)[(i + 1) % R])),
            ("dgrad qkv 1152->384", 2.0 * rows * 1152 * 384, lambda i: m._conv_dgrad(X(1152)[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", X(384, torch.float32)[i % R])),
            ("dgrad out_proj 384->384", 2.0 * rows * 384 * 384, lambda i: m._conv_dgrad(X(384)[i % R], B, T, lay(i) + ".self_att.att.out_proj.weight", X(384)[(i + 1) % R], c_bf16=True)),
            ("dgrad postnet conv5 512", 2.0 * rows * 512 * 512 * 5, lambda i: m._conv_dgrad(X(512)[i % R], B, T, pn(i) + ".weight", X(512, torch.float32, 2)[i % 2])),
            ("wgrad conv9", 2.0 * rows * 1536 * 384 * 9, lambda i: m._conv_wgrad(X(1536)[i % R], X(384)[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", lay(i) + ".pos_ffn.0.conv.weight")),
            ("wgrad conv1", 2.0 * rows * 1536 * 384, lambda i: m._conv_wgrad(X(384)[i % R], X(1536)[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", lay(i) + ".pos_ffn.2.conv.weight")),
            ("wgrad qkv", 2.0 * rows * 1152 * 384, lambda i: m._conv_wgrad(X(1152)[i % R], X(384)[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", lay(i) + ".self_att.att.in_proj_weight")),
            ("wgrad out_proj", 2.0 * rows * 384 * 384, lambda i: m._conv_wgrad(X(384)[i % R], X(384)[(i + 1) % R], B, T, lay(i) + ".self_att.att.out_proj.weight", lay(i) + ".self_att.att.out_proj.weight")),
            ("wgrad postnet conv5 512", 2.0 * rows * 512 * 512 * 5, lambda i: m._conv_wgrad(X(512)[i % R], X(512)[(i + 1) % R], B, T, pn(i) + ".weight", pn(i) + ".weight")),
        ]
        if os.environ.get("ISOLATE"):
            # pipeline isolation of the CTA-pair kernel: full kernel / MMA issue only / TMA only
            print(f"T={T}: us per launch: full | MMA-only | TMA-only   for pair/256, pair/384, single-CTA/256, single-CTA/384")
            for name, flops, fn in cases:
                line = f"{name:30s}"
                for kind in (1, 3):
                    for cfg in (0, 1):
                        for mode in (0, 1, 2):
                            L.gemm_tc_tune(kind | (mode << 4), cfg)
                            line += f"{timeit(fn):8.1f}"
                        line += "  |"
                print(line, flush=True)
            L.gemm_tc_tune(2, -1)
            continue
        configs = [("1cta/256", 0, 0), ("1cta/192", 0, 1), ("1cta/128", 0, 2), ("pair/256", 1, 0), ("pair/384", 1, 1), ("pair/128", 1, 2),
                   ("heuristic", 2, -1)]
        print(f"T={T} rows={rows}   us per launch (TFLOP/s)")
        print(f"{'':30s}" + "".join(f"{c[0]:>17s}" for c in configs))
        for name, flops, fn in cases:
            line = f"{name:30s}"
            clk = []
            for _, pair, cfg in configs:
                L.gemm_tc_tune(pair, cfg)
                try:
                    us = timeit(fn)
                    line += f"{us:9.1f} ({flops / us / 1e6:5.0f})"
                    if pair == 1:
                        torch.cuda.synchronize()
                        c, ns = dbg.tolist()
                        clk.append(f"{c / max(ns, 1):.2f}")
                except RuntimeError as e:
                    line += f"{'err':>17s}"
            print(line + "   pair-kernel SM GHz: " + "/".join(clk), flush=True)
        L.gemm_tc_tune(2, -1)
    print("flag", L.gemm_tc_error_flag())


if __name__ == "__main__":
    main()
