"""Micro-benchmark of the GEMM launches of one decoder FFT block + PostNet conv at BASELINE cfg-3 size
(B=32, Tm=800): TFLOP/s per launch, CUDA events, rotating buffers.  Run on the GPU box."""
import importlib
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200")
L = importlib.import_module("fine-grained-emotional-control-of-tts_b200._lib")
PAD = 4


def timeit(fn, iters=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B, T = int(os.environ.get("B", 32)), int(os.environ.get("T", 800))
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4).cuda()
    m.store.pack(True)
    m.store.ensure_grads()
    D, Fd, H, E = 384, 1536, 2, 512
    rows = B * (T + 2 * PAD)
    bf = torch.bfloat16
    R = 3
    xD = [torch.randn(rows, D, device="cuda").to(bf) for _ in range(R)]
    xF = [torch.randn(rows, Fd, device="cuda").to(bf) for _ in range(R)]
    xE = [torch.randn(rows, E, device="cuda").to(bf) for _ in range(R)]
    oD32 = [torch.empty(rows, D, device="cuda") for _ in range(R)]
    oF = [torch.empty(rows, Fd, device="cuda", dtype=bf) for _ in range(R)]
    oE = [torch.empty(rows, E, device="cuda", dtype=bf) for _ in range(R)]
    o32E = [torch.empty(rows, E, device="cuda") for _ in range(2)]
    qkv = [torch.randn(rows, 3 * D, device="cuda").to(bf) for _ in range(R)]
    ldk = (T + 7) // 8 * 8
    S = [torch.empty(B * H, T, ldk, device="cuda") for _ in range(2)]
    P = [torch.rand(B * H, T, ldk, device="cuda").to(bf) for _ in range(2)]
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    lay = lambda i: f"decoder.layers.{i % 6}"
    res = []

    only = os.environ.get("ONLY")

    def rec(name, flops, fn):
        if only and only not in name:
            return
        ms = timeit(fn, iters=int(os.environ.get("ITERS", 10)), warm=int(os.environ.get("WARM", 3)))
        res.append((name, flops, ms))
        print(f"{name:34s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)

    f9 = 2.0 * rows * Fd * D * 9
    f1 = 2.0 * rows * Fd * D
    rec("fwd conv9 384->1536 relu bf16", f9, (lambda i: m._conv(xD[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", oF[i % R], c_bf16=True, bias=m._P(lay(i) + ".pos_ffn.0.conv.bias"), relu=1)))
    rec("fwd conv1 1536->384 f32", f1, (lambda i: m._conv(xF[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", oD32[i % R], c_bf16=False, bias=m._P(lay(i) + ".pos_ffn.2.conv.bias"))))
    rec("fwd qkv 384->1152 bf16", 2.0 * rows * 3 * D * D, (lambda i: m._conv(xD[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", qkv[i % R], c_bf16=True)))
    rec("fwd out_proj 384->384 f32", 2.0 * rows * D * D, (lambda i: m._conv(xD[i % R], B, T, lay(i) + ".self_att.att.out_proj.weight", oD32[i % R], c_bf16=False)))
    rec("dgrad conv9 1536->384 f32", f9, (lambda i: m._conv_dgrad(xF[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", oD32[i % R])))
    # experiment: the same dgrad with a K-major ([Cin][tap][Cout]) weight copy, i.e. through the forward-mode kernel
    wT = torch.randn(D, 9 * Fd, device="cuda").to(bf)
    oS = [torch.empty(2, rows, D, device="cuda") for _ in range(2)]
    for split in (1, 2):
        rec(f"dgrad conv9 K-major weights split{split}", f9, (lambda i: L.gemm(
            mode=0, M=rows, N=D, K=Fd, taps=9, A=xF[i % R], lda=Fd, a_rows=rows, a_inner=Fd, a_row_off=4, a_tap_step=-1,
            B=wT, ldb=9 * Fd, b_rows=D, b_inner=9 * Fd, b_tap_step=Fd, Cout=oS[i % 2], ldc=D, c_bf16=False, ab_bf16=True,
            split_k=split, c_split_stride=rows * D if split > 1 else 0)))
        rec(f"dgrad conv9 MN-major (model) split{split}", f9, (lambda i: m._conv_dgrad(xF[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", oS[i % 2] if split > 1 else oD32[i % R], split=split)))
    rec("dgrad conv1 384->1536 bf16 relu_aux", f1, (lambda i: m._conv_dgrad(xD[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", oF[i % R], c_bf16=True, relu_aux=xF[(i + 1) % R])))
    rec("dgrad qkv 1152->384 f32", 2.0 * rows * 3 * D * D, (lambda i: m._conv_dgrad(qkv[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", oD32[i % R])))
    rec("wgrad conv9 (+colsum)", f9, (lambda i: m._conv_wgrad(xF[i % R], xD[i % R], B, T, lay(i) + ".pos_ffn.0.conv.weight", lay(i) + ".pos_ffn.0.conv.weight", lay(i) + ".pos_ffn.0.conv.bias")))
    rec("wgrad conv1 (+colsum)", f1, (lambda i: m._conv_wgrad(xD[i % R], xF[i % R], B, T, lay(i) + ".pos_ffn.2.conv.weight", lay(i) + ".pos_ffn.2.conv.weight", lay(i) + ".pos_ffn.2.conv.bias")))
    rec("wgrad qkv (+colsum)", 2.0 * rows * 3 * D * D, (lambda i: m._conv_wgrad(qkv[i % R], xD[i % R], B, T, lay(i) + ".self_att.att.in_proj_weight", lay(i) + ".self_att.att.in_proj_weight", lay(i) + ".self_att.att.in_proj_bias")))
    fa = 2.0 * B * H * T * T * (D // H)
    hd, TP, ld = D // H, T + 2 * PAD, 3 * D
    rec("attn QK^T", fa, (lambda i: L.gemm(mode=0, M=T, N=T, K=hd, A=qkv[i % R], A_off=PAD * ld, lda=ld, a_rows=T, a_inner=hd, a_s1=hd, a_s2=TP * ld, B=qkv[i % R], B_off=PAD * ld + D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B, Cout=S[i % 2], ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=True)))
    rec("attn softmax fwd (bytes GB/s)", 0.0, (lambda i: L.call("fs2_softmax_fwd", S[i % 2], lens, B, H, T, ldk, 0.07, 0.0, 0, None, P[i % 2], None, 1)))
    rec("attn PV", fa, (lambda i: L.gemm(mode=1, M=T, N=hd, K=T, A=P[i % 2], lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk, B=qkv[i % R], B_off=PAD * ld + 2 * D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B, Cout=xD[i % R], C_off=PAD * D, ldc=D, c_s1=hd, c_s2=TP * D, c_bf16=True, ab_bf16=True)))
    rec("attn dV = P^T dO", fa, (lambda i: L.gemm(mode=2, M=T, N=hd, K=T, A=P[i % 2], lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk, B=xD[i % R], B_off=PAD * D, ldb=D, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * D, batch1=H, batch2=B, Cout=qkv[i % R], C_off=PAD * ld + 2 * D, ldc=ld, c_s1=hd, c_s2=TP * ld, c_bf16=True, ab_bf16=True)))
    fp = 2.0 * rows * E * E * 5
    rec("postnet conv5 512->512 bf16", fp, (lambda i: m._conv(xE[i % R], B, T, f"postnet.convs_intermedite.{i % 3}.conv.weight", oE[i % R], c_bf16=True, halo=2)))
    rec("postnet dgrad conv5", fp, (lambda i: m._conv_dgrad(xE[i % R], B, T, f"postnet.convs_intermedite.{i % 3}.conv.weight", o32E[i % 2])))
    rec("postnet wgrad conv5", fp, (lambda i: m._conv_wgrad(xE[i % R], xE[(i + 1) % R], B, T, f"postnet.convs_intermedite.{i % 3}.conv.weight", f"postnet.convs_intermedite.{i % 3}.conv.weight", f"postnet.convs_intermedite.{i % 3}.conv.bias")))
    tot = sum(r[2] for r in res)
    print("flag", L.gemm_tc_error_flag())


if __name__ == "__main__":
    main()
