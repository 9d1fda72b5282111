"""GPU parity of the two GEMM kernels (SIMT exact path and tcgen05/TMA path) against a float64 torch
restatement of the Fs2Gemm descriptor (tests/gemm_ref.py).  Shapes are the ones the FastSpeech2 path
uses: Conv1d k=9/5/3 as shifted-row implicit GEMM, its dgrad and wgrad, and the batched attention GEMMs."""
import pytest
import torch

from gemm_ref import ref_gemm

pytestmark = pytest.mark.gpu


def _rand(shape, seed, dtype):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to("cuda").to(dtype)


def _run(lib, use_tc, ab_dtype, mode, M, N, K, taps, A, B, c_bf16=False, **kw):
    ab_bf16 = ab_dtype == torch.bfloat16
    Nout = N * taps if mode == 2 else N
    C = torch.full((M + kw.get("c_row_off", 0), Nout), float("nan"), device="cuda",
                   dtype=torch.bfloat16 if c_bf16 else torch.float32)
    if kw.get("split_k", 1) != 1 or kw.get("accumulate", 0):
        C.zero_()
    lib.gemm(mode=mode, M=M, N=N, K=K, taps=taps, A=A, lda=A.shape[1], a_rows=A.shape[0], a_inner=A.shape[1],
             B=B, ldb=B.shape[1], b_rows=B.shape[0], b_inner=B.shape[1], Cout=C, ldc=Nout, c_bf16=c_bf16,
             ab_bf16=ab_bf16, c_tap_stride=N, use_tc=use_tc, **kw)
    torch.cuda.synchronize()
    if use_tc:
        assert lib.gemm_tc_error_flag() == 0, "tcgen05 kernel reported an mbarrier timeout"
    return C


CASES = [
    # name, mode, M, N, K, taps, A rows, A inner, B rows, B inner, extra
    ("linear", 0, 300, 200, 384, 1, 300, 384, 200, 384, {}),
    ("qkv", 0, 1000, 1152, 384, 1, 1000, 384, 1152, 384, {}),
    ("conv9_fwd", 0, 392, 256, 128, 9, 400, 128, 256, 9 * 128, dict(a_tap_step=1, b_tap_step=128)),
    ("conv5_k80", 0, 392, 512, 80, 5, 400, 80, 512, 5 * 80, dict(a_row_off=2, a_tap_step=1, b_tap_step=80)),
    ("n80", 0, 392, 80, 512, 5, 400, 512, 80, 5 * 512, dict(a_row_off=2, a_tap_step=1, b_tap_step=512)),
    ("dgrad", 1, 400, 128, 256, 9, 400, 256, 256, 9 * 128, dict(a_row_off=4, a_tap_step=-1, b_tap_step=128)),
    ("dgrad_k80", 1, 400, 384, 80, 1, 400, 80, 80, 384, {}),
    ("wgrad", 2, 256, 128, 400, 9, 400, 256, 400, 128, dict(b_row_off=-4, b_tap_step=1)),
    ("wgrad_split", 2, 256, 128, 2000, 3, 2000, 256, 2000, 128, dict(b_row_off=-1, b_tap_step=1, split_k=4)),
    ("wgrad_m80", 2, 80, 384, 700, 1, 700, 80, 700, 384, {}),
    # shapes that exercise every tile configuration of the CTA-pair (cta_group::2) kernel, ragged M / N, several waves
    ("pair_n384_k512", 0, 1100, 384, 512, 1, 1100, 512, 384, 512, {}),
    ("pair_n1536_conv9", 0, 700, 1536, 128, 9, 708, 128, 1536, 9 * 128, dict(a_tap_step=1, b_tap_step=128)),
    ("pair_n1152", 0, 520, 1152, 384, 1, 520, 384, 1152, 384, {}),
    ("pair_waves", 0, 40000, 384, 128, 1, 40000, 128, 384, 128, {}),
    ("pair_dgrad_n384", 1, 900, 384, 320, 9, 908, 320, 320, 9 * 384, dict(a_row_off=4, a_tap_step=-1, b_tap_step=384)),
    ("pair_dgrad_n1536", 1, 600, 1536, 384, 1, 600, 384, 384, 1536, {}),
    ("pair_dgrad_n80", 1, 600, 80, 512, 5, 608, 512, 512, 5 * 80, dict(a_row_off=2, a_tap_step=-1, b_tap_step=80)),
    ("pair_wgrad_auto", 2, 384, 384, 3000, 9, 3000, 384, 3000, 384, dict(b_row_off=-4, b_tap_step=1, split_k=0)),
    ("pair_wgrad_n1536", 2, 384, 1536, 1500, 1, 1500, 384, 1500, 1536, dict(split_k=0)),
    ("pair_wgrad_m1536", 2, 1536, 384, 1500, 3, 1500, 1536, 1500, 384, dict(b_row_off=-1, b_tap_step=1, split_k=0)),
    ("pair_wgrad_n80", 2, 512, 80, 1300, 5, 1300, 512, 1300, 80, dict(b_row_off=-2, b_tap_step=1, split_k=0)),
    # dgrad with fewer tiles than SMs: split_k = 0 lets the library split the reduction (vector atomics into a C it zeroes)
    ("dgrad_autosplit", 1, 600, 384, 512, 9, 608, 512, 512, 9 * 384, dict(a_row_off=4, a_tap_step=-1, b_tap_step=384, split_k=0)),
    ("dgrad_autosplit_k1", 1, 1000, 384, 1152, 1, 1000, 1152, 1152, 384, dict(split_k=0)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("path", ["simt_f32", "simt_bf16", "tc"])
def test_gemm_modes(lib, case, path):
    name, mode, M, N, K, taps, ar, ai, br, bi, extra = case
    dt = torch.float32 if path == "simt_f32" else torch.bfloat16
    A = _rand((ar, ai), 1, dt)
    B = _rand((br, bi), 2, dt)
    C = _run(lib, path == "tc", dt, mode, M, N, K, taps, A, B, **extra)
    kw = {k: v for k, v in extra.items() if k in ("a_row_off", "a_tap_step", "b_row_off", "b_tap_step")}
    ref = ref_gemm(mode, M, N, K, taps, A.float(), B.float(), **kw)
    got = C.double().view(ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if path == "simt_f32" else 2e-4      # operands are exact in both; only fp32 accumulation order differs
    assert err <= tol * max(scale, 1.0), f"{name}/{path}: max err {err} (scale {scale})"


@pytest.mark.parametrize("path", ["simt_bf16", "tc"])
def test_gemm_epilogue_rowspace(lib, path):
    """bias + ReLU + per-item length mask + reflect-halo mirror writes + bf16 output, in the padded row space."""
    Bn, T, Cin, Cout, k = 3, 40, 128, 256, 3
    Tp = T + 8
    rows = Bn * Tp
    x = _rand((rows, Cin), 3, torch.bfloat16)
    w = _rand((Cout, k * Cin), 4, torch.bfloat16)
    bias = _rand((Cout,), 5, torch.float32)
    lens = torch.tensor([40, 17, 1], dtype=torch.int32, device="cuda")
    M = rows - 8
    C = torch.zeros(rows, Cout, device="cuda", dtype=torch.bfloat16)
    lib.gemm(mode=0, M=M, N=Cout, K=Cin, taps=k, A=x, lda=Cin, a_rows=rows, a_inner=Cin, a_row_off=3, a_tap_step=1,
             B=w, ldb=k * Cin, b_rows=Cout, b_inner=k * Cin, b_tap_step=Cin, Cout=C, ldc=Cout, c_bf16=True,
             ab_bf16=True, c_row_off=4, bias=bias, relu=1, rs_T=T, rs_Tp=Tp, lens=lens, halo=2,
             use_tc=(path == "tc"))
    torch.cuda.synchronize()
    ref = ref_gemm(0, M, Cout, Cin, k, x.float(), w.float(), a_row_off=3, a_tap_step=1, b_tap_step=Cin)
    ref = torch.relu(ref + bias.double()[None])
    full = torch.zeros(rows, Cout, dtype=torch.float64, device="cuda")
    full[4:4 + M] = ref
    exp = torch.zeros_like(full)
    for b in range(Bn):
        base = b * Tp + 4
        L = int(lens[b])
        exp[base:base + L] = full[base:base + L]
        for i in range(1, 3):           # reflect halo of width 2 at the rectangle edge
            exp[base - i] = exp[base + i]
            exp[base + T - 1 + i] = exp[base + T - 1 - i]
    err = (C.double() - exp).abs().max().item()
    assert err <= 2e-2 * max(exp.abs().max().item(), 1.0), f"{path}: {err}"   # bf16 output rounding


@pytest.mark.parametrize("path", ["simt_bf16", "tc"])
def test_gemm_attention_batched(lib, path):
    """QK^T (mode 0), PV (mode 1), dV = P^T dO (mode 2) batched over (head, item) with the strides the model uses."""
    Bn, H, T, D = 3, 2, 100, 192
    Tp = T + 8
    ld = 3 * H * D
    rows = Bn * Tp
    qkv = _rand((rows, ld), 6, torch.bfloat16)
    use_tc = path == "tc"
    ldk = (T + 7) // 8 * 8
    S = torch.zeros(Bn * H, T, ldk, device="cuda", dtype=torch.float32)
    # S[b,h] = Q K^T
    lib.gemm(mode=0, M=T, N=T, K=D, A=qkv, A_off=4 * ld, lda=ld, a_rows=T, a_inner=D, a_s1=D, a_s2=Tp * ld,
             B=qkv, B_off=4 * ld + H * D, ldb=ld, b_rows=T, b_inner=D, b_s1=D, b_s2=Tp * ld,
             batch1=H, batch2=Bn, Cout=S, ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=True,
             use_tc=use_tc)
    torch.cuda.synchronize()
    q3 = qkv.float().view(Bn, Tp, 3, H, D)[:, 4:4 + T]
    Q, K, V = q3[:, :, 0], q3[:, :, 1], q3[:, :, 2]          # (B,T,H,D)
    Sref = torch.einsum("bthd,bshd->bhts", Q.double(), K.double())
    err = (S.view(Bn, H, T, ldk)[..., :T].double() - Sref).abs().max().item()
    assert err <= 1e-3 * Sref.abs().max().item(), f"QK^T {path}: {err}"
    # O = P V
    P = torch.softmax(Sref.float() / D ** 0.5, -1)
    Pb = torch.zeros(Bn * H, T, ldk, device="cuda", dtype=torch.bfloat16)
    Pb.view(Bn, H, T, ldk)[..., :T] = P.to(torch.bfloat16)
    O = torch.zeros(rows, H * D, device="cuda", dtype=torch.bfloat16)
    lib.gemm(mode=1, M=T, N=D, K=T, A=Pb, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
             B=qkv, B_off=4 * ld + 2 * H * D, ldb=ld, b_rows=T, b_inner=D, b_s1=D, b_s2=Tp * ld,
             batch1=H, batch2=Bn, Cout=O, C_off=4 * H * D, ldc=H * D, c_s1=D, c_s2=Tp * H * D, c_bf16=True,
             ab_bf16=True, use_tc=use_tc)
    torch.cuda.synchronize()
    Pf = Pb.view(Bn, H, T, ldk)[..., :T].double()
    Oref = torch.einsum("bhts,bshd->bthd", Pf, V.double())
    got = O.view(Bn, Tp, H, D)[:, 4:4 + T].double()
    err = (got - Oref).abs().max().item()
    assert err <= 1e-2 * Oref.abs().max().item(), f"PV {path}: {err}"
    # dV = P^T dO   (written into the V slot of a dQKV buffer)
    dO = _rand((rows, H * D), 7, torch.bfloat16)
    dqkv = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
    lib.gemm(mode=2, M=T, N=D, K=T, A=Pb, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
             B=dO, B_off=4 * H * D, ldb=H * D, b_rows=T, b_inner=D, b_s1=D, b_s2=Tp * H * D,
             batch1=H, batch2=Bn, Cout=dqkv, C_off=4 * ld + 2 * H * D, ldc=ld, c_s1=D, c_s2=Tp * ld,
             c_tap_stride=0, c_bf16=True, ab_bf16=True, use_tc=use_tc)
    torch.cuda.synchronize()
    dOf = dO.view(Bn, Tp, H, D)[:, 4:4 + T].double()
    dVref = torch.einsum("bhts,bthd->bshd", Pf, dOf)
    got = dqkv.view(Bn, Tp, 3, H, D)[:, 4:4 + T, 2].double()
    err = (got - dVref).abs().max().item()
    assert err <= 1e-2 * dVref.abs().max().item(), f"dV {path}: {err}"
    if use_tc:
        assert lib.gemm_tc_error_flag() == 0


@pytest.mark.parametrize("M,rows", [(600, 608), (9000, 9008)])      # single-CTA tiles / CTA-pair 256 x 384 tiles
def test_dgrad_deterministic_split_k(lib, M, rows):
    """c_split_stride: every split of the reduction is stored (no atomics) to its own copy of C; the copies add up to
    the full product and two runs are bit-identical."""
    N, K, taps = 384, 512, 9
    A = _rand((rows, K), 1, torch.bfloat16)
    B = _rand((K, taps * N), 2, torch.bfloat16)
    ref = ref_gemm(1, M, N, K, taps, A.float(), B.float(), a_row_off=4, a_tap_step=-1, b_tap_step=N)
    outs = []
    for _ in range(2):
        C = torch.full((2, M, N), float("nan"), device="cuda")
        lib.gemm(mode=1, M=M, N=N, K=K, taps=taps, A=A, lda=K, a_rows=rows, a_inner=K, a_row_off=4, a_tap_step=-1,
                 B=B, ldb=taps * N, b_rows=K, b_inner=taps * N, b_tap_step=N, Cout=C, ldc=N, c_bf16=False, ab_bf16=True,
                 split_k=2, c_split_stride=M * N, use_tc=True)
        torch.cuda.synchronize()
        assert lib.gemm_tc_error_flag() == 0
        outs.append(C)
    assert torch.equal(outs[0], outs[1])                                      # no atomics: run-to-run identical
    got = outs[0].double().sum(0)
    assert (got - ref).abs().max().item() <= 2e-4 * max(ref.abs().max().item(), 1.0)
    assert outs[0][0].abs().max() > 0 and outs[0][1].abs().max() > 0           # both halves carry a partial sum


@pytest.mark.parametrize("name,M,N,K,taps,split", [
    ("ffn_conv9_fold", 1536, 384, 2100, 9, 0),      # 128 x 192 tiles: the bias gradient rides inside the GEMM (ones-tile MMA), auto split-K
    ("in_proj_fold", 1152, 384, 900, 1, 0),
    ("ragged_m_fold", 200, 384, 777, 1, 3),         # M not a multiple of 128, fixed split
    ("postnet_fallback", 512, 512, 1500, 5, 0),     # 256-wide tiles have no spare tensor-memory columns: separate column-sum kernel
    ("n80_fallback", 512, 80, 1300, 5, 0),
])
def test_wgrad_bias_gradient_column_sums(lib, name, M, N, K, taps, split):
    """Fs2Gemm.a_colsum: db[m] += sum_k dy[k, m] next to dW = dy^T x (mode 2) -- the bias gradient of a Conv1d / Linear
    (reference: autograd of speechbrain Conv1d / nn.Linear biases, train.py:80)."""
    p = (taps - 1) // 2
    dy = _rand((K, M), 11, torch.bfloat16)
    x = _rand((K, N), 12, torch.bfloat16)
    dW = torch.zeros(M, taps * N, device="cuda")
    db = torch.full((M,), 0.5, device="cuda")                      # accumulates (+=) like the flat gradient buffer
    lib.gemm(mode=2, M=M, N=N, K=K, taps=taps, A=dy, lda=M, a_rows=K, a_inner=M, B=x, ldb=N, b_rows=K, b_inner=N,
             b_row_off=-p, b_tap_step=1, Cout=dW, ldc=taps * N, c_tap_stride=N, c_col_stride=1, c_bf16=False, ab_bf16=True,
             accumulate=1, split_k=split, a_colsum=db)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    ref_db = 0.5 + dy.double().sum(0)
    assert (db.double() - ref_db).abs().max().item() <= 2e-4 * max(1.0, ref_db.abs().max().item()), name
    ref = ref_gemm(2, M, N, K, taps, dy.float(), x.float(), b_row_off=-p, b_tap_step=1)
    err = (dW.double().view(ref.shape) - ref).abs().max().item()
    assert err <= 2e-4 * max(ref.abs().max().item(), 1.0), name    # the weight gradient itself is unchanged
