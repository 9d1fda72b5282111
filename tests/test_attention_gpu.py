"""GPU parity of the fused tcgen05 attention kernels (fs2_attn_fwd / fs2_attn_bwd) against a float64 torch
restatement of nn.MultiheadAttention's math path with the reference's mask quirk (model.py:338-343),
and against the unfused kernels (QK^T GEMM + fs2_softmax_fwd + PV GEMM) whose dropout stream they share."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
PAD = 4
HD = 192


def _setup(B, H, T, lens, seed):
    g = torch.Generator().manual_seed(seed)
    D = H * HD
    TP = T + 2 * PAD
    qkv = (torch.randn(B * TP, 3 * D, generator=g) * 0.7).cuda().to(torch.bfloat16)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    ldk = (T + 7) // 8 * 8
    return qkv, lens_t, D, TP, ldk


def _ref(qkv, lens, B, H, T, D, TP):
    """float64 attention per (b,h) with kv = min(len[b], len[(b*H+h) % B]) valid keys.  Returns P (B*H,T,T), O (B,T,D)."""
    x = qkv.double().view(B, TP, 3 * D)[:, PAD:PAD + T]
    P = torch.zeros(B * H, T, T, dtype=torch.float64, device="cuda")
    O = torch.zeros(B, T, D, dtype=torch.float64, device="cuda")
    for b in range(B):
        for h in range(H):
            kv = min(lens[b], lens[(b * H + h) % B])
            q = x[b, :, h * HD:(h + 1) * HD]
            k = x[b, :kv, D + h * HD:D + (h + 1) * HD]
            v = x[b, :kv, 2 * D + h * HD:2 * D + (h + 1) * HD]
            p = torch.softmax(q @ k.t() / math.sqrt(HD), -1)
            P[b * H + h, :, :kv] = p
            O[b, :, h * HD:(h + 1) * HD] = p @ v
    return P, O


CASES = [(2, 2, 40, [40, 23]), (3, 2, 128, [128, 77, 5]), (2, 2, 333, [333, 200]), (4, 2, 800, [800, 640, 411, 64])]


@pytest.mark.parametrize("B,H,T,lens", CASES)
def test_fused_attention_forward(lib, B, H, T, lens):
    qkv, lens_t, D, TP, ldk = _setup(B, H, T, lens, 1)
    P = torch.full((B * H, T, ldk), float("nan"), device="cuda", dtype=torch.bfloat16)
    O = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_attn_fwd", qkv, lens_t, B, H, T, D, ldk, 1.0 / math.sqrt(HD), 0.0, 0, None, P, None, O)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    Pr, Or = _ref(qkv, lens, B, H, T, D, TP)
    assert not torch.isnan(P.float()).any(), "every column < ldk of P must be written"
    assert (P[:, :, T:].float() == 0).all()
    err_p = (P[:, :, :T].double() - Pr).abs().max().item()
    assert err_p < 4e-3, err_p                      # bf16 rounding of probabilities <= 1
    got = O.double().view(B, TP, D)[:, PAD:PAD + T]
    err_o = (got - Or).abs().max().item()
    assert err_o < 2e-2 * max(Or.abs().max().item(), 1.0), err_o     # bf16 P and bf16 output rounding
    halo = O.view(B, TP, D)
    assert (halo[:, :PAD] == 0).all() and (halo[:, PAD + T:] == 0).all(), "halo rows must not be touched"


@pytest.mark.parametrize("B,H,T,lens", CASES[:3])
def test_fused_attention_matches_unfused_with_dropout(lib, B, H, T, lens):
    """Same dropout stream as fs2_softmax_fwd: Pd of the fused kernel == Pd of the unfused path (up to bf16 ulps of P),
    including which elements are dropped."""
    qkv, lens_t, D, TP, ldk = _setup(B, H, T, lens, 2)
    scale, p_drop, seed = 1.0 / math.sqrt(HD), 0.1, 0xABCDEF
    ld = 3 * D
    S = torch.zeros(B * H, T, ldk, device="cuda")
    lib.gemm(mode=0, M=T, N=T, K=HD, A=qkv, A_off=PAD * ld, lda=ld, a_rows=T, a_inner=HD, a_s1=HD, a_s2=TP * ld,
             B=qkv, B_off=PAD * ld + D, ldb=ld, b_rows=T, b_inner=HD, b_s1=HD, b_s2=TP * ld, batch1=H, batch2=B,
             Cout=S, ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=True)
    P0 = torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16)
    Pd0 = torch.zeros_like(P0)
    lib.call("fs2_softmax_fwd", S, lens_t, B, H, T, ldk, scale, p_drop, seed, None, P0, Pd0, 1)
    O0 = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lib.gemm(mode=1, M=T, N=HD, K=T, A=Pd0, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
             B=qkv, B_off=PAD * ld + 2 * D, ldb=ld, b_rows=T, b_inner=HD, b_s1=HD, b_s2=TP * ld, batch1=H, batch2=B,
             Cout=O0, C_off=PAD * D, ldc=D, c_s1=HD, c_s2=TP * D, c_bf16=True, ab_bf16=True)
    P1 = torch.zeros_like(P0)
    Pd1 = torch.zeros_like(P0)
    O1 = torch.zeros_like(O0)
    lib.call("fs2_attn_fwd", qkv, lens_t, B, H, T, D, ldk, scale, p_drop, seed, None, P1, Pd1, O1)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    assert (P0.float() - P1.float()).abs().max().item() < 4e-3
    kept0, kept1 = Pd0 != 0, Pd1 != 0
    differ = (kept0 != kept1) & (P0.float() > 1e-6)          # tiny probabilities may round to 0 in one path only
    assert not differ.any(), "dropout masks differ between the fused and the unfused kernels"
    assert (Pd0.float() - Pd1.float()).abs().max().item() < 6e-3
    assert (O0.float() - O1.float()).abs().max().item() < 3e-2 * max(O0.float().abs().max().item(), 1.0)


@pytest.mark.parametrize("B,H,T,lens", CASES)
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_fused_attention_backward(lib, B, H, T, lens, p_drop):
    """dS and dQ of fs2_attn_bwd against float64 autograd through the reference softmax (same dropout mask, taken
    from the forward kernel's Pd)."""
    qkv, lens_t, D, TP, ldk = _setup(B, H, T, lens, 3)
    scale, seed = 1.0 / math.sqrt(HD), 77
    P = torch.zeros(B * H, T, ldk, device="cuda", dtype=torch.bfloat16)
    Pd = torch.zeros_like(P)
    O = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_attn_fwd", qkv, lens_t, B, H, T, D, ldk, scale, p_drop, seed, None, P, Pd if p_drop > 0 else None, O)
    g = torch.Generator().manual_seed(9)
    dO = torch.zeros(B, TP, D)
    dO[:, PAD:PAD + T] = torch.randn(B, T, D, generator=g) * 0.5
    dO = dO.view(B * TP, D).cuda().to(torch.bfloat16)
    dS = torch.full((B * H, T, ldk), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv = torch.zeros(B * TP, 3 * D, device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_attn_bwd", dO, O, qkv, P, lens_t, B, H, T, D, ldk, scale, p_drop, seed, None, dS, dqkv)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    assert not torch.isnan(dS.float()).any()
    # float64 reference: keep-mask recovered from the forward kernel (Pd != 0 where P != 0)
    x = qkv.double().view(B, TP, 3 * D)[:, PAD:PAD + T]
    g_o = dO.double().view(B, TP, D)[:, PAD:PAD + T]
    max_ds = max_dq = 0.0
    ref_ds = ref_dq = 0.0
    for b in range(B):
        for h in range(H):
            kv = min(lens[b], lens[(b * H + h) % B])
            q = x[b, :, h * HD:(h + 1) * HD].clone().requires_grad_(True)
            k = x[b, :kv, D + h * HD:D + (h + 1) * HD]
            v = x[b, :kv, 2 * D + h * HD:2 * D + (h + 1) * HD]
            s = (q @ k.t() * scale)
            s.retain_grad()
            p = torch.softmax(s, -1)
            if p_drop > 0:
                keep = (Pd[b * H + h, :, :kv] != 0) | (P[b * H + h, :, :kv] == 0)
                pd = p * keep.double() / (1 - p_drop)
            else:
                pd = p
            o = pd @ v
            (o * g_o[b, :, h * HD:(h + 1) * HD]).sum().backward()
            ds_ref = s.grad * scale            # kernel's dS includes the 1/sqrt(d) factor: dQ = dS K
            got_ds = dS[b * H + h, :, :kv].double()
            max_ds = max(max_ds, (got_ds - ds_ref).abs().max().item())
            ref_ds = max(ref_ds, ds_ref.abs().max().item())
            got_dq = dqkv.double().view(B, TP, 3 * D)[b, PAD:PAD + T, h * HD:(h + 1) * HD]
            max_dq = max(max_dq, (got_dq - q.grad).abs().max().item())
            ref_dq = max(ref_dq, q.grad.abs().max().item())
            assert (dS[b * H + h, :, kv:].float() == 0).all()
    assert max_ds < 3e-2 * ref_ds, (max_ds, ref_ds)      # bf16 P, bf16 O in the row term, bf16 dS
    assert max_dq < 3e-2 * ref_dq, (max_dq, ref_dq)
