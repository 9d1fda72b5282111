"""GPU parity of the flash-style tcgen05 attention kernels (fs2_flash_attn_fwd / fs2_flash_attn_bwd) against a float64
torch restatement of nn.MultiheadAttention's math path with the reference's mask quirk (model.py:338-347, 414-427;
SURVEY Q1).  Dropout: the kernels' keep mask is read back with fs2_flash_attn_mask and applied to the float64 reference,
and its statistics (rate, independence across rows / columns / heads / seeds) are checked on their own."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
PAD = 4
HD = 192
LOG2E = 1.4426950408889634


def _setup(B, H, T, seed, amp=0.7):
    g = torch.Generator().manual_seed(seed)
    D = H * HD
    TP = T + 2 * PAD
    qkv = (torch.randn(B * TP, 3 * D, generator=g) * amp).cuda().to(torch.bfloat16)
    return qkv, D, TP


def _lse_len(lib, T):
    return int(lib.load().fs2_flash_attn_lse_len(ctypes.c_int(T)))


def _mask(lib, B, H, T, p_drop, seed):
    keep = torch.zeros(B * H, T, T, dtype=torch.uint8, device="cuda")
    lib.call("fs2_flash_attn_mask", B * H, T, p_drop, seed, None, keep)
    return keep.bool()


def _kv(lens, B, H, b, h, plain):
    return lens[b] if plain else min(lens[b], lens[(b * H + h) % B])


def _ref_fwd(qkv, lens, B, H, T, D, TP, keep=None, p_drop=0.0, plain=False):
    x = qkv.double().view(B, TP, 3 * D)[:, PAD:PAD + T]
    O = torch.zeros(B, T, D, dtype=torch.float64, device="cuda")
    lse = torch.zeros(B * H, T, dtype=torch.float64, device="cuda")
    for b in range(B):
        for h in range(H):
            kv = _kv(lens, B, H, b, h, plain)
            if kv == 0:
                continue
            q = x[b, :, h * HD:(h + 1) * HD]
            k = x[b, :kv, D + h * HD:D + (h + 1) * HD]
            v = x[b, :kv, 2 * D + h * HD:2 * D + (h + 1) * HD]
            s = q @ k.t() / math.sqrt(HD)
            lse[b * H + h] = torch.logsumexp(s, -1) * LOG2E
            p = torch.softmax(s, -1)
            if keep is not None:
                p = p * keep[b * H + h, :, :kv].double() / (1 - p_drop)
            O[b, :, h * HD:(h + 1) * HD] = p @ v
    return O, lse


CASES = [(2, 2, 40, [40, 23]), (3, 2, 128, [128, 77, 5]), (2, 2, 333, [333, 200]), (4, 2, 800, [800, 640, 411, 64]),
         (2, 2, 129, [129, 64]), (3, 2, 200, [200, 0, 97])]


@pytest.mark.parametrize("p_in_tmem", [1, 0])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
@pytest.mark.parametrize("B,H,T,lens", CASES)
def test_flash_forward(lib, B, H, T, lens, p_drop, p_in_tmem):
    qkv, D, TP = _setup(B, H, T, 1)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    Tl = _lse_len(lib, T)
    O = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H, Tl, device="cuda")
    seed = 0xABCDEF
    lib.load().fs2_flash_attn_tune(ctypes.c_int(p_in_tmem))
    try:
        lib.call("fs2_flash_attn_fwd", qkv, lens_t, B, H, T, D, 1.0 / math.sqrt(HD), p_drop, seed, None, lse, O, 0)
        torch.cuda.synchronize()
    finally:
        lib.load().fs2_flash_attn_tune(ctypes.c_int(1))
    assert lib.gemm_tc_error_flag() == 0
    keep = _mask(lib, B, H, T, p_drop, seed) if p_drop > 0 else None
    Or, lr = _ref_fwd(qkv, lens, B, H, T, D, TP, keep, p_drop)
    got = O.double().view(B, TP, D)[:, PAD:PAD + T]
    err_o = (got - Or).abs().max().item()
    assert err_o < 2e-2 * max(Or.abs().max().item(), 1.0), err_o     # bf16 probabilities and bf16 output rounding
    live = torch.tensor([[_kv(lens, B, H, b, h, False) > 0 for h in range(H)] for b in range(B)], device="cuda").view(-1)
    err_l = (lse[:, :T].double() - lr)[live].abs().max().item()
    assert err_l < 2e-3, err_l
    halo = O.view(B, TP, D)
    assert (halo[:, :PAD] == 0).all() and (halo[:, PAD + T:] == 0).all(), "halo rows must not be touched"


def test_flash_forward_plain_mask_no_lse(lib):
    """The intensity extractor's call: ordinary key-padding mask, inference (no statistics kept)."""
    B, H, T, lens = 3, 2, 150, [150, 90, 31]
    qkv, D, TP = _setup(B, H, T, 5)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    O = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_flash_attn_fwd", qkv, lens_t, B, H, T, D, 1.0 / math.sqrt(HD), 0.0, 0, None, None, O, 1)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    Or, _ = _ref_fwd(qkv, lens, B, H, T, D, TP, plain=True)
    got = O.double().view(B, TP, D)[:, PAD:PAD + T]
    assert (got - Or).abs().max().item() < 2e-2 * max(Or.abs().max().item(), 1.0)


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
@pytest.mark.parametrize("B,H,T,lens", CASES)
def test_flash_backward(lib, B, H, T, lens, p_drop):
    """dQ, dK, dV against float64 autograd through softmax (+ the kernels' own dropout mask)."""
    qkv, D, TP = _setup(B, H, T, 3)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    scale, seed = 1.0 / math.sqrt(HD), 77
    Tl = _lse_len(lib, T)
    O = torch.zeros(B * TP, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H, Tl, device="cuda")
    lib.call("fs2_flash_attn_fwd", qkv, lens_t, B, H, T, D, scale, p_drop, seed, None, lse, O, 0)
    g = torch.Generator().manual_seed(9)
    dO = torch.zeros(B, TP, D)
    dO[:, PAD:PAD + T] = torch.randn(B, T, D, generator=g) * 0.5
    dO = dO.view(B * TP, D).cuda().to(torch.bfloat16)
    dvec = torch.zeros(B * H, Tl, device="cuda")
    dqkv = torch.full((B * TP, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv.view(B, TP, 3 * D)[:, :PAD] = 0
    dqkv.view(B, TP, 3 * D)[:, PAD + T:] = 0
    lib.call("fs2_flash_attn_bwd", dO, O, qkv, lse, lens_t, B, H, T, D, scale, p_drop, seed, None, dvec, dqkv, 0)
    torch.cuda.synchronize()
    assert lib.gemm_tc_error_flag() == 0
    got = dqkv.double().view(B, TP, 3 * D)
    assert not torch.isnan(got).any(), "every valid row of dQ, dK, dV must be written"
    assert (got[:, :PAD] == 0).all() and (got[:, PAD + T:] == 0).all(), "halo rows must not be touched"
    keep = _mask(lib, B, H, T, p_drop, seed) if p_drop > 0 else None
    x = qkv.double().view(B, TP, 3 * D)[:, PAD:PAD + T].clone().requires_grad_(True)
    g_o = dO.double().view(B, TP, D)[:, PAD:PAD + T]
    tot = 0.0
    for b in range(B):
        for h in range(H):
            kv = _kv(lens, B, H, b, h, False)
            if kv == 0:
                continue
            q = x[b, :, h * HD:(h + 1) * HD]
            k = x[b, :kv, D + h * HD:D + (h + 1) * HD]
            v = x[b, :kv, 2 * D + h * HD:2 * D + (h + 1) * HD]
            p = torch.softmax(q @ k.t() * scale, -1)
            if keep is not None:
                p = p * keep[b * H + h, :, :kv].double() / (1 - p_drop)
            tot = tot + ((p @ v) * g_o[b, :, h * HD:(h + 1) * HD]).sum()
    tot.backward()
    ref = x.grad
    got = got[:, PAD:PAD + T]
    for name, lo in (("dQ", 0), ("dK", D), ("dV", 2 * D)):
        a, r = got[..., lo:lo + D], ref[..., lo:lo + D]
        err, mag = (a - r).abs().max().item(), r.abs().max().item()
        assert err < 3e-2 * max(mag, 1e-3), (name, err, mag)          # bf16 P / dS operands, bf16 O in the row term
        rl2 = ((a - r).norm() / r.norm().clamp_min(1e-30)).item()
        assert rl2 < 1e-2, (name, "relative L2", rl2)


def test_flash_dropout_mask_statistics(lib):
    """Train-mode statistical check of the attention dropout (the reference's is torch's Philox stream, which cannot be
    matched bit for bit): drop rate p within 4 sigma overall and per (item, head); no correlation between neighbouring
    keys, neighbouring queries, heads or seeds; and the forward output is unbiased: mean over seeds -> the p = 0 output."""
    B, H, T, p = 2, 2, 512, 0.1
    keep = _mask(lib, B, H, T, p, 1234).float()
    n = keep.numel()
    rate = 1.0 - keep.mean().item()
    assert abs(rate - p) < 4 * math.sqrt(p * (1 - p) / n), rate
    per = 1.0 - keep.view(B * H, -1).mean(1)
    assert (per - p).abs().max().item() < 5 * math.sqrt(p * (1 - p) / (T * T))
    d = keep - keep.mean()
    var = d.pow(2).mean().item()
    lim = 5 / math.sqrt(n)
    assert abs((d[:, :, 1:] * d[:, :, :-1]).mean().item() / var) < lim       # neighbouring keys
    assert abs((d[:, 1:] * d[:, :-1]).mean().item() / var) < lim             # neighbouring queries
    assert abs((d[1:] * d[:-1]).mean().item() / var) < lim                   # heads / items
    keep2 = _mask(lib, B, H, T, p, 1235).float()
    assert abs(((keep2 - keep2.mean()) * d).mean().item() / var) < lim       # seeds
    # unbiasedness of the forward
    T2, lens = 96, [96, 50]
    qkv, D, TP = _setup(2, 2, T2, 11)
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    O0 = torch.zeros(2 * TP, D, device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_flash_attn_fwd", qkv, lens_t, 2, 2, T2, D, 1.0 / math.sqrt(HD), 0.0, 0, None, None, O0, 0)
    acc = torch.zeros_like(O0, dtype=torch.float32)
    n_seeds = 400
    O = torch.zeros_like(O0)
    for s in range(n_seeds):
        lib.call("fs2_flash_attn_fwd", qkv, lens_t, 2, 2, T2, D, 1.0 / math.sqrt(HD), p, 1000 + s, None, None, O, 0)
        acc += O.float()
    torch.cuda.synchronize()
    mean = acc / n_seeds
    # per-element standard error of the seed mean: sqrt(p/(1-p) * sum_k P_k^2 v_k^2 / n) <= sqrt(p/(1-p)/n) * max|v|
    bound = 6 * math.sqrt(p / (1 - p) / n_seeds) * qkv.float().abs().max().item()
    assert (mean - O0.float()).abs().max().item() < bound
    assert ((mean - O0.float()).norm() / O0.float().norm()).item() < 0.05
