"""GPU parity of the row / loss kernels, called through the C ABI (ctypes), against torch restatements of the
reference ops (the oracle's own functions where they exist).  Integer artefacts (mel_lens, frame->phoneme map,
duration cumsums, non-zero counts) and the fp32 length-regulator copy are bit-exact; floating point is compared
at the tolerance written next to each assert."""
import importlib
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
PAD = 4


def pad_rows(x, halo=0):
    """(B,T,C) -> padded row space [B*(T+8), C] with reflect halo of width `halo` (zeros beyond)."""
    B, T, C = x.shape
    out = torch.zeros(B, T + 2 * PAD, C, device=x.device, dtype=x.dtype)
    out[:, PAD:PAD + T] = x
    for i in range(1, halo + 1):
        out[:, PAD - i] = x[:, i]
        out[:, PAD + T - 1 + i] = x[:, T - 1 - i]
    return out.reshape(B * (T + 2 * PAD), C).contiguous()


def unpad(x, B, T):
    return x.view(B, T + 2 * PAD, -1)[:, PAD:PAD + T]


def randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


# ------------------------------------------------------------------------------------------ LengthRegulator
@pytest.fixture(params=[8, 0, 16, 32, 64, 128], ids=lambda r: f"bulk{r}")
def lr_bulk_rows(request, lib):
    """plain fp32 expansion: bulk-copy-engine kernel at each CTA size, and the SIMT kernel (0)"""
    raw = lib.load()
    raw.fs2_lr_bulk_rows.argtypes = [lib.C.c_int]
    assert raw.fs2_lr_bulk_rows(request.param) == 0
    yield request.param
    raw.fs2_lr_bulk_rows(8)


@pytest.mark.parametrize("B,Tp,pace", [(2, 5, 1.0), (64, 128, 1.0), (64, 128, 0.8), (32, 128, 1.2), (3, 7, 1.0)])
def test_length_regulator_bit_exact(lib, lr_bulk_rows, B, Tp, pace):
    g = torch.Generator().manual_seed(B * 131 + Tp)
    dur = torch.exp(torch.randn(B, Tp, generator=g) * 0.6 + 1.6).round().clamp(0, 40).long()
    dur[torch.rand(B, Tp, generator=g) < 0.05] = 0
    if B > 2:
        dur[1, Tp // 2:] = 0                       # a short (padded) utterance
        dur[2] = 0
        dur[2, 0] = 1
    D = 384
    feats = torch.randn(B, Tp, D, generator=g)
    ref, ref_lens = O.upsample(feats, dur, pace=pace)            # speechbrain upsample restated (oracle)
    Tm = ref.shape[1]
    dur_c, feats_c = dur.cuda(), feats.cuda()
    ends = torch.zeros(B, Tp, dtype=torch.int32, device="cuda")
    mel_lens = torch.zeros(B, dtype=torch.int32, device="cuda")
    lib.call("fs2_lr_prepare", dur_c, None, pace, B, Tp, ends, mel_lens)
    assert mel_lens.cpu().tolist() == ref_lens                   # bit-exact
    frames = (pace * dur).long()
    assert torch.equal(ends.cpu().long(), torch.cumsum(frames, 1))
    out = torch.full((B, Tm, D), float("nan"), device="cuda")
    f2p = torch.full((B, Tm), -7, dtype=torch.int32, device="cuda")
    lib.call("fs2_lr_expand", feats_c, Tp, 0, ends, mel_lens, None, B, Tp, Tm, D, out, None, 0, Tm, 0, f2p)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)                            # fp32 gather-copy: bit-exact incl. zero padding
    idx_ref = torch.full((B, Tm), -1, dtype=torch.int32)
    for b in range(B):
        m = torch.repeat_interleave(torch.arange(Tp), frames[b])
        idx_ref[b, : m.numel()] = m.int()
    assert torch.equal(f2p.cpu(), idx_ref)                        # frame -> phoneme index map: bit-exact
    # masks (get_mask_from_lengths): True = pad
    mask_ref = O.get_mask_from_lengths(torch.tensor(ref_lens))
    assert torch.equal((f2p.cpu() < 0), mask_ref)
    # backward = segment sums: sum over a phoneme's frames
    df = torch.randn(B, Tm, D, generator=g).cuda()
    dph = torch.zeros(B, Tp, D, device="cuda")
    lib.call("fs2_lr_bwd", df, None, Tm, 0, ends, mel_lens, B, Tp, Tm, D, dph, Tp, 0)
    ref_b = torch.zeros(B, Tp, D, dtype=torch.float64)
    for b in range(B):
        m = torch.repeat_interleave(torch.arange(Tp), frames[b])
        ref_b[b].index_add_(0, m, df[b, : m.numel()].double().cpu())
    assert (dph.double().cpu() - ref_b).abs().max() <= 1e-4       # fp32 sums of <= 48 terms
    # the model's call: padded row space on both sides, two gradient inputs, an output buffer full of garbage --
    # every row (halo rows included) is written, deterministically, by both forms of the kernel
    PAD = 4
    dfp = torch.full((B, Tm + 2 * PAD, D), float("nan"), device="cuda")
    dfp2 = torch.full((B, Tm + 2 * PAD, D), float("nan"), device="cuda")
    dfp[:, PAD:PAD + Tm] = df
    dfp2[:, PAD:PAD + Tm] = 0.5 * df
    outs = []
    raw = lib.load()
    raw.fs2_lr_tune_bwd.argtypes = [lib.C.c_int]
    for form in (1, 0, 1):
        raw.fs2_lr_tune_bwd(form)
        dpp = torch.full((B, Tp + 2 * PAD, D), float("nan"), device="cuda")
        lib.call("fs2_lr_bwd", dfp, dfp2, Tm + 2 * PAD, PAD, ends, mel_lens, B, Tp, Tm, D, dpp, Tp + 2 * PAD, PAD)
        torch.cuda.synchronize()
        assert (dpp[:, :PAD] == 0).all() and (dpp[:, PAD + Tp:] == 0).all()
        assert (dpp[:, PAD:PAD + Tp].double().cpu() - 1.5 * ref_b).abs().max() <= 2e-4
        outs.append(dpp)
    raw.fs2_lr_tune_bwd(0)
    assert torch.equal(outs[0], outs[2])                          # segment-sum form: bit-reproducible


@pytest.mark.parametrize("B,Tp,maxd", [(4, 9, 6), (6, 61, 24)])
def test_length_regulator_with_posenc_and_padded_rows(lib, B, Tp, maxd):
    """The fused form the model uses: padded row space in/out, + decoder pos-enc, bf16 operand copy."""
    D = 384
    g = torch.Generator().manual_seed(3 + Tp)
    dur = torch.randint(0, maxd, (B, Tp), generator=g)
    dur[:, 0] += 1
    if B > 4:
        dur[1, 2:] = 0                               # a short utterance: most of its rows are zero rows
        dur[2, 1:40] = 0                             # a long run of empty phonemes
        dur[3, 5] = 70                               # one phoneme across many CTAs
    feats = torch.randn(B, Tp, D, generator=g)
    pe = O.PositionalEncoding(D).pe[0]
    ref, lens = O.upsample(feats, dur)
    Tm = ref.shape[1]
    mask = (~O.get_mask_from_lengths(torch.tensor(lens))).unsqueeze(-1)
    ref = (ref + pe[:Tm]) * mask
    ends = torch.zeros(B, Tp, dtype=torch.int32, device="cuda")
    ml = torch.zeros(B, dtype=torch.int32, device="cuda")
    lib.call("fs2_lr_prepare", dur.cuda(), None, 1.0, B, Tp, ends, ml)
    inp = pad_rows(feats.cuda())
    of = torch.full((B * (Tm + 8), D), float("nan"), device="cuda")
    oa = torch.full((B * (Tm + 8), D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.call("fs2_lr_expand", inp, Tp + 8, PAD, ends, ml, pe.cuda(), B, Tp, Tm, D, of, oa, 1, Tm + 8, PAD, None)
    assert torch.equal(unpad(of, B, Tm).cpu(), ref)               # same fp32 add, same order: bit-exact
    assert torch.equal(oa.float().cpu(), of.to(torch.bfloat16).float().cpu())
    full = of.view(B, Tm + 8, D)
    assert full[:, :PAD].abs().sum() == 0 and full[:, PAD + Tm:].abs().sum() == 0   # zero halo rows


def test_length_regulator_plain_copy_in_padded_rows(lib, lr_bulk_rows):
    """Plain fp32 expansion (bulk-copy-engine form) with row pitch / offset on both sides and long zero-duration runs:
    halo rows and rows past an item's length must come out as zeros, phoneme runs straddle CTA boundaries."""
    B, Tp, D = 5, 40, 384
    g = torch.Generator().manual_seed(11)
    dur = torch.randint(0, 30, (B, Tp), generator=g)
    dur[0, 3:30] = 0
    dur[1, 1:] = 0
    dur[1, 0] = 200                                               # one phoneme across several CTAs
    dur[2] = 0
    dur[2, -1] = 7                                                # everything before the last phoneme is empty
    feats = torch.randn(B, Tp, D, generator=g)
    ref, lens = O.upsample(feats, dur)
    Tm = ref.shape[1]
    ends = torch.zeros(B, Tp, dtype=torch.int32, device="cuda")
    ml = torch.zeros(B, dtype=torch.int32, device="cuda")
    lib.call("fs2_lr_prepare", dur.cuda(), None, 1.0, B, Tp, ends, ml)
    inp = pad_rows(feats.cuda())
    of = torch.full((B * (Tm + 8), D), float("nan"), device="cuda")
    f2p = torch.full((B, Tm), -7, dtype=torch.int32, device="cuda")
    lib.call("fs2_lr_expand", inp, Tp + 8, PAD, ends, ml, None, B, Tp, Tm, D, of, None, 0, Tm + 8, PAD, f2p)
    torch.cuda.synchronize()
    assert torch.equal(unpad(of, B, Tm).cpu(), ref)
    full = of.view(B, Tm + 8, D)
    assert full[:, :PAD].abs().sum() == 0 and full[:, PAD + Tm:].abs().sum() == 0
    for b in range(B):
        m = torch.repeat_interleave(torch.arange(Tp), dur[b]).int()
        assert torch.equal(f2p[b, : m.numel()].cpu(), m) and (f2p[b, m.numel():] == -1).all()


def test_dur_decode_inference(lib):
    x = randn(4, 50, seed=1) * 2
    out = torch.empty_like(x)
    lib.call("fs2_dur_decode", x, x.numel(), out)
    ref = torch.clamp(torch.special.expm1(x), 0)
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------- average_over_durations
@pytest.mark.parametrize("B,Tp,Tm", [(2, 5, 15), (64, 128, 800), (8, 40, 333)])
def test_average_over_durations(lib, B, Tp, Tm):
    g = torch.Generator().manual_seed(Tm)
    w = torch.rand(B, Tp, generator=g)
    dur = torch.floor(w / w.sum(1, keepdim=True) * Tm).long()
    dur[:, 0] += Tm - dur.sum(1)
    dur[0, 1] = 0 if Tp > 2 else dur[0, 1]
    dur[0, 0] = Tm - dur[0, 1:].sum()
    vals = torch.randn(B, Tm, generator=g)
    vals[torch.rand(B, Tm, generator=g) < 0.3] = 0.0             # unvoiced frames are exactly zero
    ref = O.average_over_durations(vals.unsqueeze(1), dur)[:, 0]  # CPU torch: the oracle
    avg = torch.empty(B, Tp, device="cuda")
    st, en, nz = (torch.empty(B, Tp, dtype=torch.int32, device="cuda") for _ in range(3))
    lib.call("fs2_avg_over_durations", vals.cuda(), dur.cuda(), B, Tp, Tm, avg, st, en, nz)
    ends_ref = torch.cumsum(dur, 1)
    assert torch.equal(en.cpu().long(), ends_ref)                                    # bit-exact ints
    assert torch.equal(st.cpu().long(), F.pad(ends_ref[:, :-1], (1, 0)))
    nzc = F.pad(torch.cumsum(vals != 0, 1), (1, 0))
    assert torch.equal(nz.cpu().long(), torch.gather(nzc, 1, ends_ref) - torch.gather(nzc, 1, F.pad(ends_ref[:, :-1], (1, 0))))
    # values: same prefix-sum-difference formulation; fp32, tolerance 1e-6 absolute (oracle fp32 itself is ~1e-6 from fp64)
    assert (avg.cpu() - ref).abs().max() <= 2e-6


# -------------------------------------------------------------------------------------------- LayerNorm
def _ln_struct(lib, cls, **kw):
    p = cls()
    for k, v in kw.items():
        setattr(p, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return p


@pytest.mark.parametrize("C,tanh,with_branch,halo", [(384, 0, True, 4), (512, 1, False, 2), (80, 0, False, 0)])
def test_layernorm_fwd_bwd(lib, C, tanh, with_branch, halo):
    B, T = 3, 21
    lens = torch.tensor([21, 13, 5], dtype=torch.int32, device="cuda")
    x = randn(B, T, C, seed=1).requires_grad_()
    br = randn(B, T, C, seed=2).requires_grad_() if with_branch else None
    gamma = (randn(C, seed=3) * 0.2 + 1).requires_grad_()
    beta = (randn(C, seed=4) * 0.1).requires_grad_()
    post = randn(B, T, C, seed=5)
    eps = 1e-5
    z = x + br if with_branch else x
    u = F.layer_norm(z, (C,), gamma, beta, eps)
    if tanh:
        u = torch.tanh(u)
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1)
    ref = u * mask + post
    dy = randn(B, T, C, seed=6)
    ref.backward(dy)
    rows = B * (T + 8)
    of = torch.full((rows, C), float("nan"), device="cuda")
    oa = torch.full((rows, C), float("nan"), device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    xp, bp, pp = pad_rows(x.detach()), pad_rows(br.detach()) if with_branch else None, pad_rows(post)
    p = _ln_struct(lib, lib.Fs2LnFwd, B=B, T=T, C=C, x=xp, gamma=gamma.detach(), beta=beta.detach(), eps=eps, tanh_act=tanh,
                   lens=lens, post_add=pp, out_f32=of, out_act=oa, act_bf16=0, halo=halo, mean=mean, rstd=rstd)
    if with_branch:
        p.branch = bp.data_ptr()
    lib.call("fs2_ln_fwd", lib.C.addressof(p))
    assert (unpad(of, B, T) - ref.detach()).abs().max() <= 2e-5
    assert torch.equal(of, pad_rows(unpad(of, B, T).contiguous(), halo))       # reflect halo rows, zeros beyond
    assert torch.equal(of, oa)
    dx = torch.full((rows, C), float("nan"), device="cuda")
    dact = torch.full((rows, C), float("nan"), device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    q = _ln_struct(lib, lib.Fs2LnBwd, B=B, T=T, C=C, dy=pad_rows(dy), x=xp, gamma=gamma.detach(), beta=beta.detach(), eps=eps,
                   tanh_act=tanh, lens=lens, mean=mean, rstd=rstd, dx_f32=dx, dact=dact, act_bf16=0, dgamma=dg, dbeta=db)
    if with_branch:
        q.branch = bp.data_ptr()
    lib.call("fs2_ln_bwd", lib.C.addressof(q))
    assert (unpad(dx, B, T) - x.grad).abs().max() <= 1e-4 * max(1.0, x.grad.abs().max().item())
    assert (dg - gamma.grad).abs().max() <= 1e-4 * max(1.0, gamma.grad.abs().max().item())
    assert (db - beta.grad).abs().max() <= 1e-4 * max(1.0, beta.grad.abs().max().item())
    full = dact.view(B, T + 8, C)
    assert full[:, :PAD].abs().sum() == 0 and full[:, PAD + T:].abs().sum() == 0      # gradient tensors: zero halo


def test_layernorm_dropout_is_consistent_between_fwd_and_bwd(lib):
    """Counter-based dropout: the backward regenerates exactly the forward's mask (train-mode parity is
    statistical, SURVEY section 7: torch's Philox stream cannot be reproduced)."""
    B, T, C, p_drop = 2, 33, 384, 0.5
    x = randn(B, T, C, seed=1)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    rows = B * (T + 8)
    of = torch.empty(rows, C, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    p = _ln_struct(lib, lib.Fs2LnFwd, B=B, T=T, C=C, x=pad_rows(x), gamma=gamma, beta=beta, eps=1e-5, drop_a_p=p_drop,
                   drop_a_seed=12345, out_f32=of, act_bf16=0, mean=mean, rstd=rstd)
    lib.call("fs2_ln_fwd", lib.C.addressof(p))
    out = unpad(of, B, T)
    ln = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    keep = out != 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p_drop)) < 0.02                                   # drop rate
    assert torch.allclose(out[keep], ln[keep] / (1 - p_drop), atol=1e-5)     # inverted-dropout scaling
    xr = x.clone().requires_grad_()
    ref = F.layer_norm(xr, (C,), gamma, beta, 1e-5) * keep / (1 - p_drop)
    dy = randn(B, T, C, seed=9)
    ref.backward(dy)
    dx = torch.empty(rows, C, device="cuda")
    q = _ln_struct(lib, lib.Fs2LnBwd, B=B, T=T, C=C, dy=pad_rows(dy), x=pad_rows(x), gamma=gamma, beta=beta, eps=1e-5,
                   drop_a_p=p_drop, drop_a_seed=12345, mean=mean, rstd=rstd, dx_f32=dx, act_bf16=0)
    lib.call("fs2_ln_bwd", lib.C.addressof(q))
    assert (unpad(dx, B, T) - xr.grad).abs().max() <= 1e-4


# ---------------------------------------------------------------------------------------------- softmax
def test_softmax_mask_quirk(lib):
    """keys valid for (b,h) = [0, min(len[b], len[(b*H+h) % B])) -- the reference's transposed attn_mask (Q1)."""
    B, H, T = 4, 2, 37
    ldk = 40
    lens = torch.tensor([37, 30, 12, 5], dtype=torch.int32, device="cuda")
    S = randn(B * H, T, ldk, seed=2)
    P = torch.full((B * H, T, ldk), float("nan"), device="cuda")
    scale = 1 / math.sqrt(192)
    lib.call("fs2_softmax_fwd", S, lens, B, H, T, ldk, scale, 0.0, 0, None, P, None, 0)
    pad = torch.arange(T, device="cuda")[None, :] >= lens[:, None]
    ref = torch.zeros(B * H, T, ldk, device="cuda")
    for b in range(B):
        for h in range(H):
            km = pad[b] | pad[(b * H + h) % B]
            s = (S[b * H + h, :, :T] * scale).masked_fill(km[None, :], float("-inf"))
            ref[b * H + h, :, :T] = torch.softmax(s, -1)
    assert (P - ref).abs().max() <= 2e-6
    dP = randn(B * H, T, ldk, seed=3)
    Pr = ref.clone().requires_grad_()
    dS = torch.empty_like(P)
    lib.call("fs2_softmax_bwd", P, dP, lens, B, H, T, ldk, scale, 0.0, 0, None, dS, 0)
    Sr = S.clone().requires_grad_()
    out = torch.zeros_like(ref)
    for b in range(B):
        for h in range(H):
            km = pad[b] | pad[(b * H + h) % B]
            s = (Sr[b * H + h, :, :T] * scale).masked_fill(km[None, :], float("-inf"))
            out[b * H + h, :, :T] = torch.softmax(s, -1)
    out.backward(dP)
    assert (dS - Sr.grad).abs().max() <= 2e-6


# ----------------------------------------------------------------------------------------------- losses
def _loss_inputs(B, Tp, Tm, seed, full=False):
    g = torch.Generator().manual_seed(seed)
    mel_len = torch.randint(11, Tm + 1, (B,), generator=g)
    mel_len[0] = Tm
    if full:
        mel_len[:] = Tm
    phon_len = torch.randint(2, Tp + 1, (B,), generator=g)
    phon_len[0] = Tp
    mel_tgt = torch.rand(B, Tm, 80, generator=g) * 13.5 - 11.5
    mel_out = mel_tgt + torch.randn(B, Tm, 80, generator=g) * 0.7
    post = mel_tgt + torch.randn(B, Tm, 80, generator=g) * 0.5
    for b in range(B):
        mel_tgt[b, mel_len[b]:] = 0
        mel_out[b, mel_len[b]:] = 0
    dur = torch.randint(0, 12, (B, Tp), generator=g)
    return dict(mel_out=mel_out, post=post, log_dur=torch.randn(B, Tp, generator=g), pitch=torch.randn(B, Tp, 1, generator=g),
                avg_pitch=torch.randn(B, Tp, 1, generator=g), energy=torch.randn(B, Tp, 1, generator=g),
                avg_energy=torch.randn(B, Tp, 1, generator=g), mel_tgt=mel_tgt, dur=dur, mel_len=mel_len, phon_len=phon_len)


@pytest.mark.parametrize("B,Tp,Tm,full", [(4, 13, 80, False), (32, 128, 800, False), (3, 6, 11, True), (2, 130, 40, False)])
def test_loss_values_and_gradients(pkg, B, Tp, Tm, full):
    """Loss drop-in vs the oracle's Loss (loss.py restated) incl. quirks Q5/Q6 and SSIM; gradients via autograd."""
    d = _loss_inputs(B, Tp, Tm, seed=B * 7 + Tm, full=full)
    names = ["mel_out", "post", "log_dur", "pitch", "energy"]
    ref_in = {k: d[k].double().requires_grad_() for k in names}
    crit_o = O.Loss(**{**O.DEFAULT_LOSS_CONFIG, "ssim_loss_weight": 0.7, "pitch_loss_weight": 1.3})
    preds_o = (ref_in["mel_out"], ref_in["post"], ref_in["log_dur"], ref_in["pitch"], d["avg_pitch"].double(), ref_in["energy"],
               d["avg_energy"].double(), d["mel_len"])
    lo = crit_o(preds_o, (d["mel_tgt"].double(), d["dur"], None, None, d["mel_len"], d["phon_len"]), 0)
    lo["total_loss"].backward()
    c = lambda t: t.cuda()
    my_in = {k: c(d[k]).requires_grad_() for k in names}
    crit = pkg.Loss(**{**pkg.DEFAULT_LOSS_CONFIG, "ssim_loss_weight": 0.7, "pitch_loss_weight": 1.3})
    preds = (my_in["mel_out"], my_in["post"], my_in["log_dur"], my_in["pitch"], c(d["avg_pitch"]), my_in["energy"],
             c(d["avg_energy"]), d["mel_len"])
    lm = crit(preds, (c(d["mel_tgt"]), c(d["dur"]), None, None, c(d["mel_len"]), c(d["phon_len"])), 0)
    lm["total_loss"].backward()
    for k in lo:
        a, b = float(lm[k]), float(lo[k])
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (k, a, b)             # fp32 reductions vs fp64 oracle
    for k in names:
        ga, gb = my_in[k].grad.double().cpu(), ref_in[k].grad
        err = (ga - gb).abs().max().item()
        assert err <= 5e-4 * gb.abs().max().item() + 1e-9, (k, err, gb.abs().max().item())


def test_ssim_clamp_matches_reference(pkg):
    """SSIMLoss replaces out-of-range values by constants (no gradient); in-range values pass through."""
    d = _loss_inputs(2, 6, 30, seed=5)
    c = lambda t: t.cuda()
    mo = c(d["mel_out"]).requires_grad_()
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    preds = (mo, c(d["post"]), c(d["log_dur"]), c(d["pitch"]), c(d["avg_pitch"]), c(d["energy"]), c(d["avg_energy"]), d["mel_len"])
    out = crit(preds, (c(d["mel_tgt"]), c(d["dur"]), None, None, c(d["mel_len"]), c(d["phon_len"])), 0)
    assert 0.0 <= float(out["ssim_loss"]) <= 1.0


# --------------------------------------------------------------------------------- small fused row kernels
def test_embed_posenc_and_masks(lib):
    B, Tp, D = 3, 11, 384
    tokens = torch.tensor([[5, 7, 9, 1, 2, 3, 4, 5, 6, 7, 8], [3, 4, 5, 6, 7, 0, 0, 0, 0, 0, 0], [9] + [0] * 10]).cuda()
    emb = randn(95, D, seed=1)
    pe = O.PositionalEncoding(D).pe[0].cuda()
    of = torch.full((B * (Tp + 8), D), float("nan"), device="cuda")
    lens = torch.zeros(B, dtype=torch.int32, device="cuda")
    lib.call("fs2_embed_posenc", tokens, emb, pe, B, Tp, D, 0, of, None, 0, lens)
    ref = (emb[tokens] + pe[:Tp]) * (tokens != 0).unsqueeze(-1)
    assert torch.equal(unpad(of, B, Tp), ref)                    # same fp32 add: bit-exact
    assert lens.tolist() == [11, 5, 1]                           # key-padding mask as lengths: bit-exact


def test_embed_add_matches_conv1d_reflect(lib):
    B, Tp, D, k = 3, 17, 384, 3
    x, contour = randn(B, Tp, D, seed=1), randn(B, Tp, seed=2)
    w, bias = randn(D, 1, k, seed=3) * 0.3, randn(D, seed=4) * 0.1
    lens = torch.tensor([17, 9, 2], dtype=torch.int32, device="cuda")
    conv = F.conv1d(F.pad(contour[:, None, :], (1, 1), mode="reflect"), w, bias).permute(0, 2, 1)
    ref = x + conv
    yf = torch.empty(B * (Tp + 8), D, device="cuda")
    ya = torch.empty(B * (Tp + 8), D, device="cuda")
    lib.call("fs2_embed_add", pad_rows(x), contour, w, bias, k, lens, B, Tp, D, yf, ya, 0, 1)
    assert (unpad(yf, B, Tp) - ref).abs().max() <= 1e-5
    mask = (torch.arange(Tp, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1)
    assert torch.equal(ya, pad_rows((unpad(yf, B, Tp) * mask).contiguous(), 1))
    dy = randn(B, Tp, D, seed=5)
    wr, br = w.clone().requires_grad_(), bias.clone().requires_grad_()
    (F.conv1d(F.pad(contour[:, None, :], (1, 1), mode="reflect"), wr, br).permute(0, 2, 1) * dy).sum().backward()
    dw, db = torch.zeros(D, 1, k, device="cuda"), torch.zeros(D, device="cuda")
    lib.call("fs2_embed_add_bwd", pad_rows(dy), contour, k, B, Tp, D, dw, db)
    assert (dw - wr.grad).abs().max() <= 1e-4 and (db - br.grad).abs().max() <= 1e-4


def test_conditioning_split_weight_equals_cat_linear(lib):
    """model.py:352-360: cat[token, speaker, intensity] @ W^T == token@Wt^T + Ws.spk + Wi.intensity."""
    B, Tp, D = 4, 10, 384
    tok, inten = randn(B, Tp, D, seed=1), randn(B, Tp, 5, seed=2)
    W = (randn(D, 2 * D + 5, seed=3) * 0.05).requires_grad_()
    spk_emb = randn(4, D, seed=4).requires_grad_()
    speakers = torch.tensor([2, 0, 3, 2]).cuda()
    lens = torch.tensor([10, 7, 3, 1], dtype=torch.int32, device="cuda")
    mask = (torch.arange(Tp, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1)
    ref = F.linear(torch.cat([tok, spk_emb[speakers][:, None].expand(-1, Tp, -1), inten], -1), W) * mask
    G = pad_rows((tok @ W.detach()[:, :D].t()).contiguous())
    yf = torch.empty(B * (Tp + 8), D, device="cuda")
    lib.call("fs2_cond_finish", G, W.detach(), spk_emb.detach(), speakers, inten, lens, B, Tp, D,
             torch.empty(B, D, device="cuda"), yf, None, 0, 1)
    assert (unpad(yf, B, Tp) - ref.detach()).abs().max() <= 2e-5
    dy = randn(B, Tp, D, seed=6) * mask
    ref.backward(dy)
    dW, dspk = torch.zeros_like(W), torch.zeros_like(spk_emb)
    lib.call("fs2_cond_bwd", pad_rows(dy), W.detach(), spk_emb.detach(), speakers, inten, B, Tp, D,
             torch.empty(B, D, device="cuda"), dW, dspk)
    assert (dW[:, D:] - W.grad[:, D:]).abs().max() <= 1e-4 * max(1.0, W.grad.abs().max().item())
    assert (dspk - spk_emb.grad).abs().max() <= 1e-4 * max(1.0, spk_emb.grad.abs().max().item())


def test_fold_halo_colsum_pad_unpad(lib):
    B, T, C, p = 3, 12, 80, 2
    src = randn(B * (T + 8), C, seed=1)
    lens = torch.tensor([12, 6, 1], dtype=torch.int32, device="cuda")
    out = torch.full_like(src, float("nan"))
    lib.call("fs2_fold_halo", src, B, T, C, p, None, None, lens, out, None, 0)
    s = src.view(B, T + 8, C)
    ref = s[:, PAD:PAD + T].clone()
    for i in range(1, p + 1):
        ref[:, i] += s[:, PAD - i]
        ref[:, T - 1 - i] += s[:, PAD + T - 1 + i]
    ref = ref * (torch.arange(T, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1)
    assert (unpad(out, B, T) - ref).abs().max() <= 1e-6
    cs = torch.zeros(C, device="cuda")
    lib.call("fs2_colsum", src, 0, src.shape[0], C, C, cs)
    assert (cs - src.sum(0)).abs().max() <= 1e-4
    # bf16 inputs: the 16-byte-load kernel (C % 8 == 0), ragged column tails and a leading dimension wider than C
    for rows_b, Cb, ldb in ((1000, 1536, 1536), (777, 1152, 1152), (300, 384, 1152), (50, 80, 80), (33, 52, 64)):
        xb = randn(rows_b, ldb, seed=5).to(torch.bfloat16)
        csb = torch.zeros(Cb, device="cuda")
        lib.call("fs2_colsum", xb, 1, rows_b, Cb, ldb, csb)
        refb = xb[:, :Cb].double().sum(0)
        assert (csb.double() - refb).abs().max() <= 1e-3 * max(1.0, refb.abs().max().item()), (rows_b, Cb, ldb)
    plain = torch.full((B, T, C), float("nan"), device="cuda")
    lib.call("fs2_unpad_mask", src, lens, B, T, C, plain, None, 0, 0)
    assert torch.equal(plain, s[:, PAD:PAD + T] * (torch.arange(T, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1))
    back = torch.full_like(src, float("nan"))
    lib.call("fs2_pad_rows", plain, None, B, T, C, 1.0, back, None, 0)
    assert torch.equal(back, pad_rows(plain))


def test_fused_adamw_matches_torch(lib):
    n = 100003
    p0, g = randn(n, seed=1), randn(n, seed=2) * 0.1
    pr = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([pr], lr=1e-3)
    p = p0.clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        pr.grad = g * step
        opt.step()
        lib.call("fs2_adamw", p, g * step, m, v, n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0)
    assert (p - pr.detach()).abs().max() <= 2e-6


def test_intensity_segment_mean_matches_train_py(lib):
    """train.py:16-51 restated in oracle.intensity_segment_mean (the 'next' row f-1)."""
    B, Tp, Tm, D = 5, 12, 90, 5
    g = torch.Generator().manual_seed(4)
    phon_len = torch.tensor([12, 10, 7, 3, 1])
    dur = torch.randint(0, 8, (B, Tp), generator=g)
    for b in range(B):
        dur[b, phon_len[b]:] = 0
    inten = torch.randn(B, Tm, D, generator=g)
    ref = O.intensity_segment_mean(inten, phon_len, dur, Tp)
    out = torch.empty(B, Tp, D, device="cuda")
    lib.call("fs2_intensity_segment_mean", inten.cuda(), dur.cuda(), phon_len.cuda(), B, Tp, Tm, D, out)
    assert (out.cpu() - ref).abs().max() <= 1e-5


def test_get_intensity_representation_drop_in(pkg):
    """Same call as train.py:69 with a stand-in extractor: the batch tuple layout of dataset.py:120-133 goes in, the
    (B, T_phon_max, 5) representation comes out and equals the reference's per-sample loop (restated in the oracle)."""
    data = importlib.import_module(pkg.__name__ + ".data")
    (batch, _), = data.synthetic_batches(6, 1, seed=5, min_tp=6, max_tp=20, max_frames=120, pool_factor=2)
    g = torch.Generator().manual_seed(9)
    Tm = batch[3].shape[1]
    frames = torch.randn(6, Tm, 5, generator=g)

    def extractor(rank_X, mel_len, emo_ids):          # stands in for RankModel.intensity_extractor (frozen)
        assert rank_X.shape == (6, 82, Tm)
        return frames.cuda()

    dev_batch = tuple(t.cuda() if torch.is_tensor(t) else t for t in batch)
    out = pkg.get_intensity_representation(extractor, dev_batch, torch.device("cuda"))
    ref = O.intensity_segment_mean(frames, batch[2], batch[6], batch[0].shape[1])
    assert out.shape == ref.shape and (out.cpu() - ref).abs().max() <= 1e-5


def test_fused_adamw_ranges_equal_one_launch(pkg):
    """FusedAdamW.step(ranges=...) -- used by the data-parallel step to update one all-reduced piece while the next
    is still on the wire -- gives bit-identical parameters to the single-launch update."""
    torch.manual_seed(0)
    ma = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4).cuda()
    mb = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4).cuda()
    mb.store.flat.copy_(ma.store.flat)
    for m in (ma, mb):
        m.store.ensure_grads()
    g = torch.Generator(device="cuda").manual_seed(1)
    ma.store.flat_grad.copy_(torch.randn(ma.store.flat.numel(), device="cuda", generator=g) * 1e-2)
    mb.store.flat_grad.copy_(ma.store.flat_grad)
    oa, ob = pkg.FusedAdamW(ma, lr=1e-3), pkg.FusedAdamW(mb, lr=1e-3)
    lo, n = ma.grad_split_lo, ma.store.flat.numel()
    assert 0 < lo < n and lo % 4 == 0
    called = []
    for _ in range(3):
        oa.step(grad_scale=0.5)
        ob.step(grad_scale=0.5, ranges=[(lo, n, lambda: called.append("tail")), (0, lo, lambda: called.append("head"))])
    assert called == ["tail", "head"] * 3
    assert torch.equal(ma.store.flat, mb.store.flat) and torch.equal(oa.m, ob.m) and torch.equal(oa.v, ob.v)


def test_intensity_prototypes_match_the_reference_binning(pkg):
    """rank_model/inference.py:80-114 restated with its own tools (python lists, list.sort on the score, numpy
    concatenate / array_split / mean; tests/prototype_case.py, itself checked against the reference's own lines in
    tests/test_reference_glue.py) against the device version -- including a (speaker, emotion) cell with fewer frames
    than buckets (numpy: NaN for the empty slices) and a cell without utterances (stays zero)."""
    import numpy as np
    import prototype_case as PC
    I, length, r, spk, emo, n_spk, n_emo, k = PC.inputs()
    ref = PC.restated(I, length, r, spk, emo, n_spk, n_emo, k)
    out = pkg.intensity_prototypes(I.cuda(), length.cuda(), r.cuda(), spk.cuda(), emo.cuda(), n_spk, n_emo, k).cpu().numpy()
    assert out.shape == ref.shape
    assert np.array_equal(np.isnan(out), np.isnan(ref)) and np.isnan(ref[2, 2]).any()
    assert np.nanmax(np.abs(out - ref)) <= 1e-5
    assert (out[:, 3] == 0).all()
