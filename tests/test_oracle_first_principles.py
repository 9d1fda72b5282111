"""First-principles checks of the oracle (oracle/fs2_oracle.py).

The reference's arithmetic lives in speechbrain, which is absent here ("parity unpinned"), so the oracle is
built from torch.nn modules wired the way speechbrain wires them.  These tests restate the individual blocks a second
time in plain numpy float64 -- explicit loops over taps, heads and frames, no torch.nn -- from the published
definitions (Vaswani-style post-norm block with a Conv1d FFN, reflect "same" padding, LayerNorm with biased variance,
piq's SSIM, speechbrain's average_over_durations), read the weights out of the oracle's state_dict with the reference's
key names, and require agreement to fp64 round-off.  They pin what the CUDA kernels are compared
against: the packed q/k/v row order of `in_proj_weight`, the 1/sqrt(head_dim) scale, key-padding as -inf before the
softmax, the reflect halo at the edge of the padded rectangle, eps placement, mean-over-non-zero frames.
CPU only, small shapes.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

F64 = torch.float64


def _np(t):
    return t.detach().double().numpy()


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)          # biased variance
    return (x - mu) / np.sqrt(var + eps) * w + b


def conv1d_same_reflect(x, w, b):
    """x (T, Cin), w (Cout, Cin, k) torch layout, reflect padding (k-1)//2 without repeating the edge sample."""
    T, k = x.shape[0], w.shape[2]
    p = (k - 1) // 2
    idx = np.arange(-p, T + p)
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= T, 2 * (T - 1) - idx, idx)
    xp = x[idx]                                            # (T + 2p, Cin)
    y = np.zeros((T, w.shape[0]))
    for j in range(k):                                     # cross-correlation, tap j reads frame t + j - p
        y += xp[j:j + T] @ w[:, :, j].T
    return y + b


def test_conv1d_same_is_reflect_padding_at_the_rectangle_edge():
    torch.manual_seed(1)
    for k in (1, 3, 5, 9):
        m = O.SBConv1d(6, 4, k).to(F64)
        x = torch.randn(2, 13, 6, dtype=F64)
        ref = _np(m(x))
        for b in range(2):
            got = conv1d_same_reflect(_np(x[b]), _np(m.conv.weight), _np(m.conv.bias))
            assert np.abs(got - ref[b]).max() < 1e-12


def test_fft_block_against_numpy():
    """One post-norm FFT block (2 heads of 192, conv k=9 -> ReLU -> conv k=1, LayerNorm eps 1e-6) with a key-padding
    mask, written out with loops over heads and taps."""
    torch.manual_seed(2)
    D, H, Fd, T, B = 384, 2, 64, 17, 3
    layer = O.TransformerEncoderLayer(Fd, H, D, D, D, 0.0, False, [9, 1]).to(F64).eval()
    sd = {k: _np(v) for k, v in layer.state_dict().items()}
    x = torch.randn(B, T, D, dtype=F64)
    lens = [17, 11, 6]
    kpm = torch.zeros(B, T, dtype=torch.bool)
    for b, n in enumerate(lens):
        kpm[b, n:] = True
    ref, attn = layer(x, src_key_padding_mask=kpm)
    ref, attn = _np(ref), _np(attn)
    Wi, bi = sd["self_att.att.in_proj_weight"], sd["self_att.att.in_proj_bias"]
    Wo, bo = sd["self_att.att.out_proj.weight"], sd["self_att.att.out_proj.bias"]
    dh = D // H
    for b in range(B):
        xb = _np(x[b])
        q = xb @ Wi[0:D].T + bi[0:D]                       # packed rows: q, k, v
        k = xb @ Wi[D:2 * D].T + bi[D:2 * D]
        v = xb @ Wi[2 * D:].T + bi[2 * D:]
        ctx = np.zeros((T, D))
        pavg = np.zeros((T, T))
        for h in range(H):
            sl = slice(h * dh, (h + 1) * dh)
            s = (q[:, sl] / math.sqrt(dh)) @ k[:, sl].T
            s[:, lens[b]:] = -np.inf                       # padded keys never attended to
            p = np.exp(s - s.max(-1, keepdims=True))
            p /= p.sum(-1, keepdims=True)
            ctx[:, sl] = p @ v[:, sl]
            pavg += p / H
        a = ctx @ Wo.T + bo
        y = layer_norm(xb + a, sd["norm1.norm.weight"], sd["norm1.norm.bias"], 1e-6)
        f = conv1d_same_reflect(y, sd["pos_ffn.0.conv.weight"], sd["pos_ffn.0.conv.bias"])
        f = np.maximum(f, 0.0)
        f = conv1d_same_reflect(f, sd["pos_ffn.2.conv.weight"], sd["pos_ffn.2.conv.bias"])
        out = layer_norm(y + f, sd["norm2.norm.weight"], sd["norm2.norm.bias"], 1e-6)
        assert np.abs(out - ref[b]).max() < 1e-10
        assert np.abs(pavg - attn[b]).max() < 1e-12       # head-averaged weights (need_weights=True)


def test_variance_predictor_against_numpy():
    """DurationPredictor: mask -> conv3 -> ReLU -> LN(1e-5) -> mask -> conv3 -> ReLU -> LN -> mask -> Linear(384 -> 1);
    padded rows come out as the linear bias, not zero."""
    torch.manual_seed(3)
    C, T, B = 32, 12, 2
    m = O.DurationPredictor(C, C, 3, 0.0).to(F64).eval()
    sd = {k: _np(v) for k, v in m.state_dict().items()}
    x = torch.randn(B, T, C, dtype=F64)
    lens = [12, 7]
    mask = torch.zeros(B, T, 1, dtype=F64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1.0
    ref = _np(m(x, mask))
    for b in range(B):
        mk = _np(mask[b])
        h = np.maximum(conv1d_same_reflect(_np(x[b]) * mk, sd["conv1.conv.weight"], sd["conv1.conv.bias"]), 0.0)
        h = layer_norm(h, sd["ln1.norm.weight"], sd["ln1.norm.bias"], 1e-5)
        h = np.maximum(conv1d_same_reflect(h * mk, sd["conv2.conv.weight"], sd["conv2.conv.bias"]), 0.0)
        h = layer_norm(h, sd["ln2.norm.weight"], sd["ln2.norm.bias"], 1e-5)
        out = (h * mk) @ sd["linear.w.weight"].T + sd["linear.w.bias"]
        assert np.abs(out - ref[b]).max() < 1e-11
        assert np.allclose(out[lens[b]:], sd["linear.w.bias"])


def test_postnet_against_numpy():
    """PostNet: conv_pre -> LN -> tanh -> three convs back to back -> LN -> tanh -> conv_post -> LN (eps 1e-5)."""
    torch.manual_seed(4)
    m = O.PostNet(8, 16, 5, 5, 0.0).to(F64).eval()
    sd = {k: _np(v) for k, v in m.state_dict().items()}
    x = torch.randn(2, 14, 8, dtype=F64)
    ref = _np(m(x))
    for b in range(2):
        h = conv1d_same_reflect(_np(x[b]), sd["conv_pre.conv.weight"], sd["conv_pre.conv.bias"])
        h = np.tanh(layer_norm(h, sd["ln1.weight"], sd["ln1.bias"], 1e-5))
        for i in range(3):
            h = conv1d_same_reflect(h, sd[f"convs_intermedite.{i}.conv.weight"], sd[f"convs_intermedite.{i}.conv.bias"])
        h = np.tanh(layer_norm(h, sd["ln2.weight"], sd["ln2.bias"], 1e-5))
        h = conv1d_same_reflect(h, sd["conv_post.conv.weight"], sd["conv_post.conv.bias"])
        h = layer_norm(h, sd["ln3.weight"], sd["ln3.bias"], 1e-5)
        assert np.abs(h - ref[b]).max() < 1e-11


def test_positional_encoding_formula():
    """pe[pos, 2i] = sin(pos * den_i), pe[pos, 2i+1] = cos(pos * den_i), den_i = exp(-2i ln(10000) / D).  speechbrain
    builds the table in fp32 (it is a persistent fp32 buffer of the checkpoint), so the argument
    carries a relative error of a few 2^-24: the tolerance grows with pos."""
    pe = _np(O.PositionalEncoding(384).pe[0])
    assert pe.shape == (2500, 384)
    den = np.exp(-np.arange(0, 384, 2) * math.log(10000.0) / 384)
    for pos in (0, 1, 2, 127, 799, 2499):
        tol = 1e-6 + pos * 2.5e-7
        assert np.abs(pe[pos, 0::2] - np.sin(pos * den)).max() < tol
        assert np.abs(pe[pos, 1::2] - np.cos(pos * den)).max() < tol


def test_average_over_durations_brute_force():
    g = torch.Generator().manual_seed(5)
    B, Tp = 3, 9
    dur = torch.randint(0, 5, (B, Tp), generator=g)
    Tm = int(dur.sum(1).max())
    vals = torch.randn(B, 1, Tm, generator=g, dtype=F64)
    vals[torch.rand(B, 1, Tm, generator=g) < 0.3] = 0.0    # unvoiced frames do not count
    ref = _np(O.average_over_durations(vals, dur))
    for b in range(B):
        f = 0
        for p in range(Tp):
            seg = _np(vals[b, 0, f:f + int(dur[b, p])])
            f += int(dur[b, p])
            nz = seg[seg != 0.0]
            want = nz.mean() if nz.size else 0.0
            assert abs(ref[b, 0, p] - want) < 1e-12


def ssim_numpy(x, y):
    """piq-style SSIM of two (H, W) images in [0, 1]: 11x11 Gaussian (sigma 1.5), valid windows, k1 0.01, k2 0.03."""
    c = np.arange(11) - 5.0
    g1 = np.exp(-c ** 2 / (2 * 1.5 ** 2))
    win = np.outer(g1, g1)
    win /= win.sum()
    H, W = x.shape
    vals = []
    for i in range(H - 10):
        for j in range(W - 10):
            a, b = x[i:i + 11, j:j + 11], y[i:i + 11, j:j + 11]
            mx, my = (win * a).sum(), (win * b).sum()
            sxx, syy, sxy = (win * a * a).sum() - mx * mx, (win * b * b).sum() - my * my, (win * a * b).sum() - mx * my
            cs = (2 * sxy + 0.03 ** 2) / (sxx + syy + 0.03 ** 2)
            vals.append((2 * mx * my + 0.01 ** 2) / (mx * mx + my * my + 0.01 ** 2) * cs)
    return float(np.mean(vals))


def test_ssim_loss_against_numpy_windows():
    """speechbrain SSIMLoss: per-sample min-max normalisation over the VALID frames (amax over masked_fill 0, amin over
    masked_fill +inf), padded frames zeroed, SSIM over the whole padded map, 1 - batch mean."""
    g = torch.Generator().manual_seed(6)
    B, T, W = 2, 19, 16
    y_hat = torch.rand(B, T, W, generator=g, dtype=F64) * 4 - 3       # mostly negative, like log-mels
    y = torch.rand(B, T, W, generator=g, dtype=F64) * 4 - 3
    lens = torch.tensor([19, 13])
    ref = float(O.SSIMLoss()(y_hat, y, lens))
    scores = []
    for b in range(B):
        n = int(lens[b])
        imgs = []
        for t in (_np(y[b]), _np(y_hat[b])):
            valid = t[:n]
            mx = valid.max() if n == T else max(valid.max(), 0.0)     # padded rows add a 0 candidate to the max only
            mn = valid.min()
            z = (t - mn) / (mx - mn + 1e-8)
            z[n:] = 0.0
            imgs.append(z)
        scores.append(ssim_numpy(imgs[0], imgs[1]))
    # piq / speechbrain build the Gaussian window in fp32 and cast it to the input dtype: 1e-7-level differences
    assert abs(ref - (1.0 - float(np.mean(scores)))) < 1e-6


def test_mse_terms_brute_force():
    """loss.py:101-160: five per-sample sliced MSEs averaged over the batch; durations in log1p space; pitch / energy
    targets are the phoneme averages returned by the model (Q6) sliced by the MEL length (Q5)."""
    g = torch.Generator().manual_seed(7)
    B, Tp, Tm, M = 3, 6, 15, 12                   # SSIM windows need >= 11 frames and >= 11 bins
    r = lambda *s: torch.randn(*s, generator=g, dtype=F64)  # noqa: E731
    mel_out, post, mel_tgt = r(B, Tm, M), r(B, Tm, M), r(B, Tm, M)
    logd, pp, ap, pe_, ae = r(B, Tp), r(B, Tp, 1), r(B, Tp, 1), r(B, Tp, 1), r(B, Tp, 1)
    dur = torch.randint(0, 4, (B, Tp), generator=g)
    mel_len, phon_len = torch.tensor([15, 12, 11]), torch.tensor([6, 5, 3])
    cfg = dict(O.DEFAULT_LOSS_CONFIG)
    out = O.Loss(**cfg)((mel_out, post, logd, pp, ap, pe_, ae, mel_len),
                        (mel_tgt, dur, r(B, Tm), r(B, Tm), mel_len, phon_len), 0)
    want = dict(mel=0.0, post=0.0, dur=0.0, pitch=0.0, energy=0.0)
    for i in range(B):
        ml, pl = int(mel_len[i]), int(phon_len[i])
        want["mel"] += float(((mel_out[i, :ml] - mel_tgt[i, :ml]) ** 2).mean()) / B
        want["post"] += float(((post[i, :ml] - mel_tgt[i, :ml]) ** 2).mean()) / B
        want["dur"] += float(((logd[i, :pl] - torch.log1p(dur[i, :pl].double())) ** 2).mean()) / B
        want["pitch"] += float(((pp[i, :ml, 0] - ap[i, :ml, 0]) ** 2).mean()) / B      # [:ml] on a Tp-long axis
        want["energy"] += float(((pe_[i, :ml, 0] - ae[i, :ml, 0]) ** 2).mean()) / B
    assert abs(float(out["mel_loss"]) - want["mel"] * cfg["mel_loss_weight"]) < 1e-12
    assert abs(float(out["postnet_mel_loss"]) - want["post"] * cfg["postnet_mel_loss_weight"]) < 1e-12
    assert abs(float(out["dur_loss"]) - want["dur"] * cfg["duration_loss_weight"]) < 1e-12
    assert abs(float(out["pitch_loss"]) - want["pitch"] * cfg["pitch_loss_weight"]) < 1e-12
    assert abs(float(out["energy_loss"]) - want["energy"] * cfg["energy_loss_weight"]) < 1e-12
    parts = sum(float(out[k]) for k in ("ssim_loss", "mel_loss", "postnet_mel_loss", "dur_loss", "pitch_loss", "energy_loss"))
    assert abs(float(out["total_loss"]) - parts) < 1e-12


# ------------------------------------------------------------------ ragged / empty inputs (property tests)
from hypothesis import given, settings, strategies as st  # noqa: E402


@st.composite
def ragged_durations(draw):
    B = draw(st.integers(1, 4))
    Tp = draw(st.integers(1, 9))
    dur = [[draw(st.integers(0, 6)) for _ in range(Tp)] for _ in range(B)]
    if sum(dur[0]) == 0:
        dur[0][0] = 1                                        # pad_sequence needs one non-empty item
    return torch.tensor(dur)


@settings(max_examples=60, deadline=None)
@given(ragged_durations(), st.sampled_from([0.8, 1.0, 1.2]))
def test_upsample_is_repeat_by_truncated_scaled_duration(dur, pace):
    """LengthRegulator: item b, phoneme p occupies (pace * dur).long() frames (fp32 product, truncation), items are
    zero-padded to the longest; utterances may be empty."""
    B, Tp = dur.shape
    feats = torch.arange(B * Tp * 3, dtype=torch.float32).reshape(B, Tp, 3) + 1.0
    out, lens = O.upsample(feats, dur, pace=pace)
    frames = (torch.tensor(pace, dtype=torch.float32) * dur.float()).long()
    assert lens == frames.sum(1).tolist() and out.shape == (B, max(lens), 3)
    for b in range(B):
        f = 0
        for p in range(Tp):
            for _ in range(int(frames[b, p])):
                assert torch.equal(out[b, f], feats[b, p])
                f += 1
        assert (out[b, f:] == 0).all()


@settings(max_examples=60, deadline=None)
@given(ragged_durations(), st.integers(0, 2 ** 31 - 1))
def test_segment_means_on_ragged_batches(dur, seed):
    """average_over_durations (mean over NON-ZERO frames, 0 for empty segments) and the intensity segment mean
    (train.py:16-51: plain mean, 0 for zero-duration phonemes) against per-phoneme loops."""
    B, Tp = dur.shape
    g = torch.Generator().manual_seed(seed)
    Tm = int(dur.sum(1).max())
    vals = torch.randn(B, 1, Tm, generator=g, dtype=F64)
    vals[torch.rand(B, 1, Tm, generator=g) < 0.25] = 0.0
    avg = O.average_over_durations(vals, dur)
    inten = torch.randn(B, Tm, 5, generator=g, dtype=F64)
    phon_len = torch.tensor([Tp] * B)
    rep = O.intensity_segment_mean(inten, phon_len, dur, Tp)
    for b in range(B):
        f = 0
        for p in range(Tp):
            d = int(dur[b, p])
            seg = vals[b, 0, f:f + d]
            nz = seg[seg != 0]
            want = float(nz.mean()) if nz.numel() else 0.0
            assert abs(float(avg[b, 0, p]) - want) < 1e-10
            want_i = inten[b, f:f + d].mean(0) if d else torch.zeros(5, dtype=F64)
            assert torch.allclose(rep[b, p], want_i, atol=1e-12)
            f += d
