"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see below).

CPU/eager restatement of the reference FastSpeech2 hot path
(`/root/reference/emo_rank_tts/fastspeech2/model.py:149-441` and
`.../loss.py:31-186`) together with the speechbrain building blocks those two
files import (model.py:13-27, loss.py:3).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this module -- as the checker (or the
thing timed as "the reference's CPU path"), never as the product path.

Why "parity unpinned": the arithmetic of the reference lives in the third-party
package `speechbrain`, which the reference leaves un-vendored and un-pinned
(`requirements.txt:3`, bare name; must be >= 1.0.0 because train.py:7 imports
`speechbrain.inference`).  speechbrain is not installed in this image and there
is no network, and the reference ships no tests / golden vectors for this path
(SURVEY.md section 4, 8c).  The speechbrain semantics below are restated from the
published speechbrain 1.0.x sources (`speechbrain/nnet/{CNN,linear,embedding,
normalization,attention}.py`, `speechbrain/lobes/models/FastSpeech2.py`,
`speechbrain/lobes/models/transformer/Transformer.py`); every numeric op is a
stock torch op exactly as in speechbrain, so the oracle's arithmetic *is*
torch's.  The pins that do exist are (a) the shape doctest in model.py:102-146
and (b) the reference's own call sites; both are exercised in
`tests/test_oracle.py`, and seeded outputs of this oracle are frozen under
`tests/golden/` by `tests/golden/make_golden.py`.  In addition the glue half
of this file (class FastSpeech2, class Loss) IS pinned: `tests/reference_glue.py`
executes the reference's model.py / loss.py unmodified with the leaf classes
below standing in for speechbrain, and `tests/test_reference_glue.py` requires
the same state_dict (keys, order, seeded init), outputs, losses and gradients.
The leaves stay unpinned; `tests/test_oracle_first_principles.py` re-derives
each of them in numpy float64 from the published definitions.

state_dict layout is key-for-key the reference's.
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# speechbrain.nnet restatements
# ----------------------------------------------------------------------------
class SBLinear(nn.Module):
    """speechbrain.nnet.linear.Linear: wraps nn.Linear under attribute `w`."""

    def __init__(self, n_neurons, input_size, bias=True):
        super().__init__()
        self.w = nn.Linear(input_size, n_neurons, bias=bias)

    def forward(self, x):
        return self.w(x)


class SBEmbedding(nn.Module):
    """speechbrain.nnet.embedding.Embedding: nn.Embedding *without* padding_idx
    under attribute `Embedding` (row `blank_id` is trainable, non-zero)."""

    def __init__(self, num_embeddings, embedding_dim=128, blank_id=0):
        super().__init__()
        self.Embedding = nn.Embedding(num_embeddings, embedding_dim)

    def forward(self, x):
        return self.Embedding(x.long())


class SBLayerNorm(nn.Module):
    """speechbrain.nnet.normalization.LayerNorm: nn.LayerNorm under `norm`."""

    def __init__(self, input_size, eps=1e-05):
        super().__init__()
        self.norm = nn.LayerNorm(input_size, eps=eps, elementwise_affine=True)

    def forward(self, x):
        return self.norm(x)


class SBConv1d(nn.Module):
    """speechbrain.nnet.CNN.Conv1d with padding="same", stride 1, dilation 1.

    Input (B, T, C) unless skip_transpose (then (B, C, T)).  "same" padding is
    `F.pad(x, (p, p), mode="reflect")` with p = (k-1)//2, applied at the edge of
    the padded *rectangle*, followed by nn.Conv1d(padding=0) with bias.
    """

    def __init__(self, in_channels, out_channels, kernel_size, skip_transpose=False):
        super().__init__()
        self.kernel_size = kernel_size
        self.skip_transpose = skip_transpose
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, padding=0)

    def forward(self, x):
        if not self.skip_transpose:
            x = x.transpose(1, -1)
        p = (self.kernel_size - 1) // 2
        if p > 0:
            x = F.pad(x, (p, p), mode="reflect")
        y = self.conv(x)
        if not self.skip_transpose:
            y = y.transpose(1, -1)
        return y


class SBMultiheadAttention(nn.Module):
    """speechbrain.nnet.attention.MultiheadAttention: nn.MultiheadAttention
    (seq-first) under `att`; called with need_weights=True (math path)."""

    def __init__(self, nhead, d_model, dropout=0.0, kdim=None, vdim=None):
        super().__init__()
        self.att = nn.MultiheadAttention(
            embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=True,
            kdim=kdim, vdim=vdim,
        )

    def forward(self, query, key, value, attn_mask=None, key_padding_mask=None):
        query = query.permute(1, 0, 2)
        key = key.permute(1, 0, 2)
        value = value.permute(1, 0, 2)
        output, w = self.att(
            query, key, value, attn_mask=attn_mask,
            key_padding_mask=key_padding_mask, need_weights=True,
        )
        return output.permute(1, 0, 2), w


# ----------------------------------------------------------------------------
# speechbrain.lobes.models.transformer.Transformer restatements
# ----------------------------------------------------------------------------
class PositionalEncoding(nn.Module):
    def __init__(self, input_size, max_len=2500):
        super().__init__()
        self.max_len = max_len
        pe = torch.zeros(self.max_len, input_size, requires_grad=False)
        positions = torch.arange(0, self.max_len).unsqueeze(1).float()
        denominator = torch.exp(
            torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size)
        )
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)].clone().detach()


def get_key_padding_mask(padded_input, pad_idx):
    return padded_input.eq(pad_idx).detach()


def get_mask_from_lengths(lengths, max_len=None):
    if max_len is None:
        max_len = torch.max(lengths).item()
    ids = torch.arange(0, max_len, device=lengths.device, dtype=lengths.dtype)
    return ~(ids < lengths.unsqueeze(1)).bool()


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_ffn, nhead, d_model, kdim, vdim, dropout,
                 normalize_before, ffn_cnn_kernel_size_list):
        super().__init__()
        self.self_att = SBMultiheadAttention(nhead, d_model, dropout, kdim, vdim)
        self.pos_ffn = nn.Sequential(
            SBConv1d(d_model, d_ffn, ffn_cnn_kernel_size_list[0]),
            nn.ReLU(),
            SBConv1d(d_ffn, d_model, ffn_cnn_kernel_size_list[1]),
        )
        self.norm1 = SBLayerNorm(d_model, eps=1e-6)
        self.norm2 = SBLayerNorm(d_model, eps=1e-6)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.normalize_before = normalize_before

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        src1 = self.norm1(src) if self.normalize_before else src
        output, self_attn = self.self_att(
            src1, src1, src1, attn_mask=src_mask, key_padding_mask=src_key_padding_mask
        )
        src = src + self.dropout1(output)
        if not self.normalize_before:
            src = self.norm1(src)
        src1 = self.norm2(src) if self.normalize_before else src
        output = self.pos_ffn(src1)
        output = src + self.dropout2(output)
        if not self.normalize_before:
            output = self.norm2(output)
        return output, self_attn


class TransformerEncoder(nn.Module):
    def __init__(self, num_layers, nhead, d_ffn, d_model, kdim, vdim, dropout,
                 normalize_before, ffn_cnn_kernel_size_list):
        super().__init__()
        self.layers = nn.ModuleList(
            [
                TransformerEncoderLayer(d_ffn, nhead, d_model, kdim, vdim, dropout,
                                        normalize_before, ffn_cnn_kernel_size_list)
                for _ in range(num_layers)
            ]
        )
        self.norm = SBLayerNorm(d_model, eps=1e-6)

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        output = src
        attention_lst = []
        for enc_layer in self.layers:
            output, attention = enc_layer(
                output, src_mask=src_mask, src_key_padding_mask=src_key_padding_mask
            )
            attention_lst.append(attention)
        output = self.norm(output)
        return output, attention_lst


# ----------------------------------------------------------------------------
# speechbrain.lobes.models.FastSpeech2 restatements
# ----------------------------------------------------------------------------
class EncoderPreNet(nn.Module):
    def __init__(self, n_vocab, blank_id, out_channels=512):
        super().__init__()
        self.token_embedding = SBEmbedding(n_vocab, out_channels, blank_id)

    def forward(self, x):
        return self.token_embedding(x)


class DurationPredictor(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, dropout=0.0, n_units=1):
        super().__init__()
        self.conv1 = SBConv1d(in_channels, out_channels, kernel_size)
        self.conv2 = SBConv1d(out_channels, out_channels, kernel_size)
        self.linear = SBLinear(n_units, out_channels)
        self.ln1 = SBLayerNorm(out_channels)
        self.ln2 = SBLayerNorm(out_channels)
        self.relu = nn.ReLU()
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, x, x_mask):
        x = self.relu(self.conv1(x * x_mask))
        x = self.ln1(x).to(x.dtype)
        x = self.dropout1(x)
        x = self.relu(self.conv2(x * x_mask))
        x = self.ln2(x).to(x.dtype)
        x = self.dropout2(x)
        return self.linear(x * x_mask)


class PostNet(nn.Module):
    def __init__(self, n_mel_channels=80, postnet_embedding_dim=512,
                 postnet_kernel_size=5, postnet_n_convolutions=5, postnet_dropout=0.5):
        super().__init__()
        self.conv_pre = SBConv1d(n_mel_channels, postnet_embedding_dim, postnet_kernel_size)
        self.convs_intermedite = nn.ModuleList(
            [
                SBConv1d(postnet_embedding_dim, postnet_embedding_dim, postnet_kernel_size)
                for _ in range(1, postnet_n_convolutions - 1)
            ]
        )
        self.conv_post = SBConv1d(postnet_embedding_dim, n_mel_channels, postnet_kernel_size)
        self.tanh = nn.Tanh()
        self.ln1 = nn.LayerNorm(postnet_embedding_dim)
        self.ln2 = nn.LayerNorm(postnet_embedding_dim)
        self.ln3 = nn.LayerNorm(n_mel_channels)
        self.dropout1 = nn.Dropout(postnet_dropout)
        self.dropout2 = nn.Dropout(postnet_dropout)
        self.dropout3 = nn.Dropout(postnet_dropout)

    def forward(self, x):
        x = self.conv_pre(x)
        x = self.ln1(x).to(x.dtype)
        x = self.tanh(x)
        x = self.dropout1(x)
        for conv in self.convs_intermedite:
            x = conv(x)
        x = self.ln2(x).to(x.dtype)
        x = self.tanh(x)
        x = self.dropout2(x)
        x = self.conv_post(x)
        x = self.ln3(x).to(x.dtype)
        x = self.dropout3(x)
        return x


def upsample(feats, durs, pace=1.0, padding_value=0.0):
    upsampled = [
        torch.repeat_interleave(feats[i], (pace * durs[i]).long(), dim=0)
        for i in range(len(durs))
    ]
    mel_lens = [m.shape[0] for m in upsampled]
    padded = torch.nn.utils.rnn.pad_sequence(
        upsampled, batch_first=True, padding_value=padding_value
    )
    return padded, mel_lens


def average_over_durations(values, durs):
    durs_cums_ends = torch.cumsum(durs, dim=1).long()
    durs_cums_starts = F.pad(durs_cums_ends[:, :-1], (1, 0))
    values_nonzero_cums = F.pad(torch.cumsum(values != 0.0, dim=2), (1, 0))
    values_cums = F.pad(torch.cumsum(values, dim=2), (1, 0))
    bs, length = durs_cums_ends.size()
    n_formants = values.size(1)
    dcs = durs_cums_starts[:, None, :].expand(bs, n_formants, length)
    dce = durs_cums_ends[:, None, :].expand(bs, n_formants, length)
    values_sums = (torch.gather(values_cums, 2, dce) - torch.gather(values_cums, 2, dcs)).to(values.dtype)
    values_nelems = (
        torch.gather(values_nonzero_cums, 2, dce) - torch.gather(values_nonzero_cums, 2, dcs)
    ).to(values.dtype)
    return torch.where(values_nelems == 0.0, values_nelems, values_sums / values_nelems)


class _SSIMLoss(nn.Module):
    """piq-style SSIM as vendored by speechbrain (kernel 11, sigma 1.5)."""

    def __init__(self, kernel_size=11, kernel_sigma=1.5, k1=0.01, k2=0.03,
                 downsample=True, data_range=1.0):
        super().__init__()
        self.kernel_size, self.kernel_sigma = kernel_size, kernel_sigma
        self.k1, self.k2, self.downsample, self.data_range = k1, k2, downsample, data_range

    @staticmethod
    def gaussian_filter(kernel_size, sigma, dtype=torch.float32):
        coords = torch.arange(kernel_size, dtype=dtype)
        coords -= (kernel_size - 1) / 2.0
        g = coords ** 2
        g = (-(g.unsqueeze(0) + g.unsqueeze(1)) / (2 * sigma ** 2)).exp()
        g /= g.sum()
        return g.unsqueeze(0)

    def forward(self, x, y):
        x = x / float(self.data_range)
        y = y / float(self.data_range)
        f = max(1, round(min(x.size()[-2:]) / 256))
        if f > 1 and self.downsample:
            x = F.avg_pool2d(x, kernel_size=f)
            y = F.avg_pool2d(y, kernel_size=f)
        kernel = self.gaussian_filter(self.kernel_size, self.kernel_sigma).repeat(x.size(1), 1, 1, 1).to(y)
        c1, c2 = self.k1 ** 2, self.k2 ** 2
        n = x.size(1)
        mu_x = F.conv2d(x, weight=kernel, stride=1, padding=0, groups=n)
        mu_y = F.conv2d(y, weight=kernel, stride=1, padding=0, groups=n)
        mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
        sigma_xx = F.conv2d(x ** 2, weight=kernel, stride=1, padding=0, groups=n) - mu_xx
        sigma_yy = F.conv2d(y ** 2, weight=kernel, stride=1, padding=0, groups=n) - mu_yy
        sigma_xy = F.conv2d(x * y, weight=kernel, stride=1, padding=0, groups=n) - mu_xy
        cs = (2.0 * sigma_xy + c2) / (sigma_xx + sigma_yy + c2)
        ss = (2.0 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
        ssim_val = ss.mean(dim=(-1, -2)).mean(1)
        score = ssim_val.mean(dim=0)
        return torch.ones_like(score) - score


class SSIMLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.loss_func = _SSIMLoss()

    @staticmethod
    def sequence_mask(sequence_length, max_len=None):
        if max_len is None:
            max_len = sequence_length.data.max()
        seq_range = torch.arange(max_len, dtype=sequence_length.dtype, device=sequence_length.device)
        return seq_range.unsqueeze(0) < sequence_length.unsqueeze(1)

    @staticmethod
    def sample_wise_min_max(x, mask):
        maximum = torch.amax(x.masked_fill(~mask, 0), dim=(1, 2), keepdim=True)
        minimum = torch.amin(x.masked_fill(~mask, np.inf), dim=(1, 2), keepdim=True)
        return (x - minimum) / (maximum - minimum + 1e-8)

    def forward(self, y_hat, y, length):
        mask = self.sequence_mask(sequence_length=length, max_len=y.size(1)).unsqueeze(2)
        y_norm = self.sample_wise_min_max(y, mask)
        y_hat_norm = self.sample_wise_min_max(y_hat, mask)
        ssim_loss = self.loss_func((y_norm * mask).unsqueeze(1), (y_hat_norm * mask).unsqueeze(1))
        if ssim_loss.item() > 1.0:
            ssim_loss = torch.tensor(1.0, device=ssim_loss.device, dtype=ssim_loss.dtype)
        if ssim_loss.item() < 0.0:
            ssim_loss = torch.tensor(0.0, device=ssim_loss.device, dtype=ssim_loss.dtype)
        return ssim_loss


# ----------------------------------------------------------------------------
# reference glue: fastspeech2/model.py:149-441
# ----------------------------------------------------------------------------
class FastSpeech2(nn.Module):
    def __init__(
        self, enc_num_layers, enc_num_head, enc_d_model, enc_ffn_dim, enc_k_dim,
        enc_v_dim, enc_dropout, dec_num_layers, dec_num_head, dec_d_model,
        dec_ffn_dim, dec_k_dim, dec_v_dim, dec_dropout, normalize_before, ffn_type,
        ffn_cnn_kernel_size_list, n_char, n_mels, postnet_embedding_dim,
        postnet_kernel_size, postnet_n_convolutions, postnet_dropout, padding_idx,
        dur_pred_kernel_size, pitch_pred_kernel_size, energy_pred_kernel_size,
        variance_predictor_dropout, n_speakers,
    ):
        super().__init__()
        assert ffn_type == "1dcnn"
        self.enc_num_head = enc_num_head
        self.dec_num_head = dec_num_head
        self.padding_idx = padding_idx
        self.sinusoidal_positional_embed_encoder = PositionalEncoding(enc_d_model)
        self.sinusoidal_positional_embed_decoder = PositionalEncoding(dec_d_model)
        self.speaker_emb = SBEmbedding(n_speakers, enc_d_model)
        self.concat_proj = SBLinear(enc_d_model, enc_d_model + enc_d_model + 5, bias=False)
        self.encPreNet = EncoderPreNet(n_char, padding_idx, out_channels=enc_d_model)
        # NB model.py:211,217,223 -- all three predictors use dur_pred_kernel_size (quirk Q4)
        self.durPred = DurationPredictor(enc_d_model, enc_d_model, dur_pred_kernel_size, variance_predictor_dropout)
        self.pitchPred = DurationPredictor(enc_d_model, enc_d_model, dur_pred_kernel_size, variance_predictor_dropout)
        self.energyPred = DurationPredictor(enc_d_model, enc_d_model, dur_pred_kernel_size, variance_predictor_dropout)
        self.pitchEmbed = SBConv1d(1, enc_d_model, pitch_pred_kernel_size, skip_transpose=True)
        self.energyEmbed = SBConv1d(1, enc_d_model, energy_pred_kernel_size, skip_transpose=True)
        self.encoder = TransformerEncoder(enc_num_layers, enc_num_head, enc_ffn_dim, enc_d_model,
                                          enc_k_dim, enc_v_dim, enc_dropout, normalize_before,
                                          ffn_cnn_kernel_size_list)
        self.decoder = TransformerEncoder(dec_num_layers, dec_num_head, dec_ffn_dim, dec_d_model,
                                          dec_k_dim, dec_v_dim, dec_dropout, normalize_before,
                                          ffn_cnn_kernel_size_list)
        self.linear = SBLinear(n_mels, dec_d_model)
        self.postnet = PostNet(n_mels, postnet_embedding_dim, postnet_kernel_size,
                               postnet_n_convolutions, postnet_dropout)
        self.trace = None  # optional dict filled with intermediates (debug aid for tests)

    def _t(self, name, value):
        if self.trace is not None:
            self.trace[name] = value.detach().clone()

    def forward(self, tokens, speakers, durations=None, pitch=None, energy=None,
                pace=1.0, pitch_rate=1.0, energy_rate=1.0, intensity=None):
        srcmask = get_key_padding_mask(tokens, pad_idx=self.padding_idx)
        srcmask_inverted = (~srcmask).unsqueeze(-1)

        token_feats = self.encPreNet(tokens)
        pos = self.sinusoidal_positional_embed_encoder(token_feats)
        token_feats = torch.add(token_feats, pos) * srcmask_inverted
        self._t("enc_in", token_feats)
        attn_mask = (
            srcmask.unsqueeze(-1).repeat(self.enc_num_head, 1, token_feats.shape[1])
            .permute(0, 2, 1).bool()
        )
        token_feats, _ = self.encoder(token_feats, src_mask=attn_mask, src_key_padding_mask=srcmask)
        token_feats = token_feats * srcmask_inverted
        self._t("enc_out", token_feats)

        B, T, D = token_feats.shape
        speaker_emb = self.speaker_emb(speakers).unsqueeze(1).expand(-1, T, -1)
        x = torch.cat([token_feats, speaker_emb, intensity], dim=-1)
        token_feats = self.concat_proj(x)
        token_feats = token_feats * srcmask_inverted
        self._t("cond", token_feats)

        predict_durations = self.durPred(token_feats, srcmask_inverted).squeeze(-1)
        if predict_durations.dim() == 1:
            predict_durations = predict_durations.unsqueeze(0)
        if durations is None:
            dur_pred_reverse_log = torch.clamp(torch.special.expm1(predict_durations), 0)

        avg_pitch = None
        predict_pitch = self.pitchPred(token_feats, srcmask_inverted)
        predict_pitch = predict_pitch * pitch_rate
        if pitch is not None:
            avg_pitch = average_over_durations(pitch.unsqueeze(1), durations)
            pitch = self.pitchEmbed(avg_pitch)
            avg_pitch = avg_pitch.permute(0, 2, 1)
        else:
            pitch = self.pitchEmbed(predict_pitch.permute(0, 2, 1))
        pitch = pitch.permute(0, 2, 1)
        token_feats = token_feats.add(pitch)
        self._t("after_pitch", token_feats)

        avg_energy = None
        predict_energy = self.energyPred(token_feats, srcmask_inverted)
        predict_energy = predict_energy * energy_rate
        if energy is not None:
            avg_energy = average_over_durations(energy.unsqueeze(1), durations)
            energy = self.energyEmbed(avg_energy)
            avg_energy = avg_energy.permute(0, 2, 1)
        else:
            energy = self.energyEmbed(predict_energy.permute(0, 2, 1))
        energy = energy.permute(0, 2, 1)
        token_feats = token_feats.add(energy)
        self._t("after_energy", token_feats)

        spec_feats, mel_lens = upsample(
            token_feats, durations if durations is not None else dur_pred_reverse_log, pace=pace
        )
        srcmask = get_mask_from_lengths(torch.tensor(mel_lens)).to(spec_feats.device)
        srcmask_inverted = (~srcmask).unsqueeze(-1)
        attn_mask = (
            srcmask.unsqueeze(-1).repeat(self.dec_num_head, 1, spec_feats.shape[1])
            .permute(0, 2, 1).bool()
        )
        pos = self.sinusoidal_positional_embed_decoder(spec_feats)
        spec_feats = torch.add(spec_feats, pos) * srcmask_inverted
        self._t("dec_in", spec_feats)

        output_mel_feats, memory, *_ = self.decoder(
            spec_feats, src_mask=attn_mask, src_key_padding_mask=srcmask
        )
        self._t("dec_out", output_mel_feats)
        mel_post = self.linear(output_mel_feats) * srcmask_inverted
        postnet_output = self.postnet(mel_post) + mel_post
        return (
            mel_post, postnet_output, predict_durations, predict_pitch, avg_pitch,
            predict_energy, avg_energy, torch.tensor(mel_lens),
        )


# ----------------------------------------------------------------------------
# reference glue: fastspeech2/loss.py:31-186
# ----------------------------------------------------------------------------
class Loss(nn.Module):
    def __init__(self, log_scale_durations, ssim_loss_weight, duration_loss_weight,
                 pitch_loss_weight, energy_loss_weight, mel_loss_weight,
                 postnet_mel_loss_weight, spn_loss_weight=1.0, spn_loss_max_epochs=8):
        super().__init__()
        self.ssim_loss = SSIMLoss()
        self.mel_loss = nn.MSELoss()
        self.postnet_mel_loss = nn.MSELoss()
        self.dur_loss = nn.MSELoss()
        self.pitch_loss = nn.MSELoss()
        self.energy_loss = nn.MSELoss()
        self.log_scale_durations = log_scale_durations
        self.ssim_loss_weight = ssim_loss_weight
        self.mel_loss_weight = mel_loss_weight
        self.postnet_mel_loss_weight = postnet_mel_loss_weight
        self.duration_loss_weight = duration_loss_weight
        self.pitch_loss_weight = pitch_loss_weight
        self.energy_loss_weight = energy_loss_weight
        self.spn_loss_weight = spn_loss_weight
        self.spn_loss_max_epochs = spn_loss_max_epochs

    def forward(self, predictions, targets, current_epoch):
        mel_target, target_durations, target_pitch, target_energy, mel_length, phon_len = targets
        assert len(mel_target.shape) == 3
        (mel_out, postnet_mel_out, log_durations, predicted_pitch, average_pitch,
         predicted_energy, average_energy, mel_lens) = predictions
        predicted_pitch = predicted_pitch.squeeze(-1)
        predicted_energy = predicted_energy.squeeze(-1)
        target_pitch = average_pitch.squeeze(-1)       # loss.py:104 -- batch targets overwritten (Q6)
        target_energy = average_energy.squeeze(-1)
        log_durations = log_durations.squeeze(-1)
        if self.log_scale_durations:
            log_target_durations = torch.log1p(target_durations.to(mel_out.dtype))
        n = mel_target.shape[0]
        mel_loss = postnet_mel_loss = dur_loss = pitch_loss = energy_loss = 0.0
        for i in range(n):
            ml, pl = int(mel_length[i]), int(phon_len[i])
            mel_loss = mel_loss + self.mel_loss(mel_out[i, :ml, :], mel_target[i, :ml, :])
            postnet_mel_loss = postnet_mel_loss + self.postnet_mel_loss(
                postnet_mel_out[i, :ml, :], mel_target[i, :ml, :])
            dur_loss = dur_loss + self.dur_loss(log_durations[i, :pl], log_target_durations[i, :pl])
            # loss.py:126-133 -- phoneme axis sliced with the *mel* length (quirk Q5)
            pitch_loss = pitch_loss + self.pitch_loss(predicted_pitch[i, :ml], target_pitch[i, :ml])
            energy_loss = energy_loss + self.energy_loss(predicted_energy[i, :ml], target_energy[i, :ml])
        ssim_loss = self.ssim_loss(mel_out, mel_target, mel_length)
        mel_loss = torch.div(mel_loss, n)
        postnet_mel_loss = torch.div(postnet_mel_loss, n)
        dur_loss = torch.div(dur_loss, n)
        pitch_loss = torch.div(pitch_loss, n)
        energy_loss = torch.div(energy_loss, n)
        total_loss = (
            ssim_loss * self.ssim_loss_weight
            + mel_loss * self.mel_loss_weight
            + postnet_mel_loss * self.postnet_mel_loss_weight
            + dur_loss * self.duration_loss_weight
            + pitch_loss * self.pitch_loss_weight
            + energy_loss * self.energy_loss_weight
        )
        return {
            "total_loss": total_loss,
            "ssim_loss": ssim_loss * self.ssim_loss_weight,
            "mel_loss": mel_loss * self.mel_loss_weight,
            "postnet_mel_loss": postnet_mel_loss * self.postnet_mel_loss_weight,
            "dur_loss": dur_loss * self.duration_loss_weight,
            "pitch_loss": pitch_loss * self.pitch_loss_weight,
            "energy_loss": energy_loss * self.energy_loss_weight,
        }


# ----------------------------------------------------------------------------
# train.py:16-51 (the "next" row f-1): duration-segment mean of frame intensities
# ----------------------------------------------------------------------------
def intensity_segment_mean(I, phon_len, duration_tgt, T_phon_max):
    B, _, D = I.shape
    out = torch.zeros((B, T_phon_max, D), dtype=I.dtype, device=I.device)
    for b in range(B):
        T_phon = int(phon_len[b])
        durations = duration_tgt[b].long()[:T_phon]
        T_mel = int(durations.sum())
        I_b = I[b, :T_mel, :]
        phon_idx = torch.repeat_interleave(torch.arange(T_phon, device=I.device), durations)
        sum_rep = torch.zeros((T_phon, D), dtype=I.dtype, device=I.device)
        sum_rep.index_add_(0, phon_idx, I_b)
        denom = durations.unsqueeze(1).to(I.dtype).clamp(min=1.0)
        out[b, :T_phon, :] = sum_rep / denom
    return out


DEFAULT_MODEL_CONFIG = dict(
    enc_num_layers=6, enc_num_head=2, enc_d_model=384, enc_ffn_dim=1536, enc_k_dim=384,
    enc_v_dim=384, enc_dropout=0.1, dec_num_layers=6, dec_num_head=2, dec_d_model=384,
    dec_ffn_dim=1536, dec_k_dim=384, dec_v_dim=384, dec_dropout=0.1, normalize_before=False,
    ffn_type="1dcnn", ffn_cnn_kernel_size_list=[9, 1], n_char=95, n_mels=80,
    postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5,
    postnet_dropout=0.5, padding_idx=0, dur_pred_kernel_size=3, pitch_pred_kernel_size=3,
    energy_pred_kernel_size=3, variance_predictor_dropout=0.5,
)  # fastspeech2/parameter.yaml:62-90

DEFAULT_LOSS_CONFIG = dict(
    log_scale_durations=True, ssim_loss_weight=1.0, duration_loss_weight=1.0,
    pitch_loss_weight=1.0, energy_loss_weight=1.0, mel_loss_weight=1.0,
    postnet_mel_loss_weight=1.0, spn_loss_weight=0.0, spn_loss_max_epochs=1,
)  # fastspeech2/parameter.yaml:96-106


def build(seed=0, n_speakers=4, dtype=torch.float32, **overrides):
    cfg = dict(DEFAULT_MODEL_CONFIG)
    cfg.update(overrides)
    torch.manual_seed(seed)
    return FastSpeech2(**cfg, n_speakers=n_speakers).to(dtype)
