"""Drop-in `Loss` for `/root/reference/emo_rank_tts/fastspeech2/loss.py:31-186`: five per-sample sliced MSE
terms + speechbrain SSIMLoss, weighted sum, 7-key dict.  Forward values AND the gradients wrt the five
predictions come out of one C-ABI call (`fs2_loss_fused`: three kernel launches, csrc/losses.cu); the backward is
one more launch that only does work when autograd's upstream gradient is not 1.  No per-sample Python loop, no
host synchronisation, no torch arithmetic kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


_WS_CACHE = {}


def _workspace(dev, B):
    """Zero-initialised scratch of the fused loss; the kernels hand it back zeroed after every call."""
    key = (dev, B)
    ws = _WS_CACHE.get(key)
    if ws is None:
        ws = torch.zeros(int(L.load().fs2_loss_ws_floats(B)), device=dev, dtype=torch.float32)
        _WS_CACHE[key] = ws
    return ws


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mel_out, post_out, log_dur, pitch_pred, energy_pred, pitch_tgt, energy_tgt, mel_tgt, dur_tgt,
                mel_len, phon_len, weights):
        if not mel_out.is_cuda:
            raise RuntimeError("fs2_b200: Loss inputs must be CUDA tensors (there is no CPU fallback)")
        dev = mel_out.device
        B, Tm, n_mels = mel_out.shape
        Tp = log_dur.shape[1]
        if mel_tgt.shape != mel_out.shape:
            raise ValueError(f"mel target shape {tuple(mel_tgt.shape)} != prediction {tuple(mel_out.shape)} "
                             "(the reference's MSELoss would fail the same way)")
        f = lambda t: t.detach().contiguous().float()
        mel_out, post_out, mel_tgt = f(mel_out), f(post_out), f(mel_tgt)
        log_dur, pitch_pred, energy_pred = f(log_dur), f(pitch_pred), f(energy_pred)
        pitch_tgt, energy_tgt = f(pitch_tgt), f(energy_tgt)
        dur_tgt = dur_tgt.contiguous().long().to(dev)
        mel_len = mel_len.contiguous().long().to(dev)
        phon_len = phon_len.contiguous().long().to(dev)
        w_ssim, w_mel, w_post, w_dur, w_pitch, w_energy = weights
        wh = (L.C.c_float * 6)(w_mel, w_post, w_dur, w_pitch, w_energy, w_ssim)
        need = any(ctx.needs_input_grad[:5])
        out = torch.empty(8, device=dev, dtype=torch.float32)
        dmel = torch.empty_like(mel_out) if need else None
        dpost = torch.empty_like(post_out) if need else None
        dph = torch.empty(3, B, Tp, device=dev) if need else None
        lib = L.load()
        P = lambda t: t.data_ptr() if t is not None else None
        rc = lib.fs2_loss_fused(P(mel_out), P(post_out), P(mel_tgt), P(log_dur), P(dur_tgt), P(pitch_pred), P(pitch_tgt),
                                P(energy_pred), P(energy_tgt), P(mel_len), P(phon_len), B, Tp, Tm, n_mels,
                                L.C.cast(wh, L.C.c_void_p), P(_workspace(dev, B)), P(out), P(dmel), P(dpost),
                                P(dph[0]) if need else None, P(dph[1]) if need else None, P(dph[2]) if need else None,
                                torch.cuda.current_stream().cuda_stream)
        L.check(rc, "fs2_loss_fused")
        ctx.grads = (dmel, dpost, dph)
        # out[0..4] = weighted mel, postnet, dur, pitch, energy; out[5] = weighted ssim; out[6] = total_loss
        total, comps = out[6], out[:6]
        ctx.mark_non_differentiable(comps)
        return total, comps

    @staticmethod
    def backward(ctx, g_total, _g_comps):
        dmel, dpost, dph = ctx.grads
        if dmel is None:
            return (None,) * 12
        # the gradients were produced for an upstream gradient of 1; anything else is applied in place (one launch that
        # returns at once when *g_total == 1, the `total_loss.backward()` case)
        L.call("fs2_loss_scale_grads", g_total.detach().contiguous().float(), dmel, dpost, dmel.numel(), dph[0], dph[1],
               dph[2], dph[0].numel())
        return (dmel, dpost, dph[0], dph[1], dph[2], None, None, None, None, None, None, None)


class Loss(nn.Module):
    """Same constructor kwargs as the reference (parameter.yaml:96-106), same call and dict keys."""

    def __init__(self, log_scale_durations, ssim_loss_weight, duration_loss_weight, pitch_loss_weight,
                 energy_loss_weight, mel_loss_weight, postnet_mel_loss_weight, spn_loss_weight=1.0,
                 spn_loss_max_epochs=8):
        super().__init__()
        if not log_scale_durations:
            raise NotImplementedError("fs2_b200: only log_scale_durations=True (parameter.yaml:97; the reference "
                                      "leaves log_target_durations undefined otherwise, loss.py:108-125)")
        self.log_scale_durations = log_scale_durations
        self.ssim_loss_weight = ssim_loss_weight
        self.mel_loss_weight = mel_loss_weight
        self.postnet_mel_loss_weight = postnet_mel_loss_weight
        self.duration_loss_weight = duration_loss_weight
        self.pitch_loss_weight = pitch_loss_weight
        self.energy_loss_weight = energy_loss_weight
        self.spn_loss_weight = spn_loss_weight           # accepted and unused, as in the reference (loss.py:40-41)
        self.spn_loss_max_epochs = spn_loss_max_epochs

    def forward(self, predictions, targets, current_epoch=0):
        mel_target, target_durations, _target_pitch, _target_energy, mel_length, phon_len = targets
        assert len(mel_target.shape) == 3
        (mel_out, postnet_mel_out, log_durations, predicted_pitch, average_pitch, predicted_energy, average_energy,
         _mel_lens) = predictions
        if average_pitch is None or average_energy is None:
            raise ValueError("Loss needs the model's phoneme-averaged pitch/energy (teacher-forced forward)")
        B, Tp = log_durations.shape[0], log_durations.shape[-1] if log_durations.dim() == 2 else log_durations.shape[1]
        sq = lambda t: t.reshape(B, -1)
        weights = (float(self.ssim_loss_weight), float(self.mel_loss_weight), float(self.postnet_mel_loss_weight),
                   float(self.duration_loss_weight), float(self.pitch_loss_weight), float(self.energy_loss_weight))
        total, comps = _LossFn.apply(mel_out, postnet_mel_out, sq(log_durations), sq(predicted_pitch),
                                     sq(predicted_energy), sq(average_pitch), sq(average_energy), mel_target,
                                     target_durations, mel_length, phon_len, weights)
        return {
            "total_loss": total,
            "ssim_loss": comps[5],
            "mel_loss": comps[0],
            "postnet_mel_loss": comps[1],
            "dur_loss": comps[2],
            "pitch_loss": comps[3],
            "energy_loss": comps[4],
        }
