"""Drop-in `Loss` for `/root/reference/emo_rank_tts/fastspeech2/loss.py:31-186`: five per-sample sliced MSE
terms + speechbrain SSIMLoss, weighted sum, 7-key dict.  Forward values and the gradients wrt the five
predictions come out of two fused CUDA passes (fs2_mse_losses, fs2_ssim_loss); no per-sample Python loop,
no host synchronisation."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


_WEIGHT_CACHE = {}


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
        wh = (L.C.c_float * 5)(w_mel, w_post, w_dur, w_pitch, w_energy)
        need = any(ctx.needs_input_grad[:5])
        out = torch.empty(8, device=dev, dtype=torch.float32)
        dmel = torch.empty_like(mel_out) if need else None
        dpost = torch.empty_like(post_out) if need else None
        ddur = torch.empty(B, Tp, device=dev) if need else None
        dpitch = torch.empty(B, Tp, device=dev) if need else None
        denergy = torch.empty(B, Tp, device=dev) if need else None
        sums = torch.empty(5 * B, device=dev, dtype=torch.float32)
        lib = L.load()
        rc = lib.fs2_mse_losses(mel_out.data_ptr(), post_out.data_ptr(), mel_tgt.data_ptr(), log_dur.data_ptr(),
                                dur_tgt.data_ptr(), pitch_pred.data_ptr(), pitch_tgt.data_ptr(), energy_pred.data_ptr(),
                                energy_tgt.data_ptr(), mel_len.data_ptr(), phon_len.data_ptr(), B, Tp, Tm, n_mels,
                                L.C.cast(wh, L.C.c_void_p), sums.data_ptr(), out.data_ptr(),
                                dmel.data_ptr() if need else None, dpost.data_ptr() if need else None,
                                ddur.data_ptr() if need else None, dpitch.data_ptr() if need else None,
                                denergy.data_ptr() if need else None, torch.cuda.current_stream().cuda_stream)
        L.check(rc, "fs2_mse_losses")
        ws = torch.empty(int(lib.fs2_ssim_ws_floats(B, Tm, n_mels)) + 4, device=dev, dtype=torch.float32)
        L.call("fs2_ssim_loss", mel_out, mel_tgt, mel_len, B, Tm, n_mels, float(w_ssim), out[5:], dmel, ws)
        ctx.grads = (dmel, dpost, ddur, dpitch, denergy)
        ctx.shapes = None
        # out[0..4] = mel, postnet, dur, pitch, energy (un-weighted); out[5] = ssim
        key = (weights, dev)
        wv = _WEIGHT_CACHE.get(key)
        if wv is None:
            wv = torch.tensor([w_mel, w_post, w_dur, w_pitch, w_energy, w_ssim], device=dev, dtype=torch.float32)
            _WEIGHT_CACHE[key] = wv
        comps = out[:6] * wv
        total = comps.sum()
        ctx.mark_non_differentiable(comps)
        return total, comps

    @staticmethod
    def backward(ctx, g_total, _g_comps):
        dmel, dpost, ddur, dpitch, denergy = ctx.grads
        return (dmel * g_total, dpost * g_total, ddur * g_total, (dpitch * g_total), (denergy * g_total),
                None, None, None, None, None, None, None)


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
