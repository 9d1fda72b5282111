"""Per-phoneme emotion-intensity representation: the step immediately upstream of every training step
(`/root/reference/emo_rank_tts/fastspeech2/train.py:16-51`).

The reference runs the frozen extractor and then loops over the batch in Python -- `.item()` syncs,
`repeat_interleave`, `index_add_`, divide by `clamp(dur, 1)`.  Here the segmented mean of the (B, Tm, D) frame
intensities over the durations is ONE kernel (`fs2_intensity_segment_mean`): no host sync, no per-sample launch."""
from __future__ import annotations

import torch

from . import _lib as L


def intensity_segment_mean(I, duration_tgt, phon_len, T_phon_max=None):
    """I: (B, Tm, D) float CUDA; duration_tgt: (B, Tp) int; phon_len: (B,) int  ->  (B, Tp, D) fp32.
    out[b, p] = mean of I[b, f] over the frames of phoneme p (zero-duration phonemes and p >= phon_len[b] give 0;
    frames beyond Tm are ignored, as the reference's slice I[b, :T_mel] does)."""
    if not I.is_cuda:
        raise RuntimeError("fs2_b200: intensity_segment_mean needs CUDA tensors (there is no CPU fallback)")
    B, Tm, D = I.shape
    Tp = int(T_phon_max) if T_phon_max is not None else int(duration_tgt.shape[1])
    if duration_tgt.shape[1] != Tp:
        raise ValueError("duration_tgt must be (B, T_phon_max)")
    I = I.detach().contiguous().float()
    dur = duration_tgt.to(I.device).contiguous().long()
    pl = phon_len.to(I.device).contiguous().long()
    out = torch.empty(B, Tp, D, device=I.device, dtype=torch.float32)
    L.call("fs2_intensity_segment_mean", I, dur, pl, B, Tp, Tm, D, out)
    return out


def get_intensity_representation(intensity_extractor, batch, device):
    """Drop-in for train.py:16-51: same arguments, same (B, T_phon_max, D) result."""
    (phoneme, _, phon_len, _, _, _, duration_tgt, mel_len, _, _, rank_X, emo_ids) = batch
    with torch.no_grad():
        # the collate hands rank_X over channels-first (B, n_mels + 2, Tm) (dataset.py:116-117): say so instead of letting
        # the extractor infer the layout from the shape (a batch padded to exactly n_mels + 2 frames would be ambiguous)
        try:
            I = intensity_extractor(rank_X, mel_len, emo_ids, channels_first=True)
        except TypeError:                      # the reference's own IntensityExtractor has no such argument
            I = intensity_extractor(rank_X, mel_len, emo_ids)
        return intensity_segment_mean(I.to(device), duration_tgt, phon_len, phoneme.shape[1])


def intensity_prototypes(I, length, scores, speakers, emotions, n_spk, n_emo, bucket_size):
    """Intensity prototype bank of `rank_model/inference.py:88-114`, on the device.

    I (N, Tmax, D) frame intensities of N utterances (the extractor's output), length (N,) valid frames, scores (N,) the
    rank model's relevance score r, speakers / emotions (N,) ids  ->  (n_spk, n_emo, bucket_size, D) fp32: per
    (speaker, emotion) the utterances are ordered by ascending score, their valid frames concatenated, cut into
    `bucket_size` contiguous runs (numpy.array_split) and averaged.  The reference appends to Python lists, sorts them
    and averages numpy slices; here the host only sorts N keys and one kernel pair does the rest."""
    if not I.is_cuda:
        raise RuntimeError("fs2_b200: intensity_prototypes needs CUDA tensors (there is no CPU fallback)")
    N, Tmax, D = I.shape
    dev = I.device
    grp = (speakers.long() * n_emo + emotions.long()).to(dev)
    scores = scores.to(dev).double()
    # stable sort by (group, score): python's list.sort in the reference is stable in the score as well
    order = torch.argsort(scores, stable=True)
    order = order[torch.argsort(grp[order], stable=True)]
    lens = length.to(dev).long()[order]
    g_sorted = grp[order]
    n_groups = n_spk * n_emo
    totals = torch.zeros(n_groups, dtype=torch.long, device=dev).index_add_(0, g_sorted, lens)
    start = torch.cumsum(lens, 0) - lens                                   # exclusive prefix over the sorted list
    group_start = torch.cumsum(totals, 0) - totals
    frame_off = start - group_start[g_sorted]                              # position inside the group's frame list
    out = torch.zeros(n_spk, n_emo, bucket_size, D, device=dev, dtype=torch.float32)
    L.call("fs2_prototype_buckets", I.detach().float()[order].contiguous(), lens.int().contiguous(), g_sorted.int().contiguous(),
           frame_off.contiguous(), totals.contiguous(), N, Tmax, D, n_groups, bucket_size, out)
    return out
