"""Synthetic batches with the 12-tuple layout of the reference collate
(`/root/reference/emo_rank_tts/fastspeech2/dataset.py:120-133`):

    (phoneme_padded i64 (B,Tp), speakers i64 (B,), input_lengths i64 (B,),
     mel_padded f32 (B,Tm,80), pitch_padded f32 (B,Tm), energy_padded f32 (B,Tm),
     duration_padded i64 (B,Tp), output_lengths i64 (B,), labels, wavs,
     rank_X f32 (B,82,Tm), emotions i64 (B,))

plus the per-phoneme intensity tensor (B,Tp,5) that train.py:69 derives from the
frozen rank model (upstream of the hot path; fed directly here).
The recipe follows SURVEY.md section 8(d): Tp ~ U{24..128}, log-normal durations with
~5 % forced zeros, sum(dur) <= 800, rows sorted by descending Tp
(dataset.py:65-67) and length-bucketed across a pool.
"""
from __future__ import annotations

import torch

N_MELS = 80
N_CHAR = 95
N_SPEAKERS = 4
N_EMOTIONS = 5


def _one_utterance(g, min_tp, max_tp, max_frames):
    tp = int(torch.randint(min_tp, max_tp + 1, (1,), generator=g))
    dur = torch.exp(torch.randn(tp, generator=g) * 0.6 + 1.6).round().clamp(0, 40).long()
    dur[torch.rand(tp, generator=g) < 0.05] = 0
    if int(dur.sum()) < 12:            # SSIM needs >= 11 frames (11x11 valid conv)
        dur[0] += 12
    while int(dur.sum()) > max_frames:  # clip so sum(dur) <= max_frames
        dur = (dur.float() * (max_frames / float(dur.sum())) * 0.98).floor().long()
        if int(dur.sum()) < 12:
            dur[0] += 12
    tm = int(dur.sum())
    return dict(
        phoneme=torch.randint(1, N_CHAR, (tp,), generator=g),
        duration=dur,
        mel=torch.rand(tm, N_MELS, generator=g) * 13.5 - 11.5,
        pitch=torch.randn(tm, generator=g),
        energy=torch.randn(tm, generator=g),
        speaker=int(torch.randint(0, N_SPEAKERS, (1,), generator=g)),
        emotion=int(torch.randint(0, N_EMOTIONS, (1,), generator=g)),
        intensity=torch.randn(tp, N_EMOTIONS, generator=g),
    )


def collate(utts):
    """Same padding/sorting rules as TextMelCollateWithAlignment (dataset.py:62-133)."""
    lens = torch.LongTensor([len(u["phoneme"]) for u in utts])
    input_lengths, order = torch.sort(lens, dim=0, descending=True)
    B = len(utts)
    Tp = int(input_lengths[0])
    Tm = max(int(u["mel"].shape[0]) for u in utts)
    phoneme = torch.zeros(B, Tp, dtype=torch.long)
    duration = torch.zeros(B, Tp, dtype=torch.long)
    mel = torch.zeros(B, Tm, N_MELS)
    pitch = torch.zeros(B, Tm)
    energy = torch.zeros(B, Tm)
    rank_X = torch.zeros(B, N_MELS + 2, Tm)
    intensity = torch.zeros(B, Tp, N_EMOTIONS)
    out_len = torch.zeros(B, dtype=torch.long)
    speakers = torch.zeros(B, dtype=torch.long)
    emotions = torch.zeros(B, dtype=torch.long)
    for i, j in enumerate(order.tolist()):
        u = utts[j]
        tp, tm = len(u["phoneme"]), u["mel"].shape[0]
        phoneme[i, :tp] = u["phoneme"]
        duration[i, :tp] = u["duration"]
        mel[i, :tm] = u["mel"]
        pitch[i, :tm] = u["pitch"]
        energy[i, :tm] = u["energy"]
        rank_X[i, :N_MELS, :tm] = u["mel"].t()
        rank_X[i, N_MELS, :tm] = u["pitch"]
        rank_X[i, N_MELS + 1, :tm] = u["energy"]
        intensity[i, :tp] = u["intensity"]
        out_len[i] = tm
        speakers[i] = u["speaker"]
        emotions[i] = u["emotion"]
    batch = (phoneme, speakers, input_lengths, mel, pitch, energy, duration, out_len,
             [""] * B, [""] * B, rank_X, emotions)
    return batch, intensity


def synthetic_batches(batch_size, n_batches, seed=1234, rank=0, min_tp=24, max_tp=128,
                      max_frames=800, pool_factor=8, world=1):
    """Length-bucketed synthetic batches: draw a pool of pool_factor*B utterances,
    sort by mel length, slice consecutive groups of B.

    world == 1 (default): one rank's private stream of batches, seeded by (seed, rank, step).
    world > 1 (data parallel): ONE pool of pool_factor*B*world utterances per draw, shared by all
    ranks (same seed everywhere); the sorted pool is cut into pool_factor "length classes" of `world * B` consecutive
    utterances which are dealt to the ranks round-robin (rank r takes utterances r, r + world, ... of its class), so the
    ranks of one step pad to the same rectangle to within a frame or two and nobody waits at the gradient all-reduce
    (taking `world` neighbouring groups instead left the slowest rank ~3 % behind the mean: 0.25 ms of an 8.4 ms step).
    The classes are visited in the order the single-GPU stream visits its groups, so the mix of short and long batches
    per rank is the same at every world size."""
    out = []
    step = 0
    while len(out) < n_batches:
        g = torch.Generator().manual_seed(seed + 1000 * (rank if world == 1 else 0) + step)
        pool = [_one_utterance(g, min_tp, max_tp, max_frames) for _ in range(pool_factor * batch_size)]
        n_groups = len(pool) // batch_size + (1 if len(pool) % batch_size else 0)
        perm = torch.randperm(n_groups, generator=g).tolist()
        fine = 1
        if world > 1:
            # a 4x larger pool: the `world` neighbouring buckets of one step then span 1/(4*pool_factor) of the length
            # distribution, which keeps the slowest rank of a step within ~1-2 % of the mean
            fine = 4
            gw = torch.Generator().manual_seed(seed + step + 7919 * world)
            pool = [_one_utterance(gw, min_tp, max_tp, max_frames) for _ in range(pool_factor * batch_size * world * fine)]
        pool.sort(key=lambda u: u["mel"].shape[0])
        groups = [pool[i:i + batch_size] for i in range(0, len(pool), batch_size)]
        for k in perm:
            # length class k of the single-GPU stream = super-groups k*fine .. k*fine+fine-1; take the middle one
            if world > 1:
                g0 = (k * fine + fine // 2) * world * batch_size
                out.append(collate(pool[g0:g0 + world * batch_size][rank::world]))
            else:
                out.append(collate(groups[k]))
            if len(out) == n_batches:
                break
        step += 1
    return out


def worst_case_batch(batch_size, tp=128, tm=800, seed=7):
    """A full (B, tp, tm) rectangle: every utterance has tp phonemes and exactly
    tm frames.  This is the shape BASELINE.json's cfg-3 quotes (<=128, <=800)."""
    g = torch.Generator().manual_seed(seed)
    utts = []
    for _ in range(batch_size):
        u = _one_utterance(g, tp, tp, tm)
        d = u["duration"]
        d[-1] += tm - int(d.sum())
        u = dict(u)
        u["duration"] = d
        u["mel"] = torch.rand(tm, N_MELS, generator=g) * 13.5 - 11.5
        u["pitch"] = torch.randn(tm, generator=g)
        u["energy"] = torch.randn(tm, generator=g)
        utts.append(u)
    return collate(utts)
