"""Seeded inputs and the reference-procedure result for the intensity prototype bucketing (SURVEY 8f row 4,
rank_model/inference.py:80-114), shared by the GPU parity test and the CPU test that checks the restated procedure
against the reference's own lines."""
import warnings

import numpy as np
import torch


def inputs(D=5):
    g = torch.Generator().manual_seed(21)
    N, Tmax, n_spk, n_emo, k = 40, 37, 3, 4, 6
    length = torch.randint(1, Tmax + 1, (N,), generator=g)
    I = torch.randn(N, Tmax, D, generator=g)
    r = torch.randn(N, generator=g)
    r[5] = r[9]                                                   # a tie: list.sort is stable
    spk = torch.randint(0, n_spk, (N,), generator=g)
    emo = torch.randint(0, n_emo - 1, (N,), generator=g)          # emotion 3 never occurs -> all-zero cells
    spk[0], emo[0], length[0] = 2, 2, 3                           # make (2, 2) a cell with 3 frames < 6 buckets
    keep = ~((spk == 2) & (emo == 2))
    keep[0] = True
    I, length, r, spk, emo = I[keep], length[keep], r[keep], spk[keep], emo[keep]
    return I, length, r, spk, emo, n_spk, n_emo, k


def restated(I, length, r, spk, emo, n_spk, n_emo, k):
    """inference.py:80-114 with its own tools: python lists, list.sort on the score, numpy concatenate / array_split /
    mean (NaN for empty slices, zeros for cells without utterances)."""
    D = I.shape[2]
    store = {(s, e): [] for s in range(n_spk) for e in range(n_emo)}
    for i in range(I.shape[0]):
        store[(int(spk[i]), int(emo[i]))].append((float(r[i]), I[i, : int(length[i])].numpy()))
    ref = np.zeros((n_spk, n_emo, k, D), dtype=np.float32)
    for (s, e), entries in store.items():
        if not entries:
            continue
        entries.sort(key=lambda x: x[0])
        feats = np.concatenate([x[1] for x in entries], axis=0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for bi, idxs in enumerate(np.array_split(np.arange(len(feats)), k)):
                ref[s, e, bi] = feats[idxs].mean(axis=0)
    return ref
