"""Fused AdamW over the model's flat parameter / gradient buffers (reference: torch.optim.AdamW(lr=1e-4) with
torch defaults, `/root/reference/emo_rank_tts/fastspeech2/train.py:232, 81`): one kernel per step instead of
a foreach pass per parameter."""
from __future__ import annotations

import torch

from . import _lib as L


class FusedAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.m = None
        self.v = None

    def zero_grad(self, set_to_none=True):
        st = self.model.store
        if st.flat_grad is not None:
            L.call("fs2_memset", st.flat_grad, 0, st.flat_grad.numel() * 4)

    def step(self, grad_scale=1.0, ranges=None):
        """One AdamW update.  `ranges` = list of (lo, hi[, before]) element ranges of the flat buffers, updated in that
        order with one launch each (`before()` is called first -- the data-parallel step waits for that piece's
        all-reduce there, so the update of one piece overlaps the reduction of the next); default: everything."""
        st = self.model.store
        if st.flat_grad is None:
            raise RuntimeError("FusedAdamW.step() before any backward()")
        if self.m is None or self.m.device != st.flat.device:
            self.m = torch.zeros_like(st.flat)
            self.v = torch.zeros_like(st.flat)
        self.step_count += 1
        n = st.flat.numel()
        for r in (ranges or [(0, n)]):
            lo, hi = int(r[0]), int(r[1])
            if len(r) > 2 and r[2] is not None:
                r[2]()
            if hi <= lo:
                continue
            if lo % 4:
                raise ValueError("FusedAdamW: range starts must be multiples of 4 elements (16-byte vector access)")
            L.call("fs2_adamw", st.flat[lo:hi], st.flat_grad[lo:hi], self.m[lo:hi], self.v[lo:hi], hi - lo, float(self.lr),
                   float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                   self.step_count, float(grad_scale))

    def state_dict(self):
        return {"step": self.step_count, "m": self.m, "v": self.v, "lr": self.lr, "betas": self.betas,
                "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.m, self.v = sd["m"], sd["v"]
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], sd["betas"], sd["eps"], sd["weight_decay"]
