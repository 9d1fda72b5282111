"""Fused AdamW over the model's flat parameter / gradient buffers (reference: torch.optim.AdamW(lr=1e-4) with
torch defaults, `/root/reference/emo_rank_tts/fastspeech2/train.py:232, 81`): one kernel per step instead of
a foreach pass per parameter.  The same pass writes the bf16 operand mirror the tensor-core GEMMs read (params.py), so
no weight re-packing pass runs between steps, and it refuses to apply a step whose tcgen05 kernels reported an mbarrier
timeout (the device-side error word): the host raises at the next periodic check instead of training on garbage."""
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
        self.error_check_interval = 32      # steps between host reads of the tcgen05 error word (one stream sync each)
        self._covered = 0

    def zero_grad(self, set_to_none=True):
        st = self.model.store
        if st.flat_grad is not None:
            L.call("fs2_memset", st.flat_grad, 0, st.flat_grad.numel() * 4)

    def step(self, grad_scale=1.0, ranges=None, advance=True):
        """One AdamW update.  `ranges` = list of (lo, hi[, before]) element ranges of the flat buffers, updated in that
        order with one launch each (`before()` is called first -- the data-parallel step waits for that piece's
        all-reduce there, so the update of one piece overlaps the reduction of the next); default: everything.
        advance = False: more pieces of the step that an earlier call (with advance = True) began -- the data-parallel
        step updates the decoder's parameters on a side stream while the encoder is still back-propagating."""
        st = self.model.store
        if st.flat_grad is None:
            raise RuntimeError("FusedAdamW.step() before any backward()")
        if self.m is None or self.m.device != st.flat.device:
            self.m = torch.zeros_like(st.flat)
            self.v = torch.zeros_like(st.flat)
        if advance:
            self.step_count += 1
            self._covered = 0
        n = st.flat.numel()
        mirror = st.mirror if (st.mirror is not None and st.mirror.device == st.flat.device) else None
        gat = st.adamw_gather() if mirror is not None else None
        covered = self._covered
        for r in (ranges or [(0, n)]):
            lo, hi = int(r[0]), int(r[1])
            if len(r) > 2 and r[2] is not None:
                r[2]()
            if hi <= lo:
                continue
            if lo % 4:
                raise ValueError("FusedAdamW: range starts must be multiples of 4 elements (16-byte vector access)")
            # the gathered operand copy is refreshed by the launch that updates its source (other launches may run
            # concurrently with kernels that read it)
            g = gat if (gat is not None and lo <= gat[0] < hi) else (0, 1, 0, 0, 0)
            L.call("fs2_adamw_fused", st.flat[lo:hi], st.flat_grad[lo:hi], self.m[lo:hi], self.v[lo:hi], hi - lo,
                   float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                   self.step_count, float(grad_scale), mirror, lo, g[0], g[1], g[2], g[3], g[4], 1)
            covered += hi - lo
        self._covered = covered
        if covered < n:
            return
        if mirror is not None and (gat is not None or st.gather_total == 0):
            st.mark_mirror_current()
        if self.error_check_interval and self.step_count % self.error_check_interval == 0:
            self.check_device_errors()

    def check_device_errors(self):
        """Read-and-clear the tcgen05 error word (synchronises the device).  While the word is set the AdamW kernel skips
        its update, so the parameters still hold the last good step when this raises."""
        if L.gemm_tc_error_flag() != 0:
            raise RuntimeError("fs2_b200: a tcgen05 kernel reported an mbarrier timeout since the last check; the "
                               "optimizer steps since then were NOT applied (parameters hold the last good state)")

    def state_dict(self):
        return {"step": self.step_count, "m": self.m, "v": self.v, "lr": self.lr, "betas": self.betas,
                "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.m, self.v = sd["m"], sd["v"]
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], sd["betas"], sd["eps"], sd["weight_decay"]
