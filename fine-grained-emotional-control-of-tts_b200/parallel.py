"""Data-parallel training of the FastSpeech2 step: one process per GPU, utterances sharded per rank
(length-bucketed), ONE all-reduce of the flat fp32 gradient buffer per step (NCCL over NVLink on the GPU box;
gloo in the CPU tests).  The reference has no distributed code (train.py:170, 214-215 are single-device); this
is the multi-GPU row of SURVEY.md section 8(e)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist



def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_for_rank(items, rank, world):
    """Round-robin shard of a list of (length-bucketed) batches: rank r takes items r, r+world, ...
    Consecutive buckets have similar padded rectangles, so every rank's step costs about the same."""
    n = len(items) // world * world
    return [items[i] for i in range(rank, n, world)]


def allreduce_flat_(flat, world=None, group=None):
    """In-place SUM all-reduce of one flat buffer (the whole gradient in a single collective)."""
    if dist.is_initialized() and (world or dist.get_world_size(group)) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_flat_(flat, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


class DataParallelStep:
    """forward + loss + backward + flat-gradient all-reduce + fused AdamW.

    The gradient lives in ONE flat fp32 buffer and becomes final back to front while the backward runs (the flat order is
    the reference's state_dict order, the backward visits it in reverse).  The model reports three pieces through
    `grad_ready_hook`: decoder / mel-linear / PostNet after part A of the backward (51 % of the buffer), the upper encoder
    layers + encoder.norm after part B, and the rest (embeddings, variance adaptor, lower encoder layers) at the end.
    Each piece is all-reduced as soon as it is final (async, NCCL's own stream, over NVLink) and its AdamW update is
    queued behind the reduction on a separate optimizer stream, so both overlap the remaining backward; only the last,
    smallest piece is exposed.  With one GPU the same schedule hides the decoder's AdamW pass behind the encoder's
    backward.  No other collective exists on the path.

    DP semantics: the reduced gradient is the MEAN over ranks of each rank's local-batch gradient,
    which equals the reference's gradient on the concatenated batch for the MSE terms (SSIM and the attn-mask
    quirk depend on local batch composition)."""

    def __init__(self, model, criterion, optimizer, group=None, overlap=True):
        self.model, self.criterion, self.optimizer, self.group = model, criterion, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.overlap = overlap
        self._pieces = []            # (lo, hi) ranges whose update has been queued on the optimizer stream this step
        self._opt_stream = None
        if self.world > 1:
            broadcast_flat_(model.store.flat, 0, group)      # identical replicas
            # independent dropout noise per rank, as over the reference's concatenated batch (the counter-based masks are
            # keyed by (seed, step, site, position): one seed on every rank would repeat the same mask N times)
            model.manual_seed(model._seed_base + 0x9E3779B1 * self.rank)

    def _piece_ready(self, lo, hi):
        """Flat-gradient range [lo, hi) is final on the main stream: start its all-reduce (world > 1) and queue its AdamW
        update behind it on the optimizer stream -- both run while the rest of the backward is still computing."""
        if hi <= lo:
            return
        main = torch.cuda.current_stream()
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=self.model.store.flat.device)
        work = None
        if self.world > 1:
            work = dist.all_reduce(self.model.store.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._opt_stream):
            self._opt_stream.wait_event(ready)           # the gradients of this piece (and, world == 1, nothing else)
            if work is not None:
                work.wait()                              # stream-side wait for the reduction; the host does not block
            self.optimizer.step(grad_scale=1.0 / self.world, ranges=[(lo, hi)], advance=not self._pieces)
            # (zeroing the consumed piece here, under the backward, instead of the 341 MB memset at the head of the next step
            #  was measured: 8.49 vs 8.44 ms -- the memset competes with the backward for HBM; not kept)
        self._pieces.append((lo, hi))

    def __call__(self, batch, intensity, epoch=0):
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
        self.optimizer.zero_grad()
        preds = self.model(tokens, speakers, dur, pitch, energy, intensity=intensity)
        losses = self.criterion(preds, (mel, dur, pitch, energy, out_lens, in_lens), epoch)
        self._pieces = []
        if self.overlap:
            self.model.grad_ready_hook = self._piece_ready
        try:
            losses["total_loss"].backward()
        finally:
            self.model.grad_ready_hook = None
        n = self.model.store.flat_grad.numel()
        # what the hooks have not covered: the head of the flat buffer (embeddings, variance adaptor, lower encoder layers)
        done_lo = min((lo for lo, _ in self._pieces), default=n)
        covered = sorted(self._pieces)
        assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1)) and (not covered or covered[-1][1] == n)
        self._piece_ready(0, done_lo)
        torch.cuda.current_stream().wait_stream(self._opt_stream)        # next forward reads the updated parameters
        return losses, preds
