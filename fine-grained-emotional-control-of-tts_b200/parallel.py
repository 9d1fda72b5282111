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

    The gradient is reduced in two pieces of ONE flat fp32 buffer: the tail that holds decoder / mel-linear / PostNet
    gradients is final when the decoder's backward ends (about 75 % into the backward), so its all-reduce is issued
    there (async, NCCL's own stream) and travels over NVLink while the variance adaptor and the encoder are still
    back-propagating; the head follows after the backward.  No other collective exists on the path.

    DP semantics: the reduced gradient is the MEAN over ranks of each rank's local-batch gradient,
    which equals the reference's gradient on the concatenated batch for the MSE terms (SSIM and the attn-mask
    quirk depend on local batch composition)."""

    def __init__(self, model, criterion, optimizer, group=None, overlap=True):
        self.model, self.criterion, self.optimizer, self.group = model, criterion, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap
        self._pending = []
        self._early_lo = None
        if self.world > 1:
            broadcast_flat_(model.store.flat, 0, group)      # identical replicas

    def _early_reduce(self, lo, hi):
        g = self.model.store.flat_grad
        if lo < hi:
            self._pending.append(dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._early_lo = lo

    def __call__(self, batch, intensity, epoch=0):
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
        self.optimizer.zero_grad()
        preds = self.model(tokens, speakers, dur, pitch, energy, intensity=intensity)
        losses = self.criterion(preds, (mel, dur, pitch, energy, out_lens, in_lens), epoch)
        self._pending, self._early_lo = [], None
        hook = self.world > 1 and self.overlap
        if hook:
            self.model.grad_ready_hook = self._early_reduce
        try:
            losses["total_loss"].backward()
        finally:
            if hook:
                self.model.grad_ready_hook = None
        g = self.model.store.flat_grad
        n = g.numel()
        if self.world > 1:
            lo = n if self._early_lo is None else self._early_lo
            early = list(self._pending)
            head = dist.all_reduce(g[:lo], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            # tail (decoder / PostNet) is updated while the head (encoder, variance adaptor) is still on the wire
            self.optimizer.step(grad_scale=1.0 / self.world,
                                ranges=[(lo, n, lambda: [w.wait() for w in early]), (0, lo, head.wait)])
        else:
            self.optimizer.step(grad_scale=1.0)
        return losses, preds
