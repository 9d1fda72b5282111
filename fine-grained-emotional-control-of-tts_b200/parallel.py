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
    """forward + loss + backward + single flat-gradient all-reduce + fused AdamW.

    DP semantics: the reduced gradient is the MEAN over ranks of each rank's local-batch gradient,
    which equals the reference's gradient on the concatenated batch for the MSE terms (SSIM and the attn-mask
    quirk depend on local batch composition)."""

    def __init__(self, model, criterion, optimizer, group=None):
        self.model, self.criterion, self.optimizer, self.group = model, criterion, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            broadcast_flat_(model.store.flat, 0, group)      # identical replicas

    def __call__(self, batch, intensity, epoch=0):
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
        self.optimizer.zero_grad()
        preds = self.model(tokens, speakers, dur, pitch, energy, intensity=intensity)
        losses = self.criterion(preds, (mel, dur, pitch, energy, out_lens, in_lens), epoch)
        losses["total_loss"].backward()
        allreduce_flat_(self.model.store.flat_grad, self.world, self.group)
        self.optimizer.step(grad_scale=1.0 / self.world)
        return losses, preds
