"""Multi-GPU plumbing: the batch of voxel grids shards across ranks (one process per GPU), every
rank holds the full 13-scalar model, and ONE all-reduce per step carries the parameter-gradient
payload (<= 96 floats) over NCCL / NVLink.  DDP semantics: mean over ranks of per-shard gradients
(what Lightning's implicit DDP does for the reference, scripts/main.py:224-236).
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world,
                                **({"device_id": device} if use_cuda else {}))
    return rank, world, device


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n units for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_grads(params: Sequence[torch.nn.Parameter], group=None, already_scaled: bool = False):
    """One collective for all parameter gradients: pack -> all_reduce(SUM) -> (/world) -> unpack.
    `already_scaled`: the backward already multiplied by 1/world (model.grad_scale), so only sum."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.stack([p.grad.reshape(()) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if not already_scaled:
        flat /= dist.get_world_size(group)
    for p, g in zip(ps, flat.unbind()):
        p.grad.copy_(g)
