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


class PeerAllReduce:
    """Sum of a small float32 vector over the ranks of `group` through NVLink peer memory (csrc/peer_allreduce.cu):
    one single-CTA kernel per call instead of an NCCL collective — the gradient payload is <= 96 floats, so the
    collective is pure latency.  torch's symmetric memory is the plumbing (allocation + exchange of the peer
    mappings); the data path is our kernel.  Raises at construction when peer memory is not available (the caller
    then keeps the NCCL all-reduce)."""

    #: models recognise this and let the parameter-Jacobian kernel do the exchange itself
    #: (sn_scenenet_param_grads_allreduce: no separate launch on the step's critical path)
    fused_with_param_grads = True

    def __init__(self, device, group=None, timeout_s: float = 600.0):
        """timeout_s: bound on the wait for a peer (0 = wait for ever).  Like a collective watchdog, running into it is
        fatal (the kernel traps and the next CUDA call raises): ranks skew by minutes around checkpoints / validation,
        so keep it in minutes.  `ok()` / `check()` read the status word at a point of the caller's choosing."""
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from ._lib import check, lib
        group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        nbytes = int(lib.sn_peer_allreduce_buffer_bytes(self.world))
        if nbytes < 0:
            raise RuntimeError(f"peer all-reduce supports at most 16 ranks, got {self.world}")
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer mappings")
        self.ptrs = (C.c_uint64 * self.world)(*ptrs)
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.timeout_ms = int(max(0.0, float(timeout_s)) * 1000)
        self._check, self._lib = check, lib
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's buffer is zeroed and mapped before the first exchange

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        """in-place sum over ranks of a contiguous float32 CUDA vector (<= 96 elements)"""
        if flat.dtype != torch.float32 or not flat.is_cuda or not flat.is_contiguous():
            raise TypeError("peer all-reduce: contiguous float32 CUDA vector expected")
        from .ops import _on_device, _stream
        with _on_device(flat.device):
            self._check(self._lib.sn_peer_allreduce(flat.data_ptr(), flat.numel(), self.rank, self.world, self.ptrs,
                                                    self.seq.data_ptr(), self.status.data_ptr(), self.timeout_ms, _stream()),
                        "sn_peer_allreduce")
        return flat

    def ok(self) -> bool:
        """False if any call ran into the bound waiting for a peer (device sync; the kernel has trapped by then, so the
        synchronisation itself raises on a live failure — this is for post-mortems and tests)."""
        return int(self.status) == 0

    def check(self) -> None:
        """raise if an exchange failed (call at a safe point, e.g. once per optimizer step or per epoch)"""
        if not self.ok():
            raise RuntimeError("scenenet_b200: a peer did not answer the gradient exchange within "
                               f"{self.timeout_ms / 1000:.0f} s; the replicas are out of step")
