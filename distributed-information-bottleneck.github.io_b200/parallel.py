"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the single
exchange step of the path -- a sum all-reduce of the flat [gradients || statistics] buffer between backward and
Adam.  The reference has no multi-GPU code; samples are independent units, every loss term is
a batch mean (models.py:111-112), so each global batch splits into contiguous row ranges per rank."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_and_rank(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(n, rank, world):
    """Rows [lo, hi) of an n-row global batch owned by ``rank``: contiguous, sizes differ by at most one,
    earlier ranks take the remainder.  Covers [0, n) exactly once over all ranks."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_sum_(flat: torch.Tensor, group=None):
    """In-place sum over the data-parallel group on the current stream; no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_sum_async(flat: torch.Tensor, group=None):
    """Asynchronous in-place sum on the CURRENT stream's NCCL channel; returns the work handle (``.wait()`` makes the then
    current stream wait for it) or None for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None
