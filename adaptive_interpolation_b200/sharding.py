"""Contiguous sharding of the point array over GPUs.

The path is embarrassingly parallel: rank r evaluates x[lo_r:hi_r], the (<= 36 KB) table is
replicated on every GPU, y stays sharded, and there is NO data-path collective.  The only
communication is bookkeeping: a start barrier and a MAX-reduction of one timing scalar
(torch.distributed, NCCL on GPUs / gloo in the CPU tests).  The reference has no counterpart
(single process, single device: generate.py:158-187).
"""


def shard_bounds(n, world_size, align=512):
    """[(lo, hi)] * world_size: contiguous, disjoint, covering [0, n).  Interior boundaries are
    multiples of `align` elements so every shard keeps the 16-byte vector path."""
    n, world_size = int(n), int(world_size)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if n < 0:
        raise ValueError("n must be >= 0")
    per = -(-n // world_size)
    per = -(-per // align) * align if n >= align * world_size else per
    bounds, lo = [], 0
    for _ in range(world_size):
        hi = min(n, lo + per)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_of(n, rank, world_size, align=512):
    return shard_bounds(n, world_size, align)[rank]


def max_over_ranks(value, dist=None, device=None):
    """MAX of a python float over all ranks (identity when not distributed)."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist=None, device=None):
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
