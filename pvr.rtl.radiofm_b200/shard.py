"""Sharding of independent units (IQ streams of a batch, stations of a wideband capture) over the GPUs of one box.

The chain has no exchange step: every stream / station is an independent recurrence with private
state, so ranks take contiguous ranges of unit ids and never communicate on the data path.  torch.distributed is
used only for the rendezvous, the barrier around the timed region and gathering small results to rank 0.
"""
from __future__ import annotations


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of unit ids owned by `rank`: id * world / n partition (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    lo = n_units * rank // world
    hi = n_units * (rank + 1) // world
    return lo, hi


def owner_of(unit: int, n_units: int, world: int) -> int:
    """Rank that owns `unit` under shard_range()."""
    if not (0 <= unit < n_units):
        raise ValueError("unit out of range")
    r = min(world - 1, (unit * world + world - 1) // max(n_units, 1))
    while shard_range(n_units, r, world)[0] > unit:
        r -= 1
    while shard_range(n_units, r, world)[1] <= unit:
        r += 1
    return r


def gather_to_root(dist, obj, dst: int = 0):
    """Gather one picklable object per rank to `dst` (list in rank order on dst, None elsewhere)."""
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def max_over_ranks(dist, value: float, device=None) -> float:
    """The multi-GPU timing rule: a step takes as long as its slowest rank."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
