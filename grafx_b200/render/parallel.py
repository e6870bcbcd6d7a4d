"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The reference is single-device; the hot path shards over the batch-of-
renders axis with NO exchange step in the math (the `out` bus sums tracks of the same item), so
the data path needs no collective.  The only optional communication is gathering the rendered
mixes; timing helpers reduce with MAX over ranks as the bench contract requires."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int | None = None, world: int | None = None, dim: int = 0) -> torch.Tensor:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def gather_batch(local: torch.Tensor, total: int, dim: int = 0) -> torch.Tensor:
    """All-gather of the per-rank shards (possibly ragged) back into one tensor of size `total`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = biggest
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p.narrow(dim, 0, hi - lo) for p, (lo, hi) in zip(parts, sizes)], dim)


def render_grafx_sharded(processors, input_signals, per_type_parameters, render_data, gather: bool = False, **kw):
    """render_grafx on this rank's slice of a 4-D `[B, |V0|, C, L]` source batch.  Parameters are
    per node (shared by every batch item), so they are replicated."""
    from .graph import render_grafx

    assert input_signals.ndim == 4, "sharding is over the batch-of-renders axis"
    total = input_signals.shape[0]
    local = shard_batch(input_signals)
    out, inter, buf = render_grafx(processors, local, per_type_parameters, render_data, **kw)
    if gather:
        out = gather_batch(out, total)
    return out, inter, buf


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
