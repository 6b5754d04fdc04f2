"""Multi-GPU plumbing: shard the batch, gather the results.

Problems are independent (no problem reads another's data), so the batch is cut into
contiguous shards, one per rank / GPU, each solved by its own handle with no data-path
communication; the only collective is the final gather of trajectories and solver
scalars.  Works with any torch.distributed backend: ``nccl`` on
CUDA tensors over NVLink on the GPU box, ``gloo`` on CPU tensors in the unit tests."""
from __future__ import annotations

import numpy as np


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of the batch owned by ``rank`` (sizes differ by at most 1)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local, batch: int, dist, device=None):
    """All-gather per-problem tensors (dim 0 = local problems) of unequal shard sizes.
    ``local`` is a dict name -> torch tensor; returns dict name -> tensor with dim 0 = batch
    on every rank.  Shards are padded to the largest shard so one all_gather per tensor
    suffices (all_gather needs equal shapes)."""
    import torch

    world = dist.get_world_size()
    sizes = [shard_bounds(batch, world, r)[1] - shard_bounds(batch, world, r)[0] for r in range(world)]
    mx = max(sizes)
    out = {}
    for name, t in local.items():
        if t.shape[0] != sizes[dist.get_rank()]:
            raise ValueError(f"{name}: dim 0 is {t.shape[0]}, shard size is {sizes[dist.get_rank()]}")
        pad = t
        if t.shape[0] < mx:
            pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad.contiguous())
        out[name] = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
    return out


def split_inputs(arrays: dict, world_size: int, rank: int) -> dict:
    """Slice host arrays (dim 0 = batch) down to this rank's shard."""
    out = {}
    for k, a in arrays.items():
        lo, hi = shard_bounds(a.shape[0], world_size, rank)
        out[k] = np.ascontiguousarray(a[lo:hi])
    return out
