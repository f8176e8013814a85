"""Data-parallel sharding of a batch across ranks.

Each image's forward pass is independent (the reference is batch-1, one global ``image[]``,
``MobileNet.c:29``), so rank ``r`` of ``g`` takes the contiguous slice
``[r*N/g, (r+1)*N/g)`` of the global batch, weights are replicated, and nothing is exchanged
during the forward pass.  The only collective is one all-gather of the per-rank logits (and
top-1) at the end — NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of the global batch owned by ``rank``; remainders go to the low ranks."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def gather_logits(local_logits, n_global: int, group=None):
    """All-gather ragged per-rank ``[n_r, classes]`` logits into ``[n_global, classes]`` (torch
    tensors, any backend).  Shards are padded to the largest shard so one collective suffices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local_logits
    classes = local_logits.shape[1]
    per = max(shard_range(n_global, r, world)[1] - shard_range(n_global, r, world)[0] for r in range(world))
    padded = torch.zeros(per, classes, dtype=local_logits.dtype, device=local_logits.device)
    padded[: local_logits.shape[0]] = local_logits
    out = torch.empty(world * per, classes, dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world):
        a, b = shard_range(n_global, r, world)
        pieces.append(out[r * per: r * per + (b - a)])
    return torch.cat(pieces, dim=0)
