"""Data-parallel sharding of a batch across ranks.

Each image's forward pass is independent (the reference is batch-1, one global ``image[]``,
``MobileNet.c:29``), so rank ``r`` of ``g`` takes the contiguous slice
``[r*N/g, (r+1)*N/g)`` of the global batch, weights are replicated, and nothing is exchanged
during the forward pass.  The only collective is one all-gather of the per-rank logits (and
top-1) at the end — NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of the global batch owned by ``rank``; remainders go to the low ranks."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def weighted_range(n: int, rank: int, world: int, weights: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """[first, last) when the shards follow per-rank weights (e.g. each GPU's measured host link, ``mnv1_dp_calibrate``):
    largest-remainder apportionment, ties to the lower rank — the Python mirror of ``mnv1_dp_shard_weighted`` (dp.cpp);
    ``weights=None`` is ``shard_range``."""
    if weights is None:
        return shard_range(n, rank, world)
    if world <= 0 or not (0 <= rank < world) or n < 0 or len(weights) != world or not all(w > 0 for w in weights):
        raise ValueError("bad shard arguments")
    import numpy as np
    w32 = [float(np.float32(w)) for w in weights]          # the C-ABI takes float weights and works in double
    total = float(sum(w32))
    share = [n * w / total for w in w32]
    cnt = [int(x) for x in share]
    frac = [x - c for x, c in zip(share, cnt)]
    for _ in range(n - sum(cnt)):
        best = max(range(world), key=lambda r: (frac[r], -r))
        cnt[best] += 1
        frac[best] = -1.0
    first = sum(cnt[:rank])
    return first, first + cnt[rank]


def gather_logits(local_logits, n_global: int, group=None, weights: Optional[Sequence[float]] = None):
    """All-gather ragged per-rank ``[n_r, classes]`` logits into ``[n_global, classes]`` (torch
    tensors, any backend).  Shards are padded to the largest shard so one collective suffices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local_logits
    classes = local_logits.shape[1]
    if weights is not None:
        spans = [weighted_range(n_global, r, world, weights) for r in range(world)]
        per = max(b - a for a, b in spans)
        padded = torch.zeros(per, classes, dtype=local_logits.dtype, device=local_logits.device)
        padded[: local_logits.shape[0]] = local_logits
        out = torch.empty(world * per, classes, dtype=local_logits.dtype, device=local_logits.device)
        dist.all_gather_into_tensor(out, padded, group=group)
        return torch.cat([out[r * per: r * per + (b - a)] for r, (a, b) in enumerate(spans)], dim=0)
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
