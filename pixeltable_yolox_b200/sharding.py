"""Batch sharding across the GPUs of one box: inference is embarrassingly parallel
over images (per-image NMS, eval-mode BN), so every rank runs the same plan on its own slice and
there is NO data-path collective; only the per-rank results are gathered on the host side."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n items for `rank` (first n % world ranks get +1)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_detection_counts(local_counts: List[int], group=None) -> List[int]:
    """All ranks learn every image's detection count (host-side metadata exchange over the default
    process group; works with gloo and nccl)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local_counts)
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, list(local_counts), group=group)
    out: List[int] = []
    for part in gathered:
        out.extend(part)
    return out
