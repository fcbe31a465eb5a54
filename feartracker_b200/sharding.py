"""Multi-GPU partitioning of the frame batch.

Frames are independent, weights are replicated, so the only collective on the path is ONE
all-gather of the fixed-size per-frame box records (48 B each) after the decode kernel.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``total`` frames owned by ``rank`` (first ranks take the remainder)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_boxes(local: torch.Tensor, total: int) -> torch.Tensor:
    """local: (n_local, 48) uint8 FearBox records of this rank's shard -> (total, 48) on every rank.
    Issued on the current stream right after the decode kernel; no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = -(-total // world)
    rec = local.shape[1]
    padded = local
    if local.shape[0] != per:
        padded = torch.zeros((per, rec), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * per, rec), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    if world * per == total:
        return out
    pieces = []
    for r in range(world):
        b, e = shard_range(total, r, world)
        pieces.append(out[r * per: r * per + (e - b)])
    return torch.cat(pieces)
