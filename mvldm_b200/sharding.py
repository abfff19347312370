"""Scene-batch sharding over the GPUs of one box: scenes never interact, so each rank takes
a contiguous slice of the scene batch and runs the unmodified single-GPU path; no data-path collective.
Only the final latents are gathered."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def scene_slice(num_scenes: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the scenes owned by `rank`; earlier ranks take the remainder."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(num_scenes, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scenes(local: torch.Tensor, num_scenes: int, group=None) -> torch.Tensor:
    """all-gather the per-rank results (dim 0 = scenes) back into scene order on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [scene_slice(num_scenes, r, ws) for r in range(ws)]
    max_n = max(b - a for a, b in sizes)
    pad = torch.zeros((max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
