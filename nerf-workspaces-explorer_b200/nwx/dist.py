"""Multi-GPU rendering: rays shard naturally, so a frame (or a multi-view batch) is cut into
contiguous ray ranges, one per rank (= row tiles of the image when the range is a multiple of W);
weights are replicated; the only exchange is one all-gather of the uint8 pixel tiles (NCCL over
NVLink on GPUs, gloo in the CPU tests).  The reference has no distributed code."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of `total` rays: the first total % world ranks get one more."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def gather_tiles(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather per-rank tiles [count_r, C] (ranges from shard_range) into [total, C] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    if len(set(counts)) == 1:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = max(counts)                      # ragged split: pad to the widest tile, then trim
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def render_sharded(total_rays: int, render_range: Callable[[int, int], torch.Tensor],
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """render_range(start, count) -> [count, C] for this rank's range; returns the full [total, C]."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    start, count = shard_range(total_rays, rank, world)
    return gather_tiles(render_range(start, count), total_rays, group)


def render_poses_sharded(handler, c2w: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                         to_host: bool = False):
    """[B,4,4] poses -> uint8 [B,H,W,3] on every rank, each rank rendering 1/world of the B*H*W rays (row tiles
    of the frame when B == 1) and one all-gather of the uint8 tiles.  to_host=True returns the frames as a host
    array through the handler's pinned read-back buffer (what render_poses returns on one GPU)."""
    eng = handler.engine
    H, W = handler._img_h, handler._img_w
    B = c2w.shape[0]
    with torch.no_grad():
        full = render_sharded(B * H * W, lambda start, count: handler.render_rays_u8(c2w, start, count), group)
    if to_host:
        return handler.frames_to_host(full, B)
    return full.view(B, H, W, 3)
