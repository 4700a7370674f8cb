"""Multi-GPU partition of one frame: tile row j of the frame belongs to rank
j % world, every rank renders its rows into a compact band image, rank 0 gathers the bands
and puts the rows back in order.  The tile stream (per-tile RNG) makes any partition made
of whole reference tiles bit-identical to the single-device frame.

Pure torch / torch.distributed code with no CUDA calls of its own, so the same functions
run under NCCL on GPUs (bench.py) and under gloo on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def tiles_y(height: int, tile_size: int) -> int:
    return (height + tile_size - 1) // tile_size


def local_tile_rows(height: int, tile_size: int, rank: int, world: int) -> list[int]:
    """Frame tile rows rendered by `rank`, in the order they sit in its band image."""
    return list(range(rank, tiles_y(height, tile_size), world))


def band_pixel_rows(height: int, tile_size: int, rank: int, world: int) -> int:
    rows = local_tile_rows(height, tile_size, rank, world)
    return sum(min(tile_size, height - r * tile_size) for r in rows)


def padded_band_rows(height: int, tile_size: int, world: int) -> int:
    """Band height every rank allocates so the gather moves equal-sized buffers."""
    return ((tiles_y(height, tile_size) + world - 1) // world) * tile_size


def deinterleave(bands: list[torch.Tensor], frame: torch.Tensor, tile_size: int) -> torch.Tensor:
    """bands[r]: [padded_rows, W, C] of rank r  ->  frame [H, W, C] with rows in image order."""
    height = frame.shape[0]
    world = len(bands)
    for r, band in enumerate(bands):
        for k, tile_row in enumerate(local_tile_rows(height, tile_size, r, world)):
            y0 = tile_row * tile_size
            h = min(tile_size, height - y0)
            frame[y0:y0 + h].copy_(band[k * tile_size:k * tile_size + h], non_blocking=True)
    return frame


def gather_frame(band: torch.Tensor, frame: torch.Tensor | None, tile_size: int,
                 gathered: list[torch.Tensor] | None = None) -> torch.Tensor | None:
    """Collective: every rank passes its padded band; rank 0 returns the assembled frame.
    The only exchange of the path: one gather, W*H*C*4/world bytes per non-root rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return deinterleave([band], frame, tile_size)
    if rank == 0 and gathered is None:
        gathered = [torch.empty_like(band) for _ in range(world)]
    dist.gather(band, gathered if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    return deinterleave(gathered, frame, tile_size)
