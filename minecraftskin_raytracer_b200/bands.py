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


def frame_row_index(height: int, tile_size: int, world: int, padded_rows: int, device=None) -> torch.Tensor:
    """For every frame row y: its row in the concatenation of all ranks' padded bands."""
    idx = torch.empty(height, dtype=torch.long)
    for r in range(world):
        for k, tile_row in enumerate(local_tile_rows(height, tile_size, r, world)):
            y0 = tile_row * tile_size
            h = min(tile_size, height - y0)
            idx[y0:y0 + h] = torch.arange(r * padded_rows + k * tile_size, r * padded_rows + k * tile_size + h)
    return idx.to(device) if device is not None else idx


def deinterleave(bands: list[torch.Tensor] | torch.Tensor, frame: torch.Tensor, tile_size: int,
                 row_index: torch.Tensor | None = None) -> torch.Tensor:
    """bands: one [padded_rows, W, C] tensor per rank (or their concatenation [world*padded_rows, W, C])
    ->  frame [H, W, C] with rows in image order, as ONE gather over rows."""
    if isinstance(bands, (list, tuple)):
        world, padded = len(bands), bands[0].shape[0]
        stacked = bands[0] if world == 1 else torch.cat(list(bands), dim=0)
    else:
        stacked = bands
        padded = padded_band_rows(frame.shape[0], tile_size, 1) if row_index is None else None
        world = stacked.shape[0] // padded if padded else None
    if row_index is None:
        row_index = frame_row_index(frame.shape[0], tile_size, world, padded, stacked.device)
    torch.index_select(stacked, 0, row_index, out=frame)
    return frame


def gather_frame(band: torch.Tensor, frame: torch.Tensor | None, tile_size: int,
                 gathered: torch.Tensor | None = None, row_index: torch.Tensor | None = None) -> torch.Tensor | None:
    """Collective: every rank passes its padded band; rank 0 returns the assembled frame.
    The only exchange of the path: one gather, W*H*C*4/world bytes per non-root rank.
    gathered (rank 0): optional preallocated [world*padded_rows, W, C] receive buffer."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return deinterleave([band], frame, tile_size, row_index)
    if rank == 0:
        if gathered is None:
            gathered = torch.empty((world * band.shape[0],) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
        dist.gather(band, list(gathered.chunk(world, dim=0)), dst=0)
        return deinterleave(gathered, frame, tile_size,
                            row_index if row_index is not None else
                            frame_row_index(frame.shape[0], tile_size, world, band.shape[0], band.device))
    dist.gather(band, None, dst=0)
    return None


class PeerFrame:
    """The root's full-frame image mapped into every rank of the box (one process per GPU).

    Rank 0 owns a plain device allocation — the frame followed by one 32-bit flag per rank — and
    broadcasts its 64-byte IPC handle; the other ranks map it (peer access over NVLink).  Each rank's
    kernels then store their tile rows straight into the root's frame (Context.render_rows_into_frame)
    and the path's only exchange step shrinks to a barrier, `fence(stream)`: every other rank releases
    its flag in the root's memory after its kernels (stream-ordered, system scope), and the root's
    stream acquires all of them; after that (in stream order) rank 0 may read `frame`.
    `fence_all()` is the two-way form (a one-element NCCL all-reduce) for when the peers must also
    wait for the root, e.g. before they overwrite a frame the root is still copying out.
    No CUDA call of its own: allocation, mapping and flags go through the C ABI."""

    FLAG_BYTES = 4096  # one page after the frame: a 32-bit flag per rank

    def __init__(self, lib, height: int, width: int, local_device: int):
        self.lib, self.device = lib, local_device
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        dev = torch.device("cuda", local_device)
        frame_bytes = height * width * 16
        self.buffer = lib.DeviceBuffer(local_device, (frame_bytes + self.FLAG_BYTES,), dtype="uint8") if self.rank == 0 else None
        box = [self.buffer.ipc_handle() if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
        self.ptr = self.buffer.ptr if self.rank == 0 else lib.ipc_open(local_device, box[0])
        self.flags_ptr = self.ptr + frame_bytes
        self.frame = None
        self._whole = None
        if self.rank == 0:
            self._whole = torch.as_tensor(self.buffer, device=dev)
            self._whole[frame_bytes:].zero_()
            self.frame = self._whole[:frame_bytes].view(torch.float32).view(height, width, 4)
            torch.cuda.synchronize(dev)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.epoch = 0
        if self.world > 1:
            dist.barrier()  # flags are zeroed before anyone signals

    def fence(self, stream: int = 0):
        """One-way barrier: after this (in the root's stream order) every rank's rows of the current frame
        are in the root's image.  Non-root ranks only signal and run ahead."""
        self.epoch += 1
        if self.world == 1:
            return
        if self.rank == 0:
            self.lib.peer_wait(self.device, self.flags_ptr + 4, self.world - 1, self.epoch, 0, stream)
        else:
            self.lib.peer_signal(self.device, self.flags_ptr + 4 * self.rank, self.epoch, stream)

    def fence_all(self):
        """Two-way barrier on the current torch stream (NCCL): nobody passes before everybody arrived."""
        if self.world > 1:
            dist.all_reduce(self._flag)

    def close(self):
        if self.rank != 0 and self.ptr:
            self.lib.ipc_close(self.device, self.ptr)
        self.ptr = 0
        if self.buffer is not None:
            self.frame = None
            self._whole = None
            self.buffer.free()
            self.buffer = None
