"""Multi-GPU partition of one frame.  The tile stream (per-tile RNG) makes any partition
made of whole reference tiles bit-identical to the single-device frame.  Two forms:

* tiles (default): the frame's tiles are dealt to the ranks by cost (lib.partition_tiles: the tiles the
  figure covers cost ~50x a background tile), every rank's kernels store its tiles straight into ONE frame —
  rank 0's device image mapped into every process over NVLink peer memory (PeerFrame), or one page-locked
  host image in shared memory that every GPU writes through its own PCIe link (HostFrame) — and the only
  exchange left is a barrier;
* rows (MCSKIN_EXCHANGE=gather, SURVEY's original plan): tile row j belongs to rank j % world, every rank
  renders its rows into a compact band image, rank 0 gathers the bands with NCCL and puts the rows back in order.

Batches shard by skin: skin i belongs to rank i % world (shard_batch), no exchange at all.

Pure torch / torch.distributed code with no CUDA calls of its own, so the same functions
run under NCCL on GPUs (bench.py) and under gloo on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the skins of a batch rendered by `rank` (skin i -> GPU i mod n)."""
    return list(range(rank, n_items, world))


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
        self._timed_out = torch.zeros(1, dtype=torch.int32, device=dev)  # set by the wait kernel if a rank never signals
        self.epoch = 0
        if self.world > 1:
            dist.barrier()  # flags are zeroed before anyone signals

    def fence(self, stream: int = 0):
        """One-way barrier: after this (in the root's stream order) every rank's tiles of the current frame
        are in the root's image.  Non-root ranks only signal.  Follow it with release() once the root is done with
        the frame: the peers then hold their NEXT frame's stores until the root has said so (they never run more
        than one frame ahead, and never overwrite a frame the root still reads).  A rank that never signals makes
        the root's wait give up after ~2 s; check_timeout() raises then."""
        self.epoch += 1
        if self.world == 1:
            return
        if self.rank == 0:
            self.lib.peer_wait(self.device, self.flags_ptr + 4, self.world - 1, self.epoch, self._timed_out.data_ptr(), stream)
        else:
            self.lib.peer_signal(self.device, self.flags_ptr + 4 * self.rank, self.epoch, stream)

    def release(self, stream: int = 0):
        """The other half of the handshake (stream-ordered like fence): the root releases flag 0 of its page when it
        is done with the frame; every other rank's stream waits for it (polling the root's memory over NVLink)."""
        if self.world == 1:
            return
        if self.rank == 0:
            self.lib.peer_signal(self.device, self.flags_ptr, self.epoch, stream)
        else:
            self.lib.peer_wait(self.device, self.flags_ptr, 1, self.epoch, 0, stream)

    def check_timeout(self):
        """Root, after synchronising: raises if a wait of this frame or an earlier one gave up on a rank."""
        if self.rank == 0 and self.world > 1 and int(self._timed_out.item()) != 0:
            raise RuntimeError("PeerFrame: a rank did not signal within the wait kernel's time limit; the frame is incomplete")

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


class HostFrame:
    """ONE page-locked host image shared by every rank of the box: a POSIX shared-memory segment created by
    rank 0, mapped by all ranks and registered with CUDA in each of them, so that every GPU's kernels store
    their tiles of the frame straight into it (zero-copy, one PCIe link per GPU instead of a single
    device-to-host copy of the whole frame from the root).  After the pixels: one 64-bit counter per rank and
    one for the root, the barrier between the processes (`publish` / `wait_all` / `release` / `wait_released`)."""

    def __init__(self, lib, height: int, width: int, register: bool = True):
        from multiprocessing import shared_memory

        import numpy as np
        self.lib = lib
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.frame_bytes = height * width * 16
        total = self.frame_bytes + 4096
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=total)
            name[0] = self.shm.name
        if self.world > 1:
            dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.bytes = np.ndarray((total,), dtype=np.uint8, buffer=self.shm.buf)
        self.frame = self.bytes[:self.frame_bytes].view(np.float32).reshape(height, width, 4)
        self.flags = self.bytes[self.frame_bytes:self.frame_bytes + 8 * (self.world + 1)].view(np.uint64)
        if self.rank == 0:
            self.flags[:] = 0
        # device address of the same pages (register=False: host-only use, e.g. the gloo tests on a CPU box)
        self.registered = bool(register)
        self.ptr = lib.host_register(self.bytes) if register else 0
        self.epoch = 0
        if self.world > 1:
            dist.barrier()

    def publish(self):
        """This rank's tiles of the current frame are in host memory (call after synchronising its stream)."""
        self.epoch += 1
        self.flags[self.rank] = self.epoch

    def wait_all(self, timeout_s: float = 20.0):
        """Root: every rank has published the current frame."""
        import time
        t0 = time.perf_counter()
        for r in range(self.world):
            while int(self.flags[r]) < self.epoch:
                if time.perf_counter() - t0 > timeout_s:
                    raise RuntimeError(f"HostFrame: rank {r} did not publish frame {self.epoch}")

    def release(self):
        """Root: the frame has been consumed, the ranks may overwrite it."""
        self.flags[self.world] = self.epoch

    def wait_released(self, timeout_s: float = 20.0):
        import time
        t0 = time.perf_counter()
        while int(self.flags[self.world]) < self.epoch:
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError(f"HostFrame: the root did not release frame {self.epoch}")

    def close(self):
        if self.registered:
            try:
                self.lib.host_unregister(self.bytes)
            except Exception:  # noqa: BLE001
                pass
        self.frame = self.flags = self.bytes = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass
