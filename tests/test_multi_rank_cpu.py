"""The N>1 path on CPU: two gloo ranks each produce their interleaved tile rows (with the
CPU oracle standing in for the kernels), gather on rank 0 through the same bands.py code the
NCCL path of bench.py uses, and must reproduce the single-process frame bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, width, height, tile_size, spp, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from minecraftskin_raytracer_b200 import bands, lib
    from minecraftskin_raytracer_b200.scene import synth_skin
    from oracle.harness import Oracle
    from tests.scenes import make_config
    orc = Oracle()
    scene = lib.build_skin_scene(synth_skin(2), "walking")
    cfg = make_config(width=width, height=height, tile_size=tile_size, samples_per_pixel=spp, max_bounces=2)
    # this rank's tile rows, rendered tile by tile into a full-size scratch image, then packed as a band
    scratch = np.zeros((height, width, 4), dtype=np.float32)
    scratch[..., 3] = 1
    tiles = orc.generate_tiles(width, height, tile_size)
    mine = set(bands.local_tile_rows(height, tile_size, rank, world))
    for t in tiles:
        if t["y"] // tile_size in mine:
            scratch = orc.render_tile(scene, cfg, tuple(t), scratch)
    band = torch.zeros((bands.padded_band_rows(height, tile_size, world), width, 4))
    for k, row in enumerate(bands.local_tile_rows(height, tile_size, rank, world)):
        y0 = row * tile_size
        h = min(tile_size, height - y0)
        band[k * tile_size:k * tile_size + h] = torch.from_numpy(scratch[y0:y0 + h])
    assert bands.band_pixel_rows(height, tile_size, rank, world) == sum(min(tile_size, height - r * tile_size) for r in mine)
    frame = torch.zeros((height, width, 4)) if rank == 0 else None
    result = bands.gather_frame(band, frame, tile_size)
    if rank == 0:
        np.save(out_path, result.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 70), (2, 64), (3, 50)])
def test_band_gather_reproduces_frame(tmp_path, oracle, mclib, world, height):
    from minecraftskin_raytracer_b200.scene import synth_skin
    from tests.scenes import make_config
    width, tile_size, spp = 48, 16, 2
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), width, height, tile_size, spp, out), nprocs=world, join=True)
    scene = mclib.build_skin_scene(synth_skin(2), "walking")
    want = oracle.render(scene, make_config(width=width, height=height, tile_size=tile_size, samples_per_pixel=spp, max_bounces=2))
    got = np.load(out)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _tile_worker(rank, world, port, width, height, tile_size, spp, out_path):
    """The tile-set split (bench.py's N>1 e2e path) on CPU: every rank renders its cost-balanced tile set — the
    CPU oracle standing in for the kernels — straight into ONE shared host frame, publishes, and the root
    waits for all of them; two frames, so the release / wait_released handshake runs as well."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from minecraftskin_raytracer_b200 import bands, lib
    from minecraftskin_raytracer_b200.scene import synth_skin
    from oracle.harness import Oracle
    from tests.scenes import make_config
    orc = Oracle()
    scene = lib.build_skin_scene(synth_skin(2), "walking")
    cfg = make_config(width=width, height=height, tile_size=tile_size, samples_per_pixel=spp, max_bounces=2)
    tiles = orc.generate_tiles(width, height, tile_size)
    mine = lib.partition_tiles(scene, cfg, world, rank)
    host = bands.HostFrame(lib, height, width, register=False)
    for frame_no in range(2):
        if rank == 0:
            host.frame[...] = -1.0
        dist.barrier()
        scratch = np.zeros((height, width, 4), dtype=np.float32)
        for i in mine:
            t = tiles[i]
            scratch = orc.render_tile(scene, cfg, tuple(t), scratch)
            host.frame[t["y"]:t["y"] + t["height"], t["x"]:t["x"] + t["width"]] = scratch[t["y"]:t["y"] + t["height"], t["x"]:t["x"] + t["width"]]
        host.publish()
        if rank == 0:
            host.wait_all()
            np.save(out_path, np.array(host.frame))
            host.release()
        else:
            host.wait_released()
    assert bands.shard_batch(10, rank, world) == list(range(rank, 10, world))
    dist.barrier()
    host.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 70), (3, 50)])
def test_tile_sets_into_shared_host_frame(tmp_path, oracle, mclib, world, height):
    from minecraftskin_raytracer_b200.scene import synth_skin
    from tests.scenes import make_config
    width, tile_size, spp = 48, 16, 2
    out = str(tmp_path / "frame.npy")
    mp.spawn(_tile_worker, args=(world, _free_port(), width, height, tile_size, spp, out), nprocs=world, join=True)
    scene = mclib.build_skin_scene(synth_skin(2), "walking")
    want = oracle.render(scene, make_config(width=width, height=height, tile_size=tile_size, samples_per_pixel=spp, max_bounces=2))
    got = np.load(out)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
