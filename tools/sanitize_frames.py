#!/usr/bin/env python
"""Workload for compute-sanitizer (tools/sanitize.sh): one small frame through every render mode the
library keeps, the overlapped copy-out path, graph capture + replay, batches, the tile-list partitions
and the peer flag kernels.  Small on purpose: the sanitizer slows kernels down by 10-100x.

    compute-sanitizer --tool memcheck python tools/sanitize_frames.py [--quick]
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    quick = "--quick" in sys.argv
    import torch

    from minecraftskin_raytracer_b200 import _abi, build
    build.build()
    from minecraftskin_raytracer_b200 import lib
    from minecraftskin_raytracer_b200.scene import synth_skin

    dev = torch.device("cuda", 0)
    scene = lib.build_skin_scene(synth_skin(1), "walking")
    standing = lib.build_skin_scene(synth_skin(2), None)
    cfgs = {
        "spp4": _abi.default_config(width=96, height=64, samples_per_pixel=4, max_bounces=2),
        "spp16": _abi.default_config(width=128, height=96, samples_per_pixel=16, max_bounces=4),
        "dof": _abi.default_config(width=64, height=64, samples_per_pixel=2, max_bounces=2, dof_enabled=1, aperture=0.3),
        "spp3_hard_ao": _abi.default_config(width=64, height=48, samples_per_pixel=3, max_bounces=2, soft_shadows=0,
                                            ao_enabled=1, ao_samples=4, tile_size=16),
        "spp1": _abi.default_config(width=64, height=64, samples_per_pixel=1, max_bounces=1),
    }
    modes = {
        "default": {},
        "all_active": {"force_all_active": 1},
        "megakernel": {"shade_mode": 1},
        "megakernel_warp": {"shade_mode": 2},
        "tiny_wave_budget": {"wave_budget_bytes": 1 << 20},
        "small_queue": {"wave_queue_pct": 20},
        "tiny_queue": {"wave_queue_pct": 1},
        "split_tiles": {"primary_blocks_per_sm": 100000},
        "one_lane_no_graph": {"frame_lanes": 1, "use_graphs": 0, "cache_tile_seeds": 0},
        "five_lanes": {"frame_lanes": 5},
    }
    if quick:
        modes = {k: modes[k] for k in ("default", "megakernel", "one_lane_no_graph")}
        cfgs = {k: cfgs[k] for k in ("spp4", "spp16", "dof")}
    done = 0
    # 1. host API: pageable, then page-locked destinations (copy-out overlapped with shading), graph capture + replay
    for name, cfg in cfgs.items():
        want, want_u8, _ = lib.render(scene, cfg, want_u8=True)
        f32 = torch.empty((cfg.height, cfg.width, 4), dtype=torch.float32, pin_memory=True).numpy()
        u8 = torch.empty((cfg.height, cfg.width, 4), dtype=torch.uint8, pin_memory=True).numpy()
        for _ in range(3):
            lib.render(scene, cfg, out_f32=f32, out_u8=u8)
            assert np.array_equal(f32.view(np.uint32), want.view(np.uint32)) and np.array_equal(u8, want_u8), name
        done += 4
    # 2. every render mode through a device-resident context: direct launches, capture, replay
    for mode, opts in modes.items():
        for name, cfg in cfgs.items():
            ctx = lib.Context(0)
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_scene(scene, cfg)
            out = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.float32, device=dev)
            out8 = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ref = None
            for _ in range(3):
                ctx.render_bands(0, 1, out.data_ptr(), out8.data_ptr(), 0)
                ctx.sync()
                got = out.cpu().numpy()
                ref = got if ref is None else ref
                assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (mode, name)
            ctx.close()
            done += 3
    # 3. partitions written into one frame (the multi-GPU layout) + the flag kernels of its barrier
    cfg = cfgs["spp4"]
    ctx = lib.Context(0)
    ctx.set_scene(scene, cfg)
    buf = lib.DeviceBuffer(0, (cfg.height * cfg.width * 16 + 4096,), dtype="uint8")
    torch.as_tensor(buf, device=dev).zero_()
    for world in (2, 3):
        for r in range(world):
            ctx.render_rows_into_frame(r, world, buf.ptr, 0, 0)
            ctx.sync()
        if hasattr(ctx, "render_tiles_into_frame"):
            from minecraftskin_raytracer_b200 import bands
            parts = bands.tile_partition(scene, cfg, world)
            for r in range(world):
                ctx.render_tiles_into_frame(parts[r], buf.ptr, 0, 0)
                ctx.sync()
    flags = buf.ptr + cfg.height * cfg.width * 16
    torch.cuda.synchronize()
    for epoch in (1, 2):
        lib.peer_signal(0, flags + 4, epoch)
        lib.peer_wait(0, flags + 4, 1, epoch, flags + 64)
    ctx.sync()
    ctx.close()
    buf.free()
    done += 8
    # 4. a batch of skins (grouped launches), twice
    n = 6
    bcfg = _abi.default_config(width=64, height=64, samples_per_pixel=4, max_bounces=2)
    scenes = [lib.build_skin_scene(synth_skin(100 + i, "legacy" if i % 3 == 0 else "64x64"), [None, "dab"][i % 2]) for i in range(n)]
    ctx = lib.Context(0)
    out = torch.zeros((n, 64, 64, 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        ctx.render_batch(scenes, bcfg, out.data_ptr(), 0, 0)
        ctx.sync()
    ctx.close()
    done += 2
    # 5. one tile, the single-ray entry points
    img = np.zeros((cfg.height, cfg.width, 4), dtype=np.float32)
    lib.render_tile(standing, cfg, (32, 32, 32, 32), img)
    rng = np.random.default_rng(1)
    rays = np.zeros(512, dtype=_abi.RAY_DTYPE)
    rays["origin"] = np.float32([0, 18, 50])
    d = rng.normal(size=(512, 3)).astype(np.float32) * 0.2 + np.float32([0, 0, -1])
    rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    hits = lib.intersect(scene, rays)
    lib.trace(scene, cfg, rays)
    keep = hits[hits["hit"] == 1]
    if len(keep):
        seeds = rng.integers(0, 2**32, size=len(keep), dtype=np.uint32)
        lib.soft_shadow(scene, keep["point"], keep["normal"], seeds, 8)
        lib.ambient_occlusion(scene, keep["point"], keep["normal"], seeds, 8, 3.0)
        lib.in_shadow(scene, keep["point"], keep["normal"], np.tile(np.float32(scene.light_pos), (len(keep), 1)))
    lib.aov(scene, cfg)
    print(f"sanitize_frames: {done} frames rendered, all outputs consistent")


if __name__ == "__main__":
    main()
