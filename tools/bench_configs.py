#!/usr/bin/env python
"""Times the BASELINE.json configurations other than the headline on one GPU (device-resident,
CUDA events) and prints a markdown table; they are parity-test cases, not bench lines, but their
sizes exercise chunking, 4K/8K frames and batches.   python tools/bench_configs.py [--quick]"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import synth_skin  # noqa: E402

CONFIGS = [
    ("C1 512x512 1spp 2b (64x64 skin)", 1, "64x64", None, dict(width=512, height=512, samples_per_pixel=1, max_bounces=2)),
    ("C2 1080p 4spp 4b (legacy skin)", 2, "legacy", None, dict(width=1920, height=1080, samples_per_pixel=4, max_bounces=4)),
    ("headline 1080p 16spp 4b", 0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)),
    ("headline, pose walking", 0, "64x64", "walking", dict(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)),
    ("headline, hard shadows", 0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=16, max_bounces=4, soft_shadows=0)),
    ("GUI defaults 1080p 64spp 4b AO16 DOF", 0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=64, max_bounces=4, ao_enabled=1, ao_samples=16, dof_enabled=1, aperture=0.3)),
    ("C3 4K 16spp 4b (slim skin)", 3, "slim", None, dict(width=3840, height=2160, samples_per_pixel=16, max_bounces=4)),
    ("C5 8K 64spp 8b", 5, "64x64", None, dict(width=7680, height=4320, samples_per_pixel=64, max_bounces=8)),
]


def time_frame(ctx, scene, cfg, steps):
    out = torch.empty((cfg.height, cfg.width, 4), dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx.set_scene(scene, cfg)
        ctx.render_bands(0, 1, out.data_ptr(), 0, stream.cuda_stream)
        ctx.sync()
        stream.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for _ in range(steps):
            ctx.render_bands(0, 1, out.data_ptr(), 0, stream.cuda_stream)
        t1.record(stream)
        stream.synchronize()
        stats = ctx.sync()
    return t0.elapsed_time(t1) / steps, stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    ctx = lib.Context(0)
    print("| config | ms/frame | Msamples/s | active pixels | launches |")
    print("|---|---|---|---|---|")
    for name, seed, kind, pose, over in CONFIGS:
        if args.quick and over["width"] > 4000:
            continue
        scene = lib.build_skin_scene(synth_skin(seed, kind), pose)
        cfg = _abi.default_config(**over)
        steps = 3 if over["width"] * over["height"] * over["samples_per_pixel"] > 2e8 else 10
        ms, stats = time_frame(ctx, scene, cfg, steps)
        samples = over["width"] * over["height"] * over["samples_per_pixel"]
        print(f"| {name} | {ms:.3f} | {samples / ms / 1e3:.0f} | {stats['n_active_pixels']} | {stats['n_kernel_launches']} |", flush=True)
    # C4: batch of skins at 256x256 4spp 2b
    n = 128 if args.quick else 512
    cfg = _abi.default_config(width=256, height=256, samples_per_pixel=4, max_bounces=2)
    scenes = [lib.build_skin_scene(synth_skin(i)) for i in range(n)]
    out = torch.empty((n, 256, 256, 4), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    ctx.render_batch(scenes[:min(n, 128)], cfg, out.data_ptr(), 0, 0)  # warm-up at the steady-state chunk size
    ctx.sync()
    t = time.perf_counter()
    ctx.render_batch(scenes, cfg, out.data_ptr(), 0, 0)
    ctx.sync()
    dt = time.perf_counter() - t
    print(f"| C4 batch of {n} skins 256x256 4spp 2b (one GPU's share) | {dt * 1e3 / n:.3f} per skin | {256 * 256 * 4 * n / dt / 1e6:.0f} | - | - |")
    ctx.close()


if __name__ == "__main__":
    main()
