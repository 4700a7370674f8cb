#!/usr/bin/env python
"""Frame times of the headline frame in every built-in pose (one GPU, device-resident).   python tools/pose_times.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import BUILTIN_POSE_ORDER, synth_skin  # noqa: E402

cfg = _abi.default_config(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)
frame = torch.zeros((1080, 1920, 4), device="cuda")
for pose in BUILTIN_POSE_ORDER:
    scene = lib.build_skin_scene(synth_skin(0), pose)
    ctx = lib.Context(0)
    ctx.set_scene(scene, cfg)
    for _ in range(5):
        ctx.render_bands(0, 1, frame.data_ptr(), 0, 0)
        st = ctx.sync()
    ctx.set_option("use_graphs", 0)
    ctx.set_option("frame_lanes", 1)
    for _ in range(3):
        ctx.render_bands(0, 1, frame.data_ptr(), 0, 0)
        s1 = ctx.sync()
    print(f"{pose:9s} frame {st['ms_device']:.3f} ms | one stream: primary {s1['ms_primary']:.3f} shade {s1['ms_shade']:.3f} | active pixels {s1['n_active_pixels']}")
    ctx.close()
