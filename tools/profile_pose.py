#!/usr/bin/env python
"""A few serial frames (one stream, direct launches) of the headline frame in one pose, for ncu:
   ncu --set full --import-source on -k regex:k_ -s 10 -c 5 -o gpurun_out/pose python tools/profile_pose.py walking"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import synth_skin  # noqa: E402

pose = sys.argv[1] if len(sys.argv) > 1 else "walking"
cfg = _abi.default_config(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)
frame = torch.zeros((1080, 1920, 4), device="cuda")
scene = lib.build_skin_scene(synth_skin(0), None if pose == "standing" else pose)
ctx = lib.Context(0)
ctx.set_option("use_graphs", 0)
ctx.set_option("frame_lanes", 1)
ctx.set_scene(scene, cfg)
for _ in range(4):
    ctx.render_bands(0, 1, frame.data_ptr(), 0, 0)
    st = ctx.sync()
print(f"{pose}: frame {st['ms_device']:.3f} ms (primary {st['ms_primary']:.3f} shade {st['ms_shade']:.3f})")
ctx.close()
