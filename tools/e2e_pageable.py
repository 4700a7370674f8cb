#!/usr/bin/env python
"""End-to-end time of the host call with PAGEABLE destinations (what TileRenderer::render's Image is): headline frame,
host scene in, float image out to a plain numpy array.   MCSKIN_STAGED_COPY=0|1 python tools/e2e_pageable.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import synth_skin  # noqa: E402

cfg = _abi.default_config(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)
scene = lib.build_skin_scene(synth_skin(0), None)
cs = scene.as_c()
out = np.empty((1080, 1920, 4), dtype=np.float32)
out8 = np.empty((1080, 1920, 4), dtype=np.uint8)
for label, kw in (("f32", dict(out_f32=out)), ("f32+u8", dict(out_f32=out, out_u8=out8)), ("u8", dict(out_u8=out8, want_f32=False))):
    for _ in range(4):
        lib.render(cs, cfg, **kw)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        st = lib.render(cs, cfg, **kw)[2]
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"staged={os.environ.get('MCSKIN_STAGED_COPY', '1')} pageable {label}: {ms:.3f} ms per frame (kernels {st['ms_device']:.3f})")
