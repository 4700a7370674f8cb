#!/usr/bin/env python
"""Where the primary pass of one part of an N-way tile split spends its time: per-block entry / exit times
(option "debug_primary_timing").   python tools/primary_block_times.py [N] [part]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import synth_skin  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    parts = [int(sys.argv[2])] if len(sys.argv) > 2 else list(range(n))
    scene = lib.build_skin_scene(synth_skin(0), None)
    cfg = _abi.default_config(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)
    frame = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.float32, device="cuda:0")
    for part in parts:
        ctx = lib.Context(0)
        ctx.set_option("use_graphs", 0)
        ctx.set_option("frame_lanes", 1)
        ctx.set_option("debug_primary_timing", 1)
        ctx.set_scene(scene, cfg)
        tiles = lib.partition_tiles(scene, cfg, n, part) if n > 1 else np.arange(60 * 34, dtype=np.int32)
        for _ in range(3):
            ctx.render_tiles_into_frame(tiles, frame.data_ptr(), 0, 0)
            st = ctx.sync()
        t = ctx.debug_block_times()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        dur = (t[:, 1] - t[:, 0]).astype(np.float64) / 1e3
        start = (t[:, 0] - t0).astype(np.float64) / 1e3
        end = (t[:, 1] - t0).astype(np.float64) / 1e3
        order = np.argsort(-end)
        print(f"part {part}/{n}: {len(t)} blocks, primary pass {st['ms_primary'] * 1e3:.1f} us, last block ends at {end.max():.1f} us; "
              f"block duration mean {dur.mean():.1f} max {dur.max():.1f} us; blocks starting after 5 us: {(start > 5).sum()}")
        for i in order[:6]:
            tile = int(t[i, 2])
            print(f"   tile ({tile % 60:2d},{tile // 60:2d}) part {int(t[i, 3]) & 0xffff}/{int(t[i, 3]) >> 16}: start {start[i]:6.1f} dur {dur[i]:6.1f} end {end[i]:6.1f} us")
        ctx.close()


if __name__ == "__main__":
    main()
