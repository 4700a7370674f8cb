#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/gpu_tests.log
timeout 600 python tools/bench_configs.py > gpurun_out/configs.md 2>&1; tail -14 gpurun_out/configs.md
python - <<'PY' 2>&1 | tee gpurun_out/batch.log
import time, torch, sys
sys.path.insert(0, ".")
from minecraftskin_raytracer_b200 import _abi, lib
from minecraftskin_raytracer_b200.scene import synth_skin
n = 512
cfg = _abi.default_config(width=256, height=256, samples_per_pixel=4, max_bounces=2)
scenes = [lib.build_skin_scene(synth_skin(i)) for i in range(n)]
out = torch.empty((n, 256, 256, 4), dtype=torch.float32, device="cuda:0")
for mode, group in ((1, 128), (1, 256), (0, 128)):
    ctx = lib.Context(0)
    ctx.set_option("batch_mode", mode)
    ctx.set_option("batch_group", group)
    torch.cuda.synchronize()
    ctx.render_batch(scenes[:group], cfg, out.data_ptr(), 0, 0); ctx.sync()
    for rep in range(2):
        t = time.perf_counter()
        ctx.render_batch(scenes, cfg, out.data_ptr(), 0, 0)
        t_launch = time.perf_counter() - t
        ctx.sync()
        dt = time.perf_counter() - t
        print(f"mode {mode} group {group}: {dt * 1e3 / n:.4f} ms per skin ({dt*1e3:.1f} ms for {n}; host part {t_launch*1e3:.1f} ms), {256 * 256 * 4 * n / dt / 1e6:.0f} Msamples/s")
    ctx.close()
PY
