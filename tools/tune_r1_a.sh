#!/usr/bin/env bash
# round-1 tuning sweep A (run on the GPU box): tests, then environment and build variants
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "old-like lanes1 div8 noseedcache" MCSKIN_FRAME_LANES=1 MCSKIN_DEEP_GRID_DIV=8 MCSKIN_CACHE_TILE_SEEDS=0
$T "lanes1 div1" MCSKIN_FRAME_LANES=1 MCSKIN_DEEP_GRID_DIV=1
$T "lanes1 div2" MCSKIN_FRAME_LANES=1 MCSKIN_DEEP_GRID_DIV=2
$T "lanes1 div1 prefetch" MCSKIN_FRAME_LANES=1 MCSKIN_SHADOW_PREFETCH=1
$T "lanes2" MCSKIN_FRAME_LANES=2
$T "lanes2 prefetch" MCSKIN_FRAME_LANES=2 MCSKIN_SHADOW_PREFETCH=1
$T "lanes3" MCSKIN_FRAME_LANES=3
$T "lanes4" MCSKIN_FRAME_LANES=4
$T "lanes6" MCSKIN_FRAME_LANES=6
$T "lanes2 shadeblocks4" MCSKIN_FRAME_LANES=2 MCSKIN_SHADE_BLOCKS=4
$T "lanes2 shadeblocks6" MCSKIN_FRAME_LANES=2 MCSKIN_SHADE_BLOCKS=6
$T "lanes2 shadeblocks12" MCSKIN_FRAME_LANES=2 MCSKIN_SHADE_BLOCKS=12
$T "lanes2 levels1" MCSKIN_FRAME_LANES=2 MCSKIN_WAVE_LEVELS=1
$T "lanes2 levels2" MCSKIN_FRAME_LANES=2 MCSKIN_WAVE_LEVELS=2
$T "lanes2 levels4" MCSKIN_FRAME_LANES=2 MCSKIN_WAVE_LEVELS=4
$T "lanes1 levels2" MCSKIN_FRAME_LANES=1 MCSKIN_WAVE_LEVELS=2
$T "lanes1 levels4" MCSKIN_FRAME_LANES=1 MCSKIN_WAVE_LEVELS=4
$T "lanes2 primaryblocks4" MCSKIN_FRAME_LANES=2 MCSKIN_PRIMARY_BLOCKS=4
$T "lanes2 primaryblocks18" MCSKIN_FRAME_LANES=2 MCSKIN_PRIMARY_BLOCKS=18
for v in lcg1 lcg2 lcg3 lcg4 shade4 hit04 shadow5; do
  $T "variant $v lanes1" MCSKIN_LIB=$PWD/minecraftskin_raytracer_b200/_lib/variants/libmcskin_cuda_$v.so MCSKIN_FRAME_LANES=1
  $T "variant $v lanes2" MCSKIN_LIB=$PWD/minecraftskin_raytracer_b200/_lib/variants/libmcskin_cuda_$v.so MCSKIN_FRAME_LANES=2
done
# launch list of the default configuration and of the one-lane configuration
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_lanes1.csv env MCSKIN_FRAME_LANES=1 python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/ncu_l1.log 2>&1
