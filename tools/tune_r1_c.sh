#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default (lanes2 graph)" 
$T "lanes1" MCSKIN_FRAME_LANES=1
$T "lanes2 primaryblocks4" MCSKIN_PRIMARY_BLOCKS=4
$T "lanes2 primaryblocks6" MCSKIN_PRIMARY_BLOCKS=6
$T "lanes3 primaryblocks4" MCSKIN_PRIMARY_BLOCKS=4 MCSKIN_FRAME_LANES=3
$T "lanes2 primaryblocks4 shadeblocks6" MCSKIN_PRIMARY_BLOCKS=4 MCSKIN_SHADE_BLOCKS=6
$T "lanes2 primaryblocks4 shadeblocks12" MCSKIN_PRIMARY_BLOCKS=4 MCSKIN_SHADE_BLOCKS=12
$T "lanes2 primaryblocks4 prefetch" MCSKIN_PRIMARY_BLOCKS=4 MCSKIN_SHADOW_PREFETCH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_lanes1.csv env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/ncu_l1.log 2>&1
