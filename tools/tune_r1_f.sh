#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
$T "default again"
for v in shade3 hit03 shade3hit03; do
  $T "variant $v" MCSKIN_LIB=$PWD/minecraftskin_raytracer_b200/_lib/variants/libmcskin_cuda_$v.so
done
$T "lanes4" MCSKIN_FRAME_LANES=4
$T "lanes2" MCSKIN_FRAME_LANES=2
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_serial.csv env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_b.log 2>&1
