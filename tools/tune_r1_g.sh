#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
$T "default again"
$T "lanes2" MCSKIN_FRAME_LANES=2
$T "lanes4" MCSKIN_FRAME_LANES=4
$T "lanes1" MCSKIN_FRAME_LANES=1
$T "lanes3 primaryblocks9" MCSKIN_PRIMARY_BLOCKS=9
