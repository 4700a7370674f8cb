#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
$T "default again"
$T "lanes2" MCSKIN_FRAME_LANES=2
$T "lanes4" MCSKIN_FRAME_LANES=4
