#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
$T "split8 default(auto levels)" MCSKIN_BENCH_SPLIT=8
$T "split8 levels4" MCSKIN_BENCH_SPLIT=8 MCSKIN_WAVE_LEVELS=4
$T "split4 default" MCSKIN_BENCH_SPLIT=4
