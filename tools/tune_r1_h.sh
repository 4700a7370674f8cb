#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
$T "heavy0 (no heavy split)" MCSKIN_HEAVY_TILES=0
$T "heavy16" MCSKIN_HEAVY_TILES=16
$T "heavy32" MCSKIN_HEAVY_TILES=32
