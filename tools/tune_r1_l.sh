#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "default"
timeout 600 python tools/bench_configs.py --quick > gpurun_out/configs.md 2>&1; tail -12 gpurun_out/configs.md
