#!/usr/bin/env bash
# One GPU-box visit of a development round: GPU tests, the bench line, a few tuning lines.
mkdir -p gpurun_out
export MCSKIN_SKIP_REF_BUILD=1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
