#!/usr/bin/env bash
# On the GPU box: the GPU test suite, smoke(), both bench arms, then the round's profile capture.
mkdir -p gpurun_out
export MCSKIN_SKIP_REF_BUILD=1
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -4 gpurun_out/pytest_gpu.log | head -1)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "reference arm rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-400
bash tools/profile_r2.sh > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
