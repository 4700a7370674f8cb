#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
for L in 2 3 4 5 6; do
  $T "lanes$L" MCSKIN_FRAME_LANES=$L
done
$T "lanes3 shadeblocks12" MCSKIN_FRAME_LANES=3 MCSKIN_SHADE_BLOCKS=12
$T "lanes3 shadeblocks16" MCSKIN_FRAME_LANES=3 MCSKIN_SHADE_BLOCKS=16
$T "lanes4 shadeblocks12" MCSKIN_FRAME_LANES=4 MCSKIN_SHADE_BLOCKS=12
$T "lanes4 shadeblocks4" MCSKIN_FRAME_LANES=4 MCSKIN_SHADE_BLOCKS=4
$T "lanes3 levels2" MCSKIN_FRAME_LANES=3 MCSKIN_WAVE_LEVELS=2
$T "lanes3 levels4" MCSKIN_FRAME_LANES=3 MCSKIN_WAVE_LEVELS=4
MCSKIN_FRAME_LANES=3 python bench.py --no-cpu-baseline > gpurun_out/bench_l3.json 2> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench_l3.json
MCSKIN_FRAME_LANES=3 MCSKIN_OVERLAP_COPY=0 python bench.py --no-cpu-baseline > gpurun_out/bench_l3_nooverlap.json 2>> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench_l3_nooverlap.json
python tools/bench_configs.py --quick > gpurun_out/configs.md 2>&1; tail -15 gpurun_out/configs.md
