#!/usr/bin/env bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
T=tools/tune_env.sh
$T "lanes1 nograph" MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0
$T "lanes1 graph" MCSKIN_FRAME_LANES=1
$T "lanes2 nograph nostagger" MCSKIN_FRAME_LANES=2 MCSKIN_GRAPHS=0 MCSKIN_STAGGER=0
$T "lanes2 nograph stagger" MCSKIN_FRAME_LANES=2 MCSKIN_GRAPHS=0
for L in 2 3 4 6 8; do
  $T "lanes$L graph stagger" MCSKIN_FRAME_LANES=$L
  $T "lanes$L graph nostagger" MCSKIN_FRAME_LANES=$L MCSKIN_STAGGER=0
done
$T "lanes4 graph stagger shadeblocks4" MCSKIN_FRAME_LANES=4 MCSKIN_SHADE_BLOCKS=4
$T "lanes4 graph stagger shadeblocks6" MCSKIN_FRAME_LANES=4 MCSKIN_SHADE_BLOCKS=6
$T "lanes4 graph stagger primaryblocks4" MCSKIN_FRAME_LANES=4 MCSKIN_PRIMARY_BLOCKS=4
$T "lanes4 graph stagger levels2" MCSKIN_FRAME_LANES=4 MCSKIN_WAVE_LEVELS=2
$T "lanes4 graph stagger levels4" MCSKIN_FRAME_LANES=4 MCSKIN_WAVE_LEVELS=4
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
