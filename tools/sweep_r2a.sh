#!/usr/bin/env bash
# lanes x soft-shadow blocks x queue levels on the headline frame, then one rank's share of an N-way split
mkdir -p gpurun_out; : > gpurun_out/tune.log
export MCSKIN_SKIP_REF_BUILD=1
for lanes in 1 2 3; do for soft in 2 3 4; do
  tools/tune_env.sh "lanes$lanes soft$soft" MCSKIN_FRAME_LANES=$lanes MCSKIN_SOFT_BLOCKS=$soft
done; done
for lv in 1 2 3; do for lanes in 1 3; do
  tools/tune_env.sh "levels$lv lanes$lanes" MCSKIN_WAVE_LEVELS=$lv MCSKIN_FRAME_LANES=$lanes
done; done
for n in 2 4 8; do
  tools/tune_env.sh "split$n" MCSKIN_BENCH_SPLIT=$n
  tools/tune_env.sh "split$n lanes1" MCSKIN_BENCH_SPLIT=$n MCSKIN_FRAME_LANES=1
  tools/tune_env.sh "split$n soft2" MCSKIN_BENCH_SPLIT=$n MCSKIN_SOFT_BLOCKS=2
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_split8_serial.csv \
   env MCSKIN_BENCH_SPLIT=8 MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_split8.log 2>&1
