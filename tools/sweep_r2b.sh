#!/usr/bin/env bash
# one part of an 8-way (and 4-way) tile split on one GPU: primary-pass splitting, grid sizes, lanes
mkdir -p gpurun_out; : > gpurun_out/tune.log
export MCSKIN_SKIP_REF_BUILD=1
for n in 8 4; do
  tools/tune_env.sh "split$n base" MCSKIN_BENCH_SPLIT=$n
  for pb in 4 8 16; do tools/tune_env.sh "split$n primary_blocks$pb" MCSKIN_BENCH_SPLIT=$n MCSKIN_PRIMARY_BLOCKS=$pb; done
  for sb in 2 3; do tools/tune_env.sh "split$n soft$sb" MCSKIN_BENCH_SPLIT=$n MCSKIN_SOFT_BLOCKS=$sb; done
  for sh in 4 16; do tools/tune_env.sh "split$n shade$sh" MCSKIN_BENCH_SPLIT=$n MCSKIN_SHADE_BLOCKS=$sh; done
  tools/tune_env.sh "split$n part3" MCSKIN_BENCH_SPLIT=$n MCSKIN_BENCH_PART=3
  tools/tune_env.sh "split$n part1" MCSKIN_BENCH_SPLIT=$n MCSKIN_BENCH_PART=1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_split8_serial.csv \
   env MCSKIN_BENCH_SPLIT=8 MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_split8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_split8_part3_serial.csv \
   env MCSKIN_BENCH_SPLIT=8 MCSKIN_BENCH_PART=3 MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_split8b.log 2>&1
