#!/usr/bin/env bash
# One rank's share of an N-way split on ONE GPU (MCSKIN_BENCH_SPLIT): event-timed frame + serial per-kernel list.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  MCSKIN_BENCH_SPLIT=$n tools/tune_env.sh "split$n" MCSKIN_BENCH_SPLIT=$n
done
for n in 1 8; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_split${n}_serial.csv \
   env MCSKIN_BENCH_SPLIT=$n MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_split$n.log 2>&1
done
