#!/usr/bin/env bash
# one rank's share of an N-way split, timed on one GPU (MCSKIN_BENCH_SPLIT): which settings suit small shares?
mkdir -p gpurun_out; : > gpurun_out/tune.log
T=tools/tune_env.sh
for N in 4 8; do
  $T "split$N default" MCSKIN_BENCH_SPLIT=$N
  $T "split$N lanes1" MCSKIN_BENCH_SPLIT=$N MCSKIN_FRAME_LANES=1
  $T "split$N lanes2" MCSKIN_BENCH_SPLIT=$N MCSKIN_FRAME_LANES=2
  $T "split$N levels1" MCSKIN_BENCH_SPLIT=$N MCSKIN_WAVE_LEVELS=1
  $T "split$N levels2" MCSKIN_BENCH_SPLIT=$N MCSKIN_WAVE_LEVELS=2
  $T "split$N levels2 lanes2" MCSKIN_BENCH_SPLIT=$N MCSKIN_WAVE_LEVELS=2 MCSKIN_FRAME_LANES=2
  $T "split$N shadeblocks4" MCSKIN_BENCH_SPLIT=$N MCSKIN_SHADE_BLOCKS=4
  $T "split$N shadeblocks2" MCSKIN_BENCH_SPLIT=$N MCSKIN_SHADE_BLOCKS=2
  $T "split$N shadeblocks2 levels2" MCSKIN_BENCH_SPLIT=$N MCSKIN_SHADE_BLOCKS=2 MCSKIN_WAVE_LEVELS=2
done
