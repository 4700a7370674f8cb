#!/usr/bin/env bash
# per-round engine states: primary-pass splitting at 1/2/4/8-way shares on one GPU
mkdir -p gpurun_out; : > gpurun_out/tune.log
export MCSKIN_SKIP_REF_BUILD=1
for n in 8 4 2 1; do
  for pb in 2 4 8 16; do tools/tune_env.sh "split$n primary_blocks$pb" MCSKIN_BENCH_SPLIT=$n MCSKIN_PRIMARY_BLOCKS=$pb; done
done
for n in 8 4; do for ht in 8 32; do tools/tune_env.sh "split$n heavy_tiles$ht" MCSKIN_BENCH_SPLIT=$n MCSKIN_HEAVY_TILES=$ht; done; done
