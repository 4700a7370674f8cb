#!/usr/bin/env bash
# Runs bench.py --kernel-only once per library variant under _lib/variants/ (build.build_variant) and prints one line each.
for so in minecraftskin_raytracer_b200/_lib/variants/libmcskin_cuda_*.so; do
  name=$(basename "$so" .so); name=${name#libmcskin_cuda_}
  tools/tune_env.sh "variant $name" MCSKIN_LIB="$PWD/$so" "$@"
done
