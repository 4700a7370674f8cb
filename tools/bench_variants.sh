#!/usr/bin/env bash
# Runs bench.py --kernel-only once per library variant under _lib/variants/ and prints pass times.
for so in minecraftskin_raytracer_b200/_lib/variants/libmcskin_cuda_*.so; do
  name=$(basename "$so" .so); name=${name#libmcskin_cuda_}
  MCSKIN_LIB="$PWD/$so" python bench.py --steps "${STEPS:-8}" --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print(f\"$name: frame {d['ms_per_step']:.3f} ms  primary {r['whole_step']['ms_primary_pass']:.3f}  shade {r['ms_per_launch']:.3f}  frac_step {r['whole_step']['frac']:.3f} clocks {d['clocks']}\")"
done
