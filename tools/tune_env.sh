#!/usr/bin/env bash
# Usage: tools/tune_env.sh "label" VAR=VAL ...   — one bench.py --kernel-only run with the given environment,
# one summary line appended to gpurun_out/tune.log
label="$1"; shift
env "$@" python bench.py --steps "${STEPS:-10}" --warmup 3 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
b=r['serial_breakdown']
print(f\"$label: frame {d['ms_per_step']:.3f} ms  kernels {r['ms_per_launch']:.3f} frac {r['frac']:.3f} | serial primary {b['ms_primary_pass']:.3f} shade {b['ms_shade_pass']:.3f} | launches {d['gpu_launches']} clk {d['clocks']['sm_mhz']}\")" | tee -a gpurun_out/tune.log
