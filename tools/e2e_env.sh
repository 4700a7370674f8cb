#!/usr/bin/env bash
# Usage: tools/e2e_env.sh "label" VAR=VAL ...   — one full bench.py run (own arm) with the given environment; one line
# with the device-timed frame, the end-to-end frame and the kernels' share of it appended to gpurun_out/tune.log
label="$1"; shift
env "$@" python bench.py --no-extra --steps "${STEPS:-20}" --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']
print(f\"$label: frame {d['ms_per_step']:.3f} ms  e2e {e['ms_per_frame']:.3f} ms ({e['how'][-40:]})\")" | tee -a gpurun_out/tune.log
