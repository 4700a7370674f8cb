#!/usr/bin/env bash
# tools/diag_n.sh N : where a step of the N-GPU split goes — every rank into a local frame (no exchange), peer stores
# without the barrier, the full step; kernel-only bench lines.
mkdir -p gpurun_out; export MCSKIN_SKIP_REF_BUILD=1
n=$1
for diag in nofence full; do
  MCSKIN_BENCH_VERBOSE=1 MCSKIN_BENCH_DIAG=$diag python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 --kernel-only 2> gpurun_out/diag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$n $diag: frame %.4f ms kernels %.4f | serial primary %.4f shade %.4f' % (d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['serial_breakdown']['ms_primary_pass'], d['roofline']['serial_breakdown']['ms_shade_pass']))" | tee -a gpurun_out/diag_n.log
  grep "per-rank" gpurun_out/diag.err | tail -1 | tee -a gpurun_out/diag_n.log
done
