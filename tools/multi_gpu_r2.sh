#!/usr/bin/env bash
# On an N-GPU box (gpurun --gpus N): the multi-device GPU tests, then bench.py under torchrun at every N given.
#   tools/multi_gpu_r2.sh "2 4 8" [extra bench args]
mkdir -p gpurun_out
export MCSKIN_SKIP_REF_BUILD=1
NS="${1:-2}"; shift
nvidia-smi -L | head -8
( time python -m pytest tests -m gpu -q -x -k "two_devices or peer_memory or sharded_by_skin or render_multi or tile_sets_into_frame" ) > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi rc=$? $(tail -4 gpurun_out/pytest_multi.log | tr '\n' ' ')"
for n in $NS; do
  port=$((29500 + n))
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 "$@" ) > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench N=$n rc=$?"; tail -c 400 gpurun_out/bench_n$n.err | tail -3
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n ms_per_step %.4f e2e %.4f clocks %s split_check %s' % (d['ms_per_step'], d['e2e']['ms_per_frame'] or -1, d['clocks'], d['config']['split_check']))
    for k,v in d.get('extra_workloads',{}).items(): print('   ',k, {a:b for a,b in v.items() if a in ('ms_per_frame','skins_per_s','seconds')})
except Exception as e: print('no line', e)
PY
done
