#!/usr/bin/env bash
# strong-scaling check on one box: bench.py at N = 2, 4 (and 8 when the box has them), a few settings each
mkdir -p gpurun_out; : > gpurun_out/scale.log
NG=$(nvidia-smi -L | wc -l)
run() {  # label N env...
  label="$1"; n="$2"; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/scale_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['roofline']['serial_breakdown']
print(f\"$label N=$n: frame {d['ms_per_step']:.3f} ms  kernels {d['roofline']['ms_per_launch']:.3f} | serial primary {b['ms_primary_pass']:.3f} shade {b['ms_shade_pass']:.3f} | e2e {d['e2e']['ms_per_frame']:.3f} | {d['config']['partition'][:60]}\")" | tee -a gpurun_out/scale.log
}
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  run "default" $n
  run "heavy0" $n MCSKIN_HEAVY_TILES=0
  run "heavy16" $n MCSKIN_HEAVY_TILES=16
  run "gather" $n MCSKIN_EXCHANGE=gather
  run "lanes2" $n MCSKIN_FRAME_LANES=2
done
tail -3 gpurun_out/scale_err.log
