#!/usr/bin/env bash
# Round-2 profile capture (on the GPU box; every ncu pass follows a plain run of the same command that exited 0):
#   launches_serial.csv  : per-launch durations of one frame on one stream with direct launches
#   launches_default.csv : the default configuration (3 lanes, CUDA graph replay)
#   frame_serial.ncu-rep : ncu --set full of every kernel of one serial frame
mkdir -p gpurun_out
export MCSKIN_SKIP_REF_BUILD=1
python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain.json 2> gpurun_out/plain.err || { cat gpurun_out/plain.err; exit 1; }
env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/plain_serial.json 2>> gpurun_out/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_serial.csv env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s ${NCU_SKIP:-15} -c ${NCU_COUNT:-6} -f -o gpurun_out/frame_serial env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
