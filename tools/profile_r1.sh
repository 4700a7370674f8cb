#!/usr/bin/env bash
# Round-1 profile capture (run on the GPU box, after bench.py has exited 0 without ncu):
#   launches_default.csv : per-launch durations of the default configuration (3 lanes, CUDA graph replay)
#   launches_serial.csv  : the same frame on one stream with direct launches
#   frame_serial.ncu-rep : ncu --set full of every kernel of one serial frame
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err || exit 1
cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_serial.csv env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 65 -c 17 -o gpurun_out/frame_serial env MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0 python bench.py --steps 3 --warmup 3 --kernel-only > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
