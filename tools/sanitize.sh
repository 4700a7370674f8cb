#!/usr/bin/env bash
# compute-sanitizer over tools/sanitize_frames.py (run on the GPU box): memcheck, racecheck, synccheck, initcheck.
# Summaries land in gpurun_out/sanitize_<tool>.log; copy the tails into profiles/rNN/sanitizer_summary.md.
mkdir -p gpurun_out
export MCSKIN_SKIP_REF_BUILD=1
QUICK=${QUICK:-}
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = initcheck ] && extra="--track-unused-memory no"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 0 \
      python tools/sanitize_frames.py $QUICK > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_frames:' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
