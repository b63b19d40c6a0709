#!/bin/bash
# compute-sanitizer over a small-N pass of every entry point: memcheck, racecheck (shared-memory
# hazards: mbarrier rings, in-kernel block reductions), synccheck.  Summaries -> gpurun_out/r2_sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout -k 5 1200 compute-sanitizer --tool $tool --print-limit 20 $([ $tool = synccheck ] && echo --num-cuda-barriers 16384) python scripts/sanitize_target.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done' gpurun_out/r2_sanitize_$tool.log | tr '\n' ' ')"
done
