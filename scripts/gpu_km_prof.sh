#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'kmeans_assign' \
  -s 1 -c 1 -f -o gpurun_out/prof_lloyd python scripts/exp/run_lloyd_once.py > gpurun_out/ncu_lloyd.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_lloyd.log
