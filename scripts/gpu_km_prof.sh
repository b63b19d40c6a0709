#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_svd_kmeans_metrics.py -m gpu -q --timeout 240 -p no:cacheprovider -k "whole_fit or free_running" > gpurun_out/t_km.log 2>&1; echo "kmeans tests exit $?"; tail -n 5 gpurun_out/t_km.log | cut -c1-300
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'kmeans_assign' \
  -s 1 -c 1 -f -o gpurun_out/prof_lloyd python scripts/exp/run_lloyd_once.py > gpurun_out/ncu_lloyd.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_lloyd.log
