#!/bin/bash
# sanity after restore: GPU tests + smoke, then source-level ncu of the k-means and Gram kernels
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "gpu tests exit $?"; tail -n 4 gpurun_out/t_all.log | cut -c1-300
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'gram_fast|kmeans_assign' \
  -s 2 -c 6 -f -o gpurun_out/prof_km_gram python scripts/exp/run_ops_once.py > gpurun_out/ncu_km_gram.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_km_gram.log
