#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_svd_kmeans_metrics.py -m gpu -q --timeout 240 -p no:cacheprovider -k "gram or init or basis" > gpurun_out/t_gram.log 2>&1; echo "gram tests exit $?"; tail -n 2 gpurun_out/t_gram.log | cut -c1-300
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'gram_fast|eig_jacobi' \
  -s 2 -c 3 -f -o gpurun_out/prof_gram python scripts/exp/run_ops_once.py > gpurun_out/ncu_gram.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_gram.log
