#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests/test_gpu_svd_kmeans_metrics.py tests/test_gpu_multi.py tests/test_gpu_eval_loop.py -x -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/r2b_tests.log 2>&1; echo "gpu tests exit $?"; grep -E "label mismatches|passed|failed|rror" gpurun_out/r2b_tests.log | tail -n 12 | cut -c1-260
bash scripts/gpu_km_quick.sh 2>&1 | tail -n 8
