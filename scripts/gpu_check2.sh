#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_svd_kmeans_metrics.py -m gpu -q --timeout 180 -p no:cacheprovider -k "eth_init or synthetic_basis" > gpurun_out/t_svd2.log 2>&1; echo "svd tests exit $?"; tail -n 5 gpurun_out/t_svd2.log | cut -c1-300
timeout -k 5 600 python scripts/bench_kernels.py --cpu > gpurun_out/kernels.log 2>&1; echo "kernels exit $?"; cut -c1-420 gpurun_out/kernels.log
timeout -k 5 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -n 3 gpurun_out/bench_default.log | cut -c1-2500
timeout -k 5 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -n 2 gpurun_out/bench_ref.log | cut -c1-1200
