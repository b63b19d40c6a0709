#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_svd_kmeans_metrics test_gpu_descriptor test_gpu_multi; do
  timeout -k 5 ${TEST_TIMEOUT:-420} python -m pytest tests/$f.py -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -n 25 gpurun_out/$f.log | cut -c1-300
done
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout -k 5 600 python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1; echo "kernels exit $?"; python - <<'PY'
import json
for line in open("gpurun_out/kernels.log"):
    try: r=json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    if "avg_ms" in r: print(f"{r['op'][:70]:70s} {1e3*r['avg_ms']:9.1f} us  {r['achieved_gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
PY
timeout -k 5 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_default.log | cut -c1-1800
timeout -k 5 300 python scripts/bench_forward.py > gpurun_out/forward_latency.log 2>&1; echo "forward latency exit $?"; cat gpurun_out/forward_latency.log | cut -c1-200
