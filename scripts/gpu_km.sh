#!/bin/bash
# k-means parity + per-kernel timings
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_svd_kmeans_metrics.py tests/test_gpu_multi.py -m gpu -q --timeout 240 -p no:cacheprovider -k "kmeans or KMeans or anchor or seed" > gpurun_out/t_km.log 2>&1; echo "kmeans tests exit $?"; tail -n 12 gpurun_out/t_km.log | cut -c1-300
timeout -k 5 600 python scripts/bench_kernels.py --out gpurun_out/kernels_km.json > gpurun_out/kernels_km.log 2>&1; echo "kernels exit $?"; python - <<'PY'
import json
for line in open("gpurun_out/kernels_km.log"):
    try: r=json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    if "avg_ms" in r: print(f"{r['op'][:70]:70s} {1e3*r['avg_ms']:9.1f} us  {r['achieved_gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
PY
