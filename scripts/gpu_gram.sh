#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_gpu_svd_kmeans_metrics.py tests/test_gpu_descriptor.py tests/test_gpu_multi.py -m gpu -q --timeout 240 -p no:cacheprovider -k "gram or svd or init or basis or eig or parameter" > gpurun_out/t_gram.log 2>&1; echo "gram tests exit $?"; tail -n 6 gpurun_out/t_gram.log | cut -c1-300
timeout -k 5 600 python scripts/bench_kernels.py --out gpurun_out/kernels_gram.json > gpurun_out/kernels_gram.log 2>&1; echo "kernels exit $?"; python - <<'PY'
import json
for line in open("gpurun_out/kernels_gram.log"):
    try: r=json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    if "avg_ms" in r: print(f"{r['op'][:70]:70s} {1e3*r['avg_ms']:9.1f} us  {r['achieved_gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
PY
