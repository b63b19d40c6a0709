#!/bin/bash
mkdir -p gpurun_out
for mode in write+read write; do
timeout -k 5 600 python scripts/bench_kernels.py --flush $mode --out gpurun_out/kernels_$mode.json > gpurun_out/kernels_$mode.log 2>&1; echo "kernels ($mode) exit $?"; python - <<PY
import json
for line in open("gpurun_out/kernels_$mode.log"):
    try: r=json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    if "avg_ms" in r: print(f"{r['op'][:78]:78s} {1e3*r['avg_ms']:9.1f} us  min {1e3*r['min_ms']:8.1f}  {r['achieved_gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
PY
done
