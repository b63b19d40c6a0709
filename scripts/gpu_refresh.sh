#!/bin/bash
# refresh of the per-kernel table (both flush modes, with CPU rows) and the ncu summary of the non-headline kernels
mkdir -p gpurun_out
timeout -k 5 600 python scripts/bench_kernels.py --cpu --flush write+read --out gpurun_out/kernels_final.json > gpurun_out/kernels_final.log 2>&1; echo "kernels exit $?"
timeout -k 5 300 python scripts/bench_kernels.py --flush write --out gpurun_out/kernels_final_dirty.json > gpurun_out/kernels_final_dirty.log 2>&1; echo "kernels (dirty) exit $?"
python - <<'PY'
import json
for line in open("gpurun_out/kernels_final.log"):
    try: r=json.loads(line)
    except Exception: continue
    if "avg_ms" in r: print(f"{r['op'][:78]:78s} {1e3*r['avg_ms']:9.1f} us  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
PY
timeout -k 5 600 ncu --set full --clock-control none --import-source on \
  -k regex:'gram_fast|kmeans_assign|ade_fde_fast|reconstruct_fast|reconstruct_bwd_fast|eig_jacobi|svd_small' \
  -s 8 -c 14 -f -o gpurun_out/prof_ops_final python scripts/exp/run_ops_once.py > gpurun_out/ncu_ops_final.log 2>&1
echo "ncu ops exit $?"; tail -1 gpurun_out/ncu_ops_final.log
