#!/bin/bash
# Round-end style verification on one B200: all GPU tests, smoke, both bench arms, per-kernel table, config 4, ncu launch
# list and full captures, sanitizer.  Everything lands in gpurun_out/ (scripts/collect_profiles.py copies the summaries
# into profiles/).  $1 = tag
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/t_all_$TAG.log 2>&1; echo "gpu tests exit $?"; grep -E "label mismatches|worst relative|passed|failed" gpurun_out/t_all_$TAG.log | tail -n 6 | cut -c1-300
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke_$TAG.log
timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference_$TAG.log 2>&1; echo "bench reference exit $?"; tail -n 1 gpurun_out/bench_reference_$TAG.log | cut -c1-300
timeout -k 5 500 python bench.py > gpurun_out/bench_default_$TAG.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_default_$TAG.log | cut -c1-3500
timeout -k 5 900 python scripts/bench_kernels.py --cpu --out gpurun_out/kernels_$TAG.json > gpurun_out/kernels_$TAG.log 2>&1; echo "kernels exit $?"; python - <<PY
import json
for line in open("gpurun_out/kernels_$TAG.log"):
    try: r=json.loads(line)
    except Exception: print(line.strip()[:300]); continue
    if "avg_ms" in r: print(f"{r['op'][:78]:78s} {1e3*r['avg_ms']:9.1f} us  {r['achieved_gbs']:8.1f} GB/s  {100*r['frac_of_measured_hbm_peak']:5.1f}%")
    elif "seconds" in r: print(f"CPU {r['op'][:70]:70s} {1e3*r['seconds']:9.1f} ms ({r['cores']} cores)")
PY
timeout -k 5 300 python scripts/bench_forward.py > gpurun_out/forward_latency_$TAG.log 2>&1; echo "forward latency exit $?"; tail -n 4 gpurun_out/forward_latency_$TAG.log | cut -c1-300
timeout -k 5 600 python scripts/run_config4.py > gpurun_out/config4_$TAG.log 2>&1; echo "config 4 exit $?"; tail -n 1 gpurun_out/config4_$TAG.log | cut -c1-900
bash scripts/gpu_km_quick.sh > gpurun_out/km_quick_$TAG.log 2>&1; tail -n 12 gpurun_out/km_quick_$TAG.log
bash scripts/gpu_init_quick.sh > gpurun_out/init_quick_$TAG.log 2>&1; tail -n 9 gpurun_out/init_quick_$TAG.log
# launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
echo "launch list exit $?"
# full capture of the headline kernel (3 launches) and of the other hot kernels
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:project_reconstruct -s 5 -c 3 -f \
  -o gpurun_out/prof_pr_$TAG python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stages --e2e-steps 1 > gpurun_out/ncu_pr_$TAG.log 2>&1
echo "ncu headline exit $?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on \
  -k regex:'gram_fast|kmeans_assign|kmeans_seed|ade_fde_fast|reconstruct_fast|reconstruct_bwd_fast|eig_jacobi|svd_small' \
  -s 9 -c 12 -f -o gpurun_out/prof_ops_$TAG python scripts/exp/run_ops_once.py > gpurun_out/ncu_ops_$TAG.log 2>&1
echo "ncu ops exit $?"; tail -2 gpurun_out/ncu_ops_$TAG.log
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -s 2 -c 1 -f -o gpurun_out/prof_lloyd_$TAG python scripts/exp/km_profile_target.py > gpurun_out/ncu_lloyd_$TAG.log 2>&1
echo "ncu lloyd exit $?"
if [ "$2" = "sanitize" ]; then bash scripts/gpu_sanitize.sh; fi
