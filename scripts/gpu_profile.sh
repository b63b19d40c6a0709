#!/bin/bash
# ncu captures of the bench command (run under gpurun, one GPU).  $1 = round tag
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
echo "launch list exit $?"
# full capture of the headline kernel (3 launches)
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:project_reconstruct -s 5 -c 3 -f \
  -o gpurun_out/prof_pr_$TAG python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
  > gpurun_out/ncu_pr_$TAG.log 2>&1
echo "ncu full exit $?"
