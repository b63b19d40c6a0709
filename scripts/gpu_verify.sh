#!/bin/bash
# what the driver runs at round end, on one B200: GPU tests, smoke(), both bench arms (+ the ncu launch list of the bench)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/t_verify.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/t_verify.log | cut -c1-300
timeout -k 5 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_verify.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke_verify.log
timeout -k 5 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference_verify.log 2>&1; echo "bench reference exit $?"; tail -n 1 gpurun_out/bench_reference_verify.log | cut -c1-200
timeout -k 5 400 python bench.py > gpurun_out/bench_verify.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_verify.log | cut -c1-2200
if [ "$1" = "launches" ]; then
  timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_verify.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu_verify.log 2>&1
  echo "launch list exit $?"
fi
