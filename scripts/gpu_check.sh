#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, bench; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
python - <<'PY' > gpurun_out/host.txt 2>&1
import os, torch
print("cpus", os.cpu_count(), "torch", torch.__version__, "threads", torch.get_num_threads(), "cap", torch.backends.cpu.get_cpu_capability())
PY
for f in test_gpu_svd_kmeans_metrics test_gpu_descriptor; do
  timeout -k 5 ${TEST_TIMEOUT:-420} python -m pytest tests/$f.py -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -n 40 gpurun_out/$f.log | cut -c1-300
done
timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
for v in 1 2 3 4; do
  timeout -k 5 200 python bench.py --steps 100 --warmup 10 --variant $v --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_v$v.log 2>&1; echo "bench v$v exit $?"; tail -n 2 gpurun_out/bench_v$v.log | cut -c1-1500
done
