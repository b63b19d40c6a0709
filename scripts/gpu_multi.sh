#!/bin/bash
# Multi-GPU checks (run with gpurun --gpus N).  $1 = N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.csv
timeout -k 5 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 240 -p no:cacheprovider > gpurun_out/test_gpu_multi_n$N.log 2>&1
echo "test_gpu_multi (NCCL) exit $?"; tail -n 3 gpurun_out/test_gpu_multi_n$N.log | cut -c1-200
for g in 1 $N; do
  if [ "$g" = "1" ]; then
    timeout -k 5 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_g1_of$N.log 2>&1
  else
    timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g --no-cpu-baseline > gpurun_out/bench_g${g}_of$N.log 2>&1
  fi
  echo "bench gpus=$g exit $?"; grep '^{' gpurun_out/bench_g${g}_of$N.log | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print({k: r[k] for k in ('n_gpus', 'value', 'ms_per_step', 'gpu_launches')}, 'roofline', round(r['roofline']['frac'], 3), 'e2e', round(r['e2e']['value'] / 1e6, 1), 'M/s', r['clocks'])
"
done
for g in 1 $N; do
  if [ "$g" = "1" ]; then
    timeout -k 5 300 python scripts/bench_multi.py > gpurun_out/multi_g1_of$N.log 2>&1
  else
    timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 \
      scripts/bench_multi.py > gpurun_out/multi_g${g}_of$N.log 2>&1
  fi
  echo "bench_multi gpus=$g exit $?"; grep '^{' gpurun_out/multi_g${g}_of$N.log | cut -c1-700
done
