#!/bin/bash
# 8-GPU refresh of the headline bench + config 5 (run with gpurun --gpus 8)
mkdir -p gpurun_out
g=8
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $g --no-cpu-baseline > gpurun_out/scale_bench_g$g.log 2>&1
echo "bench gpus=$g exit $?"; grep '^{' gpurun_out/scale_bench_g$g.log | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print({k: r[k] for k in ('n_gpus', 'value', 'ms_per_step')}, 'roofline', round(r['roofline']['frac'], 3), 'e2e', round(r['e2e']['value'] / 1e6, 1), 'M/s', r['clocks'])
"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29528 \
  scripts/bench_multi.py > gpurun_out/scale_multi_g$g.log 2>&1
echo "bench_multi gpus=$g exit $?"; grep '^{' gpurun_out/scale_multi_g$g.log | cut -c1-700
