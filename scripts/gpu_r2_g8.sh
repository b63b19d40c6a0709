#!/bin/bash
# 8-GPU box: fused / sharded paths against the oracle at 8 ranks, then the bench with its stages at N = 8 and 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_g8.txt 2>&1
timeout -k 5 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --timeout 500 -p no:cacheprovider -k "[8]" > gpurun_out/r2_multi_tests_g8.log 2>&1; echo "multi tests exit $?"; tail -n 3 gpurun_out/r2_multi_tests_g8.log | cut -c1-300
for N in 8 4; do
  timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_g$N.log 2>&1; echo "bench $N exit $?"; tail -n 1 gpurun_out/r2_bench_g$N.log | cut -c1-200
done
