#!/bin/bash
# peer-memory fused sharded k-means: NCCL test at world 2 + config-5 timings (run with gpurun --gpus 2 or more).  $1 = ranks
N=${1:-2}
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/test_gpu_multi_peer.log 2>&1
echo "test_gpu_multi (NCCL + peer) exit $?"; tail -n 15 gpurun_out/test_gpu_multi_peer.log | cut -c1-250
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  scripts/bench_multi.py > gpurun_out/multi_peer_g$N.log 2>&1
echo "bench_multi gpus=$N exit $?"; grep '^{' gpurun_out/multi_peer_g$N.log | cut -c1-900; tail -n 5 gpurun_out/multi_peer_g$N.log | grep -v '^{' | cut -c1-300
