#!/bin/bash
# N-GPU call: multi-rank tests and the bench with its config-5 stages.  usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_g$N.txt 2>&1
timeout -k 5 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_comm.py tests/test_gpu_descriptor.py -x -q -m gpu --timeout 600 -p no:cacheprovider -k "sharded or collectives or non_current" > gpurun_out/r2_multi_tests_g$N.log 2>&1; echo "multi tests exit $?"; tail -n 3 gpurun_out/r2_multi_tests_g$N.log | cut -c1-300
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_g$N.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/r2_bench_g$N.log | cut -c1-3500
