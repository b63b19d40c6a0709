#!/bin/bash
# 8-GPU validation: fused / sharded paths against the oracle at 8 ranks, the bench with its stages, raw PCIe ceiling
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_g$N.txt 2>&1
timeout -k 5 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --timeout 500 -p no:cacheprovider -k "[$N]" > gpurun_out/r2_multi_tests_g$N.log 2>&1; echo "multi tests exit $?"; tail -n 3 gpurun_out/r2_multi_tests_g$N.log | cut -c1-300
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/exp/pcie_probe.py > gpurun_out/r2_pcie_g$N.log 2>&1; echo "pcie probe exit $?"; tail -n 1 gpurun_out/r2_pcie_g$N.log | cut -c1-800
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_g$N.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/r2_bench_g$N.log | cut -c1-3800
