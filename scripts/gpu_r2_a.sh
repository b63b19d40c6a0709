#!/bin/bash
# round 2, call A: full GPU test-suite, smoke(), both bench arms, k-means whole-fit timing, host / PCIe topology
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
(lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); c=$(cat $d/class 2>/dev/null); if [ "${c:0:6}" = "0x0302" ]; then echo "$d numa=$(cat $f)"; fi; done; free -g | head -2) >> gpurun_out/r2_topo.txt 2>&1
timeout -k 5 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider -s > gpurun_out/r2a_tests.log 2>&1; echo "gpu tests exit $?"; grep -E "label mismatches|worst relative|passed|failed|error" gpurun_out/r2a_tests.log | tail -n 12 | cut -c1-260
timeout -k 5 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2a_smoke.log
bash scripts/gpu_km_quick.sh 2>&1 | tail -n 1
timeout -k 5 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2a_bench_reference.log 2>&1; echo "bench reference exit $?"; tail -n 1 gpurun_out/r2a_bench_reference.log | cut -c1-300
timeout -k 5 500 python bench.py > gpurun_out/r2a_bench.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/r2a_bench.log | cut -c1-3000
