#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_descriptor.py -m gpu -q --timeout 240 -p no:cacheprovider -x > gpurun_out/t_desc_quick.log 2>&1; echo "desc tests exit $?"; tail -n 3 gpurun_out/t_desc_quick.log | cut -c1-300
for f in "" "--no-pdl" "" "--no-pdl"; do
  timeout -k 5 200 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --e2e-steps 1 $f 2>&1 | grep '^{' | python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line); print('launch', r['roofline']['launch'], 'ms/step', round(r['ms_per_step'],5), 'frac', round(r['roofline']['frac'], 4), 'min launch', round(r['roofline']['min_launch_ms'],5), r['clocks'])
"
done
