#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "tanh or c2_full or shared_denominator" > gpurun_out/c2r3_tests.log 2>&1; echo "exit code $?"; tail -2 gpurun_out/c2r3_tests.log
for cfg in "2 3" "1 3" "2 2"; do
  set -- $cfg
  B200AG_KERNEL_CACHE=0 B200AG_HEAVY_VEC=$1 B200AG_FUSED_MINBLOCKS=$2 timeout 300 python bench.py --steps 10 --no-extras 2>gpurun_out/c2r3_err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('vec $1 minblocks $2', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'eager', round(d['execution']['eager_ms_per_step'],3), 'parity', d['parity_max_rel_err_vs_reference'])
"
done 2>&1 | tee gpurun_out/c2r3_sweep.log
