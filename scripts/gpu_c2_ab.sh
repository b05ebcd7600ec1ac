#!/bin/bash
mkdir -p gpurun_out
for mb in 3 4 5; do
  B200AG_KERNEL_CACHE=0 B200AG_FUSED_MINBLOCKS=$mb timeout 300 python bench.py --steps 10 --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('minblocks', $mb, 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'parity', d['parity_max_rel_err_vs_reference'])
"
done
