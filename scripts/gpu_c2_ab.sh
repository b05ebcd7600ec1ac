#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 4" "2 1" "2 2" "2 3"; do
  set -- $cfg
  B200AG_KERNEL_CACHE=0 B200AG_HEAVY_VEC=$1 B200AG_FUSED_MINBLOCKS=$2 timeout 300 python bench.py --steps 10 --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('vec $1 minblocks $2', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'parity', d['parity_max_rel_err_vs_reference'])
"
done
