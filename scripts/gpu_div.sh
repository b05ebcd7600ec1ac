#!/bin/bash
# shared-reciprocal division: bit-exactness test + tanh cases, then C2 timing with launch-bound knobs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "shared_denominator or tanh or c2_full or binary_f64 or fused_chain" > gpurun_out/div_tests.log 2>&1; echo "exit code $?"
tail -5 gpurun_out/div_tests.log
for cfg in "1 4" "1 3" "1 5"; do
  set -- $cfg
  B200AG_KERNEL_CACHE=0 B200AG_HEAVY_VEC=$1 B200AG_FUSED_MINBLOCKS=$2 timeout 300 python bench.py --steps 10 --no-extras 2>gpurun_out/div_bench_err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('vec $1 minblocks $2', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'parity', d['parity_max_rel_err_vs_reference'], 'e2e', d['e2e']['ms_per_step'])
"
done 2>&1 | tee gpurun_out/div_bench.log
