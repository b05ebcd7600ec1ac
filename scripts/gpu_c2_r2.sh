#!/bin/bash
# C2 after the shared-reciprocal change: knob sweep, bench line with graph replay, launch list + full ncu capture
mkdir -p gpurun_out
for cfg in "1 2" "2 3" "2 2"; do
  set -- $cfg
  B200AG_KERNEL_CACHE=0 B200AG_HEAVY_VEC=$1 B200AG_FUSED_MINBLOCKS=$2 timeout 300 python bench.py --steps 10 --no-extras 2>gpurun_out/c2r2_err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('vec $1 minblocks $2', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), d['execution'])
"
done 2>&1 | tee gpurun_out/c2r2_sweep.log
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/c2r2_plain.log 2>gpurun_out/c2r2_plain.err && tail -c 1500 gpurun_out/c2r2_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2r2_launches.csv $CMD > gpurun_out/c2r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 1 -o gpurun_out/prof_c2r2 $CMD > gpurun_out/c2r2_ncu_full.log 2>&1
tail -2 gpurun_out/c2r2_ncu_full.log
