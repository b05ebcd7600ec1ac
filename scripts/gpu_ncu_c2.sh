#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 2 -o gpurun_out/prof_c2 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail; tail -2 gpurun_out/ncu_full.log; head -c 600 gpurun_out/plain_c2.log
