#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/jacobi_sweep.log
python scripts/jacobi_probe.py >> gpurun_out/jacobi_sweep.log 2>&1
for t in 256,1 256,2 128,1 128,2 64,1 64,2 32,2; do for gm in 8 16 32; do
  B200AG_ROW_TUNE=$t B200AG_GRID_MULT=$gm timeout 100 python scripts/jacobi_probe.py 2>&1 | tail -1 >> gpurun_out/jacobi_sweep.log
done; done
cat gpurun_out/jacobi_sweep.log
