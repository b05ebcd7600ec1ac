#!/bin/bash
# sweep tile / split-K settings of the DMMA GEMM on the LSTM shapes
mkdir -p gpurun_out; : > gpurun_out/gemm_sweep.log
for tile in 1 3 2; do for sp in 1 2 3 4 6; do
  echo "== tile=$tile splits=$sp" >> gpurun_out/gemm_sweep.log
  B200AG_GEMM_TILE=$tile B200AG_GEMM_SPLITS=$sp timeout 120 python scripts/gemm_raw_probe.py >> gpurun_out/gemm_sweep.log 2>&1
done; done
echo "== auto" >> gpurun_out/gemm_sweep.log
timeout 120 python scripts/gemm_raw_probe.py >> gpurun_out/gemm_sweep.log 2>&1
cat gpurun_out/gemm_sweep.log
