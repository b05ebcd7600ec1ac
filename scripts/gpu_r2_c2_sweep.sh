#!/bin/bash
# C2 kernel: resident CTAs per SM x elements per thread x grid size
mkdir -p gpurun_out
{
for MB in 1 2 3; do for HV in 1 2 4; do
  B200AG_FUSED_MINBLOCKS=$MB B200AG_HEAVY_VEC=$HV python scripts/replay_config.py c2 268435456 10
done; done
for GM in 4 8 32; do
  B200AG_GRID_MULT=$GM python scripts/replay_config.py c2 268435456 10
done
} > gpurun_out/c2_sweep.log 2>&1
grep -v Warn gpurun_out/c2_sweep.log | cut -c1-200
