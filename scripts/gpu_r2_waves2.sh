#!/bin/bash
mkdir -p gpurun_out
{
for GM in 3 6 64 128; do
B200AG_GRID_MULT=$GM python scripts/replay_config.py c2 268435456 10
done
for GM in 32 64; do
B200AG_GRID_MULT=$GM python scripts/replay_config.py c4 200 2
B200AG_GRID_MULT=$GM python scripts/replay_config.py c5 65536 5
done
} > gpurun_out/waves2.log 2>&1
grep -v Warn gpurun_out/waves2.log | cut -c1-200
