#!/bin/bash
# resident CTAs per SM (register cap) of ALL generated kernels, C3 and C5
mkdir -p gpurun_out
{
for MB in 0 2 3 4; do B200AG_FUSED_MINBLOCKS=$MB python scripts/replay_config.py c3 256 3; done
for MB in 0 2 3 4; do B200AG_FUSED_MINBLOCKS=$MB python scripts/replay_config.py c5 65536 5; done
} > gpurun_out/mb.log 2>&1
grep -v Warn gpurun_out/mb.log | cut -c1-200
