#!/bin/bash
# C4: resident CTAs per SM of the generated kernels with lookups (register cap) - the advection gathers are latency-bound
mkdir -p gpurun_out
{
for MB in 1 4 5 6 8; do B200AG_TABLE_MINBLOCKS=$MB python scripts/replay_config.py c4 100 2; done
} > gpurun_out/c4_occ.log 2>&1
grep -v Warn gpurun_out/c4_occ.log | cut -c1-200
