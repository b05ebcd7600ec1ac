#!/bin/bash
mkdir -p gpurun_out
{
for E in 3.0e7 1e9 1e12; do
  echo "== EXPAND_WORK=$E"
  B200AG_EXPAND_WORK=$E python scripts/replay_config.py c5 65536 5
  B200AG_EXPAND_WORK=$E python scripts/c5_fused_profile.py
done
} > gpurun_out/c5_expand.log 2>&1
grep -v Warn gpurun_out/c5_expand.log | cut -c1-260
