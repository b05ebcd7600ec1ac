#!/bin/bash
# zero-stride / 32-bit offset specialisation of flat general-layout kernels: parity + C5 / C4 timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
for E in 3.0e7 1e12; do
  echo "== EXPAND_WORK=$E"
  B200AG_EXPAND_WORK=$E python scripts/replay_config.py c5 65536 5
  B200AG_EXPAND_WORK=$E python scripts/c5_fused_profile.py
done
python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c3 256 3
} > gpurun_out/zstr.log 2>&1
grep -v Warn gpurun_out/zstr.log | cut -c1-330
