#!/bin/bash
# one-thread-per-window kernels (innermost window dim through 16-byte vectors): parity + C5 timings with / without
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
for U in 0 1; do
  echo "== UNROLL_INNER=$U"
  B200AG_UNROLL_INNER=$U python scripts/replay_config.py c5 65536 5
  B200AG_UNROLL_INNER=$U python scripts/c5_fused_profile.py
done
python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c1 0 200
} > gpurun_out/window.log 2>&1
grep -v Warn gpurun_out/window.log | cut -c1-330
