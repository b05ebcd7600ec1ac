#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
B200AG_UNROLL_WINDOWS=0 python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c5 65536 5
B200AG_EXPAND_WORK=1e12 python scripts/replay_config.py c5 65536 5
python scripts/c5_fused_profile.py
echo "== EXPAND off"
B200AG_EXPAND_WORK=1e12 python scripts/c5_fused_profile.py
python scripts/replay_config.py c4 20 2
B200AG_UNROLL_WINDOWS=0 python scripts/replay_config.py c4 20 2
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c1 0 200
} > gpurun_out/unroll.log 2>&1
grep -v Warn gpurun_out/unroll.log | cut -c1-250
