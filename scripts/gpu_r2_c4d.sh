#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fluid or golden or roll or stencil or row_mode or views" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
{
B200AG_KEEP_WIDE=0 python scripts/replay_config.py c4 20 2
python scripts/replay_config.py c4 20 2
python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c3 64 3
python scripts/replay_config.py c1 0 200
python scripts/fused_profile.py c4 2
} > gpurun_out/c4d.log 2>&1
grep -v Warn gpurun_out/c4d.log | cut -c1-200
