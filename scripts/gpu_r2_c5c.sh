#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or convnet or conv2d or max_pool or short_axis or views" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
{
B200AG_UNROLL_WINDOWS=0 python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c5 65536 5
python scripts/c5_fused_profile.py
python scripts/replay_config.py c4 20 2
} > gpurun_out/c5c.log 2>&1
grep -v Warn gpurun_out/c5c.log | cut -c1-220
