#!/bin/bash
# whole-wave grids from the occupancy calculator: all five configs
mkdir -p gpurun_out
{
python scripts/replay_config.py c2 268435456 10
B200AG_GRID_MULT=32 python scripts/replay_config.py c2 268435456 10
B200AG_GRID_MULT=8 python scripts/replay_config.py c2 268435456 10
python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c1 0 200
} > gpurun_out/waves.log 2>&1
grep -v Warn gpurun_out/waves.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x -k "unary or binary or golden or fused or roll or views" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -2 gpurun_out/pytest_gpu_sub.log
