#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or convnet or conv2d or max_pool or convolve or cabi" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
{
B200AG_CONV_DIRECT_CMAX=4 python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c5 65536 5
B200AG_EXPAND_WORK=1e12 python scripts/replay_config.py c5 65536 5
} > gpurun_out/c5b.log 2>&1
grep -v Warn gpurun_out/c5b.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c5_launches.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c5_launches.csv 14
