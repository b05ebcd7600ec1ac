#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or convnet or max_pool or views or concat or short_axis" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
{
python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c5 65536 5
for P in 0 8 12 24 40; do B200AG_ROLL_PUSH=$P python scripts/replay_config.py c4 20 2; done
} > gpurun_out/c4c5.log 2>&1
grep -v Warn gpurun_out/c4c5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c5_launches.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c5_launches.csv 16
