#!/bin/bash
# lanes-are-images input-gradient kernel: parity (conv / golden / convnet subsets, then everything) + C5 timing A/B + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "conv2d or convnet or golden" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -3 gpurun_out/pytest_gpu_sub.log
{
B200AG_CONV_FULL_LANES=0 python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c5 65536 5
} > gpurun_out/dgrad.log 2>&1
grep -v Warn gpurun_out/dgrad.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c5_launches.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c5_launches.csv 16
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
