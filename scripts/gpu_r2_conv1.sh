#!/bin/bash
# conv1 direct kernels: rounds of work items per staged group (forward), cotangent staging layout (weight gradient)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "conv2d or convnet or golden" > gpurun_out/pytest_gpu_sub.log 2>&1; tail -2 gpurun_out/pytest_gpu_sub.log
{
for R in 1 2 3 4; do B200AG_CONV_DIRECT_ROUNDS=$R python scripts/replay_config.py c5 65536 5; done
} > gpurun_out/conv1.log 2>&1
grep -v Warn gpurun_out/conv1.log | cut -c1-200
