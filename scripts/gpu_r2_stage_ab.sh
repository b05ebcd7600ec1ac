#!/bin/bash
mkdir -p gpurun_out
{
python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c1 0 200
cp autograd_b200/libb200ag.so /tmp/lib3.so
B200AG_NVCC_EXTRA="-DB200AG_GEMM_STAGES=2" python -m autograd_b200.build --force > gpurun_out/build2.log 2>&1
B200AG_NVCC_EXTRA="-DB200AG_GEMM_STAGES=2" python scripts/replay_config.py c5 65536 5
B200AG_NVCC_EXTRA="-DB200AG_GEMM_STAGES=2" python scripts/replay_config.py c3 256 3
B200AG_NVCC_EXTRA="-DB200AG_GEMM_STAGES=2" python scripts/replay_config.py c1 0 200
} > gpurun_out/stage_ab.log 2>&1
grep -v Warn gpurun_out/stage_ab.log | tail -12
