#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/cfg.log 2>&1
cat gpurun_out/cfg.log | grep -v Warn
python scripts/gemm_sweep.py > gpurun_out/gemm_sweep_r2b.log 2>&1; cut -c1-420 gpurun_out/gemm_sweep_r2b.log
