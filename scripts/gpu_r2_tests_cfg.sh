#!/bin/bash
# GPU suite + replay timings of C1/C3/C5 (and C4 when asked)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
[ -n "$WITH_C4" ] && python scripts/replay_config.py c4 200 2
} > gpurun_out/cfg.log 2>&1
cat gpurun_out/cfg.log | grep -v Warn
