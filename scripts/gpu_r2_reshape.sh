#!/bin/bash
# pending reshape nodes: parity + C4 timing with / without + the other configs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
B200AG_LAZY_RESHAPE=0 python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c4 200 2
python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c1 0 200
} > gpurun_out/reshape.log 2>&1
grep -v Warn gpurun_out/reshape.log | cut -c1-330
