#!/bin/bash
# round 2: cluster split-K + balanced splits + mixed concat — GPU suite (x-fail-fast), C1/C3/C5 replay, C3 launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 32 3
B200AG_GEMM_TILE=1 python scripts/replay_config.py c3 32 3
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/gemm_ab.log 2>&1
cat gpurun_out/gemm_ab.log | grep -v Warn
python scripts/probe_extras.py kernels > gpurun_out/kernels.log 2>&1; grep -A2 "dgemm" gpurun_out/kernels.log | grep -v "^--" | paste - - -
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c3_launches.csv python scripts/profile_config.py c3 4 > gpurun_out/ncu_c3.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c3_launches.csv 12
grep "gemm_grouped" gpurun_out/r02_c3_launches.csv | awk -F'","' '{print $5, $8, $9, $15}' | sed 's/void b200ag:://' | cut -c1-150 | tail -12
