#!/bin/bash
# GEMM edge-fragment skipping: parity subset, replay timings, launch lists of C3 / C5, per-kernel ncu sections of C5
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/cfg.log 2>&1
grep -v Warn gpurun_out/cfg.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c3_launches.csv python scripts/profile_config.py c3 4 > gpurun_out/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c5_launches.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c5_launches.csv 24
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section ComputeWorkloadAnalysis --clock-control none -c 400 --csv --page raw --log-file gpurun_out/r02_c5_sections.csv python scripts/profile_config.py c5 65536 > gpurun_out/ncu_c5_sections.log 2>&1
echo "sections rc=$?"
