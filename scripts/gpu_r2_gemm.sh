#!/bin/bash
# round 2: grouped / deferred GEMM — GPU suite, then C3/C1 replay timings with and without deferral, tile knobs, launch list of C3
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
for T in 32; do
  B200AG_DEFER_GEMM=0 python scripts/replay_config.py c3 $T 3
  python scripts/replay_config.py c3 $T 3
  B200AG_GEMM_TILE=1 python scripts/replay_config.py c3 $T 3
  B200AG_GEMM_TILE=2 python scripts/replay_config.py c3 $T 3
done
B200AG_DEFER_GEMM=0 python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/gemm_ab.log 2>&1
cat gpurun_out/gemm_ab.log | grep -v Warn
python scripts/probe_extras.py kernels > gpurun_out/kernels.log 2>&1; grep -i "dgemm" gpurun_out/kernels.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c3_launches.csv python scripts/profile_config.py c3 4 > gpurun_out/ncu_c3.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c3_launches.csv 16
