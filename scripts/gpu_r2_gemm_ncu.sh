#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_config.py c3 4 > gpurun_out/plain_c3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -s 13 -c 6 -o gpurun_out/r02_prof_gemm python scripts/profile_config.py c3 4 > gpurun_out/ncufull_gemm.log 2>&1
tail -2 gpurun_out/ncufull_gemm.log
