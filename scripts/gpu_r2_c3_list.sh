#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_c3_launches.csv python scripts/profile_config.py c3 4 > gpurun_out/ncu_c3.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_c3_launches.csv 40
