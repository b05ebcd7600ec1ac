#!/bin/bash
# usage: gpu_ncu_list.sh <tag> <profile_config args...>
TAG=$1; shift
mkdir -p gpurun_out
python scripts/profile_config.py "$@" > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_config.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log; wc -l gpurun_out/launches_$TAG.csv
