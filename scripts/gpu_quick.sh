#!/bin/bash
# usage: gpu_quick.sh "<pytest -k expr>" <script and args...>
K=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
"$@" > gpurun_out/quick.log 2>&1; tail -40 gpurun_out/quick.log
