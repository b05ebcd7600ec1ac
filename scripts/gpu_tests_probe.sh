#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1500 python scripts/probe_extras.py "$@" > gpurun_out/probe.log 2>&1; echo "rc=$?" >> gpurun_out/probe.log
grep -v '^  "' gpurun_out/probe.log | tail -c 7000
