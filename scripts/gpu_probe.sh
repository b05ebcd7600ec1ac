#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/probe_extras.py "$@" > gpurun_out/probe.log 2>&1; echo "rc=$?" >> gpurun_out/probe.log
tail -c 6000 gpurun_out/probe.log
