#!/bin/bash
# first GPU contact: smoke, GPU parity suite, bench at reduced and full size
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --n $((1<<24)) --steps 5 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/bench_small.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_small.log; tail -3 gpurun_out/bench_full.log
