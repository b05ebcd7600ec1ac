#!/bin/bash
# usage: gpu_bench.sh <ngpus> [bench args]
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
fi
tail -5 gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.log
