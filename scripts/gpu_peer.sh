#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/_peer_worker.py > gpurun_out/peer_$N.log 2>&1; echo "rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/peer_$N.log | tail -25
