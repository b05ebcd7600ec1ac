#!/bin/bash
# multi-GPU validation: peer / NVLS allreduce test (pytest -m gpu -k peer), then the bench at N GPUs like the driver runs it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -q -x > gpurun_out/pytest_peer_$N.log 2>&1; echo "peer pytest rc=$?"; tail -15 gpurun_out/pytest_peer_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/_peer_worker.py > gpurun_out/peer_$N.log 2>&1; echo "peer worker x$N rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|Warn" gpurun_out/peer_$N.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench x$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", json.dumps(d["e2e"])[:700])
    for k in ("c1_dp", "c5_dp"):
        print(k, json.dumps(d["other_configs"][k])[:1200])
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/bench_n$N.err
