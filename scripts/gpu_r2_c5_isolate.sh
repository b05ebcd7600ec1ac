#!/bin/bash
mkdir -p gpurun_out
{
python scripts/replay_config.py c5 65536 5
B200AG_DEFER_GEMM=0 python scripts/replay_config.py c5 65536 5
B200AG_EXACT=1 python scripts/replay_config.py c5 65536 5
B200AG_SHORT_REDUCE=0 python scripts/replay_config.py c5 65536 5
python scripts/replay_config.py c3 256 3
} > gpurun_out/c5_isolate.log 2>&1
grep -v Warn gpurun_out/c5_isolate.log | tail -8
python - <<'PY'
import os, sys
sys.path[:0] = ["baseline/_ref", "."]
import autograd_b200 as b200
b200.install()
from autograd import grad
from examples import workloads as w
be = b200.backend.get()
n, _p, loss = w.convnet_make(); W, X, T = w.convnet_data(65536, n)
a = (b200.asarray(W), b200.asarray(X), b200.asarray(T)); f = grad(loss)
f(*a).materialize(); be.synchronize()
be.profile = []
f(*a).materialize(); be.synchronize()
tot = 0
for prog, npts, e0, e1 in be.profile:
    ms = be.event_elapsed_ms(e0, e1); tot += ms
    print(f"fused n={npts} ops={len(prog.instrs)} leaves={len(prog.leaves)} mode={prog.mode} vec={prog.vec} relax={len(prog.relax)} {ms*1e3:.1f} us")
print("fused total ms", tot)
PY
