"""Per fused-kernel timings of one eager C5 gradient (CUDA events around every generated-kernel launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "baseline", "_ref"), ROOT]
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
    if ms > 0.05:
        print(f"fused n={npts} ops={len(prog.instrs)} leaves={[(l[0], l[1], l[3]) for l in prog.leaves]} mode={prog.mode} vec={prog.vec} nd={prog.ndim} {ms*1e3:.1f} us")
print("fused total ms", tot, "launches", len(be.profile))
