"""Per generated-kernel timings of one eager gradient of a BASELINE config (CUDA events around every fused launch),
aggregated per program: calls, total time, elements, operands, achieved bandwidth on the algorithmic bytes.

    python scripts/fused_profile.py c4 [size]
"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "baseline", "_ref"), ROOT]
import numpy as onp
import autograd_b200 as b200
b200.install()
from autograd import grad, value_and_grad
from benchmarks.extras import tree_to_dev, materialize
from examples import workloads as w
be = b200.backend.get()
cfg = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
if cfg == "c1":
    params, X, T = w.mlp_data(); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(T), 1.0); f = grad(w.mlp_objective)
elif cfg == "c3":
    params, X, Y = w.lstm_data(seq_len=size or 4); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(X)); f = grad(w.lstm_objective)
elif cfg == "c4":
    p, s0, tg = w.fluid_data(2048, 2048); a = (b200.asarray(p), b200.asarray(s0), b200.asarray(tg), 2048, 2048, size or 2); f = value_and_grad(w.fluid_objective)
else:
    n, _p, loss = w.convnet_make(); W, X, T = w.convnet_data(size or 65536, n); a = (b200.asarray(W), b200.asarray(X), b200.asarray(T)); f = grad(loss)
materialize(f(*a)); be.synchronize()
be.profile = []
l0 = be.launch_count()
materialize(f(*a)); be.synchronize()
launches = be.launch_count() - l0
agg = collections.OrderedDict()
tot = 0.0
for prog, npts, e0, e1 in be.profile:
    ms = be.event_elapsed_ms(e0, e1); tot += ms
    item = {"d": 8, "f": 4, "l": 8, "q": 8, "i": 4, "?": 1}
    # algorithmic bytes: every non-scalar operand read once (general-layout operands may be smaller than n), outputs written once
    nbytes = sum(item[ch] for (ch, cls, *_r) in prog.leaves if cls != "S") * npts + sum(item[ch] for _, ch in prog.outputs) * npts
    rec = agg.setdefault(prog.key, [0, 0.0, npts, len(prog.instrs), len(prog.leaves), len(prog.outputs), prog.mode, prog.vec, prog.unroll, len(prog.tables), nbytes])
    rec[0] += 1; rec[1] += ms
print(f"{cfg}: {launches} launches, {len(be.profile)} generated-kernel launches, {tot:.3f} ms in generated kernels")
for key, (calls, ms, npts, ops, nleaves, nouts, mode, vec, unroll, tables, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {ms*1e3:9.1f} us  x{calls:<3d} {ms/calls*1e3:8.1f} us each  n={npts:<10d} ops={ops:<3d} leaves={nleaves:<2d} outs={nouts} tables={tables} {mode} vec={vec} unroll={unroll}  <= {nbytes/ (ms/calls*1e-3) / 1e9:7.0f} GB/s upper-bound bytes")
