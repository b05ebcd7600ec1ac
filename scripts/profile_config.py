"""Run ONE eager grad of a config after a warm-up (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
b200.install()
from autograd import grad, value_and_grad
from benchmarks.extras import tree_to_dev, materialize
from examples import workloads as w
be = b200.backend.get()
cfg = sys.argv[1]
if cfg == "c1":
    params, X, T = w.mlp_data(); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(T), 1.0); f = grad(w.mlp_objective)
elif cfg == "c3":
    T_ = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    params, X, Y = w.lstm_data(seq_len=T_); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(X)); f = grad(w.lstm_objective)
elif cfg == "c4":
    st = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p, s0, tg = w.fluid_data(2048, 2048); a = (b200.asarray(p), b200.asarray(s0), b200.asarray(tg), 2048, 2048, st); f = value_and_grad(w.fluid_objective)
elif cfg == "c5":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    n, _p, loss = w.convnet_make(); W, X, T = w.convnet_data(N, n); a = (b200.asarray(W), b200.asarray(X), b200.asarray(T)); f = grad(loss)
materialize(f(*a)); be.synchronize()
l0 = be.launch_count()
materialize(f(*a)); be.synchronize()
print("launches in measured call:", be.launch_count() - l0)
