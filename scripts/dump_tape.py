"""Print, in launch order, what one eager gradient of a config launches (host emulation; no GPU needed):
python scripts/dump_tape.py c5 [batch].  For every generated kernel: points, recorded ops, leaves with their
distinct bytes, outputs; for every library call its name and sizes.  Used to find avoidable HBM traffic."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")]
import numpy as onp
import autograd_b200 as b200
from fakedev import FakeBackend
fb = FakeBackend(compile_generated=False)
b200.backend.set_backend(fb)
b200.install()
from autograd import grad, value_and_grad
from benchmarks.extras import tree_to_dev, materialize
from examples import workloads as w

log = []
orig_run = fb.run_program
def run_program(program, leaf_info, const_values, out_ptrs, n):
    dims = leaf_info[1]
    leaves = []
    total = 0
    for (ch, cls, has_shift, flag), (ptr, strides, shifts) in zip(program.leaves, leaf_info[0]):
        isz = onp.dtype(ch).itemsize
        if cls == "S": distinct = 1
        elif cls == "C": distinct = n
        else:
            distinct = 1
            for d, s in zip(dims, strides):
                if s != 0: distinct *= d
        total += distinct * isz
        leaves.append(f"{ch}{cls}{'s' if has_shift else ''}{flag}:{distinct*isz/1e6:.1f}MB")
    obytes = sum(onp.dtype(ch).itemsize * n for _j, ch in program.outputs)
    ops = collections.Counter(i[0] for i in program.instrs)
    log.append(f"FUSED n={n} dims={tuple(dims[:program.ndim])} read={total/1e6:.1f}MB write={obytes/1e6:.1f}MB outs={len(program.outputs)} leaves=[{', '.join(leaves)}] ops={dict(ops)}")
    return orig_run(program, leaf_info, const_values, out_ptrs, n)
fb.run_program = run_program
def wrap(name):
    f = getattr(fb, name)
    def g(*a, **k):
        small = [x for x in a if isinstance(x, (int, str, tuple)) and not (isinstance(x, int) and x > 1 << 40)]
        log.append(f"{name.upper()} {small}")
        return f(*a, **k)
    setattr(fb, name, g)
for name in ("copy_strided", "add_strided", "concat", "concat_mixed", "reduce", "argreduce", "logsumexp", "gemm_grouped",
             "gather", "scatter_add", "scatter_assign", "cumulate", "conv2d", "conv2d_wgrad", "memset"):
    wrap(name)

cfg = sys.argv[1]
if cfg == "c5":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    n, _p, loss = w.convnet_make(); W, X, T = w.convnet_data(N, n); a = (b200.asarray(W), b200.asarray(X), b200.asarray(T)); f = grad(loss)
elif cfg == "c1":
    params, X, T = w.mlp_data(); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(T), 1.0); f = grad(w.mlp_objective)
elif cfg == "c4":
    st = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    G_ = int(sys.argv[3]) if len(sys.argv) > 3 else 256; p, s0, tg = w.fluid_data(G_, G_); a = (b200.asarray(p), b200.asarray(s0), b200.asarray(tg), G_, G_, st); f = value_and_grad(w.fluid_objective)
elif cfg == "c3":
    T_ = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    params, X, Y = w.lstm_data(seq_len=T_); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(X)); f = grad(w.lstm_objective)
materialize(f(*a)); fb.synchronize()
log.clear()
materialize(f(*a)); fb.synchronize()
for i, l in enumerate(log):
    print(i, l)
