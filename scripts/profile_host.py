"""cProfile of the eager (re-tracing) path on the GPU box: where does host time go?"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import autograd_b200 as b200
b200.install()
from autograd import grad
from benchmarks.extras import tree_to_dev, materialize
from examples import workloads as w
be = b200.backend.get()
params, X, T = w.mlp_data()
a = (tree_to_dev(params), b200.asarray(X), b200.asarray(T), 1.0); f = grad(w.mlp_objective)
for _ in range(5): materialize(f(*a))
be.synchronize()
t = time.perf_counter()
for _ in range(100): materialize(f(*a))
be.synchronize()
print("C1 eager ms/grad:", (time.perf_counter() - t) * 10)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): materialize(f(*a))
be.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(35)
