"""Worker for tests/test_dist_gloo.py: one rank of a world_size-N gloo job on the host-emulated backend."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as onp

import autograd_b200 as b200
from fakedev import FakeBackend

b200.backend.set_backend(FakeBackend(compile_generated=False))
b200.install()
from autograd import grad

from autograd_b200 import dist
from examples import workloads as w

import cases

dist.init("gloo")
world, rank = dist.world_size(), dist.rank()
assert world == int(os.environ["WORLD_SIZE"])

# full batch (identical on every rank), sharded by rows; prior split evenly so the sum is the full gradient
params, X, T = w.mlp_data(batch=48, layer_sizes=(20, 16, 12, 10))
sl = dist.shard_batch(X.shape[0])
g_local = grad(w.mlp_objective)(cases.to_dev(params), b200.asarray(X[sl]), b200.asarray(T[sl]), 1.0 / world)
g_sum = dist.allreduce_grads(g_local)
with b200.host_arrays():
    want = grad(w.mlp_objective)(params, X, T, 1.0)
for a, r in zip(cases.leaves(cases.to_host(g_sum)), cases.leaves(want)):
    cases.assert_close(a, r, cases.TOL_REDUCE, f"rank {rank} data-parallel MLP gradient")

# convnet: strong-scaling split of one batch
n, _pred, loss = w.convnet_make()
W, Xc, Tc = w.convnet_data(8, n)
sl = dist.shard_batch(8)
gl = grad(loss)(b200.asarray(W), b200.asarray(Xc[sl]), b200.asarray(Tc[sl]), 1.0 / world)
gs = dist.allreduce_grads(gl)
with b200.host_arrays():
    want = grad(loss)(W, Xc, Tc, 1.0)
# the example's logsumexp shifts by the shard's max instead of the global one: rounding-level differences only
cases.assert_close(gs.get(), want, 1e-9, f"rank {rank} data-parallel convnet gradient")
print(f"rank {rank}/{world} ok", flush=True)
