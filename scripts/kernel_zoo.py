"""One launch (after one warm-up) of every kernel family at a roofline-relevant size, labelled on stdout in launch
order — the command behind profiles/r01_kernel_zoo_*.csv (ncu --set full)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
b200.install()
import autograd.numpy as np
from autograd import elementwise_grad as egrad
from examples import workloads as w
be = b200.backend.get()
rs = onp.random.RandomState(0)
n = 1 << 27
x = b200.asarray(onp.linspace(-3.0, 3.0, n)); y = b200.asarray(onp.linspace(1.0, 2.0, n))
m2 = b200.asarray(rs.rand(8192, 8192)); d2 = b200.asarray(rs.rand(8192, 8192))
ri = b200.asarray(rs.randint(0, 8192, 8192 * 8192)); ci = b200.asarray(rs.randint(0, 8192, 8192 * 8192))
A = b200.asarray(rs.rand(4096, 4096)); B = b200.asarray(rs.rand(4096, 4096))
X = b200.asarray(rs.rand(1024, 641)); W32 = b200.asarray(rs.rand(641, 512).astype(onp.float32))
img = b200.asarray(rs.rand(8192, 1, 28, 28)); flt = b200.asarray(rs.rand(1, 6, 5, 5))
gout = b200.asarray(rs.rand(8192, 6, 24, 24))
lse_in = b200.asarray(rs.rand(1 << 20, 128))

def scatter():
    z = b200.array.zeros((8192, 8192)); onp.add.at(z, (ri, ci), y[: 8192 * 8192]); return z

zoo = [
    ("elementwise vjp_multiply y*g (24 B/pt)", lambda: y * x),
    ("elementwise vjp_tanh g/cosh(x)**2 (24 B/pt)", lambda: y / np.cosh(x) ** 2),
    ("fused egrad(tanh) order 1 (16 B/pt)", lambda: egrad(w.tanh_fn)(x)),
    ("fused egrad^4(tanh) (16 B/pt, 348 ops)", lambda: w.tanh_nth_derivative(4)(x)),
    ("reduce sum all (8 B/pt)", lambda: np.sum(x)),
    ("reduce column sum 8192^2", lambda: np.sum(m2, axis=0)),
    ("reduce row max 8192^2", lambda: np.max(m2, axis=1)),
    ("logsumexp [2^20,128] axis 1", lambda: b200.array.logsumexp(lse_in, axis=1)),
    ("transpose copy 8192^2", lambda: m2.T.copy()),
    ("jacobi roll stencil 8192^2 (24 B/pt)", lambda: (d2 + np.roll(m2, 1, 0) + np.roll(m2, -1, 0) + np.roll(m2, 1, 1) + np.roll(m2, -1, 1)) / 4.0),
    ("fused gather m2[ri,ci]*2 (random)", lambda: m2[ri, ci] * 2.0),
    ("fused scatter-add (random)", scatter),
    ("dgemm 4096^3", lambda: np.dot(A, B)),
    ("dgemm 1024x512x641 f64xf32 (LSTM gate)", lambda: np.dot(X, W32)),
    ("conv2d fwd 8192x1x28x28 * 1x6x5x5", lambda: b200.array.conv2d(img, flt, "valid", True)),
    ("conv2d wgrad", lambda: b200.array.conv2d_filter_grad(gout, img, (1, 6, 5, 5), "valid")),
]
for name, fn in zoo:
    fn().materialize(); be.synchronize()
l0 = be.launch_count()
for name, fn in zoo:
    a = be.launch_count()
    fn().materialize(); be.synchronize()
    print(f"launches {a - l0}..{be.launch_count() - l0 - 1}: {name}", flush=True)
