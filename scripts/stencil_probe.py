"""Roll-stencil kernels in isolation (for ncu): Jacobi sweep of fluidsim at 2048^2 and a 4-point stencil at 8192^2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
b200.install()
import autograd.numpy as np
from benchmarks.extras import time_device
be = b200.backend.get()
rs = onp.random.RandomState(0)
for n in (2048, 8192):
    p = b200.asarray(rs.rand(n, n)); div = b200.asarray(rs.rand(n, n))
    f = lambda: (div + np.roll(p, 1, axis=0) + np.roll(p, -1, axis=0) + np.roll(p, 1, axis=1) + np.roll(p, -1, axis=1)) / 4.0
    r = time_device(be, f, warmup=3, reps=20)
    print(f"jacobi sweep {n}^2: {r['ms']*1e3:.1f} us, {3*8*n*n/(r['ms']*1e-3)/1e9:.0f} GB/s algorithmic (2 reads + 1 write)", flush=True)
    g = lambda: p * 2.0 + div
    r = time_device(be, g, warmup=3, reps=20)
    print(f"contiguous 2 reads + 1 write {n}^2: {r['ms']*1e3:.1f} us, {3*8*n*n/(r['ms']*1e-3)/1e9:.0f} GB/s", flush=True)
