"""Per-sweep time of the fluidsim Jacobi projection (fluidsim.py:21-30) inside a replayed CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
b200.install()
import autograd.numpy as np
be = b200.backend.get()
rs = onp.random.RandomState(0)
n = 2048
p0 = b200.asarray(rs.rand(n, n)); div = b200.asarray(rs.rand(n, n))


def sweeps(p, div):
    for _ in range(20):
        p = (div + np.roll(p, 1, axis=0) + np.roll(p, -1, axis=0) + np.roll(p, 1, axis=1) + np.roll(p, -1, axis=1)) / 4.0
    return p


fast = b200.capture(sweeps)
for _ in range(3):
    fast(p0, div)
be.synchronize()
e0, e1 = be.event_create(), be.event_create()
be.event_record(e0)
for _ in range(20):
    fast(p0, div)
be.event_record(e1)
be.synchronize()
us = be.event_elapsed_ms(e0, e1) * 1e3 / 400
print(f"ROW_TUNE={os.environ.get('B200AG_ROW_TUNE','-')} GRID_MULT={os.environ.get('B200AG_GRID_MULT','16')}: "
      f"{us:.2f} us per sweep, {3 * 8 * n * n / us / 1e3:.0f} GB/s algorithmic", flush=True)
