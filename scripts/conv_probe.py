"""A/B of the float64 convolution kernel families (direct CUDA-core vs implicit-GEMM DMMA) on LeNet's shapes."""
import sys

import numpy as onp

sys.path.insert(0, ".")
import autograd_b200 as b200  # noqa: E402
from autograd_b200 import array as A  # noqa: E402
from autograd_b200 import backend  # noqa: E402

be = backend.get()
rs = onp.random.RandomState(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    be.synchronize()
    e0, e1 = be.event_create(), be.event_create()
    be.event_record(e0)
    for _ in range(reps):
        r = fn()
        r._mat()
    be.event_record(e1)
    be.synchronize()
    return be.event_elapsed_ms(e0, e1) / reps, r


x1 = b200.asarray(rs.rand(N, 1, 28, 28))
w1 = b200.asarray(rs.randn(1, 6, 5, 5))
g1 = b200.asarray(rs.randn(N, 6, 24, 24))
p1 = b200.asarray(rs.rand(N, 6, 12, 12))
w2 = b200.asarray(rs.randn(6, 16, 5, 5))
g2 = b200.asarray(rs.randn(N, 16, 8, 8))
cases = [
    ("conv1 forward  [N,1,28,28]*[1,6,5,5]", lambda: A.conv2d(x1, w1, mode="valid", flip=True), 2.0 * N * 6 * 576 * 25),
    ("conv1 weight-grad", lambda: A.conv2d_filter_grad(g1, x1, (1, 6, 5, 5)), 2.0 * N * 6 * 576 * 25),
    ("conv2 forward  [N,6,12,12]*[6,16,5,5]", lambda: A.conv2d(p1, w2, mode="valid", flip=True), 2.0 * N * 16 * 64 * 150),
    ("conv2 input-grad", lambda: A.conv2d_input_grad(g2, w2, (N, 6, 12, 12)), 2.0 * N * 16 * 64 * 150),
    ("conv2 weight-grad", lambda: A.conv2d_filter_grad(g2, p1, (6, 16, 5, 5)), 2.0 * N * 16 * 64 * 150),
]
for name, fn, flops in cases:
    res = {}
    for mode in (0, 1, 0, 1):
        be.conv_mode(mode)
        ms, r = timeit(fn)
        res[mode] = (ms, r.get())
        del r
    be.conv_mode(1)
    err = onp.abs(res[0][1] - res[1][1]).max() / onp.abs(res[0][1]).max()
    print(f"{name:42s} DMMA {res[0][0]:7.3f} ms ({flops / res[0][0] / 1e9:5.1f} TF)   direct {res[1][0]:7.3f} ms "
          f"({flops / res[1][0] / 1e9:5.1f} TF)   max rel diff {err:.1e}", flush=True)
