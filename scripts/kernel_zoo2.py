"""Second kernel zoo (ncu --set full): the kernels added after profiles/r01_kernel_zoo_ncu.csv was taken — direct
convolutions, scans, sort, single-launch concatenation, the fused max-pool forward / VJP kernels, the float32
tcgen05 GEMM with its split pre-pass.  One launch of each after a warm-up, labelled in launch order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
b200.install()
import autograd.numpy as np
from autograd import grad
be = b200.backend.get()
rs = onp.random.RandomState(0)
N = 16384
img = b200.asarray(rs.rand(N, 1, 28, 28)); flt = b200.asarray(rs.rand(1, 6, 5, 5))
gout = b200.asarray(rs.rand(N, 6, 24, 24)); bias = b200.asarray(rs.randn(1, 6, 1, 1))
big = b200.asarray(rs.rand(64, 1 << 20)); keys = b200.asarray(rs.randn(4096, 4096))
pieces = [b200.asarray(rs.rand(k)) for k in (1 << 24, 1 << 22, 1 << 23, 1 << 20, 12345, 1 << 24)]
A32 = b200.asarray(rs.rand(4096, 4096).astype(onp.float32)); B32 = b200.asarray(rs.rand(4096, 4096).astype(onp.float32))
act = b200.asarray(rs.randn(N, 6, 24, 24))


def pool_fwd():
    y6 = np.reshape(act + bias, (N, 6, 12, 2, 12, 2))
    return np.tanh(np.max(np.max(y6, axis=3), axis=4))


pool_loss = lambda a: np.sum(np.tanh(np.max(np.max(np.reshape(a + bias, (N, 6, 12, 2, 12, 2)), axis=3), axis=4)) ** 2)

zoo = [
    ("conv2d direct fwd 16384x1x28x28 * 1x6x5x5", lambda: b200.array.conv2d(img, flt, "valid", True)),
    ("conv2d direct weight-grad", lambda: b200.array.conv2d_filter_grad(gout, img, (1, 6, 5, 5), "valid")),
    ("cumsum last axis [64, 2^20] (segmented scan: totals, scan, apply)", lambda: np.cumsum(big, axis=1)),
    ("cumsum axis 0 [64, 2^20] (column walk)", lambda: np.cumsum(big, axis=0)),
    ("sort rows [4096, 4096] (bitonic)", lambda: np.sort(keys, axis=1)),
    ("concatenate 6 contiguous pieces (one launch)", lambda: np.concatenate(pieces)),
    ("max-pool forward: (x+b) 6-D max max tanh (materialise + fused)", pool_fwd),
    ("max-pool block gradient (fwd + grad_chooser VJPs fused)", lambda: grad(pool_loss)(act)),
    ("sgemm 4096^3 on tcgen05: split A, split B, 3xTF32 CTA-pair GEMM", lambda: np.dot(A32, B32)),
]
for name, fn in zoo:
    fn().materialize(); be.synchronize()
l0 = be.launch_count()
for name, fn in zoo:
    a = be.launch_count()
    fn().materialize(); be.synchronize()
    print(f"launches {a - l0}..{be.launch_count() - l0 - 1}: {name}", flush=True)
