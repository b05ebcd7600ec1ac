"""Run the float32 tensor-core GEMM a few times at one shape (target for ncu)."""
import sys

import numpy as onp

sys.path.insert(0, ".")
import autograd_b200 as b200  # noqa: E402
from autograd_b200 import backend  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 8192, 4096)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
be = backend.get()
rs = onp.random.RandomState(0)
a = b200.asarray(rs.randn(M, K).astype(onp.float32))
b = b200.asarray(rs.randn(K, N).astype(onp.float32))
c = a @ b
be.synchronize()
e0, e1 = be.event_create(), be.event_create()
be.event_record(e0)
for _ in range(reps):
    c = a @ b
    c._mat()
be.event_record(e1)
be.synchronize()
ms = be.event_elapsed_ms(e0, e1) / reps
print(f"{M}x{N}x{K}: {ms:.3f} ms/product  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent, dtype {c.dtype}")
