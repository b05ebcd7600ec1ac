"""Kernel-level timing of b200ag_gemm on the LSTM shapes: raw C-ABI calls back to back (no Python array layer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as onp
import autograd_b200 as b200
from autograd_b200 import backend
be = backend.get()
rs = onp.random.RandomState(0)
f64, f32 = onp.dtype("float64"), onp.dtype("float32")
cases = []
M, N, K = 1024, 512, 641      # forward gate: cat[1024,641] (f64) @ W[641,512] (f32)
cases.append(("fwd  f64 x f32", M, N, K, (K, 1), f64, (N, 1), f32))
M, N, K = 1024, 641, 512      # dX = G[1024,512] @ W^T: B is a transposed view of W[641,512] (f32)
cases.append(("dX   G x W^T  ", M, N, K, (K, 1), f64, (1, K), f32))
M, N, K = 641, 512, 1024      # dW = cat^T @ G: A is a transposed view of cat[1024,641]
cases.append(("dW   cat^T x G", M, N, K, (1, M), f64, (N, 1), f64))
M, N, K = 1024, 128, 513
cases.append(("readout fwd   ", M, N, K, (K, 1), f64, (N, 1), f32))
for name, M, N, K, a_str, a_dt, b_str, b_dt in cases:
    A = b200.asarray(rs.rand(M * K).astype(a_dt))
    B = b200.asarray(rs.rand(K * N).astype(b_dt))
    C = b200.array.zeros((M, N))
    am, bm, cm = A._mat(), B._mat(), C._mat()
    fn = lambda: be.gemm(cm.ptr, f64, (N, 1, 0), am.ptr, a_dt, (a_str[0], a_str[1], 0), bm.ptr, b_dt, (b_str[0], b_str[1], 0), M, N, K, 1)
    for _ in range(5):
        fn()
    be.synchronize()
    e0, e1 = be.event_create(), be.event_create()
    be.event_record(e0)
    for _ in range(100):
        fn()
    be.event_record(e1)
    be.synchronize()
    us = be.event_elapsed_ms(e0, e1) * 10
    print(f"{name} {M}x{N}x{K}: {us:6.1f} us  {2.0 * M * N * K / us / 1e6:5.2f} TFLOP/s", flush=True)
