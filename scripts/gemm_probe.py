"""GEMM shapes of the LSTM / MLP configs, a few launches each (for ncu) + cuBLAS DGEMM yardstick via torch."""
import os, sys, time
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
shapes = [(1024, 512, 641, "f64xf32"), (1024, 641, 512, "f64xf64T"), (641, 512, 1024, "f64Txf64"), (4096, 4096, 4096, "f64xf64"), (256, 200, 784, "f64xf64"), (65536, 120, 256, "f64xf64")]
if "--c3" in sys.argv:
    shapes = shapes[:3]
for M, N, K, kind in shapes:
    A = b200.asarray(rs.rand(M, K)); B = b200.asarray(rs.rand(K, N))
    if kind == "f64xf32": B = b200.asarray(rs.rand(K, N).astype(onp.float32))
    if kind == "f64xf64T": B = b200.asarray(rs.rand(N, K)).T
    if kind == "f64Txf64": A = b200.asarray(rs.rand(K, M)).T
    r = time_device(be, lambda: np.dot(A, B), warmup=2, reps=10)
    print(f"{kind} {M}x{N}x{K}: {r['ms']*1e3:.1f} us  {2.0*M*N*K/(r['ms']*1e-3)/1e12:.2f} TFLOP/s", flush=True)
if "--cublas" in sys.argv:
    import torch
    for M, N, K in [(8192, 8192, 8192), (4096, 4096, 4096), (1024, 512, 641)]:
        a = torch.rand(M, K, dtype=torch.float64, device="cuda"); b = torch.rand(K, N, dtype=torch.float64, device="cuda")
        for _ in range(3): a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): a @ b
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"cuBLAS dgemm yardstick {M}x{N}x{K}: {ms*1e3:.1f} us {2.0*M*N*K/(ms*1e-3)/1e12:.2f} TFLOP/s", flush=True)
