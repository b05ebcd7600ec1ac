"""GPU probe for the tcgen05 3xTF32 float32 GEMM: accuracy against a float64 product and timing (run under gpurun)."""
import sys
import time

import numpy as onp

sys.path.insert(0, ".")
import autograd_b200 as b200  # noqa: E402
from autograd_b200 import backend, lazy  # noqa: E402

be = backend.get()
rs = onp.random.RandomState(0)


def run_tc(a, b, ta=False, tb=False):
    """a: host [M,K] (or [K,M] if ta), b: host [K,N] (or [N,K] if tb)."""
    da, db = b200.asarray(a), b200.asarray(b)
    M, K = (a.shape[1], a.shape[0]) if ta else a.shape
    N = b.shape[0] if tb else b.shape[1]
    a_str = (1, a.shape[1]) if ta else (a.shape[1], 1)
    b_str = (1, b.shape[1]) if tb else (b.shape[1], 1)
    out = b200.array.zeros((M, N), onp.float32)
    ma, mb, mo = da._mat(), db._mat(), out._mat()
    be.gemm_f32_tc(mo.ptr, N, ma.ptr, a_str, mb.ptr, b_str, M, N, K)
    be.synchronize()
    return out.get()


def check(M, N, K, ta=False, tb=False, scale=1.0):
    a = (rs.randn(*((K, M) if ta else (M, K))) * scale).astype(onp.float32)
    b = rs.randn(*((N, K) if tb else (K, N))).astype(onp.float32)
    got = run_tc(a, b, ta, tb)
    A = (a.T if ta else a).astype(onp.float64)
    B = (b.T if tb else b).astype(onp.float64)
    want = A @ B
    ref32 = (a.T if ta else a) @ (b.T if tb else b)
    nrm = onp.abs(want).max()
    e_tc = onp.abs(got - want).max() / nrm
    e_np = onp.abs(ref32 - want).max() / nrm
    e_vs = onp.abs(got - ref32).max() / nrm
    print(f"M={M} N={N} K={K} ta={ta} tb={tb}: tc-vs-f64 {e_tc:.2e}  numpy-sgemm-vs-f64 {e_np:.2e}  tc-vs-numpy {e_vs:.2e}", flush=True)
    return e_tc


def bench(M, N, K, reps=10):
    a = b200.asarray(rs.randn(M, K).astype(onp.float32))
    b = b200.asarray(rs.randn(K, N).astype(onp.float32))
    out = b200.array.zeros((M, N), onp.float32)
    ma, mb, mo = a._mat(), b._mat(), out._mat()
    for mode, name in ((201, "tcgen05 3xTF32 1-CTA"), (202, "tcgen05 3xTF32 CTA pair"), (200, "tcgen05 3xTF32 auto"),
                       (101, "1-CTA MMA-only (no loads)"), (0, "DMMA f64-accumulate")):
        be.gemm_f32_mode(100)
        be.gemm_f32_mode(200)
        be.gemm_f32_mode(0 if mode == 0 else 1)
        be.gemm_f32_mode(mode)
        if mode == 101:
            be.gemm_f32_mode(201)
        fn = lambda: be.gemm(mo.ptr, onp.dtype("float32"), (N, 1, 0), ma.ptr, onp.dtype("float32"), (K, 1, 0), mb.ptr,
                             onp.dtype("float32"), (N, 1, 0), M, N, K, 1)
        for _ in range(3):
            fn()
        be.synchronize()
        e0, e1 = be.event_create(), be.event_create()
        be.event_record(e0)
        for _ in range(reps):
            fn()
        be.event_record(e1)
        be.synchronize()
        ms = be.event_elapsed_ms(e0, e1) / reps
        print(f"  {name:28s} {M}x{N}x{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)", flush=True)
    be.gemm_f32_mode(100)
    be.gemm_f32_mode(200)
    be.gemm_f32_mode(1)


if __name__ == "__main__":
    t0 = time.time()
    if "--quick" not in sys.argv:
        be.gemm_f32_mode(202)
        print("CTA-pair kernel forced:")
        check(256, 256, 32)
        check(300, 200, 70)
        check(1000, 777, 333, ta=True)
        check(515, 1030, 257, tb=True)
        check(2048, 2048, 2048)
        check(256, 384, 6144, ta=True)
        check(1024, 1024, 8192)
        be.gemm_f32_mode(200)
        print("automatic choice:")
        check(128, 128, 32)
        check(128, 256, 64)
        check(256, 512, 128)
        check(300, 200, 70)
        check(1000, 777, 333, ta=True)
        check(515, 1030, 257, tb=True)
        check(640, 384, 96, ta=True, tb=True)
        check(2048, 2048, 2048)
        check(1024, 1024, 8192)
        check(4096, 4096, 4096, scale=3.0)
    for shape in ((2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 4096), (1024, 512, 16384)):
        bench(*shape)
    print("done in %.1f s" % (time.time() - t0))
