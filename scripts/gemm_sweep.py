"""Calibrate the K-split / tile choice of the grouped DMMA GEMM (csrc/gemm.cu plan_split): for the product shapes of the
BASELINE configs, time every forced (tile, split) inside a replayed CUDA graph and print the table.

    python scripts/gemm_sweep.py            # on a B200
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as onp
import autograd_b200 as b200
from autograd_b200.lazy import Mat, F64, F32

be = b200.backend.get()
rs = onp.random.RandomState(0)


def operand(rows, cols, dt, transposed=False):
    """(Mat, (row stride, col stride)) of a rows x cols operand; transposed = stored as cols x rows (a swapaxes view)."""
    shape = (cols, rows) if transposed else (rows, cols)
    a = b200.asarray(rs.rand(*shape).astype(dt))
    m = a._mat()
    return a, (m.ptr, m.dtype, (1, shape[1]) if transposed else (shape[1], 1))


def problem(M, N, K, a_dt=onp.float64, b_dt=onp.float64, ta=False, tb=False, c_dt=onp.float64):
    keep_a, (ap, adt, astr) = operand(M, K, a_dt, ta)
    keep_b, (bp, bdt, bstr) = operand(K, N, b_dt, tb)
    c = Mat.empty((M, N), onp.dtype(c_dt))
    return (keep_a, keep_b, c), (c.ptr, c.dtype, (N, 1), ap, adt, astr, bp, bdt, bstr, M, N, K)


def time_group(probs, tile, splits, reps=20):
    be.gemm_tune(tile, splits)
    be.gemm_grouped(probs)          # warm-up, attribute setup
    be.synchronize()
    be.graph_begin()
    for _ in range(reps):
        be.gemm_grouped(probs)
    g = be.graph_end()
    be.graph_launch(g)
    be.synchronize()
    e0, e1 = be.event_create(), be.event_create()
    be.event_record(e0)
    for _ in range(3):
        be.graph_launch(g)
    be.event_record(e1)
    be.synchronize()
    ms = be.event_elapsed_ms(e0, e1) / (3 * reps)
    be.graph_destroy(g)
    be.gemm_tune(0, 0)
    return ms * 1e3


f32 = onp.float32
GROUPS = {
    "c1 fwd1 256x200x784": [dict(M=256, N=200, K=784)],
    "c1 fwd2 256x100x200": [dict(M=256, N=100, K=200)],
    "c1 dW1 784x200x256 (A^T)": [dict(M=784, N=200, K=256, ta=True)],
    "c1 dW2+dX2": [dict(M=200, N=100, K=256, ta=True), dict(M=256, N=200, K=100, tb=True)],
    "c3 gates fwd 4x(1024x512x641)": [dict(M=1024, N=512, K=641, b_dt=f32)] * 4,
    "c3 dA 4x(1024x641x512)+(1024x513x128)": [dict(M=1024, N=641, K=512, b_dt=f32, tb=True)] * 4 + [dict(M=1024, N=513, K=128, b_dt=f32, tb=True)],
    "c3 dW 4x(641x512x1024)+(513x128x1024)": [dict(M=641, N=512, K=1024, ta=True)] * 4 + [dict(M=513, N=128, K=1024, ta=True)],
    "c3 predict fwd 1024x128x513": [dict(M=1024, N=128, K=513, b_dt=f32)],
    "c5 fc1 fwd 65536x120x256": [dict(M=65536, N=120, K=256)],
    "c5 fc1 dW 256x120x65536 (A^T)": [dict(M=256, N=120, K=65536, ta=True)],
    "c5 fc2 dW 120x84x65536 (A^T)": [dict(M=120, N=84, K=65536, ta=True)],
    "c5 fc1 dX 65536x256x120 (B^T)": [dict(M=65536, N=256, K=120, tb=True)],
    "single 1024x512x641": [dict(M=1024, N=512, K=641)],
    "single 4096^3": [dict(M=4096, N=4096, K=4096)],
}
only = sys.argv[1:] or None
for name, specs in GROUPS.items():
    if only and not any(o in name for o in only):
        continue
    keeps, probs = [], []
    for sp in specs:
        k, p = problem(**sp)
        keeps.append(k)
        probs.append(p)
    flops = sum(2.0 * sp["M"] * sp["N"] * sp["K"] for sp in specs)
    auto = time_group(probs, 0, 0)
    row = [f"auto {auto:7.1f}us {flops / auto / 1e6:5.1f}TF"]
    best = (auto, "auto")
    for tile in (1, 2):
        if tile == 2 and all(sp["M"] * sp["N"] < 128 * 128 * 8 for sp in specs):
            continue
        for S in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64):
            if min(sp["K"] for sp in specs) // S < 16 or (tile == 2 and S > 8) or (S > 8 and sum(sp["M"] * sp["N"] for sp in specs) > 64 * 64 * 600):
                continue
            t = time_group(probs, tile, S)
            row.append(f"t{tile}s{S}:{t:.1f}")
            if t < best[0]:
                best = (t, f"t{tile}s{S}")
    print(f"{name:42s} best {best[1]:6s} {best[0]:7.1f}us {flops / best[0] / 1e6:5.1f}TF | " + " ".join(row), flush=True)
    del keeps
