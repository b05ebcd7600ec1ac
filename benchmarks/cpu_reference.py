"""The reference's own CPU path for the headline workload (C2), timed on the box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE: this module runs the UNMODIFIED reference (autograd 1.9.1 from baseline/_ref) on
NumPy; nothing here is on the product path.  NumPy's elementwise loops are single-threaded, so one process uses one
core; the workload is elementwise (every output depends on its own input only), so — exactly like the multi-GPU
sharding of the device arm — the host can use all its cores by giving every worker process its own contiguous chunk.
Workers are spawned (not forked: the parent may hold a CUDA context) and each imports autograd once.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = 4
_f = None


def _init():
    global _f
    for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from examples import workloads as w

    _f = w.tanh_nth_derivative(ORDER)


def _chunk(args):
    import numpy as onp

    lo, hi, n = args
    x = onp.linspace(lo, hi, n, endpoint=False)
    t0 = time.perf_counter()
    y = _f(x)
    dt = time.perf_counter() - t0
    return dt, float(y[0]), float(y[-1])


class HostPool:
    """`procs` worker processes, each evaluating egrad^4(tanh) (examples/tanh.py:22-33) on its own chunk."""

    def __init__(self, procs: int | None = None):
        self.procs = max(1, procs or min(os.cpu_count() or 1, 64))
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init)

    def step(self, sample_points: int) -> float:
        """Seconds (wall) to evaluate `sample_points` points of the linspace over [-7, 7] with all workers."""
        per = sample_points // self.procs
        jobs = [(-7.0 + 14.0 * i / self.procs, -7.0 + 14.0 * (i + 1) / self.procs, per) for i in range(self.procs)]
        t0 = time.perf_counter()
        self.pool.map(_chunk, jobs, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
