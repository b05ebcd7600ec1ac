"""Run / time a named BASELINE workload through REFERENCE autograd on NumPy (TEST INFRASTRUCTURE).

Imported only by tests, smoke() and bench.py's cpu_baseline / --impl reference legs.
"""

from __future__ import annotations

import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def _host_ctx():
    """If the device backend is installed in this process, make autograd.numpy creation functions host again."""
    mod = sys.modules.get("autograd_b200")
    if mod is not None and mod.is_installed():
        return mod.host_arrays()
    return contextlib.nullcontext()


def blas_info() -> str:
    import numpy as onp

    try:
        cfg = onp.show_config(mode="dicts")
        b = cfg["Build Dependencies"]["blas"]
        return f"{b.get('name')} {b.get('version')}"
    except Exception:
        return "unknown"


def time_call(fn, *args, warmup=1, repeats=2):
    """Best and median wall time of fn(*args) on the reference path."""
    with _host_ctx():
        for _ in range(warmup):
            fn(*args)
        ts = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = fn(*args)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return {"best_s": ts[0], "median_s": ts[len(ts) // 2], "out": out}


def evaluate(fn, *args):
    import numpy as onp

    with _host_ctx(), onp.errstate(all="ignore"):
        return fn(*args)
