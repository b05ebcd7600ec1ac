"""np.linalg.solve / inv / det / slogdet / cholesky on device arrays, COMPOSED from the array operations of
:mod:`autograd_b200.array` (fused elementwise kernels, argmax, row gathers): Gauss-Jordan elimination with partial
pivoting whose pivot search and row exchange stay on the device (the pivot row is a device index, the exchange a row
gather), and a column-by-column Cholesky factorisation.  No LAPACK-class kernels: these exist so that the reference's
linalg-based code (autograd/numpy/linalg.py rules, autograd/scipy/stats/multivariate_normal.py, the Gaussian-process
examples) runs on device values; they launch O(n) small kernels per matrix and are meant for the small systems those
programs solve (SURVEY.md section 2 rows 16-17 keep linalg out of the measured hot path).
"""

from __future__ import annotations

import numpy as np

from . import array as A
from .array import DeviceArray


def _as_float(a):
    a = A.asarray(a)
    return a if a.dtype.kind == "f" else a.astype(np.float64)


def _eliminate(a, b=None):
    """Gauss-Jordan with partial pivoting on one square matrix.  Returns (reduced right-hand side or None, pivots [n],
    permutation sign as a 0-d device value).  After the sweep the left block is diagonal, so x = rhs / pivots[:, None]."""
    n = a.shape[0]
    m = a if b is None else A.concatenate([a, b], axis=1)
    rows = A.asarray(np.arange(n, dtype=np.int64))
    sign = A.asarray(np.array(1.0, dtype=a.dtype))
    pivots = []
    for k in range(n):
        if k < n - 1:
            p = A.argmax(abs(m[k:, k])) + k                               # device 0-d index of the pivot row
            order = A.where(rows == k, p, A.where(rows == p, k, rows))    # exchange rows k and p
            m = m[order]
            sign = A.where(p != k, -sign, sign)
        piv = m[k, k]
        pivots.append(piv)
        factors = A.where(rows == k, 0.0, m[:, k] / piv)                  # eliminate column k from every other row
        m = m - factors[:, None] * m[k][None, :]
        m = m.materialize()
    return (None if b is None else m[:, n:]), A.stack(pivots), sign


def _batched(fn, a):
    """Apply a single-matrix routine over the leading (batch) axes of ``a``."""
    lead = a.shape[:-2]
    if not lead:
        return fn(a)
    outs = [fn(a[idx]) for idx in np.ndindex(*lead)]
    if isinstance(outs[0], tuple):
        return tuple(A.reshape(A.stack([o[i] for o in outs]), lead + outs[0][i].shape) for i in range(len(outs[0])))
    return A.reshape(A.stack(outs), lead + outs[0].shape)


def solve(a, b):
    a, b = _as_float(a), _as_float(b)
    if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")
    vector = b.ndim == 1                   # NumPy 2: b is a vector only when it is exactly 1-D, else a stack of matrices
    lead = a.shape[:-2]
    if lead:
        bb = b[..., None] if vector else b
        bb = A.broadcast_to(bb, lead + bb.shape[-2:]) if bb.shape[:-2] != lead else bb
        outs = [solve(a[idx], bb[idx]) for idx in np.ndindex(*lead)]
        res = A.reshape(A.stack(outs), lead + outs[0].shape)
        return res[..., 0] if vector else res
    rhs = b[:, None] if vector else b
    x, pivots, _ = _eliminate(a, rhs)
    x = x / pivots[:, None]
    return x[:, 0] if vector else x


def inv(a):
    a = _as_float(a)
    if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")

    def one(m):
        x, pivots, _ = _eliminate(m, A.asarray(np.eye(m.shape[0], dtype=m.dtype)))
        return x / pivots[:, None]

    return _batched(one, a)


def slogdet(a):
    a = _as_float(a)
    if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")

    def one(m):
        _, pivots, sign = _eliminate(m)
        sgn = sign * A.reduce_(np.sign(pivots), "prod", None, False)
        return sgn, A.reduce_(np.log(abs(pivots)), "sum", None, False)

    sgn, logabs = _batched(one, a)
    try:
        from numpy.linalg._linalg import SlogdetResult
        return SlogdetResult(sgn, logabs)
    except Exception:
        return sgn, logabs


def det(a):
    a = _as_float(a)
    if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")

    def one(m):
        _, pivots, sign = _eliminate(m)
        return sign * A.reduce_(pivots, "prod", None, False)

    return _batched(one, a)


def cholesky(a, upper=False):
    """Lower-triangular L with L L^T = a (symmetric positive definite), built column by column."""
    a = _as_float(a)
    if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")

    def one(m):
        n = m.shape[0]
        cols = []                                            # cols[j] = column j of L (length n, zeros above the diagonal)
        for j in range(n):
            if j:
                lj = A.stack(cols, axis=1)                   # n x j
                resid = m[:, j] - A.dot(lj, lj[j])           # a[:, j] - L[:, :j] L[j, :j]^T
            else:
                resid = m[:, j]
            d = np.sqrt(resid[j])
            col = A.where(A.asarray(np.arange(n)) >= j, resid / d, 0.0)
            cols.append(col.materialize())
        low = A.stack(cols, axis=1)
        return low.T if upper else low

    return _batched(one, a)


FUNCTIONS = {np.linalg.solve: solve, np.linalg.inv: inv, np.linalg.det: det, np.linalg.slogdet: slogdet,
             np.linalg.cholesky: cholesky}
A.FUNCTIONS.update(FUNCTIONS)
__all__ = ["solve", "inv", "det", "slogdet", "cholesky", "DeviceArray"]
