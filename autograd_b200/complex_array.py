"""Complex device arrays as a PAIR of real device arrays (split / structure-of-arrays layout).

autograd's complex support (ComplexArrayVSpace, numpy_vspaces.py:40-66; ``match_complex``, numpy_vjps.py:872-880; the
complex halves of tests/test_systematic.py) needs a complex value type.  On the device a complex array is kept as two
float arrays of equal shape, ``re`` and ``im``: every complex ufunc is composed from the real ufuncs of
:mod:`autograd_b200.array` with NumPy's own complex formulas (numpy/_core/src/umath/funcs.inc.src ``nc_*`` and
loops.c.src: Smith's division, ``nc_expm1``, ``nc_log1p``, csqrt), so the lazy engine fuses a whole complex expression
into ONE generated kernel that reads the two planes with coalesced 128-bit accesses; linear structural functions
(reshape, transpose, indexing, concatenate, sum ...) act on both planes; bilinear contractions (dot, matmul, tensordot,
einsum, outer, kron) are four real contractions on the DMMA GEMM.  No new kernels are involved.
"""

from __future__ import annotations

import numbers

import numpy as np

from . import array as A
from .array import DeviceArray

_LN2 = 0.693147180559945309417232121458176568
_LN10 = 2.302585092994045684017991454684364208


def _is_complex_scalar(x) -> bool:
    return isinstance(x, (complex, np.complexfloating))


def is_complex_operand(x) -> bool:
    return type(x) is ComplexDeviceArray or _is_complex_scalar(x) or (isinstance(x, np.ndarray) and x.dtype.kind == "c")


class ComplexDeviceArray:
    __slots__ = ("re", "im", "__weakref__")
    __array_priority__ = 2000.0

    def __init__(self, re, im):
        re, im = A.asarray(re), A.asarray(im)
        if re.dtype.kind != "f":
            re = re.astype(np.float64)
        if im.dtype != re.dtype:
            dt = np.result_type(re.dtype, im.dtype)
            re, im = re.astype(dt), im.astype(dt)
        if re.shape != im.shape:
            shape = np.broadcast_shapes(re.shape, im.shape)
            re, im = A.broadcast_to(re, shape), A.broadcast_to(im, shape)
        self.re, self.im = re, im

    # ------------------------------------------------------------ metadata
    @property
    def shape(self):
        return self.re.shape

    @property
    def ndim(self):
        return self.re.ndim

    @property
    def size(self):
        return self.re.size

    @property
    def dtype(self):
        return np.dtype(np.complex64) if self.re.dtype == np.float32 else np.dtype(np.complex128)

    @property
    def real(self):
        return self.re

    @property
    def imag(self):
        return self.im

    @property
    def T(self):
        return ComplexDeviceArray(self.re.T, self.im.T)

    def __len__(self):
        return len(self.re)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self):
        return "ComplexDeviceArray(" + repr(self.get())[6:]

    def __hash__(self):
        return id(self)

    def __array__(self, dtype=None, copy=None):
        raise TypeError("autograd_b200.ComplexDeviceArray refuses implicit conversion to a host numpy array; call .get()")

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype)
        out.real = self.re.get()
        out.imag = self.im.get()
        return out

    def item(self):
        return self.get().item()

    def __complex__(self):
        return complex(self.get())

    def __bool__(self):
        return bool(self.get())

    def materialize(self):
        self.re.materialize()
        self.im.materialize()
        return self

    # ------------------------------------------------------------ structure
    def __getitem__(self, idx):
        return ComplexDeviceArray(self.re[idx], self.im[idx])

    def conj(self):
        return ComplexDeviceArray(self.re, -self.im)

    conjugate = conj

    def copy(self, order="C"):
        return ComplexDeviceArray(self.re.copy(), self.im.copy())

    def astype(self, dtype, **kw):
        dtype = np.dtype(dtype)
        if dtype.kind == "c":
            part = np.float32 if dtype == np.complex64 else np.float64
            return ComplexDeviceArray(self.re.astype(part), self.im.astype(part))
        return self.re.astype(dtype)          # NumPy discards the imaginary part (with a ComplexWarning)

    def reshape(self, *shape, order="C"):
        return _both(A.reshape)(self, shape[0] if len(shape) == 1 else shape)

    def ravel(self, order="C"):
        return _both(A.ravel)(self)

    def flatten(self, order="C"):
        return self.ravel().copy()

    def transpose(self, *axes):
        axes = axes[0] if len(axes) == 1 and not isinstance(axes[0], numbers.Integral) else (axes or None)
        return _both(A.transpose)(self, axes)

    def swapaxes(self, a, b):
        return _both(A.swapaxes)(self, a, b)

    def squeeze(self, axis=None):
        return _both(A.squeeze)(self, axis)

    def sum(self, axis=None, dtype=None, out=None, keepdims=False):
        return c_sum(self, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, dtype=None, out=None, keepdims=False):
        return c_mean(self, axis=axis, keepdims=keepdims)

    def prod(self, axis=None, dtype=None, out=None, keepdims=False):
        return c_prod(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, out=None, keepdims=False):
        return c_extreme(self, "max", axis, keepdims)

    def min(self, axis=None, out=None, keepdims=False):
        return c_extreme(self, "min", axis, keepdims)

    def dot(self, other):
        return c_dot(self, other)

    # ------------------------------------------------------------ NumPy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        for x in inputs:
            if A._is_box(x):
                return NotImplemented
        if method == "__call__":
            kwargs.pop("casting", None)
            out = kwargs.pop("out", None)
            if kwargs.pop("dtype", None) is not None or kwargs:
                raise TypeError(f"autograd_b200: numpy.{ufunc.__name__} keywords are not supported on complex device arrays")
            res = ufunc_call(ufunc, inputs)
            if out is not None:
                (target,) = out if isinstance(out, tuple) else (out,)
                if isinstance(target, np.ndarray):
                    # `x += y` on a HOST accumulator (the cotangent of a host argument, core.py:262-264) receiving a
                    # device contribution: an explicit device->host read, as for real values (array.__array_ufunc__)
                    target[...] = res.get() if hasattr(res, "get") else res
                    return target
                if type(target) is not ComplexDeviceArray:
                    raise TypeError("autograd_b200: out= of a complex result must be a ComplexDeviceArray")
                res = as_complex(res)
                target.re._assign(res.re)
                target.im._assign(res.im)
                return target
            return res
        if method == "reduce":
            axis, keepdims = kwargs.get("axis", 0), kwargs.get("keepdims", False)
            if ufunc is np.add:
                return c_sum(inputs[0], axis=axis, keepdims=keepdims)
            if ufunc is np.multiply:
                return c_prod(inputs[0], axis=axis, keepdims=keepdims)
            if ufunc in (np.maximum, np.minimum):
                return c_extreme(inputs[0], "max" if ufunc is np.maximum else "min", axis, keepdims)
        if method == "at" and ufunc is np.add:
            target, idx, values = inputs
            if type(target) is not ComplexDeviceArray:
                raise TypeError("autograd_b200: np.add.at with complex values needs a complex device target")
            values = as_complex(values)
            target.re._add_at(idx, values.re)
            target.im._add_at(idx, values.im)
            return None
        raise TypeError(f"autograd_b200: numpy.{ufunc.__name__}.{method} has no complex device implementation")

    def __array_function__(self, func, types, args, kwargs):
        impl = C_FUNCTIONS.get(func)
        if impl is None:
            if func in A.DECOMPOSE:
                return func._implementation(*args, **kwargs)
            raise TypeError(f"autograd_b200: numpy.{func.__name__} has no complex device implementation")
        return impl(*args, **kwargs)

    # ------------------------------------------------------------ operators
    def _bin(self, ufunc, other, reflected=False):
        if A._is_box(other):
            return NotImplemented
        return ufunc_call(ufunc, (other, self) if reflected else (self, other))

    def __add__(self, o): return self._bin(np.add, o)
    def __radd__(self, o): return self._bin(np.add, o, True)
    def __sub__(self, o): return self._bin(np.subtract, o)
    def __rsub__(self, o): return self._bin(np.subtract, o, True)
    def __mul__(self, o): return self._bin(np.multiply, o)
    def __rmul__(self, o): return self._bin(np.multiply, o, True)
    def __truediv__(self, o): return self._bin(np.true_divide, o)
    def __rtruediv__(self, o): return self._bin(np.true_divide, o, True)
    def __pow__(self, o): return self._bin(np.power, o)
    def __rpow__(self, o): return self._bin(np.power, o, True)
    def __matmul__(self, o): return NotImplemented if A._is_box(o) else c_matmul(self, o)
    def __rmatmul__(self, o): return NotImplemented if A._is_box(o) else c_matmul(o, self)
    def __neg__(self): return ComplexDeviceArray(-self.re, -self.im)
    def __pos__(self): return self
    def __abs__(self): return ufunc_call(np.absolute, (self,))
    def __eq__(self, o): return self._bin(np.equal, o)
    def __ne__(self, o): return self._bin(np.not_equal, o)

    def __iadd__(self, o):
        res = as_complex(self + o)
        self.re._assign(res.re)
        self.im._assign(res.im)
        return self


# =================================================================== conversion helpers
def as_complex(x) -> ComplexDeviceArray:
    if type(x) is ComplexDeviceArray:
        return x
    if isinstance(x, DeviceArray):
        return ComplexDeviceArray(x, A.zeros(x.shape, x.dtype if x.dtype.kind == "f" else np.float64))
    h = np.asarray(x)
    if h.dtype.kind == "c":
        part = np.float32 if h.dtype == np.complex64 else np.float64
        return ComplexDeviceArray(A.asarray(np.array(h.real, dtype=part, order="C")),       # (np.ascontiguousarray would
                                  A.asarray(np.array(h.imag, dtype=part, order="C")))       #  turn a 0-d value into 1-d)
    d = A.asarray(h if h.dtype.kind == "f" else h.astype(np.float64))
    return ComplexDeviceArray(d, A.zeros(d.shape, d.dtype))


def _parts(x):
    """(re, im, is_complex): parts are DeviceArrays, host arrays or Python floats; ``im`` is None for real operands."""
    if type(x) is ComplexDeviceArray:
        return x.re, x.im, True
    if _is_complex_scalar(x):
        return float(x.real), float(x.imag), True
    if isinstance(x, np.ndarray) and x.dtype.kind == "c":
        c = as_complex(x)
        return c.re, c.im, True
    if isinstance(x, (list, tuple)):
        return _parts(np.asarray(x))
    return x, None, False


def _wrap(re, im):
    if not isinstance(re, DeviceArray) and not isinstance(im, DeviceArray):
        return complex(re, im) if np.ndim(re) == 0 and np.ndim(im) == 0 else as_complex(np.asarray(re) + 1j * np.asarray(im))
    return ComplexDeviceArray(re, im)


def _both(fn):
    """Lift a LINEAR structural function of one array argument (first positional) to both planes."""
    def lifted(a, *args, **kwargs):
        a = as_complex(a)
        return ComplexDeviceArray(fn(a.re, *args, **kwargs), fn(a.im, *args, **kwargs))
    return lifted


def _both_seq(fn):
    """Lift a function of a SEQUENCE of arrays (concatenate, stack) to both planes."""
    def lifted(arrays, *args, **kwargs):
        cs = [as_complex(a) for a in arrays]
        return ComplexDeviceArray(fn([c.re for c in cs], *args, **kwargs), fn([c.im for c in cs], *args, **kwargs))
    return lifted


# =================================================================== elementwise formulas
def _c_mul(a, b, c, d):
    return a * c - b * d, a * d + b * c


def _c_div(a, b, c, d):
    """(a + ib) / (c + id) by Smith's algorithm, as NumPy's complex division loop does (loops.c.src)."""
    big_c = np.abs(c) >= np.abs(d)
    c1 = np.where(big_c, c, d)
    d1 = np.where(big_c, d, c)
    safe = np.where(c1 == 0, 1.0, c1)
    rat = d1 / safe
    scl = 1.0 / (c1 + d1 * rat)
    re = np.where(big_c, (a + b * rat) * scl, (a * rat + b) * scl)
    im = np.where(big_c, (b - a * rat) * scl, (b * rat - a) * scl)
    zero = (c == 0) & (d == 0)
    return np.where(zero, a / np.abs(c), re), np.where(zero, b / np.abs(c), im)


def _c_exp(a, b):
    e = np.exp(a)
    return e * np.cos(b), e * np.sin(b)


def _c_log(a, b):
    return np.log(np.hypot(a, b)), np.arctan2(b, a)


def _c_sqrt(a, b):
    """Principal square root (C99 csqrt as used by npy_csqrt): t = sqrt((|a| + |z|) / 2)."""
    t = np.sqrt((np.abs(a) + np.hypot(a, b)) * 0.5)
    t_safe = np.where(t == 0, 1.0, t)
    re = np.where(a >= 0, t, np.abs(b) / (2.0 * t_safe))
    im = np.where(a >= 0, b / (2.0 * t_safe), np.copysign(t, b))
    zero = (a == 0) & (b == 0)
    return np.where(zero, 0.0, re), np.where(zero, b, im)


def _c_pow_int(a, b, n: int):
    if n == 0:
        return a * 0.0 + 1.0, b * 0.0
    neg = n < 0
    n = abs(n)
    rr, ri, pr, pi = None, None, a, b
    while n:                                  # binary powering, like npy_cpow for small integer exponents
        if n & 1:
            rr, ri = (pr, pi) if rr is None else _c_mul(rr, ri, pr, pi)
        n >>= 1
        if n:
            pr, pi = _c_mul(pr, pi, pr, pi)
    if neg:
        return _c_div(1.0, 0.0, rr, ri)
    return rr, ri


def _unary(ufunc, a, b):
    if ufunc is np.negative:
        return -a, -b
    if ufunc is np.positive:
        return a, b
    if ufunc in (np.conjugate,):
        return a, -b
    if ufunc is np.exp:
        return _c_exp(a, b)
    if ufunc is np.exp2:
        return _c_exp(a * _LN2, b * _LN2)
    if ufunc is np.expm1:                     # nc_expm1: expm1(a) cos(b) - 2 sin^2(b/2), exp(a) sin(b)
        s = np.sin(b * 0.5)
        return np.expm1(a) * np.cos(b) - 2.0 * s * s, np.exp(a) * np.sin(b)
    if ufunc is np.log:
        return _c_log(a, b)
    if ufunc is np.log2:
        lr, li = _c_log(a, b)
        return lr / _LN2, li / _LN2
    if ufunc is np.log10:
        lr, li = _c_log(a, b)
        return lr / _LN10, li / _LN10
    if ufunc is np.log1p:                     # nc_log1p
        return np.log(np.hypot(a + 1.0, b)), np.arctan2(b, a + 1.0)
    if ufunc is np.sqrt:
        return _c_sqrt(a, b)
    if ufunc is np.square:
        return _c_mul(a, b, a, b)
    if ufunc is np.reciprocal:
        return _c_div(1.0, 0.0, a, b)
    if ufunc is np.sin:
        return np.sin(a) * np.cosh(b), np.cos(a) * np.sinh(b)
    if ufunc is np.cos:
        return np.cos(a) * np.cosh(b), -(np.sin(a) * np.sinh(b))
    if ufunc is np.tan:
        den = np.cos(2.0 * a) + np.cosh(2.0 * b)
        return np.sin(2.0 * a) / den, np.sinh(2.0 * b) / den
    if ufunc is np.sinh:
        return np.sinh(a) * np.cos(b), np.cosh(a) * np.sin(b)
    if ufunc is np.cosh:
        return np.cosh(a) * np.cos(b), np.sinh(a) * np.sin(b)
    if ufunc is np.tanh:
        den = np.cosh(2.0 * a) + np.cos(2.0 * b)
        return np.sinh(2.0 * a) / den, np.sin(2.0 * b) / den
    if ufunc is np.arcsinh:                   # log(z + sqrt(z^2 + 1)), evaluated on the half plane Re z >= 0 (odd function)
        flip = a < 0
        a1, b1 = np.where(flip, -a, a), np.where(flip, -b, b)
        zr, zi = _c_mul(a1, b1, a1, b1)
        sr, si = _c_sqrt(zr + 1.0, zi)
        lr, li = _c_log(a1 + sr, b1 + si)
        return np.where(flip, -lr, lr), np.where(flip, -li, li)
    if ufunc is np.arccosh:                   # log(z + sqrt(z + 1) sqrt(z - 1))
        pr, pi = _c_sqrt(a + 1.0, b)
        mr, mi = _c_sqrt(a - 1.0, b)
        qr, qi = _c_mul(pr, pi, mr, mi)
        return _c_log(a + qr, b + qi)
    if ufunc is np.arctanh:                   # (log(1 + z) - log(1 - z)) / 2
        l1r, l1i = np.log(np.hypot(1.0 + a, b)), np.arctan2(b, 1.0 + a)
        l2r, l2i = np.log(np.hypot(1.0 - a, b)), np.arctan2(-b, 1.0 - a)
        return 0.5 * (l1r - l2r), 0.5 * (l1i - l2i)
    raise TypeError(f"autograd_b200: numpy.{ufunc.__name__} has no complex device implementation")


def ufunc_call(ufunc, inputs):
    """``ufunc(*inputs)`` where at least one input is complex (device pair, Python/NumPy complex scalar or array)."""
    if ufunc is np.matmul:
        return c_matmul(*inputs)
    parts = [_parts(x) for x in inputs]
    if ufunc.nin == 1:
        a, b, _ = parts[0]
        if ufunc in (np.absolute,):
            return np.hypot(a, b)
        if ufunc is np.isnan:
            return np.isnan(a) | np.isnan(b)
        if ufunc is np.isfinite:
            return np.isfinite(a) & np.isfinite(b)
        if ufunc is np.isinf:
            return np.isinf(a) | np.isinf(b)
        if ufunc is np.sign:                  # NumPy 2: z / |z|
            r = np.hypot(a, b)
            safe = np.where(r == 0, 1.0, r)
            return _wrap(a / safe, b / safe)
        return _wrap(*_unary(ufunc, a, b))
    (a, b, ca), (c, d, cc) = parts
    b = 0.0 if b is None else b
    d = 0.0 if d is None else d
    if ufunc is np.add:
        return _wrap(a + c, (b + d) if (ca and cc) else (b if ca else d))
    if ufunc is np.subtract:
        return _wrap(a - c, (b - d) if (ca and cc) else (b if ca else -d))
    if ufunc is np.multiply:
        if not cc:
            return _wrap(a * c, b * c)
        if not ca:
            return _wrap(a * c, a * d)
        return _wrap(*_c_mul(a, b, c, d))
    if ufunc in (np.true_divide, np.divide):
        if not cc:
            return _wrap(a / c, b / c)
        return _wrap(*_c_div(a, b, c, d))
    if ufunc in (np.power, np.float_power):
        if not cc and isinstance(c, (int, float, np.integer, np.floating)) and float(c) == int(c) and abs(int(c)) <= 64:
            return _wrap(*_c_pow_int(a, b, int(c)))
        lr, li = _c_log(a, b)                 # z ** w = exp(w log z); 0 ** w = 0 for Re w > 0
        er, ei = _c_exp(*_c_mul(c, d, lr, li))
        zero = (a == 0) & (b == 0)
        return _wrap(np.where(zero, 0.0, er), np.where(zero, 0.0, ei))
    if ufunc is np.equal:
        return (a == c) & (b == d)
    if ufunc is np.not_equal:
        return (a != c) | (b != d)
    raise TypeError(f"autograd_b200: numpy.{ufunc.__name__} has no complex device implementation")


# =================================================================== reductions
def c_sum(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    a = as_complex(a)
    return ComplexDeviceArray(A.reduce_(a.re, "sum", axis, keepdims), A.reduce_(a.im, "sum", axis, keepdims))


def _count(shape, axis):
    if axis is None:
        return int(np.prod(shape)) if shape else 1
    axes = (axis,) if isinstance(axis, numbers.Integral) else tuple(axis)
    return int(np.prod([shape[ax] for ax in axes]))


def c_mean(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    a = as_complex(a)
    n = _count(a.shape, axis)
    s = c_sum(a, axis=axis, keepdims=keepdims)
    return ComplexDeviceArray(s.re / n, s.im / n)


def c_var(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw):
    a = as_complex(a)
    m = c_mean(a, axis=axis, keepdims=True)
    dr, di = a.re - m.re, a.im - m.im
    return A.reduce_(dr * dr + di * di, "sum", axis, keepdims) / (_count(a.shape, axis) - ddof)


def c_std(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw):
    return np.sqrt(c_var(a, axis=axis, ddof=ddof, keepdims=keepdims))


def c_prod(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    """Product over ``axis`` by pairwise halving (log2(n) fused complex multiplies)."""
    a = as_complex(a)
    if axis is None:
        out = c_prod(_both(A.ravel)(a), axis=0, keepdims=False) if a.ndim else a
        return _both(A.reshape)(out, (1,) * a.ndim) if keepdims else out
    axes = (axis,) if isinstance(axis, numbers.Integral) else tuple(sorted(ax % a.ndim for ax in axis))
    re, im = a.re, a.im
    for ax in reversed(axes):
        ax = ax % re.ndim
        n = re.shape[ax]
        sl = lambda lo, hi: tuple(slice(lo, hi) if i == ax else slice(None) for i in range(re.ndim))  # noqa: E731
        while n > 1:
            half = n // 2
            pr, pi = _c_mul(re[sl(0, half)], im[sl(0, half)], re[sl(half, 2 * half)], im[sl(half, 2 * half)])
            if n % 2:
                pr = A.concatenate([pr, re[sl(2 * half, n)]], axis=ax)
                pi = A.concatenate([pi, im[sl(2 * half, n)]], axis=ax)
            re, im, n = pr, pi, half + (n % 2)
        if n == 0:
            shape = re.shape[:ax] + (1,) + re.shape[ax + 1:]
            re, im = A.ones(shape, re.dtype), A.zeros(shape, re.dtype)
    out = ComplexDeviceArray(re, im)
    if not keepdims:
        out = _both(A.squeeze)(out, axes)
    return out


def c_extreme(a, which, axis=None, keepdims=False):
    """np.max / np.min of complex values: NumPy orders them lexicographically (real part, then imaginary part)."""
    a = as_complex(a)
    op = "max" if which == "max" else "min"
    best_re = A.reduce_(a.re, op, axis, True)
    fill = -np.inf if which == "max" else np.inf
    best_im = A.reduce_(A.where(a.re == best_re, a.im, fill), op, axis, True)
    out = ComplexDeviceArray(best_re, best_im)
    if not keepdims:
        out = _both(A.squeeze)(out, axis) if axis is not None else _both(A.reshape)(out, ())
    return out


# =================================================================== contractions (bilinear: four real products)
def _bilinear(fn):
    def lifted(a, b, *args, **kwargs):
        ar, ai, ca = _parts(a)
        br, bi, cb = _parts(b)
        if not ca:
            return ComplexDeviceArray(fn(ar, br, *args, **kwargs), fn(ar, bi, *args, **kwargs))
        if not cb:
            return ComplexDeviceArray(fn(ar, br, *args, **kwargs), fn(ai, br, *args, **kwargs))
        return ComplexDeviceArray(fn(ar, br, *args, **kwargs) - fn(ai, bi, *args, **kwargs),
                                  fn(ar, bi, *args, **kwargs) + fn(ai, br, *args, **kwargs))
    return lifted


c_dot = _bilinear(lambda a, b: A.dot(a, b))
c_matmul = _bilinear(lambda a, b: A.matmul(a, b))
c_tensordot = _bilinear(lambda a, b, axes=2: A.tensordot(a, b, axes))
c_inner = _bilinear(lambda a, b: A.inner(a, b))
c_outer = _bilinear(lambda a, b: A.outer(a, b))
c_kron = _bilinear(lambda a, b: A.kron(a, b))


def c_einsum(*operands, **kwargs):
    """np.einsum with complex operands: multilinear, so the result is the sum over the 2^k real/imaginary choices of the
    k complex operands with the sign / plane given by the number of imaginary factors."""
    if not isinstance(operands[0], str):
        raise NotImplementedError("autograd_b200: complex einsum supports the subscript-string form")
    spec, ops = operands[0], list(operands[1:])
    parts = [_parts(o) for o in ops]
    re = im = None
    idx = [i for i, p in enumerate(parts) if p[2]]
    for mask in range(1 << len(idx)):
        chosen, n_im = [], 0
        for i, p in enumerate(parts):
            if p[2] and (mask >> idx.index(i)) & 1:
                chosen.append(p[1])
                n_im += 1
            else:
                chosen.append(p[0])
        term = A.einsum(spec, *chosen, **kwargs)
        sign = -1.0 if (n_im // 2) % 2 else 1.0          # i^n: 1, i, -1, -i
        if n_im % 2 == 0:
            re = term * sign if re is None else re + term * sign
        else:
            im = term * sign if im is None else im + term * sign
    if im is None:
        im = re * 0.0
    if re is None:
        re = im * 0.0
    return ComplexDeviceArray(re, im)


def c_real_if_close(a, tol=100):
    """np.real_if_close: the real part when every imaginary part is within ``tol`` machine epsilons of zero (this is a
    data-dependent result TYPE, so it reads one boolean back from the device, as NumPy inspects the values)."""
    a = as_complex(a)
    eps = np.finfo(a.re.dtype).eps
    close = bool(A.reduce_(abs(a.im) < (tol * eps if tol > 1 else tol), "all", None, False).get())
    return a.re if close else a


def c_where(cond, x=None, y=None):
    if type(cond) is ComplexDeviceArray:          # truthiness of a complex number: non-zero real or imaginary part
        cond = (cond.re != 0) | (cond.im != 0)
    elif _is_complex_scalar(cond):
        cond = bool(cond)
    xr, xi, cx = _parts(x)
    yr, yi, cy = _parts(y)
    xi = 0.0 if xi is None else xi
    yi = 0.0 if yi is None else yi
    return ComplexDeviceArray(A.where(cond, xr, yr), A.where(cond, xi, yi))


def c_angle(z, deg=False):
    z = as_complex(z)
    out = np.arctan2(z.im, z.re)
    return out * (180.0 / np.pi) if deg else out


def c_full_like(a, fill_value, dtype=None, **kw):
    r, i, _ = _parts(fill_value)
    return ComplexDeviceArray(A.full(np.shape(a), r, np.float64), A.full(np.shape(a), 0.0 if i is None else i, np.float64))


def full(shape, fill_value, dtype=np.complex128):
    part = np.float32 if np.dtype(dtype) == np.complex64 else np.float64
    v = complex(fill_value)
    return ComplexDeviceArray(A.full(shape, v.real, part), A.full(shape, v.imag, part))


C_FUNCTIONS = {
    np.shape: lambda a: a.shape, np.ndim: lambda a: a.ndim, np.size: lambda a, axis=None: a.size if axis is None else a.shape[axis],
    np.result_type: lambda *args: np.result_type(*[x.dtype if hasattr(x, "dtype") else x for x in args]),
    np.iscomplexobj: lambda x: True, np.isrealobj: lambda x: False,
    np.real: lambda z: as_complex(z).re, np.imag: lambda z: as_complex(z).im, np.conj: lambda z: as_complex(z).conj(),
    np.conjugate: lambda z: as_complex(z).conj(), np.angle: c_angle,
    np.reshape: _both(A.reshape), np.ravel: _both(A.ravel), np.transpose: _both(A.transpose), np.swapaxes: _both(A.swapaxes),
    np.moveaxis: _both(A.moveaxis), np.expand_dims: _both(A.expand_dims), np.squeeze: _both(A.squeeze),
    np.broadcast_to: _both(A.broadcast_to), np.flip: _both(A.flip), np.roll: _both(A.roll), np.repeat: _both(A.repeat),
    np.tile: _both(A.tile), np.pad: _both(A.pad), np.take: _both(A.take), np.diag: _both(A.diag),
    np.diagonal: _both(A.diagonal), np.trace: _both(A.trace), np.tril: _both(A.tril), np.triu: _both(A.triu),
    np.diff: _both(A.diff), np.cumsum: _both(A.cumsum), np.copy: lambda a, **kw: as_complex(a).copy(),
    np.atleast_1d: _both(A.atleast_1d), np.atleast_2d: _both(A.atleast_2d), np.atleast_3d: _both(A.atleast_3d),
    np.concatenate: _both_seq(A.concatenate), np.stack: _both_seq(A.stack),
    np.sum: c_sum, np.mean: c_mean, np.var: c_var, np.std: c_std, np.prod: c_prod,
    np.max: lambda a, axis=None, out=None, keepdims=False, **kw: c_extreme(a, "max", axis, keepdims),
    np.amax: lambda a, axis=None, out=None, keepdims=False, **kw: c_extreme(a, "max", axis, keepdims),
    np.min: lambda a, axis=None, out=None, keepdims=False, **kw: c_extreme(a, "min", axis, keepdims),
    np.amin: lambda a, axis=None, out=None, keepdims=False, **kw: c_extreme(a, "min", axis, keepdims),
    np.dot: c_dot, np.matmul: c_matmul, np.tensordot: c_tensordot, np.inner: c_inner, np.outer: c_outer, np.kron: c_kron,
    np.einsum: c_einsum, np.where: c_where, np.real_if_close: c_real_if_close,
    np.zeros_like: lambda a, dtype=None, **kw: full(np.shape(a), 0.0) if dtype is None or np.dtype(dtype).kind == "c" else A.zeros(np.shape(a), dtype),
    np.ones_like: lambda a, dtype=None, **kw: full(np.shape(a), 1.0) if dtype is None or np.dtype(dtype).kind == "c" else A.ones(np.shape(a), dtype),
    np.full_like: c_full_like,
    np.astype: lambda a, dtype, copy=True, **kw: as_complex(a).astype(dtype),
    np.absolute: lambda z: ufunc_call(np.absolute, (z,)),
}

A._Complex = ComplexDeviceArray
A._complex_ufunc = ufunc_call
A._complex_full = full
A._complex_where = c_where
