"""Parity cases shared by the CPU (fake backend: host logic) and GPU (CUDA backend: kernels) suites.

Every case builds seeded host inputs, runs the SAME expression on numpy.ndarray (the oracle: NumPy is
what the reference executes at autograd/tracer.py:94) and on DeviceArray, and compares with the
north_star tolerances: bit-exact for bool/int results, 1e-12 for f64 elementwise, 1e-10 for f64
reductions/contractions, 1e-5 for f32 — relative to the max-norm of the reference.
"""

from __future__ import annotations

import numpy as onp

import autograd_b200 as b200

TOL_ELEMENTWISE = 1e-12
TOL_REDUCE = 1e-10
TOL_F32 = 1e-5


def to_dev(x):
    if isinstance(x, onp.ndarray):
        return b200.asarray(x)
    if isinstance(x, (list, tuple)):
        return type(x)(to_dev(e) for e in x)
    if isinstance(x, dict):
        return {k: to_dev(v) for k, v in x.items()}
    return x


def to_host(x):
    if isinstance(x, b200.DeviceArray):
        return x.get()
    if isinstance(x, (list, tuple)):
        return type(x)(to_host(e) for e in x)
    if isinstance(x, dict):
        return {k: to_host(v) for k, v in x.items()}
    return x


def leaves(t):
    if isinstance(t, (list, tuple)):
        return [l for x in t for l in leaves(x)]
    if isinstance(t, dict):
        return [l for k in sorted(t) for l in leaves(t[k])]
    return [t]


def assert_close(got, want, tol, what=""):
    got, want = onp.asarray(got), onp.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    if want.dtype.kind in "biu" or tol == 0:
        assert onp.array_equal(got, want, equal_nan=want.dtype.kind == "f"), f"{what}: not bit-exact"
        return
    if want.dtype == onp.float32:
        tol = max(tol, TOL_F32)
    finite = onp.isfinite(want)
    assert onp.array_equal(finite, onp.isfinite(got)), f"{what}: non-finite pattern differs"
    assert onp.array_equal(got[~finite], want[~finite], equal_nan=True), f"{what}: inf/nan values differ"
    if finite.any():
        scale = onp.max(onp.abs(want[finite]))
        err = onp.max(onp.abs(got[finite].astype(onp.float64) - want[finite].astype(onp.float64)))
        assert err <= tol * max(scale, 1e-300), f"{what}: max abs err {err:.3e} > {tol:g} * {scale:.3e}"


def check(fn, *host_args, tol=TOL_ELEMENTWISE, what=""):
    """fn(np_module_like, *args) evaluated on host and device."""
    with onp.errstate(all="ignore"):
        want = fn(*host_args)
    got = fn(*to_dev(list(host_args)))
    for g, w in zip(leaves(to_host(got)), leaves(want)):
        assert_close(g, w, tol, what or getattr(fn, "__name__", "case"))


def rs(seed=0):
    return onp.random.RandomState(seed)


# =============================================================== elementwise
UNARY = ["negative", "exp", "log", "log1p", "expm1", "tanh", "sinh", "cosh", "sin", "cos", "tan", "sqrt", "square",
         "abs", "sign", "floor", "ceil", "trunc", "rint", "reciprocal", "exp2", "log2", "log10", "arctan", "arcsin",
         "arccos", "arcsinh", "isfinite", "isnan", "isinf", "logical_not"]
BINARY = ["add", "subtract", "multiply", "true_divide", "power", "maximum", "minimum", "fmax", "fmin", "remainder",
          "floor_divide", "arctan2", "hypot", "logaddexp", "equal", "not_equal", "less", "less_equal", "greater",
          "greater_equal", "logical_and", "logical_or", "logical_xor"]


def case_unary(name, dtype=onp.float64, n=4099):
    x = rs(1).randn(n).astype(dtype) * 2
    if name in ("log", "sqrt", "log2", "log10", "reciprocal"):
        x = onp.abs(x) + 0.1
    if name in ("log1p",):
        x = onp.abs(x)
    if name in ("arcsin", "arccos"):
        x = onp.clip(x / 4, -1, 1)
    x[:4] = [0.0, -0.0, 1.0, -1.0] if name not in ("log", "sqrt", "log2", "log10", "reciprocal") else [1.0, 2.0, 0.5, 4.0]
    f = getattr(onp, name)
    # CUDA libm functions are within 1-2 ulp of glibc's; transcendental results get a few ulp of slack
    exact = name in ("negative", "square", "abs", "sign", "floor", "ceil", "trunc", "rint", "isfinite", "isnan",
                     "isinf", "logical_not", "sqrt", "reciprocal")
    check(lambda a: f(a), x, tol=0 if exact else TOL_ELEMENTWISE, what=f"{name}[{onp.dtype(dtype).name}]")


def case_binary(name, dtype=onp.float64):
    r = rs(2)
    a = (r.randn(37, 53) * 3).astype(dtype)
    b = (r.randn(37, 53) * 3).astype(dtype)
    if name == "power":
        a = onp.abs(a) + 0.1
    if name in ("remainder", "floor_divide"):
        b[onp.abs(b) < 0.3] = 0.7
    f = getattr(onp, name)
    exact = name in ("add", "subtract", "multiply", "true_divide", "maximum", "minimum", "fmax", "fmin", "remainder",
                     "floor_divide", "equal", "not_equal", "less", "less_equal", "greater", "greater_equal",
                     "logical_and", "logical_or", "logical_xor")
    check(lambda x, y: f(x, y), a, b, tol=0 if exact else TOL_ELEMENTWISE, what=f"{name}[{onp.dtype(dtype).name}]")


def case_nan_inf_propagation():
    x = onp.array([onp.nan, onp.inf, -onp.inf, 0.0, -0.0, 1.5, -2.5, 1e308, 5e-324])
    y = onp.array([1.0, onp.nan, 2.0, -0.0, 0.0, onp.inf, -onp.inf, 1e308, 2.0])
    check(lambda a, b: (onp.maximum(a, b), onp.minimum(a, b), onp.fmax(a, b), onp.fmin(a, b), a + b, a * b, a / b,
                        onp.sign(a), onp.abs(a), onp.isfinite(a), a == b, a < b), x, y, tol=0, what="nan/inf")


def case_broadcasting():
    r = rs(3)
    a, b, c, s = r.randn(7, 1, 5), r.randn(3, 5), r.randn(5), r.randn()
    check(lambda a, b, c: a * b + c - 2.5 * a / (b * b + 1.0), a, b, c, tol=0, what="broadcast chain")
    check(lambda a, c: (a + c) * s, a, c, tol=0, what="python-float scalar")
    check(lambda a: a + onp.float64(1.5), a, tol=0, what="numpy scalar")
    check(lambda b: 1 - b, b, tol=0, what="int scalar rsub")
    check(lambda b: 2 ** b, b, what="rpow")


def case_dtype_promotion():
    r = rs(4)
    f32, f64 = r.randn(33).astype(onp.float32), r.randn(33)
    i64, i32 = r.randint(-50, 50, 33), r.randint(-50, 50, 33).astype(onp.int32)
    bl = r.rand(33) > 0.5
    check(lambda a, b: a + b, f32, f64, tol=0, what="f32+f64")
    check(lambda a: a * 2.0, f32, tol=0, what="f32*pyfloat stays f32")
    check(lambda a: a * 2, i64, tol=0, what="i64*pyint")
    check(lambda a: a * 2.5, i64, tol=0, what="i64*pyfloat")
    check(lambda a, b: a + b, i32, i64, tol=0, what="i32+i64")
    check(lambda a, b: a / b, i64, i64 * 0 + 3, tol=0, what="int true_divide")
    check(lambda a, b: a * b, f64, bl, tol=0, what="f64*bool")
    check(lambda a: onp.exp(a), i64 % 5, what="exp(int)")
    check(lambda a, b: a == b, i64, f64, tol=0, what="int==float")
    check(lambda a: a.astype(onp.float32), f64, tol=0, what="astype f32")
    check(lambda a: (a * 7.3).astype(int), f64, tol=0, what="astype int truncation")
    check(lambda a: a.astype(bool), f64 * (r.rand(33) > 0.3), tol=0, what="astype bool")
    check(lambda a, b: onp.where(b, a, 0.0), f32, bl, tol=0, what="where f32 / py float")
    check(lambda a, b: onp.where(b, a, a * 2), i64, bl, tol=0, what="where int")


def case_integer_semantics():
    r = rs(5)
    a = r.randint(-100, 100, 257)
    b = r.randint(1, 9, 257) * onp.where(r.rand(257) > 0.5, 1, -1)
    check(lambda a, b: (a % b, a // b, onp.mod(a, 7), onp.mod(a, -7), a ** 2, onp.abs(a), -a, onp.sign(a)), a, b, tol=0,
          what="floored int mod / floordiv")
    x = r.randn(257) * 20
    check(lambda x: (onp.mod(x, 3.0), onp.mod(x, -3.0), onp.floor(x), onp.floor(x).astype(int) % 16), x, tol=0,
          what="floored float mod")


def case_power_special_exponents():
    x = onp.abs(rs(6).randn(301)) + 0.05
    check(lambda x: (x ** 2, x ** 1, x ** 0, x ** -1, x ** 0.5, x ** 3, x ** 2.5, x ** -2), x, what="pow const")
    check(lambda x: x ** 2, -x, tol=0, what="square of negative")


def case_fused_chain_matches_unfused():
    x = rs(7).randn(100003)
    def f(x):
        t = onp.exp(-2 * x)
        u = (1.0 - t) / (1.0 + t)
        return u * u - x * 0.25 + onp.tanh(x) / (1.0 + x * x)
    check(f, x, what="fused chain")


def case_large_vector_tail():
    for n in (1, 2, 3, 255, 256, 257, 4095, 4097, 100001):
        x = rs(n).randn(n)
        y = rs(n + 1).randn(n)
        check(lambda a, b: a * b + a, x, y, tol=0, what=f"n={n}")
        check(lambda a: a[1:] * 2.0, x, tol=0, what=f"misaligned view n={n}") if n > 1 else None


# =============================================================== layout
def case_views():
    assert b200.asarray(onp.float64(3.5)).shape == () and b200.asarray(onp.array(2.0)).get().shape == ()
    assert b200.asarray(onp.zeros((0, 3))).shape == (0, 3) and b200.asarray(onp.zeros((0, 3))).get().shape == (0, 3)
    r = rs(8)
    a = r.randn(6, 8, 10)
    check(lambda a: a.T + 1.0, a, tol=0, what="T")
    check(lambda a: onp.swapaxes(a, 0, 2) * 2, a, tol=0, what="swapaxes")
    check(lambda a: onp.transpose(a, (1, 2, 0)) - 1, a, tol=0, what="transpose perm")
    check(lambda a: a[1:5, ::2, ::-1] + 0.5, a, tol=0, what="slices")
    check(lambda a: a[2], a, tol=0, what="int index view")
    check(lambda a: a[:, None, 3, :] * 1.0, a, tol=0, what="newaxis")
    check(lambda a: a[..., -1], a, tol=0, what="ellipsis")
    check(lambda a: onp.reshape(a, (48, 10)) + 0, a, tol=0, what="reshape")
    check(lambda a: onp.reshape(a.T, (10, 48)) + 0, a, tol=0, what="reshape of transposed (copy)")
    check(lambda a: onp.ravel(a[:, ::2]), a, tol=0, what="ravel strided")
    check(lambda a: onp.expand_dims(a, 1).shape == (6, 1, 8, 10) and onp.squeeze(a[:, :1], 1), a, tol=0, what="squeeze")
    check(lambda a: onp.broadcast_to(a[:, :1, :], (6, 8, 10)) + a, a, tol=0, what="broadcast_to")
    check(lambda a: onp.moveaxis(a, 0, -1) + 0, a, tol=0, what="moveaxis")
    check(lambda a: onp.flip(a, 1) + 0, a, tol=0, what="flip")
    check(lambda a: (a + 1.0).reshape(-1) * 2, a, tol=0, what="lazy reshape stays fused")
    check(lambda a: a.copy().T.copy(), a, tol=0, what="copy of transpose (tiled transpose kernel)")
    big = r.randn(130, 70)
    check(lambda a: a.T.copy(), big, tol=0, what="2-D transpose copy")
    check(lambda a: a.astype(onp.float32).T.copy(), big, tol=0, what="2-D transpose copy f32")


def case_roll():
    r = rs(9)
    a = r.randn(16, 24)
    for shift, axis in ((1, 0), (-1, 0), (1, 1), (-1, 1), (5, 0), (-7, 1), (16, 0), (0, 1)):
        check(lambda a: onp.roll(a, shift, axis=axis), a, tol=0, what=f"roll {shift} axis {axis}")
    check(lambda a: onp.roll(a, 1, axis=0) + onp.roll(a, -1, axis=0) + onp.roll(a, 1, axis=1) + onp.roll(a, -1, axis=1),
          a, tol=0, what="stencil of rolls")
    check(lambda a: onp.roll(onp.roll(a, 2, axis=0), -3, axis=1) * 2.0, a, tol=0, what="roll of roll")
    check(lambda a: onp.roll(a, 3), a, tol=0, what="flat roll")
    check(lambda a: onp.roll(a + 1.0, 1, axis=1), a, tol=0, what="roll of pending expression")
    check(lambda a: onp.roll(a, 1, axis=0).T + 0, a, tol=0, what="transpose of rolled")
    check(lambda a: onp.roll(a, 1, axis=0)[2:5], a, tol=0, what="slice of rolled")


def case_row_mode_layouts():
    """General-layout operands with a long inner dim: exercises the row-structured generated kernels."""
    r = rs(30)
    a = r.randn(48, 384)
    b = r.randn(384)
    c = r.randn(48, 1)
    check(lambda a, b, c: a * b + c, a, b, c, tol=0, what="row/col broadcast, inner 384")
    check(lambda a: onp.roll(a, 1, axis=0) + onp.roll(a, -1, axis=0) + onp.roll(a, 1, axis=1) + onp.roll(a, -1, axis=1),
          a, tol=0, what="4-point stencil, inner 384")
    check(lambda a: (a + onp.roll(a, 7, axis=1)) / 4.0 - onp.roll(a, -5, axis=0), a, tol=0, what="rolls + contiguous")
    sq = r.randn(256, 256)
    check(lambda s: s.T * 2.0 + s, sq, tol=0, what="transposed operand, inner 256")
    check(lambda s: s[::2, ::-1] + 1.0, sq, tol=0, what="strided/negative-stride view")
    t3 = r.randn(6, 10, 130)
    check(lambda t: onp.roll(t, 1, axis=1) - t + onp.roll(t, -3, axis=2) * onp.roll(t, 2, axis=0), t3, tol=0,
          what="3-D rolls (outer coords decomposed per row)")
    check(lambda t, v: t * v[:130] + onp.swapaxes(t, 0, 1)[:6, :6, :].sum(), t3, b, tol=TOL_REDUCE, what="3-D broadcast")
    odd = r.randn(33, 131)
    check(lambda o: onp.roll(o, 1, axis=1) * 2.0 + o, odd, tol=0, what="odd inner size (no vector path)")
    f32 = r.randn(40, 512).astype(onp.float32)
    check(lambda x: onp.roll(x, -1, axis=1) + x * 2, f32, tol=0, what="f32 row mode (4-wide vectors)")
    idx = onp.arange(2048)
    cx, cy = onp.meshgrid(idx[:40], idx[:160])
    v = r.randn(160, 40)
    check(lambda cx, v: (cx - v, onp.floor(cx - v).astype(int) % 40), cx, v, tol=0, what="meshgrid broadcast views")


def case_tri_diag():
    r = rs(31)
    a, v = r.randn(6, 7), r.randn(5)
    check(lambda a: (onp.tril(a), onp.triu(a), onp.tril(a, -1), onp.triu(a, 2)), a, tol=0, what="tril/triu")
    check(lambda a: (onp.diag(a), onp.diag(a, 1), onp.diag(a, -2), onp.diagonal(a.T, 1)), a, tol=0, what="diag of matrix")
    check(lambda v: (onp.diag(v), onp.diag(v, 2), onp.diag(v, -1)), v, tol=0, what="diag of vector")
    check(lambda a: (onp.trace(a), onp.trace(a, 1)), a, tol=TOL_REDUCE, what="trace")
    check(lambda a: (onp.diff(a), onp.diff(a, axis=0), onp.diff(a, n=2, axis=1)), a, tol=0, what="diff")
    t = r.randn(3, 4, 5)
    check(lambda t: onp.diagonal(t, 0, 1, 2), t, tol=0, what="diagonal of 3-D")


def case_concat_repeat_pad():
    r = rs(10)
    a, b, c = r.randn(4, 5), r.randn(3, 5), r.randn(4, 2)
    check(lambda a, b: onp.concatenate([a, b], 0), a, b, tol=0, what="concat0")
    check(lambda a, c: onp.concatenate([a, c], axis=1), a, c, tol=0, what="concat1")
    check(lambda a, c: onp.hstack((a, c, onp.ones((4, 1)))), a, c, tol=0, what="hstack with ones")
    check(lambda a, b: onp.vstack((a, b)), a, b, tol=0, what="vstack")
    check(lambda a: onp.stack([a, a * 2], axis=1), a, tol=0, what="stack")
    check(lambda a: onp.concatenate([a.astype(onp.float32), a]), a, tol=0, what="concat promotes")
    check(lambda a: onp.repeat(a[:1], 7, axis=0), a, tol=0, what="repeat row")
    check(lambda a: onp.repeat(a, 3, axis=1), a, tol=0, what="repeat general")
    check(lambda a: onp.repeat(a, 2), a, tol=0, what="repeat flat")
    check(lambda a: onp.tile(a, (2, 3)), a, tol=0, what="tile")
    check(lambda a: onp.pad(a, ((1, 2), (0, 3)), mode="constant"), a, tol=0, what="pad")
    check(lambda a: onp.concatenate([onp.ravel(a), onp.ravel(a.T)]), a, tol=0, what="flatten-style concat")


def case_indexing():
    r = rs(11)
    f = r.randn(12, 9)
    i = r.randint(0, 12, 200)
    j = r.randint(0, 9, 200)
    check(lambda f, i, j: f[i, j], f, i, j, tol=0, what="pair gather")
    check(lambda f, i: f[i], f, i, tol=0, what="row gather")
    check(lambda f, i: f[i - 12], f, i, tol=0, what="negative indices")
    check(lambda f, i, j: f[i.reshape(10, 20), j.reshape(10, 20)], f, i, j, tol=0, what="2-D index arrays")
    check(lambda f: f[[0, 0, 3]], f, tol=0, what="list index with duplicates")
    check(lambda f, i: onp.take(f, i, axis=0), f, i, tol=0, what="take axis 0")
    check(lambda f, j: onp.take(f, j, axis=1), f, j, tol=0, what="take axis 1")
    v = r.randn(200)

    def scatter(f, i, j, v):
        out = onp.zeros(f.shape) if isinstance(f, onp.ndarray) else b200.array.zeros(f.shape)
        onp.add.at(out, (i, j), v)
        return out

    check(scatter, f, i, j, v, tol=TOL_REDUCE, what="add.at with duplicates")

    def scatter_rows(f, i):
        out = onp.zeros(f.shape) if isinstance(f, onp.ndarray) else b200.array.zeros(f.shape)
        onp.add.at(out, i, f[i] * 0.5)
        return out

    check(scatter_rows, f, i, tol=TOL_REDUCE, what="add.at rows")

    def scatter_slice(f):
        out = onp.zeros(f.shape) if isinstance(f, onp.ndarray) else b200.array.zeros(f.shape)
        onp.add.at(out, (slice(2, 8), slice(None, None, 2)), f[2:8, ::2])
        onp.add.at(out, 3, f[3])
        return out

    check(scatter_slice, f, tol=0, what="add.at slices")

    def setitem(f):
        out = f * 1.0
        out[1:3] = 7.0
        out[:, 0] = f[:, 1]
        return out

    check(setitem, f, tol=0, what="setitem")


# =============================================================== reductions
def case_reductions():
    r = rs(12)
    a = r.randn(13, 7, 11)
    for axis in (None, 0, 1, 2, -1, (0, 1), (1, 2), (0, 2), (0, 1, 2)):
        for kd in (False, True):
            check(lambda a: onp.sum(a, axis=axis, keepdims=kd), a, tol=TOL_REDUCE, what=f"sum {axis} {kd}")
            check(lambda a: onp.max(a, axis=axis, keepdims=kd), a, tol=0, what=f"max {axis} {kd}")
            check(lambda a: onp.min(a, axis=axis, keepdims=kd), a, tol=0, what=f"min {axis} {kd}")
            check(lambda a: onp.mean(a, axis=axis, keepdims=kd), a, tol=TOL_REDUCE, what=f"mean {axis} {kd}")
    check(lambda a: onp.prod(a[:, :3, :2], axis=1), a, tol=TOL_REDUCE, what="prod")
    check(lambda a: (onp.var(a, axis=1), onp.std(a, axis=(0, 2), ddof=1)), a, tol=TOL_REDUCE, what="var/std")
    check(lambda a: onp.sum(a > 0.3, axis=1), a, tol=0, what="sum of bool -> int64")
    check(lambda a: onp.sum((a * 10).astype(int)), a, tol=0, what="int sum")
    check(lambda a: (onp.all(a > -10), onp.any(a > 10), onp.any(a > 1, axis=2)), a, tol=0, what="all/any")
    check(lambda a: a.sum(axis=0).max(), a, tol=TOL_REDUCE, what="methods")
    check(lambda a: onp.sum(a.astype(onp.float32), axis=1), a, tol=TOL_F32, what="f32 sum")
    check(lambda a: (onp.argmax(a, axis=1), onp.argmin(a, axis=0), onp.argmax(a)), a, tol=0, what="argmax")
    t = onp.array([[1.0, 3.0, 3.0, 2.0], [5.0, 5.0, 1.0, 5.0]])
    check(lambda t: (onp.argmax(t, axis=1), onp.max(t, axis=1)), t, tol=0, what="argmax ties: first wins")


def case_big_reductions():
    r = rs(13)
    v = r.randn(1 << 20)
    check(lambda v: onp.sum(v), v, tol=TOL_REDUCE, what="sum 1M")
    check(lambda v: onp.max(v), v, tol=0, what="max 1M")
    m = r.randn(256, 200)
    check(lambda m: onp.sum(m, axis=0), m, tol=TOL_REDUCE, what="column sum (bias grad)")
    check(lambda m: onp.sum(m, axis=1), m, tol=TOL_REDUCE, what="row sum")
    w = r.randn(33, 6, 12, 2, 12, 2)
    check(lambda w: onp.max(onp.max(w, axis=3), axis=4), w, tol=0, what="maxpool-style double max")
    w2 = r.randn(300, 6, 12, 2, 12, 2)
    check(lambda w: (onp.max(w, axis=3), onp.sum(w > 0.5, axis=3, keepdims=True), onp.sum(w, axis=(3, 5))), w2,
          tol=TOL_REDUCE, what="short reduced axis, many outputs (pooling windows)")
    check(lambda m: onp.sum(m * m), m, tol=TOL_REDUCE, what="sum of squares")
    check(lambda v: onp.dot(v, v), v, tol=TOL_REDUCE, what="dot 1-D 1M")


def case_logsumexp():
    from scipy.special import logsumexp as sp_lse

    r = rs(14)
    a = r.randn(64, 10) * 5

    def lse(x, **kw):
        return sp_lse(x, **kw) if isinstance(x, onp.ndarray) else b200.array.logsumexp(x, **kw)

    check(lambda a: lse(a, axis=1, keepdims=True), a, tol=TOL_REDUCE, what="lse rows keepdims")
    check(lambda a: lse(a, axis=0), a, tol=TOL_REDUCE, what="lse cols")
    check(lambda a: lse(a), a, tol=TOL_REDUCE, what="lse all")
    b = r.randn(5, 128, 3)
    check(lambda b: lse(b, axis=1), b, tol=TOL_REDUCE, what="lse middle axis")
    c = onp.array([[1000.0, 1000.0, -1000.0], [-onp.inf, -onp.inf, -onp.inf], [0.0, onp.inf, 1.0], [3.0, 3.0, 3.0]])
    check(lambda c: lse(c, axis=1), c, tol=TOL_REDUCE, what="lse extremes / ties / -inf rows")
    check(lambda a: lse(a.astype(onp.float32), axis=1), a, tol=TOL_F32, what="lse f32")


# =============================================================== contractions
def case_gemm():
    r = rs(15)
    for (m, k, n) in ((1, 1, 1), (7, 5, 3), (64, 64, 64), (65, 33, 129), (256, 784, 200), (130, 17, 260)):
        a, b = r.randn(m, k), r.randn(k, n)
        check(lambda a, b: onp.dot(a, b), a, b, tol=TOL_REDUCE, what=f"dot {m}x{k}x{n}")
    a, b = r.randn(70, 90), r.randn(70, 50)
    check(lambda a, b: onp.dot(a.T, b), a, b, tol=TOL_REDUCE, what="A^T B (weight grad)")
    check(lambda a, b: onp.dot(b, b.T), a, b, tol=TOL_REDUCE, what="B B^T")
    check(lambda a, b: onp.tensordot(a, b, [[0], [0]]), a, b, tol=TOL_REDUCE, what="tensordot axes0")
    check(lambda a, b: onp.dot(a[:, ::2].T, b[:, 1::2]), a, b, tol=TOL_REDUCE, what="strided operands")
    v = r.randn(90)
    check(lambda a, v: (onp.dot(a, v), onp.dot(v, a.T)), a, v, tol=TOL_REDUCE, what="mat-vec / vec-mat")
    x64, w32 = r.randn(40, 29), r.randn(29, 16).astype(onp.float32)
    check(lambda x, w: onp.dot(x, w), x64, w32, tol=TOL_REDUCE, what="f64 @ f32 -> f64 (LSTM promotion)")
    check(lambda x, w: onp.dot(x.astype(onp.float32), w), x64, w32, tol=TOL_F32, what="f32 @ f32")
    t3, t4 = r.randn(4, 5, 6), r.randn(6, 5, 7)
    check(lambda a, b: onp.tensordot(a, b, [[1, 2], [1, 0]]), t3, t4, tol=TOL_REDUCE, what="tensordot 2 axes")
    check(lambda a, b: onp.tensordot(a, b, 1), t3, t4, tol=TOL_REDUCE, what="tensordot int axes")
    check(lambda a, b: onp.dot(a, b), t3, t4[:, :, 0].T.copy()[:, :6].T if False else r.randn(6, 3), tol=TOL_REDUCE,
          what="3-D dot 2-D")
    b3a, b3b = r.randn(5, 8, 9), r.randn(5, 9, 4)
    check(lambda a, b: onp.matmul(a, b), b3a, b3b, tol=TOL_REDUCE, what="batched matmul")
    check(lambda a, b: a @ b[0], b3a, b3b, tol=TOL_REDUCE, what="matmul broadcast batch")
    check(lambda a, b: (onp.inner(a, a), onp.outer(a[0], b[0])), r.randn(3, 4), r.randn(2, 5), tol=TOL_REDUCE,
          what="inner/outer")


def case_einsum():
    r = rs(50)
    a, b, c = r.randn(5, 7), r.randn(7, 4), r.randn(4, 6)
    t3, t4 = r.randn(3, 5, 7), r.randn(3, 7, 2)
    check(lambda a, b: onp.einsum("ij,jk->ik", a, b), a, b, tol=TOL_REDUCE, what="einsum matmul")
    check(lambda a, b: onp.einsum("ij,jk", a, b), a, b, tol=TOL_REDUCE, what="einsum implicit output")
    check(lambda a, b, c: onp.einsum("ij,jk,kl->il", a, b, c), a, b, c, tol=TOL_REDUCE, what="einsum 3 operands")
    check(lambda x, y: onp.einsum("bij,bjk->bik", x, y), t3, t4, tol=TOL_REDUCE, what="einsum batched")
    check(lambda x, y: onp.einsum("bij,bjk->kb", x, y), t3, t4, tol=TOL_REDUCE, what="einsum batched + sum + transpose")
    check(lambda a: onp.einsum("ij->ji", a), a, tol=0, what="einsum transpose")
    check(lambda a: onp.einsum("ij->", a), a, tol=TOL_REDUCE, what="einsum total sum")
    check(lambda a, b: onp.einsum("ij,jk->", a, b), a, b, tol=TOL_REDUCE, what="einsum full contraction")
    check(lambda x, a: onp.einsum("...ij,kj->...ik", x, a), t3, a, tol=TOL_REDUCE, what="einsum ellipsis")
    check(lambda a: onp.einsum("ij,ij->i", a, a), a, tol=TOL_REDUCE, what="einsum row dots")
    check(lambda a, v: onp.einsum("ij,j", a, v), a, r.randn(7), tol=TOL_REDUCE, what="einsum mat-vec")
    check(lambda a, b: onp.einsum(a, [0, 1], b, [1, 2], [0, 2]), a, b, tol=TOL_REDUCE, what="einsum sublist form")


def case_conv2d():
    from autograd import grad
    from autograd.scipy.signal import convolve

    import autograd.numpy as np

    r = rs(16)
    shapes = [((5, 3, 12, 10), (3, 4, 5, 3)), ((9, 1, 28, 28), (1, 6, 5, 5)), ((17, 6, 12, 12), (6, 16, 5, 5)),
              ((3, 2, 7, 9), (2, 8, 3, 3)), ((4, 5, 6, 6), (5, 3, 1, 1)), ((2, 3, 9, 8), (3, 20, 3, 2)),
              ((33, 4, 11, 13), (4, 9, 4, 6)), ((1, 1, 5, 5), (1, 1, 5, 5)),
              # batches large enough for the lanes-are-images input-gradient kernel (conv_direct.cu: full correlation with
              # exact skipping of the padding taps): LeNet's conv2 (its VJP wrt A), and a direct 'full' call with a
              # channel count that is not a multiple of the staging chunk and a ragged last group of images
              ((300, 6, 12, 12), (6, 16, 5, 5)), ((290, 5, 8, 8), (5, 6, 3, 3))]
    for a_shape, b_shape in shapes:
        A, B = r.randn(*a_shape), r.randn(*b_shape)
        for mode in ("valid", "full"):
            check(lambda A, B: convolve(A, B, axes=([2, 3], [2, 3]), dot_axes=([1], [0]), mode=mode), A, B,
                  tol=TOL_REDUCE, what=f"convolve {mode} {a_shape}*{b_shape}")
        # both VJPs (input-gradient = full correlation, weight-gradient = batch-summed valid correlation)
        Wt = r.randn(a_shape[0], b_shape[1], a_shape[2] - b_shape[2] + 1, a_shape[3] - b_shape[3] + 1)
        loss = lambda A, B, Wt: np.sum(convolve(A, B, axes=([2, 3], [2, 3]), dot_axes=([1], [0]), mode="valid") * Wt)
        for argnum in (0, 1):
            check_grad(loss, (A, B, Wt), TOL_REDUCE, f"convolve vjp arg{argnum} {a_shape}*{b_shape}", argnum)


# =============================================================== registries exposed to the test modules
OP_CASES = {
    "nan_inf": case_nan_inf_propagation,
    "broadcasting": case_broadcasting,
    "dtype_promotion": case_dtype_promotion,
    "integer_semantics": case_integer_semantics,
    "power_special": case_power_special_exponents,
    "fused_chain": case_fused_chain_matches_unfused,
    "vector_tail": case_large_vector_tail,
    "views": case_views,
    "roll": case_roll,
    "tri_diag": case_tri_diag,
    "row_mode_layouts": case_row_mode_layouts,
    "concat_repeat_pad": case_concat_repeat_pad,
    "indexing": case_indexing,
    "reductions": case_reductions,
    "big_reductions": case_big_reductions,
    "logsumexp": case_logsumexp,
    "gemm": case_gemm,
    "einsum": case_einsum,
    "conv2d": case_conv2d,
}


# =============================================================== autograd-level cases (grad through the device type)
def _ref(fn, *args, **kw):
    """Run `fn` on the reference path (autograd on NumPy) inside this process."""
    with b200.host_arrays(), onp.errstate(all="ignore"):
        return fn(*args, **kw)


def check_grad(make_fn, host_args, tol, what, argnum=0):
    """grad(f)(*args) on DeviceArrays vs the same call on ndarrays through unmodified autograd."""
    from autograd import grad

    want = _ref(grad(make_fn, argnum), *host_args)
    got = to_host(grad(make_fn, argnum)(*to_dev(list(host_args))))
    for i, (g, w) in enumerate(zip(leaves(got), leaves(want))):
        assert_close(g, w, tol, f"{what} leaf {i}")


def grad_case_mlp(batch=32, sizes=(20, 16, 12, 10)):
    from examples import workloads as w

    params, X, T = w.mlp_data(batch=batch, layer_sizes=sizes)
    check_grad(w.mlp_objective, (params, X, T, 1.0), TOL_REDUCE, "C1 mlp")


def grad_case_tanh(n=1001, max_order=4):
    from examples import workloads as w

    x = onp.linspace(-7, 7, n)
    for order in range(max_order + 1):
        f = w.tanh_nth_derivative(order)
        want = _ref(f, x)
        got = f(b200.asarray(x)).get()
        assert_close(got, want, TOL_ELEMENTWISE, f"C2 tanh derivative order {order}")


def grad_case_lstm(T=3, B=8, H=16, C=12):
    from examples import workloads as w

    params, X, Y = w.lstm_data(seq_len=T, batch=B, hidden=H, chars=C)
    check_grad(w.lstm_objective, (params, X, Y), TOL_F32, "C3 lstm")
    # the reference computes activations in float64 (NumPy promotion through np.ones, rnn.py:21) and
    # hands float32 gradients back (numpy_vjps.py:624,638): dtype is part of the contract
    from autograd import grad

    g = grad(w.lstm_objective)(to_dev(params), b200.asarray(X), b200.asarray(Y))
    assert all(v.dtype == onp.float32 for v in g.values())


def grad_case_fluid(rows=16, cols=16, steps=2):
    from autograd import value_and_grad

    from examples import workloads as w

    p, s0, tg = w.fluid_data(rows, cols)
    vw, gw = _ref(value_and_grad(w.fluid_objective), p, s0, tg, rows, cols, steps)
    vg, gg = value_and_grad(w.fluid_objective)(b200.asarray(p), b200.asarray(s0), b200.asarray(tg), rows, cols, steps)
    assert_close(vg.get(), onp.asarray(vw), TOL_REDUCE, "C4 loss")
    assert_close(gg.get(), gw, TOL_REDUCE, "C4 grad")
    # forward smoke field must be bit-identical (every forward op is one correctly rounded IEEE op)
    sw = _ref(w.fluid_simulate, p[: rows * cols].reshape(rows, cols), p[rows * cols:].reshape(rows, cols), s0, steps)
    dp = b200.asarray(p)
    sg = w.fluid_simulate(dp[: rows * cols].reshape(rows, cols), dp[rows * cols:].reshape(rows, cols),
                          b200.asarray(s0), steps)
    assert_close(sg.get(), sw, 0, "C4 forward bit-exact")


def grad_case_convnet(batch=6):
    from examples import workloads as w

    n, _pred, loss = w.convnet_make()
    W, X, T = w.convnet_data(batch, n)
    check_grad(loss, (W, X, T), TOL_REDUCE, "C5 convnet")


def grad_case_operators():
    """make_vjp / jacobian / value_and_grad / elementwise_grad / nesting on small functions
    (mirrors reference tests/test_wrappers.py, tests/test_jacobian.py)."""
    import autograd.numpy as np
    from autograd import elementwise_grad, grad, jacobian, make_vjp, value_and_grad

    r = rs(20)
    x = r.randn(7)
    A = r.randn(5, 7)

    f = lambda x: np.sum(np.sin(x) * x ** 2) + np.dot(x, x)
    check_grad(f, (x,), TOL_REDUCE, "grad sin*x^2")
    want = _ref(grad(grad(f)), x) if False else None  # grad of non-scalar is invalid; use hessian-vector below
    hv = lambda x, v: np.dot(grad(f)(x), v)
    v = r.randn(7)
    check_grad(hv, (x, v), TOL_REDUCE, "hessian-vector product (nested grad)")

    g = lambda x: np.tanh(np.dot(A_, x))
    for A_ in (A,):
        want = _ref(jacobian(lambda x: np.tanh(np.dot(A_, x))), x)
        Ad = b200.asarray(A_)
        got = jacobian(lambda x: np.tanh(np.dot(Ad, x)))(b200.asarray(x)).get()
        assert_close(got, want, TOL_REDUCE, "jacobian")

    vw, gw = _ref(value_and_grad(f), x)
    vg, gg = value_and_grad(f)(b200.asarray(x))
    assert_close(vg.get(), onp.asarray(vw), TOL_REDUCE, "value_and_grad value")
    assert_close(gg.get(), gw, TOL_REDUCE, "value_and_grad grad")

    vjp_w, ans_w = _ref(make_vjp(lambda x: np.exp(x) / (1 + x ** 2)), x)
    vjp_g, ans_g = make_vjp(lambda x: np.exp(x) / (1 + x ** 2))(b200.asarray(x))
    assert_close(ans_g.get(), ans_w, TOL_ELEMENTWISE, "make_vjp value")
    assert_close(vjp_g(b200.asarray(v)).get(), _ref(vjp_w, v), TOL_ELEMENTWISE, "make_vjp cotangent")

    h = lambda x: np.log1p(x ** 2) * np.cos(x)
    e3 = elementwise_grad(elementwise_grad(elementwise_grad(h)))
    assert_close(e3(b200.asarray(x)).get(), _ref(e3, x), 1e-11, "3x nested elementwise_grad")


def grad_case_vjp_rules():
    """VJPs of the op families on the hot path vs the reference rules (numpy_vjps.py) on NumPy."""
    import autograd.numpy as np

    r = rs(21)
    a, b, c = r.randn(6, 5), r.randn(5), r.randn(6, 1)
    pos = onp.abs(r.randn(6, 5)) + 0.5
    fns = {
        "broadcast add/mul": (lambda a, b, c: np.sum((a + b) * c - a / (b * b + 1.0)), (a, b, c)),
        "power": (lambda p, b: np.sum(p ** b + p ** 2.0 + 2.0 ** p), (pos, b)),
        "unary chain": (lambda a: np.sum(np.exp(np.tanh(a)) + np.log(a * a + 1) - np.sqrt(a * a + 2) + np.abs(a)), (a,)),
        "maximum/minimum": (lambda a, c: np.sum(np.maximum(a, c) * np.minimum(a, 0.1)), (a, c)),
        "sum axis": (lambda a: np.sum(np.sum(a, axis=0) ** 2), (a,)),
        "mean keepdims": (lambda a: np.sum((a - np.mean(a, axis=1, keepdims=True)) ** 2), (a,)),
        "max": (lambda a: np.sum(np.max(a, axis=1) * np.array([1., 2., 3., 4., 5., 6.])), (a,)),
        "max ties": (lambda a: np.max(np.round(a)), (a,)),
        "roll": (lambda a: np.sum(np.roll(a, 1, axis=0) * a + np.roll(a, -2, axis=1) ** 2), (a,)),
        "reshape/transpose": (lambda a: np.sum(np.reshape(a.T, (6, 5)) * a), (a,)),
        "concatenate": (lambda a, c: np.sum(np.concatenate([a, c], axis=1) ** 2 * np.arange(6.0)), (a, c)),
        "getitem slice": (lambda a: np.sum(a[1:4, ::2] ** 2) + np.sum(a[2]), (a,)),
        "getitem fancy": (lambda a: np.sum(a[onp.array([0, 0, 3, 5]), onp.array([1, 1, 2, 4])] ** 2), (a,)),
        "where": (lambda a: np.sum(np.where(a > 0, a * a, -a)), (a,)),
        "dot": (lambda a, b: np.sum(np.dot(a, b) ** 2), (a, b)),
        "dot T": (lambda a: np.sum(np.dot(a.T, a)), (a,)),
        "tensordot": (lambda a: np.sum(np.tensordot(a, a, [[0], [0]]) ** 2), (a,)),
        "matmul": (lambda a: np.sum((a @ a.T) ** 2), (a,)),
        "repeat": (lambda c: np.sum(np.repeat(c, 4, axis=1) * np.arange(4.0)), (c,)),
        "einsum": (lambda a, b: np.sum(np.einsum("ij,j->i", a, b) ** 2) + np.sum(np.einsum("ij,kj->ik", a, a)), (a, b)),
        "einsum naked sum": (lambda a: np.sum(np.einsum("ij->j", a) ** 2), (a,)),
        "astype f32": (lambda a: np.sum(a.astype(onp.float32) ** 2), (a,)),
        "var/std": (lambda a: np.sum(np.var(a, axis=0)) + np.sum(np.std(a, axis=1)), (a,)),
        "triu/tril/diag/trace": (lambda a: np.sum(np.triu(a) ** 2) + np.sum(np.tril(a, -1) * 3) + np.sum(np.diag(a[:5, :5]) ** 3)
                                 + np.trace(a[:5, :5]) ** 2, (a,)),
        "diff": (lambda a: np.sum(np.diff(a, axis=0) ** 2), (a,)),
    }
    for name, (fn, args) in fns.items():
        for argnum in range(len(args)):
            check_grad(fn, args, TOL_REDUCE, f"vjp {name} arg{argnum}", argnum)


def grad_case_check_grads():
    """The reference's own finite-difference checker (autograd/test_util.py:60-78) driven through the
    device VSpace: modes rev+fwd, order 2."""
    import autograd.numpy as np
    from autograd.test_util import check_grads

    r = rs(22)
    x = b200.asarray(r.randn(4, 3))
    y = b200.asarray(r.randn(3))
    check_grads(lambda x: np.tanh(x) * x, modes=["rev"], order=2)(x)
    check_grads(lambda x, y: np.sum(np.dot(x, y) ** 2) + np.sum(x * y), modes=["rev"], order=2)(x, y)
    check_grads(lambda x: np.sum(x, axis=0) / (1 + np.max(x, axis=0) ** 2), modes=["rev"], order=2)(x)
    check_grads(lambda x: np.roll(x, 1, axis=0) * x, modes=["rev", "fwd"], order=2)(x)
    check_grads(lambda x: np.exp(x) / (1.0 + np.exp(x)), modes=["fwd"], order=2)(x)


def grad_case_vspace_axioms():
    """VSpace contract of reference tests/test_vspaces.py:10-131 on DeviceVSpace."""
    from autograd.extend import vspace

    r = rs(23)
    x = b200.asarray(r.randn(3, 2))
    vs = vspace(x)
    assert vs.size == 6 and vs.shape == (3, 2) and vs.dtype == onp.float64 and vs.ndim == 2
    assert vs == vspace(b200.asarray(r.randn(3, 2)))
    y, z = vs.randn(), vs.randn()
    assert vspace(y) == vs and vspace(vs.zeros()) == vs and vspace(vs.ones()) == vs
    assert_close(vs.add(y, z).get(), y.get() + z.get(), 0, "add")
    assert_close(vs.scalar_mul(y, 2.5).get(), y.get() * 2.5, 0, "scalar_mul")
    ip = float(vs.inner_prod(y, z))
    assert abs(ip - float(onp.sum(y.get() * z.get()))) < 1e-12
    assert abs(ip - float(vs.inner_prod(z, y))) < 1e-12
    basis = list(vs.standard_basis())
    assert len(basis) == vs.size
    for i, bi in enumerate(basis):
        for j, bj in enumerate(basis):
            assert float(vs.inner_prod(bi, bj)) == (1.0 if i == j else 0.0)
    acc = vs.zeros()
    acc = vs.mut_add(acc, y)
    acc = vs.mut_add(acc, z)
    assert_close(acc.get(), y.get() + z.get(), 0, "mut_add")


GRAD_CASES = {
    "c1_mlp": grad_case_mlp,
    "c2_tanh": grad_case_tanh,
    "c3_lstm": grad_case_lstm,
    "c4_fluid": grad_case_fluid,
    "c5_convnet": grad_case_convnet,
    "operators": grad_case_operators,
    "vjp_rules": grad_case_vjp_rules,
    "check_grads": grad_case_check_grads,
    "vspace_axioms": grad_case_vspace_axioms,
}


def case_gemm_split_k():
    """Few output tiles, long K: the split-K path (FP64 atomics) incl. f32 outputs and batched operands."""
    r = rs(40)
    a, b = r.randn(96, 4100), r.randn(4100, 40)
    check(lambda a, b: onp.dot(a, b), a, b, tol=TOL_REDUCE, what="split-K 96x40x4100")
    check(lambda a, b: onp.dot(a.T[:1000].T.T, b[:, :1].T.T) if False else onp.dot(a[:, :1000], b[:1000]), a, b,
          tol=TOL_REDUCE, what="split-K strided A")
    check(lambda a, b: onp.dot(a.astype(onp.float32), b.astype(onp.float32)), a, b, tol=TOL_F32, what="split-K f32 out")
    g, x = r.randn(3000, 24), r.randn(3000, 50)
    check(lambda g, x: onp.dot(x.T, g), g, x, tol=TOL_REDUCE, what="weight-gradient shape (A^T G)")
    ba, bb = r.randn(3, 20, 1500), r.randn(3, 1500, 30)
    check(lambda a, b: onp.matmul(a, b), ba, bb, tol=TOL_REDUCE, what="batched split-K")


OP_CASES["gemm_split_k"] = case_gemm_split_k


def grad_case_systematic_ufuncs():
    """Finite-difference checks (reference autograd/test_util.check_grads, rev + fwd mode, 2nd order) of the
    elementwise rule table on the device type — the analogue of reference tests/test_systematic.py:51-243."""
    import autograd.numpy as np
    from autograd.test_util import check_grads

    r = rs(60)
    x = b200.asarray(r.randn(3, 4))
    pos = b200.asarray(onp.abs(r.randn(3, 4)) + 0.5)
    unit = b200.asarray(r.rand(3, 4) * 1.6 - 0.8)
    y = b200.asarray(r.randn(4) + 3.0)
    unary_any = [np.negative, np.exp, np.tanh, np.sinh, np.cosh, np.sin, np.cos, np.square, np.arctan, np.arcsinh,
                 np.expm1, np.exp2, np.abs, lambda v: v ** 3, lambda v: np.where(v > 0, v, 0.5 * v)]
    unary_pos = [np.log, np.log1p, np.log2, np.log10, np.sqrt, np.reciprocal, lambda v: v ** 0.5, lambda v: v ** -1.5,
                 lambda v: 2.0 ** v, lambda v: np.mean(v, axis=0) * np.max(v, axis=1, keepdims=True)[:3, 0].sum()]
    unary_unit = [np.arcsin, np.arccos, np.arctanh, np.tan]
    for f in unary_any:
        check_grads(f, modes=["rev", "fwd"], order=2)(x)
    for f in unary_pos:
        check_grads(f, modes=["rev", "fwd"], order=2)(pos)
    for f in unary_unit:
        check_grads(f, modes=["rev", "fwd"], order=2)(unit)
    for f in (np.hypot, np.logaddexp):  # the reference defines no JVP for these (numpy_jvps.py): reverse mode only
        check_grads(f, modes=["rev"], order=2)(pos, y)
    binary = [np.add, np.subtract, np.multiply, np.true_divide, np.maximum, np.minimum, np.arctan2,
              lambda a, b: a ** 2.0 * b, lambda a, b: np.dot(a, b), lambda a, b: np.sum(a * b, axis=1),
              lambda a, b: np.einsum("ij,j->i", a, b), lambda a, b: np.tensordot(a, b, [[1], [0]])]
    for f in binary:
        check_grads(f, modes=["rev", "fwd"], order=2)(pos, y)
    check_grads(np.power, modes=["rev", "fwd"], order=2)(pos, y / 3.0)
    shape_fns = [lambda v: np.reshape(v, (4, 3)), lambda v: np.ravel(v), lambda v: v.T, lambda v: np.swapaxes(v, 0, 1),
                 lambda v: np.roll(v, 2, axis=1), lambda v: np.concatenate([v, 2 * v], axis=0), lambda v: v[1:, ::2],
                 lambda v: np.repeat(v, 2, axis=0), lambda v: np.expand_dims(v, 1), lambda v: np.sum(v, axis=(0, 1)),
                 lambda v: np.pad(v, ((1, 0), (0, 2)), mode="constant"), lambda v: np.tile(v, (2, 1)),
                 lambda v: np.hstack((v, v * v)), lambda v: np.broadcast_to(v[:1], (5, 4)), lambda v: np.clip(v, -0.5, 0.5)]
    for f in shape_fns:
        check_grads(f, modes=["rev", "fwd"], order=2)(x)


GRAD_CASES["systematic_ufuncs"] = grad_case_systematic_ufuncs


def grad_case_optimizers():
    """autograd.misc.optimizers on device parameters — the Adam/SGD/RMSprop updates are plain
    elementwise expressions, so each iteration's update is fused into the kernels that consume it."""
    from autograd.misc.optimizers import adam, rmsprop, sgd

    from examples import workloads as w

    params, X, T = w.mlp_data(batch=16, layer_sizes=(12, 10, 8, 5))

    def run(opt, to_dev_fn, **kw):
        p0 = to_dev_fn(params)
        Xd, Td = to_dev_fn(X), to_dev_fn(T)
        from autograd import grad

        g = grad(lambda p, i: w.mlp_objective(p, Xd, Td, 0.1))
        return opt(g, p0, num_iters=5, **kw)

    for opt, kw in ((adam, {"step_size": 0.01}), (sgd, {"step_size": 0.001, "mass": 0.9}), (rmsprop, {"step_size": 0.01})):
        want = _ref(run, opt, lambda t: t, **kw)
        got = to_host(run(opt, to_dev, **kw))
        for a, r in zip(leaves(got), leaves(want)):
            assert_close(a, r, 1e-9, f"{opt.__name__} after 5 iterations")


def grad_case_forward_mode_operators():
    """make_jvp / deriv / hessian-vector / Gauss-Newton products on the device type."""
    import autograd.numpy as np
    from autograd import deriv, hessian_tensor_product, make_ggnvp, make_jvp, tensor_jacobian_product
    from autograd.differential_operators import make_jvp_reversemode

    r = rs(70)
    x, v = r.randn(6), r.randn(6)
    A = r.randn(4, 6)
    f = lambda x: np.tanh(np.dot(A_, x)) * np.sum(x ** 2)

    def both(fn):
        global A_
        return None

    def eval_pair(builder):
        """builder(A, x, v) -> value; evaluated on host (reference) and on device."""
        want = _ref(builder, A, x, v)
        got = builder(b200.asarray(A), b200.asarray(x), b200.asarray(v))
        return to_host(got), want

    def jvp_case(A, x, v):
        val, tangent = make_jvp(lambda x: np.tanh(np.dot(A, x)) * np.sum(x ** 2))(x)(v)
        return [val, tangent]

    def jvp_rev_case(A, x, v):
        return make_jvp_reversemode(lambda x: np.tanh(np.dot(A, x)))(x)(v)

    def hvp_case(A, x, v):
        return hessian_tensor_product(lambda x: np.sum(np.tanh(np.dot(A, x)) ** 2))(x, v)

    def ggnvp_case(A, x, v):
        return make_ggnvp(lambda x: np.tanh(np.dot(A, x)))(x)(v)

    def tjp_case(A, x, v):
        return tensor_jacobian_product(lambda x: np.tanh(np.dot(A, x)))(x, np.dot(A, v))

    def deriv_case(A, x, v):
        return deriv(lambda t: np.sum(np.sin(t * x)))(0.7)

    for name, case in (("make_jvp", jvp_case), ("make_jvp_reversemode", jvp_rev_case), ("hessian_tensor_product", hvp_case),
                       ("make_ggnvp", ggnvp_case), ("tensor_jacobian_product", tjp_case), ("deriv", deriv_case)):
        got, want = eval_pair(case)
        for a, r_ in zip(leaves(got), leaves(want)):
            assert_close(onp.asarray(a), onp.asarray(r_), TOL_REDUCE, name)


GRAD_CASES["optimizers"] = grad_case_optimizers
GRAD_CASES["forward_mode_operators"] = grad_case_forward_mode_operators


def grad_case_primitives_imported_before_install():
    """`from autograd.scipy.special import logsumexp` / `...signal import convolve` done at module import time (as the
    reference examples do, neural_net.py:12, convnet.py:11) must reach the device too: install() swaps the raw
    function inside the SAME primitive objects."""
    import autograd.scipy.signal as asignal
    import autograd.scipy.special as aspecial
    import autograd_b200.install as inst
    from autograd import grad

    assert aspecial.logsumexp.fun.__name__ == "_logsumexp_raw" and asignal.convolve.fun.__name__ == "_convolve_raw"
    r = rs(80)
    x = r.randn(6, 5)
    f = lambda x: aspecial.logsumexp(x, axis=1).sum() + aspecial.logsumexp(x * 2.0)
    check_grad(lambda x: f(x), (x,), TOL_REDUCE, "logsumexp primitive object")


GRAD_CASES["primitives_preimported"] = grad_case_primitives_imported_before_install


def grad_case_f32_tensor_core_contractions():
    """All-float32 contractions large enough for the tcgen05 3xTF32 kernel (csrc/gemm_tc.cu): forward dot and both
    dot VJPs (G B^T and A^T G, numpy_vjps.py:615-651) of a float32 two-layer net, against float32 NumPy."""
    import autograd.numpy as np

    r = rs(70)
    X = (r.randn(640, 512) * 0.5).astype(onp.float32)
    W1 = (r.randn(512, 384) * 0.05).astype(onp.float32)
    W2 = (r.randn(384, 256) * 0.05).astype(onp.float32)

    def loss(W1, W2, X):
        h = np.tanh(np.dot(X, W1))
        return np.sum(np.dot(h, W2) ** 2) / X.shape[0]

    check(lambda a, b: onp.dot(a, b), X, W1, tol=TOL_F32, what="f32 dot on tensor cores")
    check(lambda a, b: onp.tensordot(a, b, [[0], [0]]), X, (r.randn(640, 300)).astype(onp.float32), tol=TOL_F32,
          what="f32 A^T B on tensor cores")
    for argnum in (0, 1, 2):
        check_grad(loss, (W1, W2, X), TOL_F32, f"f32 net grad wrt arg {argnum}", argnum=argnum)
    dev = to_dev([W1, W2, X])
    from autograd import grad

    assert grad(loss)(*dev).dtype == onp.float32


GRAD_CASES["f32_tensor_core_contractions"] = grad_case_f32_tensor_core_contractions


def case_scan_sort_surface():
    """cumsum/cumprod, sort/argsort (indices bit-exact), item assignment with index arrays, and the
    composed functions kron / rot90 / rollaxis / average / select / sinc / cross / gradient / deg2rad / logaddexp2."""
    r = rs(80)
    a = r.randn(7, 300)
    check(lambda a: onp.cumsum(a, axis=1), a, tol=TOL_REDUCE, what="cumsum last axis")
    check(lambda a: onp.cumsum(a, axis=0), a, tol=TOL_ELEMENTWISE, what="cumsum axis 0 (sequential, same order)")
    check(lambda a: onp.cumsum(a), a, tol=TOL_REDUCE, what="cumsum flattened")
    check(lambda a: onp.cumprod(1.0 + 0.01 * a, axis=1), a, tol=TOL_REDUCE, what="cumprod")
    big = r.randn(3, 40000)
    check(lambda a: onp.cumsum(a, axis=1), big, tol=TOL_REDUCE, what="cumsum long rows (segmented scan)")
    check(lambda a: onp.cumsum(a.astype(onp.float32), axis=1), big[:, :5000], tol=TOL_F32, what="cumsum f32")
    ints = r.randint(-50, 50, (5, 1000))
    check(lambda a: onp.cumsum(a, axis=1), ints, what="cumsum int64 exact")
    # sort / argsort: distinct keys -> indices bit-exact for any kind; ties + NaN -> compare against kind='stable'
    for shape in ((1000,), (6, 257), (3, 5000), (2, 3, 64)):
        x = r.randn(*shape)
        check(lambda x: onp.sort(x, axis=-1), x, what=f"sort {shape}")
        check(lambda x: onp.argsort(x, axis=-1), x, what=f"argsort {shape}")
    x = r.randn(50, 40)
    check(lambda x: (onp.sort(x, axis=0), onp.argsort(x, axis=0)), x, what="sort axis 0")
    check(lambda x: onp.sort(x, axis=None), x, what="sort flattened")
    t = onp.round(r.randn(4, 600) * 3.0)
    t[1, 5] = onp.nan
    t[1, 77] = onp.nan
    t[2, 0] = onp.inf
    t[2, 9] = -onp.inf
    check(lambda x: onp.argsort(x, axis=1, kind="stable"), t, what="argsort ties/NaN/inf (stable)")
    check(lambda x: onp.sort(x, axis=1), t, what="sort ties/NaN/inf")
    check(lambda x: onp.argsort(x, kind="stable"), r.randint(0, 20, 5000), what="argsort int64 stable")
    check(lambda x: onp.sort(x.astype(onp.float32)), r.randn(3000), what="sort f32")
    # item assignment with an index array (unpermuter)
    perm = r.permutation(500)
    dev = b200.array.zeros(500, dtype=onp.int64)
    dev[b200.asarray(perm)] = list(range(500))
    want = onp.zeros(500, dtype=onp.int64)
    want[perm] = list(range(500))
    assert onp.array_equal(dev.get(), want)
    m, v3, w3 = r.randn(4, 6), r.randn(5, 3), r.randn(5, 3)
    check(lambda a, b: onp.kron(a, b), m, r.randn(3, 2), what="kron")
    check(lambda a: onp.kron(a, a[0]), m, what="kron 2-D x 1-D")
    for k in range(4):
        check(lambda a: onp.rot90(a, k), m, what=f"rot90 k={k}")
    check(lambda a: onp.rollaxis(a.reshape(2, 2, 6), 2, 0), m, what="rollaxis")
    check(lambda a: (onp.average(a), onp.average(a, axis=1), onp.average(a, axis=0, weights=onp.arange(1.0, 5.0))), m,
          tol=TOL_REDUCE, what="average")
    check(lambda a: onp.select([a > 1, a < -1], [a * 2, a * a], 0.5), m, what="select")
    check(lambda a: onp.sinc(a), onp.concatenate([m.ravel(), [0.0]]), what="sinc")
    check(lambda a, b: onp.cross(a, b), v3, w3, what="cross")
    check(lambda a: onp.gradient(a, axis=1), m, what="gradient axis")
    check(lambda a: onp.gradient(a, 0.5), m[0], what="gradient 1-D spacing")
    check(lambda a: onp.gradient(a), m, what="gradient all axes")
    check(lambda a: (onp.deg2rad(a), onp.rad2deg(a), onp.radians(a), onp.degrees(a)), m, what="deg/rad")
    check(lambda a: onp.deg2rad(a.astype(onp.float32)), m, tol=TOL_F32, what="deg2rad f32")
    check(lambda a, b: onp.logaddexp2(a, b), m, m[::-1] * 2.0, what="logaddexp2")


OP_CASES["scan_sort_surface"] = case_scan_sort_surface


def grad_case_scan_sort_surface():
    """Reference VJP rules of the widened surface on the device type: cumsum (numpy_vjps.py:539-549), sort
    (:774-789, through argsort + unpermuter's index-array assignment), kron, cross, gradient, sinc, ..."""
    import autograd.numpy as np

    r = rs(81)
    a = r.randn(6, 40)
    v = r.randn(200)
    check_grad(lambda x: np.sum(np.cumsum(x, axis=1) ** 2), (a,), TOL_REDUCE, "cumsum axis 1")
    check_grad(lambda x: np.sum(np.cumsum(x, axis=0) * x), (a,), TOL_REDUCE, "cumsum axis 0")
    check_grad(lambda x: np.sum(np.cumsum(x) ** 2), (v,), TOL_REDUCE, "cumsum flat")
    check_grad(lambda x: np.sum(np.sort(x) * np.arange(200.0)), (v,), TOL_REDUCE, "sort 1-D")
    check_grad(lambda x: np.sum(np.kron(x, x[:2, :3]) ** 2), (a,), TOL_REDUCE, "kron")
    check_grad(lambda x: np.sum(np.cross(x[:, :3], x[:, 3:6]) ** 2), (a,), TOL_REDUCE, "cross")
    check_grad(lambda x: np.sum(np.gradient(x, axis=1) ** 2), (a,), TOL_REDUCE, "gradient")
    check_grad(lambda x: np.sum(np.sinc(x) * x), (a,), TOL_REDUCE, "sinc")
    check_grad(lambda x: np.sum(np.deg2rad(x) ** 2 + np.rad2deg(x)), (a,), TOL_REDUCE, "deg2rad/rad2deg")
    check_grad(lambda x: np.sum(np.logaddexp2(x, 2.0 * x)), (a,), TOL_REDUCE, "logaddexp2")
    check_grad(lambda x: np.sum(np.rot90(x)[:5, :5] * x[:5, :5]), (a,), TOL_REDUCE, "rot90")
    check_grad(lambda x: np.sum(np.rollaxis(x, 1)[3] ** 2), (a,), TOL_REDUCE, "rollaxis")
    check_grad(lambda x: np.sum(np.select([x > 0.5], [x ** 2], x)), (a,), TOL_REDUCE, "select")


GRAD_CASES["scan_sort_surface"] = grad_case_scan_sort_surface


def case_scipy_special_ufuncs():
    """scipy.special ufuncs reached through autograd.scipy.special (special.py:42-117) on the device: CUDA libm twins
    (erf family, lgamma/tgamma, normcdf, Bessel j0/j1/y0/y1/i0/i1) plus generated expit/logit/digamma/log_ndtr."""
    import scipy.special as sps

    r = rs(90)
    x = r.randn(3000) * 2.0
    pos = onp.abs(x) + 0.05
    unit = r.rand(3000) * 0.98 + 0.01
    tol = 1e-12
    for name, arg in (("erf", x), ("erfc", x), ("erfcx", x), ("ndtr", x * 2), ("log_ndtr", x * 4), ("expit", x * 5),
                      ("j0", x * 5), ("j1", x * 5), ("i0", x), ("i1", x), ("y0", pos * 5), ("y1", pos * 5),
                      ("gammaln", pos * 6), ("gamma", pos * 4), ("psi", pos * 6), ("psi", -pos * 3 - 0.013),
                      ("erfinv", unit * 2 - 1), ("erfcinv", unit * 2 * 0.99 + 0.005), ("ndtri", unit), ("logit", unit)):
        fn = getattr(sps, name)
        check(lambda a, fn=fn: fn(a), arg, tol=tol, what=f"scipy.special.{name}")
    check(lambda a: sps.erf(a.astype(onp.float32)), x, tol=TOL_F32, what="erf f32")
    edge = onp.array([0.0, -0.0, onp.inf, -onp.inf, onp.nan, 1.0, 2.0, -1.0, 1e-300, 170.0])
    check(lambda a: (sps.erf(a), sps.expit(a), sps.ndtr(a)), edge, tol=tol, what="special edge values")
    check(lambda a: sps.gammaln(a), onp.array([1.0, 2.0, 0.5, 1e-8, 100.0, 1e6, onp.inf]), tol=tol, what="gammaln edges")
    check(lambda a: sps.psi(a), onp.array([1.0, 0.5, 1e-6, 100.0, 1e8, -0.5, -2.5, onp.inf]), tol=tol, what="psi edges")


OP_CASES["scipy_special_ufuncs"] = case_scipy_special_ufuncs


def grad_case_scipy_stats_norm():
    """autograd.scipy.stats.norm on device operands (norm.py:9-82) — every function, every argument — and a
    black-box-SVI style objective (examples/black_box_svi.py:16-38: Gaussian entropy + Monte-Carlo log-density with
    host-side random draws mixed into device computations)."""
    import autograd.numpy as np
    import autograd.scipy.stats.norm as norm
    from autograd.scipy.special import erf, expit, gammaln

    r = rs(91)
    x, loc, scale = r.randn(5, 40), r.randn(40) * 0.3, onp.abs(r.randn(40)) + 0.5
    for name in ("pdf", "logpdf", "cdf", "sf", "logcdf", "logsf"):
        fn = getattr(norm, name)
        for argnum in (0, 1, 2):
            check_grad(lambda x, l, s, fn=fn: np.sum(fn(x, l, s) * np.cos(x)), (x, loc, scale), TOL_REDUCE,
                       f"norm.{name} wrt arg {argnum}", argnum=argnum)
    check_grad(lambda x: np.sum(erf(x) * expit(x) + gammaln(x ** 2 + 1.0)), (x,), TOL_REDUCE, "special-function chain")
    D, S = 40, 64
    eps = r.randn(S, D)          # the example draws its samples on the host (npr.RandomState); only params are traced

    def logprob(z):
        return norm.logpdf(z[:, 0], 0.0, 1.35) + np.sum(norm.logpdf(z[:, 1:], 0.0, np.exp(z[:, :1])), axis=1)

    def neg_elbo(params, eps):
        mean, log_std = params[:D], params[D:]
        samples = eps * np.exp(log_std) + mean
        entropy = 0.5 * D * (1.0 + np.log(2 * np.pi)) + np.sum(log_std)
        return -(entropy + np.mean(logprob(samples)))

    check_grad(neg_elbo, (onp.concatenate([r.randn(D) * 0.1, -1.0 + 0.1 * r.randn(D)]), eps), TOL_REDUCE,
               "black-box SVI objective")


GRAD_CASES["scipy_stats_norm"] = grad_case_scipy_stats_norm


def case_short_axis_reductions():
    """Reductions over short axes (pooling windows) are recorded as pending elementwise combinations of slices
    (array._reduce_short_axes / _select_lazy / _insert_axis_lazy): values must match NumPy for every op, for keepdims,
    for broadcasting operands inside the pending expression, for rolled and strided views, ints and bools."""
    r = rs(95)
    x = r.randn(40, 6, 12, 2, 12, 2)
    b = r.randn(1, 6, 1, 1, 1, 1)
    for op in (onp.max, onp.min, onp.sum, onp.prod):
        check(lambda a: op(a, axis=3), x, tol=TOL_REDUCE, what=f"{op.__name__} axis 3")
        check(lambda a: op(op(a, axis=3), axis=4), x, tol=TOL_REDUCE, what=f"{op.__name__} chained (pooling)")
        check(lambda a: op(a, axis=(3, 5)), x, tol=TOL_REDUCE, what=f"{op.__name__} two axes at once")
        check(lambda a: op(a, axis=5, keepdims=True), x, tol=TOL_REDUCE, what=f"{op.__name__} keepdims")
        check(lambda a, b: op(onp.tanh(a + b) * 2.0, axis=3), x, b, tol=TOL_REDUCE, what=f"{op.__name__} of a pending expression")
    check(lambda a, b: onp.sum((a + b) == onp.max(a + b, axis=3, keepdims=True), axis=3, keepdims=True), x, b,
          what="count of argmax locations (int64, bit-exact)")
    check(lambda a: onp.max(onp.roll(a, 1, axis=2), axis=3), x, tol=TOL_REDUCE, what="max of a rolled view")
    check(lambda a: onp.sum(onp.roll(a, 1, axis=3), axis=3), x, tol=TOL_REDUCE, what="sum over the rolled axis")
    check(lambda a: onp.max(a[:, :, ::2, :, 1:, :], axis=5), x, tol=TOL_REDUCE, what="max of a strided view")
    check(lambda a: onp.sum(a > 0, axis=5), x, what="bool sum -> int64")
    check(lambda a: onp.max((a * 10).astype(onp.int64), axis=3), x, what="int64 max")
    check(lambda a, b: onp.expand_dims(onp.exp(a + b), 2), x, b, tol=TOL_ELEMENTWISE, what="expand_dims of pending expression")
    check(lambda a, b: onp.squeeze((a + b)[:, :, :1] * 3.0, axis=2), x, b, tol=TOL_ELEMENTWISE, what="squeeze of pending expression")
    big = r.randn(200, 6, 12, 2, 12, 2)          # > 2^20 elements: flat general-layout kernels with baked divisor dims
    bb = r.randn(1, 6, 1, 1, 1, 1)
    check(lambda a, b: onp.max(onp.max(a + b, axis=3), axis=4), big, bb, tol=TOL_REDUCE, what="large chained max")
    check(lambda a, b: (a + b) * ((a + b) == onp.max(a + b, axis=5, keepdims=True)), big, bb, tol=TOL_ELEMENTWISE,
          what="large broadcast-back of a short-axis max")
    check(lambda a: onp.transpose(a, (0, 2, 4, 1, 3, 5)) * 2.0 + a.reshape(200, 12, 12, 6, 2, 2), big, tol=TOL_ELEMENTWISE,
          what="large transposed operand (flat general layout)")
    nanx = x.copy()
    nanx[0, 0, 0, 1, 0, 0] = onp.nan
    check(lambda a: onp.max(a, axis=3), nanx, tol=TOL_REDUCE, what="max propagates NaN like NumPy")


OP_CASES["short_axis_reductions"] = case_short_axis_reductions


def case_pending_reshape():
    """reshape / ravel of a PENDING expression whose operands cannot be re-viewed one by one (a broadcast row being
    flattened: ``(cell_xs - vx).ravel()`` of examples/fluidsim.py:37-41) stays pending as a ``reshape`` node;
    lazy._plan_reshapes picks one iteration space for both sides, or computes the reshaped value first when no single
    space serves every operand.  Values must match NumPy on every path."""
    r = rs(123)
    x = r.randn(64, 96)
    row, col = r.randn(96), r.randn(64, 1)
    table = r.randn(64, 96)
    check(lambda a, b: (a + b).ravel() * 2.0, x, row, tol=TOL_ELEMENTWISE, what="broadcast row flattened")
    check(lambda a, b, c: onp.floor((a - b).ravel()) + (a * c).ravel(), x, row, col, tol=TOL_ELEMENTWISE,
          what="two flattened expressions over one 2-d space")
    check(lambda a, b: ((a + b).ravel() * 2.0).reshape(96, 64) + 1.0, x, row, tol=TOL_ELEMENTWISE,
          what="reshape of a reshape (levels (96,64) <- (6144,) <- (64,96))")
    check(lambda a, b: ((a + b).ravel() * 2.0).reshape(96, 64) + onp.arange(64.0), x, row, tol=TOL_ELEMENTWISE,
          what="two levels that both need coordinates (falls back to computing the inner one first)")
    check(lambda a, b: (a + b).ravel()[None, :] * onp.array([[1.0], [2.0], [3.0]]), x, row, tol=TOL_ELEMENTWISE,
          what="reshaped value broadcast into a larger space")
    check(lambda a, b: onp.roll((a + b).ravel(), 5) - (a + b).ravel(), x, row, tol=TOL_ELEMENTWISE, what="roll of a reshaped value")
    check(lambda a, b: onp.sum((a * b).ravel()), x, row, tol=TOL_REDUCE, what="reduction of a reshaped value")
    check(lambda a, b: ((a * 10).astype(onp.int64) + onp.arange(96)).ravel() % 7, x, row, what="int64 through a reshape")
    check(lambda a, b, t: t[onp.mod(onp.floor((a * 20 + b).ravel()).astype(int), 64),
                            onp.mod(onp.floor((a * 30 - b).ravel()).astype(int), 96)] * (a + b).ravel(), x, row, table,
          tol=TOL_ELEMENTWISE, what="gather through indices computed from flattened broadcast expressions (advect)")
    check(lambda a, b: ((a + b).reshape(32, 2, 96) * 2.0).reshape(64, 96) - a, x, row, tol=TOL_ELEMENTWISE,
          what="there and back again")
    check(lambda a, b: (a + b).ravel().reshape(64, 96)[::2, 1:] * 3.0, x, row, tol=TOL_ELEMENTWISE, what="slice of a reshaped value")
    check(lambda a, b: onp.concatenate([(a + b).ravel(), (a - b).ravel()]), x, row, tol=TOL_ELEMENTWISE, what="concatenate")
    y = x.copy()

    def inplace(a, b):
        v = (a + b).ravel()
        a = a * 1.0
        a[0, 0] = 100.0              # the operand behind the pending reshape is a different buffer: v keeps its values
        return v + a.ravel()

    check(inplace, y, row, tol=TOL_ELEMENTWISE, what="pending reshape unaffected by later writes to a copy")

    def inplace_operand(a, b):
        a = a.copy()
        v = (a + b).ravel()
        a[0, 0] = 100.0              # written AFTER v was recorded: v must hold the old value (eager semantics)
        a[5:7] = -1.0
        return v * 1.0 + a.ravel()

    check(inplace_operand, y, row, tol=TOL_ELEMENTWISE, what="pending reshape reads its operand before a later in-place write")


OP_CASES["pending_reshape"] = case_pending_reshape


def case_reshape_is_a_view():
    """NumPy's reshape / ravel / expand_dims / squeeze of a contiguous array are VIEWS: a later in-place write through either
    handle is seen through the other.  A reshape of a PENDING expression is recorded symbolically (two independent
    expressions, so that it can fuse); the two become one buffer the moment either gets a payload (lazy.link_alias), and
    values that were computed - or recorded - from either before the write keep the old contents."""
    r = rs(321)
    x, row = r.randn(64, 96), r.randn(96)

    def write_through_the_reshape(a0, b0, make):
        a = make(a0, b0)
        v = a.reshape(-1)
        v[5] = 555.0
        v[200:300] = -1.0
        return a * 1.0, v * 1.0

    def write_through_the_source(a0, b0, make):
        a = make(a0, b0)
        v = a.reshape(-1)
        w = v * 2.0                     # recorded BEFORE the write: old values
        a[0, 5] = 555.0
        a[3] = -1.0
        return a * 1.0, v * 1.0, w

    def chain(a0, b0, make):
        a = make(a0, b0)
        c = onp.expand_dims(onp.ravel(a), 0).reshape(32, 192)
        c[1, 2] = 7.0
        d = onp.squeeze(a.reshape(64, 1, 96), 1)
        d[63, 95] = 9.0
        return a * 1.0, c * 1.0, d * 1.0

    def write_through_a_slice_of_the_source(a0, b0, make):
        a = make(a0, b0)
        v = a.reshape(96, 64)
        s_ = a[0:2]
        s_[...] = 7.0
        return a * 1.0, v * 1.0

    def flatten_is_a_copy(a0, b0, make):
        a = make(a0, b0)
        f = a.flatten()
        f[0] = 555.0
        return a * 1.0, f * 1.0

    for what, make in (("contiguous operands", lambda a, b: a + 1.0), ("a broadcast operand", lambda a, b: a + b),
                       ("a materialised array", lambda a, b: a.copy())):
        for fn in (write_through_the_reshape, write_through_the_source, chain, write_through_a_slice_of_the_source,
                   flatten_is_a_copy):
            check(lambda a, b: fn(a, b, make), x, row, tol=TOL_ELEMENTWISE, what=f"{fn.__name__} ({what})")


OP_CASES["reshape_is_a_view"] = case_reshape_is_a_view


def grad_case_max_pool():
    """Max-pool block of examples/convnet.py:109-118 with the grad_chooser VJP (numpy_vjps.py:515-530), ties included
    (equal maxima share the cotangent)."""
    import autograd.numpy as np

    r = rs(96)
    X = r.randn(40, 6, 24, 24)
    X[0, 0, 0, 0] = X[0, 0, 0, 1] = X[0, 0, 1, 0] = 7.0       # a tie inside one pooling window
    bias = r.randn(1, 6, 1, 1)

    def pooled(x, b):
        y6 = np.reshape(x + b, (40, 6, 12, 2, 12, 2))
        p = np.max(np.max(y6, axis=3), axis=4)
        return np.sum(np.tanh(p) ** 2 * np.arange(1.0, 13.0))

    for argnum in (0, 1):
        check_grad(pooled, (X, bias), TOL_REDUCE, f"max-pool grad wrt arg {argnum}", argnum=argnum)
    check_grad(lambda x: np.sum(np.mean(np.reshape(x, (40, 6, 12, 2, 12, 2)), axis=(3, 5)) ** 2), (X,), TOL_REDUCE,
               "average-pool grad")
    # > 2^20 elements: the window kernels (lazy.Program.unroll: one thread per 2 x 2 window, 16-byte vector accesses along
    # the innermost window dim, zero-stride terms left out of the offsets) - same values, ties and NaN-free maxima included
    Xl = r.randn(320, 6, 24, 24)
    Xl[0, 0, 0, 0] = Xl[0, 0, 0, 1] = Xl[0, 0, 1, 0] = Xl[0, 0, 1, 1] = 3.0      # a four-way tie
    Xl[5, 2, 6, 8] = Xl[5, 2, 7, 9] = 9.0                                           # a diagonal tie

    def pooled_large(x, b):
        y6 = np.reshape(x + b, (320, 6, 12, 2, 12, 2))
        p = np.max(np.max(y6, axis=3), axis=4)
        return np.sum(p ** 2 * np.arange(1.0, 13.0))

    for argnum in (0, 1):
        check_grad(pooled_large, (Xl, bias), TOL_REDUCE, f"large max-pool grad wrt arg {argnum}", argnum=argnum)
    check_grad(lambda x: np.sum(np.max(np.reshape(x, (320, 6, 24, 12, 2)), axis=4) ** 3), (Xl,), TOL_REDUCE,
               "large 1-d max-pool grad (window = innermost dim only)")
    check_grad(lambda x: np.sum(np.max(np.reshape(x, (320, 6, 12, 2, 24)), axis=3) ** 3), (Xl,), TOL_REDUCE,
               "large row-pair max-pool grad (window = a middle dim only)")


GRAD_CASES["max_pool"] = grad_case_max_pool


def grad_case_linalg_norm():
    """np.linalg.norm (vector norms, Frobenius) and its reference VJP (autograd/numpy/linalg.py:91-144) on the device."""
    import autograd.numpy as np
    import autograd.numpy.linalg as la

    r = rs(97)
    x = r.randn(6, 40)
    for o in (None, 1, 2, 3, onp.inf, -onp.inf, 0):
        check(lambda a: onp.linalg.norm(a[0], o), x, tol=TOL_REDUCE, what=f"vector norm ord={o}")
        check(lambda a: onp.linalg.norm(a, o, axis=1), x, tol=TOL_REDUCE, what=f"norm ord={o} axis=1")
    check(lambda a: (onp.linalg.norm(a), onp.linalg.norm(a, "fro"), onp.linalg.norm(a, axis=0, keepdims=True)), x,
          tol=TOL_REDUCE, what="Frobenius / keepdims")
    check_grad(lambda a: la.norm(a), (x,), TOL_REDUCE, "grad of the 2-norm")
    check_grad(lambda a: np.sum(la.norm(a, axis=1) ** 2), (x,), TOL_REDUCE, "grad of row norms")
    check_grad(lambda a: np.sum(la.norm(a, 3, axis=0)), (x,), TOL_REDUCE, "grad of the 3-norm")
    check_grad(lambda a: la.norm(a, "fro") * np.sum(a), (x,), TOL_REDUCE, "grad of the Frobenius norm")


GRAD_CASES["linalg_norm"] = grad_case_linalg_norm


# =============================================================== gaps found by running the reference's own test files
# (tests/test_systematic.py, test_numpy.py, test_graphs.py, test_wrappers.py ... of the reference) with device inputs
def case_mixed_indexing():
    """Index arrays combined with slices / integers / newaxis, and index arrays that are not on the leading axes:
    NumPy's placement rule for the broadcast index dimensions (test_numpy.py:315-321 test_index_mixed)."""
    r = rs(41)
    x = r.randn(5, 6, 4)
    cols = onp.array([1, 3])
    rows = onp.array([[0, 2], [4, 4]])
    check(lambda x: x[3, 2:, [1, 3]], x, tol=0, what="int, slice, list (separated: index dims first)")
    check(lambda x, c: x[:, c], x, cols, tol=0, what="columns")
    check(lambda x, c: x[:, :, c], x, cols, tol=0, what="last axis")
    check(lambda x, c: x[1:4, c, ::2], x, cols, tol=0, what="slice, array, strided slice")
    check(lambda x, r_, c: x[r_, :, c], x, rows, cols, tol=0, what="separated arrays broadcast")
    check(lambda x, r_, c: x[:, r_, c], x, rows, cols, tol=0, what="adjacent arrays stay in place")
    check(lambda x, c: x[..., c], x, cols, tol=0, what="ellipsis then array")
    check(lambda x, c: x[None, :, c], x, cols, tol=0, what="newaxis")
    check(lambda x, c: x[c, 1:3], x, cols, tol=0, what="leading array then partial slice")

    def scatter(x, r_, c):
        out = onp.zeros(x.shape) if isinstance(x, onp.ndarray) else b200.array.zeros(x.shape).copy()
        onp.add.at(out, (slice(None), r_, c), x[:, r_, c] * 0.5)
        onp.add.at(out, (3, slice(2, None), [1, 3]), x[3, 2:, [1, 3]])
        onp.add.at(out, (slice(None), slice(None), c), 1.5)
        return out

    check(scatter, x, rows, cols, tol=TOL_REDUCE, what="add.at with mixed indices (duplicates)")


OP_CASES["mixed_indexing"] = case_mixed_indexing


def case_reference_suite_surface():
    r = rs(42)
    a3, a2, v = r.randn(3, 4, 4), r.randn(4, 4, 3), r.randn(7)
    check(lambda a, b, c: onp.einsum("ijk,lji,lli->lki", a, b, c), a3, a2, a2, tol=TOL_REDUCE, what="einsum three operands, diagonal")
    check(lambda a: onp.einsum("ii->i", a[0]), a3, tol=0, what="einsum diagonal view")
    # a size-1 axis broadcasts against the other operand's extent (examples/bayesian_neural_net.py:31 "mnd,mdo->mno")
    check(lambda a, b: onp.einsum("mnd,mdo->mno", a[:1], b), r.randn(2, 5, 4), r.randn(3, 4, 2), tol=TOL_REDUCE, what="einsum broadcast batch")
    check(lambda a, b: onp.einsum("mnd,mdo->mno", a, b[:1]), r.randn(3, 5, 4), r.randn(2, 4, 2), tol=TOL_REDUCE, what="einsum broadcast batch rhs")
    check(lambda a: onp.einsum("ii", a[0]), a3, tol=TOL_REDUCE, what="einsum trace")
    check(lambda v: onp.atleast_3d(v), v, tol=0, what="atleast_3d 1-D")
    check(lambda a: onp.atleast_3d(a[0]), a3, tol=0, what="atleast_3d 2-D")
    check(lambda v: onp.sort(v[:1]), v, tol=0, what="sort of one element")


OP_CASES["reference_suite_surface"] = case_reference_suite_surface


def grad_case_reference_suite_gaps():
    import autograd.numpy as np
    from autograd import grad
    from autograd.core import make_jvp
    from autograd.test_util import check_grads

    r = rs(43)
    x3, x2, v = r.randn(5, 6, 4), r.randn(5, 5), r.randn(7)
    cols = onp.array([1, 3, 3])
    fns = {
        "mixed index": (lambda x: np.sum(x[3, 2:, [1, 3]] ** 2) + np.sum(x[:, cols] ** 3) + np.sum(x[1:4, cols, ::2]), (x3,)),
        "gradient all axes": (lambda x: np.sum(np.gradient(x)[0] ** 2) + np.sum(np.gradient(x)[1] ** 3), (x2,)),
        "gradient one axis": (lambda x: np.sum(np.gradient(x, axis=1) ** 2), (x2,)),
        "einsum repeated operand index": (lambda x: np.sum(np.einsum("ijk,lji,lli->lki", x[:3, :4, :4], x3[:4, :4, :3], x3[1:5, :4, :3]) ** 2), (x3,)),
        "atleast_3d": (lambda v: np.sum(np.atleast_3d(v) ** 2), (v,)),
        "sort": (lambda v: np.sum(np.sort(v) * np.arange(7.0)) + np.sum(np.sort(v[:1])), (v,)),
        "r_/c_": (lambda v: np.sum(np.r_[v, v * 2.0, 1:4] ** 2) + np.sum(np.c_[v, v * v] ** 3), (v,)),
        "make_diagonal": (lambda v: np.sum(np.make_diagonal(v, axis1=-1, axis2=-2) ** 2 * np.arange(7.0)), (v,)),
        "int cast index": (lambda x: np.sum(x[np.int64(np.abs(x[:, 0]) * 2.0)] ** 2), (x2,)),
    }
    for name, (fn, args) in fns.items():
        check_grad(fn, args, TOL_REDUCE, f"gap {name}")

    # --- the cotangent of a HOST argument differentiated through device data is a host value (test_util.py:38)
    xd = b200.asarray(x2)
    g = grad(lambda s: np.sum(np.tanh(s * xd)))(0.5)
    want = _ref(grad(lambda s: np.sum(np.tanh(s * x2))), 0.5)
    assert isinstance(g, (float, onp.floating, onp.ndarray)), type(g)
    assert_close(onp.float64(g), onp.float64(want), TOL_REDUCE, "host scalar argument")
    gw = grad(lambda w: np.sum(np.dot(xd, w) ** 2))(x2[0])
    assert isinstance(gw, onp.ndarray), type(gw)
    assert_close(gw, _ref(grad(lambda w: np.sum(np.dot(x2, w) ** 2)), x2[0]), TOL_REDUCE, "host vector argument")
    gi = grad(lambda w: np.sum(w[cols] * xd[0, :3]))(x2[1])            # sparse untake into a host accumulator
    assert isinstance(gi, onp.ndarray), type(gi)
    assert_close(gi, _ref(grad(lambda w: np.sum(w[cols] * x2[0, :3])), x2[1]), TOL_REDUCE, "host argument, untake")
    check_grads(lambda s: np.sum(np.exp(s) * xd), modes=["rev", "fwd"], order=2)(0.7)
    check_grads(lambda a, b: np.dot(a, b), modes=["rev", "fwd"], order=2)(0.3, xd)

    # --- host-only code keeps host results while the backend is installed (where/abs rules call anp.zeros)
    j = make_jvp(lambda x: make_jvp(np.abs, x)(1.3)[1], 0.4)(0.7)[1]
    assert not isinstance(j, b200.DeviceArray), type(j)
    gh = grad(lambda x: np.sum(np.where(x > 0, x * x, -x)))(x2)
    assert isinstance(gh, onp.ndarray), type(gh)
    assert isinstance(np.zeros(3) + onp.ones(3), onp.ndarray)
    assert isinstance(np.zeros(3) + 1.0, b200.DeviceArray)
    assert isinstance(np.zeros(3) + xd[0, :3], b200.DeviceArray)

    # --- linspace with traced end points (numpy_vjps.py:240-244, numpy_jvps.py:169-173) on device scalars
    s0 = b200.asarray(onp.array(1.2))
    gl = grad(lambda a: np.sum(np.linspace(a, 3.4, 5) ** 2))(s0)
    assert_close(to_host(gl), _ref(grad(lambda a: np.sum(np.linspace(a, 3.4, 5) ** 2)), onp.array(1.2)), TOL_REDUCE, "linspace start")
    assert_close(to_host(np.linspace(s0, s0 * 3.0, 7)), onp.linspace(1.2, 1.2 * 3.0, 7), TOL_ELEMENTWISE, "device linspace")


GRAD_CASES["reference_suite_gaps"] = grad_case_reference_suite_gaps


def case_shared_denominator_division_bit_exact():
    """Quotients over one denominator share the reciprocal refinement (codegen PRELUDE b200_rcp/b200_div); the
    result must be the SAME bits as NumPy's IEEE division for every class of operand: normal, subnormal, zero of
    either sign, huge, infinite, NaN, quotients that underflow/overflow, negated denominators, constant denominators."""
    r = rs(7)
    n = 1 << 16
    specials = onp.array([0.0, -0.0, 1.0, -1.0, onp.inf, -onp.inf, onp.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
                          1.7976931348623157e308, -1.7976931348623157e308, 1e-310, 3.0, 1.0 / 3.0, 1e300, 1e-300,
                          6.5827683646048100446e-37, 2.0 ** -1022, 2.0 ** 1023, 2.0 ** -537, 2.0 ** 512])

    def draw():
        x = r.randn(n) * onp.exp(r.uniform(-40, 40, n))
        wide = r.randn(n) * 2.0 ** r.randint(-1074, 1023, n).astype(onp.float64)
        x[n // 2:] = wide[n // 2:]
        x[: specials.size * specials.size] = 0  # overwritten below by the cross product of specials
        return x

    a, b, c, d = draw(), draw(), draw(), draw()
    k = specials.size
    blk = k * k + 7                      # 491: coprime with 10, so block p sees selector (p + idx) % 10
    assert blk % 10 == 1
    for p in range(10):
        sl = slice(p * blk, p * blk + k * k)
        a[sl] = onp.repeat(specials, k)
        d[sl] = onp.tile(specials, k)
        b[sl] = onp.tile(specials[::-1], k)
        c[sl] = 1.5
    sel = onp.arange(n) % 10

    def f(a, b, c, d, sel):
        # every quotient lives in ONE fused kernel (so the reciprocal really is shared); element i keeps quotient sel[i]
        e = d * d
        qs = [a / d, b / d, c / (-d), (-a) / d, 1.0 / d, a / e, b / e, a / 3.0, b / 3.0, c / -3.0]
        out = qs[0]
        for i in range(1, 10):
            out = onp.where(sel == i, qs[i], out)
        return out, qs[0] + qs[1] * qs[2] - qs[5] / qs[6]

    check(f, a, b, c, d, sel, tol=0, what="shared-denominator division")


OP_CASES["shared_denominator_division_bit_exact"] = case_shared_denominator_division_bit_exact


def case_convolve_general():
    """autograd.scipy.signal.convolve beyond the NCHW pattern (scipy/signal.py:10-76; reference tests
    tests/test_scipy.py:267-330): 1-D / 2-D / N-d convolved axes, contracted and untouched axes in any position,
    operands in either size order, modes valid / full / same — values and both VJPs against the reference on NumPy."""
    from autograd.scipy.signal import convolve

    import autograd.numpy as np

    r = rs(51)
    configs = [
        ((5,), (3,), None, None),
        ((3,), (7,), None, None),                                   # filter larger than signal: roles swap
        ((6, 5), (3, 2), None, None),
        ((3, 5), (3, 4), ([1], [0]), None),                         # untouched axes on both sides
        ((2, 5, 4, 3), (3, 4, 2), ([1, 2], [1, 0]), None),
        ((2, 4, 2, 3, 2), (2, 5, 4, 3), ([1, 3], [2, 3]), ([0], [0])),
        ((4, 3, 9, 8), (3, 5, 3, 3), ([2, 3], [2, 3]), ([1], [0])),  # NCHW (dedicated kernels for valid/full)
        ((4, 3, 3, 3), (3, 5, 6, 7), ([2, 3], [2, 3]), ([1], [0])),  # NCHW with the filter operand LARGER
    ]
    for a_shape, b_shape, axes, dot_axes in configs:
        A, B = r.randn(*a_shape), r.randn(*b_shape)
        for mode in ("valid", "full", "same"):
            if mode == "same" and axes is None and len(a_shape) == 1 and b_shape[0] > a_shape[0]:
                pass  # still defined: 'same' keeps A's extent
            kw = dict(axes=axes, dot_axes=dot_axes, mode=mode)
            what = f"convolve {a_shape}*{b_shape} axes={axes} dot={dot_axes} {mode}"
            check(lambda A, B: convolve(A, B, **kw), A, B, tol=TOL_REDUCE, what=what)
            loss = lambda A, B: np.sum(np.sin(convolve(A, B, **kw)))
            for argnum in (0, 1):
                check_grad(loss, (A, B), TOL_REDUCE, f"{what} vjp arg{argnum}", argnum)


OP_CASES["convolve_general"] = case_convolve_general


def case_logsumexp_weighted():
    """scipy.special.logsumexp with weights ``b`` / ``return_sign`` (reference tests test_scipy.py:237-255) and the
    reference VJP/JVP (scipy/special.py:123-147) running on device values."""
    from autograd.scipy.special import logsumexp

    import autograd.numpy as np
    from autograd import grad

    r = rs(52)
    for shape in ((4,), (3, 4), (2, 3, 4)):
        x, w = r.randn(*shape) * 3, onp.exp(r.randn(*shape))
        for axis in (None, 0, len(shape) - 1):
            for keepdims in (False, True):
                kw = dict(axis=axis, keepdims=keepdims)
                check(lambda x, w: logsumexp(x, b=w, **kw), x, w, tol=TOL_REDUCE, what=f"lse b {shape} {kw}")
                check_grad(lambda x, w: np.sum(logsumexp(x, b=w, **kw) ** 2), (x, w), TOL_REDUCE, f"lse b grad {shape} {kw}")
    x, w = r.randn(5, 6), r.randn(5, 6)
    w[0, :] = 0.0                                       # a row without any weight: -inf, shift falls back to 0
    w[1, 2] = 0.0
    x[1, 2] = 50.0                                      # the largest entry carries zero weight: not the shift
    check(lambda x, w: logsumexp(x, b=w, axis=1, return_sign=True), x, w, tol=TOL_REDUCE, what="lse signed weights")
    check(lambda x: logsumexp(x, axis=0, b=2.5), x, tol=TOL_REDUCE, what="lse scalar weight")


OP_CASES["logsumexp_weighted"] = case_logsumexp_weighted


def case_scipy_special_integer_orders():
    """jn / yn / iv / ive / polygamma of a host integer order on device arguments — the functions the reference's
    derivative rules of j1, y1, i1, psi, gammaln call (autograd/scipy/special.py:43-48,76-90) — values against SciPy,
    then first and second derivatives against the reference on NumPy."""
    import scipy.special as sps
    from autograd.scipy import special as asp

    import autograd.numpy as np

    r = rs(91)
    x = r.randn(2000) * 6.0
    pos = onp.abs(x) + 0.05
    tol = 1e-12
    for n in (0, 1, 2, 3, 7, -2, -3):
        check(lambda a, n=n: sps.jv(n, a), x, tol=tol, what=f"jv({n}, x)")
        check(lambda a, n=n: sps.yn(n, a), pos, tol=tol, what=f"yn({n}, x)")
        check(lambda a, n=n: sps.iv(n, a), x, tol=tol, what=f"iv({n}, x)")
        check(lambda a, n=n: sps.ive(n, a), x * 40, tol=tol, what=f"ive({n}, x)")
    check(lambda a: sps.ive(2, a), onp.array([0.0, 1e-8, 750.0, 5000.0, -5000.0, 1e7, onp.inf, onp.nan])[:6], tol=tol, what="ive range")
    for n in (1, 2, 3, 5):
        check(lambda a, n=n: asp.polygamma(n, a), pos * 5, tol=tol, what=f"polygamma({n}, x>0)")
        check(lambda a, n=n: asp.polygamma(n, a), -pos * 2 - 0.0137, tol=1e-10, what=f"polygamma({n}, x<0)")
    check(lambda a: asp.polygamma(1, a), onp.array([1.0, 0.5, 1e-3, 50.0, 1e5, 0.0, -1.0]), tol=tol, what="trigamma edges")
    check(lambda a: asp.multigammaln(a, 3), pos + 1.5, tol=tol, what="multigammaln")
    check(lambda a: asp.rgamma(a), x, tol=tol, what="rgamma")
    check(lambda a: asp.gammasgn(a), onp.concatenate([x, [0.0, -0.0, -1.0, -2.0, onp.inf, -onp.inf, onp.nan, 1e-300, -1e-300]]),
          tol=0, what="gammasgn")
    small = pos[:40] * 3
    for name, arg in (("j1", small), ("y1", small), ("i1", small[:40] * 0.5), ("j0", small), ("i0", small * 0.5), ("gammaln", small),
                      ("psi", small), ("gamma", small * 0.5)):
        fn = getattr(asp, name)
        check_grad(lambda a, fn=fn: np.sum(fn(a) ** 2), (arg,), TOL_REDUCE, f"d {name}")
        from autograd import grad
        check_grad(lambda a, fn=fn: np.sum(grad(lambda b: np.sum(fn(b)))(a) * np.cos(a)), (arg,), TOL_REDUCE, f"d2 {name}")
    check_grad(lambda a: np.sum(asp.jn(3, a) * asp.yn(2, a) + asp.iv(2, a) + asp.ive(3, a) + asp.polygamma(2, a)), (small,),
               TOL_REDUCE, "d jn/yn/iv/ive/polygamma")


OP_CASES["scipy_special_integer_orders"] = case_scipy_special_integer_orders


def case_partition():
    """np.partition / argpartition: the kth element is in sorted position and the two sides hold the same SETS as
    NumPy's result (the order inside the sides is unspecified by NumPy); grad_partition agrees with the reference."""
    import autograd.numpy as np

    r = rs(53)
    x = r.randn(37)
    for kth in (0, 5, 18, 36, -3):
        want = onp.partition(x, kth)
        got = to_host(onp.partition(to_dev(x), kth))
        k = kth % x.size
        assert got[k] == want[k]
        assert onp.array_equal(onp.sort(got[:k]), onp.sort(want[:k])) and onp.array_equal(onp.sort(got[k:]), onp.sort(want[k:]))
        perm = to_host(onp.argpartition(to_dev(x), kth))
        assert perm.dtype == onp.int64 and onp.array_equal(x[perm], got)
    m = r.randn(4, 9)
    got = to_host(onp.partition(to_dev(m), 3, axis=1))
    assert onp.array_equal(got[:, 3], onp.partition(m, 3, axis=1)[:, 3])
    check_grad(lambda v: np.sum(np.partition(v, 5) * np.arange(37.0)), (x,), TOL_REDUCE, "grad_partition")


OP_CASES["partition"] = case_partition


def grad_case_scipy_stats_densities():
    """autograd.scipy.stats gamma / chi2 / poisson / t / beta / dirichlet densities and special.beta / betaln on device
    operands: values against scipy.stats on the host, gradients against the reference rules on NumPy
    (autograd/scipy/stats/*.py; reference tests tests/test_scipy.py:48-224)."""
    import autograd.scipy.stats as st
    from autograd.scipy import special as asp

    import autograd.numpy as np

    r = rs(61)
    x = r.randn(40) ** 2 + 0.3
    a = r.randn(40) ** 2 + 1.1
    b = r.randn(40) ** 2 + 1.1
    u = r.rand(40) * 0.96 + 0.02
    k = onp.round(r.randn(40) ** 2 * 3)
    loc, scale, df = r.randn(40), r.randn(40) ** 2 + 0.5, r.randn(40) ** 2 + 2.1
    tol = 1e-12
    check(lambda x, a: (st.gamma.logpdf(x, a), st.gamma.pdf(x, a)), x, a, tol=tol, what="gamma")
    check(lambda x: (st.chi2.logpdf(x, 3), st.chi2.pdf(x, 3)), x, tol=tol, what="chi2")
    check(lambda k, mu: (st.poisson.logpmf(k, mu), st.poisson.pmf(k, mu)), k, a, tol=tol, what="poisson")
    check(lambda k, mu: st.poisson.logpmf(k + 0.5, mu), k, a, tol=tol, what="poisson non-integer k")
    check(lambda x, df, loc, scale: (st.t.logpdf(x, df, loc, scale), st.t.pdf(x, df, loc, scale)), loc * 2, df, loc, scale,
          tol=tol, what="t")
    check(lambda u, a, b: (st.beta.logpdf(u, a, b), st.beta.pdf(u, a, b)), u, a, b, tol=tol, what="beta")
    check(lambda a, b: (asp.betaln(a, b), asp.beta(a, b)), a, b, tol=tol, what="betaln / beta")
    xs = r.rand(5, 7) + 0.1
    xs = xs / xs.sum(0)
    alpha = r.rand(5) * 3 + 0.5
    check(lambda x, al: (st.dirichlet.logpdf(x, al), st.dirichlet.pdf(x, al)), xs, alpha, tol=1e-11, what="dirichlet")
    check(lambda x, al: st.dirichlet.logpdf(x, al), xs[:, 0].copy(), alpha, tol=1e-11, what="dirichlet one point")

    for argnum in (0, 1):
        check_grad(lambda x, a: np.sum(st.gamma.logpdf(x, a) + st.gamma.pdf(x, a)), (x, a), TOL_REDUCE, f"gamma d{argnum}", argnum)
    check_grad(lambda x: np.sum(st.chi2.logpdf(x, 4) * np.cos(x)), (x,), TOL_REDUCE, "chi2 dx")
    check_grad(lambda mu, k: np.sum(st.poisson.logpmf(k, mu) + st.poisson.pmf(k, mu)), (a, k), TOL_REDUCE, "poisson dmu")
    for argnum in (0, 1, 2, 3):
        check_grad(lambda x, df, loc, scale: np.sum(st.t.logpdf(x, df, loc, scale) + st.t.pdf(x, df, loc, scale)),
                   (loc * 2, df, loc, scale), TOL_REDUCE, f"t d{argnum}", argnum)
    for argnum in (0, 1, 2):
        check_grad(lambda u, a, b: np.sum(st.beta.logpdf(u, a, b)), (u, a, b), TOL_REDUCE, f"beta d{argnum}", argnum)
        if argnum:
            check_grad(lambda u, a, b: np.sum(asp.betaln(a, b) * u + asp.beta(a, b)), (u, a, b), TOL_REDUCE, f"betaln d{argnum}", argnum)
    x1 = xs[:, 0].copy()                        # the reference rule handles one point (dirichlet.py:12-15)
    check_grad(lambda al, x: st.dirichlet.logpdf(x, al) + st.dirichlet.pdf(x, al), (alpha, x1), TOL_REDUCE, "dirichlet dalpha")
    check_grad(lambda x, al: st.dirichlet.logpdf(x, al), (x1, alpha), TOL_REDUCE, "dirichlet dx")


GRAD_CASES["scipy_stats_densities"] = grad_case_scipy_stats_densities


def grad_case_scipy_incomplete_functions_and_cdfs():
    """scipy.special.gammainc / gammaincc / betainc as device functions (series + Lentz continued fraction) and the CDFs
    of autograd.scipy.stats gamma / chi2 / poisson / beta / t composed from them: values against SciPy, derivatives
    against the reference rules (autograd/scipy/special.py:57-63, stats/*.py)."""
    import scipy.special as sps
    import autograd.scipy.stats as st
    from autograd.scipy import special as asp

    import autograd.numpy as np

    r = rs(62)
    n = 400
    a = r.rand(n) * 30 + 0.05
    b = r.rand(n) * 30 + 0.05
    x = r.rand(n) * 60
    u = r.rand(n)
    tol = 1e-12
    check(lambda a, x: (sps.gammainc(a, x), sps.gammaincc(a, x)), a, x, tol=tol, what="gammainc / gammaincc")
    check(lambda x: (sps.gammainc(1, x), sps.gammaincc(2.5, x)), x, tol=tol, what="gammainc scalar a")
    check(lambda a, b, u: sps.betainc(a, b, u), a, b, u, tol=tol, what="betainc")
    edge_a = onp.array([0.5, 1.0, 2.0, 3.0, 1.0, 5.0])
    edge_x = onp.array([0.0, 0.0, 1e-300, onp.inf, 1e-12, 1e3])
    check(lambda a, x: (sps.gammainc(a, x), sps.gammaincc(a, x)), edge_a, edge_x, tol=tol, what="gammainc edges")
    check(lambda a, u: sps.betainc(a, a + 0.5, u), edge_a, onp.array([0.0, 1.0, 1e-12, 1 - 1e-12, 0.5, 0.999]), tol=tol, what="betainc edges")

    xs = r.randn(60) ** 2 + 0.05
    aa = r.randn(60) ** 2 + 1.1
    bb = r.randn(60) ** 2 + 1.1
    uu = r.rand(60) * 0.98 + 0.01
    kk = onp.round(r.randn(60) ** 2 * 3)
    loc, scale, df = r.randn(60), r.randn(60) ** 2 + 0.5, r.randn(60) ** 2 + 2.1
    check(lambda x, a: st.gamma.cdf(x, a), xs, aa, tol=tol, what="gamma.cdf")
    check(lambda x: st.chi2.cdf(x, 3), xs, tol=tol, what="chi2.cdf")
    check(lambda k, mu: st.poisson.cdf(k, mu), kk, aa, tol=tol, what="poisson.cdf")
    check(lambda u, a, b: st.beta.cdf(u, a, b), uu, aa, bb, tol=tol, what="beta.cdf")
    check(lambda x, df, loc, scale: (st.t.cdf(x, df, loc, scale), st.t.logcdf(x, df, loc, scale)), loc * 2.5, df, loc, scale,
          tol=tol, what="t.cdf / logcdf")
    check_grad(lambda x, a: np.sum(asp.gammainc(a, x) - 2.0 * asp.gammaincc(a, x)), (xs, aa), TOL_REDUCE, "gammainc dx")
    check_grad(lambda u, a, b: np.sum(asp.betainc(a, b, u)), (uu, aa, bb), TOL_REDUCE, "betainc dx")
    check_grad(lambda x, a: np.sum(st.gamma.cdf(x, a) + st.chi2.cdf(x, 4)), (xs, aa), TOL_REDUCE, "gamma/chi2 cdf dx")
    check_grad(lambda mu, k: np.sum(st.poisson.cdf(k, mu)), (aa, kk), TOL_REDUCE, "poisson cdf dmu")
    check_grad(lambda u, a, b: np.sum(st.beta.cdf(u, a, b)), (uu, aa, bb), TOL_REDUCE, "beta cdf dx")
    for argnum in (0, 2):
        check_grad(lambda x, df, loc, scale: np.sum(st.t.cdf(x, df, loc, scale) + st.t.logcdf(x, df, loc, scale)),
                   (loc * 2.5, df, loc, scale), TOL_REDUCE, f"t cdf d{argnum}", argnum)


GRAD_CASES["scipy_incomplete_functions_and_cdfs"] = grad_case_scipy_incomplete_functions_and_cdfs


def _cdev(z):
    """Host complex ndarray -> ComplexDeviceArray (pair of real device arrays)."""
    from autograd_b200 import complex_array as CA

    return CA.as_complex(z)


def _ccheck(fn, *host_args, tol=1e-11, what=""):
    with onp.errstate(all="ignore"):
        want = fn(*host_args)
    dev_args = [_cdev(a) if isinstance(a, onp.ndarray) and a.dtype.kind == "c" else to_dev(a) for a in host_args]
    got = fn(*dev_args)
    for g, w in zip(leaves(got), leaves(want)):
        g = g.get() if hasattr(g, "get") else g
        g, w = onp.asarray(g), onp.asarray(w)
        assert g.shape == w.shape and g.dtype == w.dtype, f"{what}: {g.shape} {g.dtype} vs {w.shape} {w.dtype}"
        scale = max(float(onp.max(onp.abs(w))), 1e-300) if w.size else 1.0
        err = float(onp.max(onp.abs(g - w))) if w.size else 0.0
        assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol:g} * {scale:.3e}"


def case_complex_values():
    """Complex values on the device are pairs of real device arrays (complex_array.py) whose ufuncs are composed from
    the real ones with NumPy's complex formulas: values against NumPy on the same inputs."""
    r = rs(71)
    z = r.randn(5, 4) + 1j * r.randn(5, 4)
    w = r.randn(5, 4) + 0.5j * r.randn(5, 4)
    x = r.randn(5, 4)
    for name in ("exp", "log", "sqrt", "sin", "cos", "tan", "sinh", "cosh", "tanh", "exp2", "expm1", "log2", "log10", "log1p",
                 "arcsinh", "arccosh", "arctanh", "square", "reciprocal", "negative", "conjugate", "absolute"):
        fn = getattr(onp, name)
        _ccheck(lambda a, fn=fn: fn(a), z, what=f"complex {name}")
    for name in ("add", "subtract", "multiply", "divide", "power"):
        fn = getattr(onp, name)
        _ccheck(lambda a, b, fn=fn: fn(a, b), z, w, what=f"complex {name}")
        _ccheck(lambda a, b, fn=fn: fn(a, b), z, x * x + 0.5, what=f"complex-real {name}")
        _ccheck(lambda a, b, fn=fn: fn(b, a), z, x * x + 0.5, what=f"real-complex {name}")
    _ccheck(lambda a: (a * 2.0, a + 1.5j, (0.5 - 2j) * a, a / (1 + 1j), a ** 3, a ** -2, 2.0 ** a, a ** (0.5 + 0.25j)), z, what="complex scalars")
    _ccheck(lambda a, b: (a * 1j, b + 0j, b * (1 + 2j)), z, x, what="promotion of real device arrays")
    _ccheck(lambda a: (onp.real(a), onp.imag(a), onp.angle(a), onp.conj(a), onp.abs(a), a.T, a.reshape(4, 5), a[1:, ::2],
                       onp.transpose(a), onp.ravel(a), onp.concatenate([a, a * 2.0], axis=1), onp.where(onp.real(a) > 0, a, 1j)), z,
            what="complex structure")
    _ccheck(lambda a: (onp.sum(a), onp.sum(a, axis=0), onp.mean(a, axis=1, keepdims=True), onp.prod(a, axis=0), onp.prod(a),
                       onp.var(a), onp.std(a, axis=0), onp.max(a), onp.min(a, axis=1)), z, what="complex reductions")
    _ccheck(lambda a, b: (onp.dot(a, b.T), onp.matmul(a.T, b), onp.tensordot(a, b, axes=([0], [0])), onp.outer(a[0], b[1]),
                          onp.inner(a[0], b[0]), onp.einsum("ij,kj->ik", a, b), onp.kron(a[:2, :2], b[:2, :3])), z, w,
            tol=1e-10, what="complex contractions")
    _ccheck(lambda a, b: (onp.dot(a, b.T), onp.dot(b, a.T)), z, x, tol=1e-10, what="complex-real contractions")
    edge = onp.array([0j, 1 + 0j, -1 + 0j, 1j, -1j, 3 - 4j, -2 + 1e-3j, 1e-8 + 1e-8j])
    _ccheck(lambda a: (onp.sqrt(a), onp.abs(a), onp.exp(a), a * a, onp.conj(a) / (a + 1.5)), edge, what="complex edge values")


OP_CASES["complex_values"] = case_complex_values


def grad_case_complex():
    """Reverse and forward mode through complex device values (ComplexDeviceVSpace mirrors ComplexArrayVSpace,
    numpy_vspaces.py:40-66): gradients against the reference on NumPy, and autograd's own check_grads on device values
    (the complex halves of the reference's tests/test_systematic.py, tests/test_complex.py)."""
    import autograd.numpy as np
    from autograd import grad
    from autograd.test_util import check_grads

    r = rs(72)
    zh = r.randn(4, 3) + 0.3j * r.randn(4, 3)
    xh = r.randn(4, 3)

    def real_loss(z, x):
        u = np.exp(z * 0.5) * np.sin(z) + x * z
        v = np.dot(u, np.conj(u).T)
        return np.real(np.sum(v * v)) + np.sum(np.abs(z) ** 2) + np.sum(np.angle(z + 3.0))

    want = _ref(grad(real_loss, 0), zh, xh)
    got = grad(real_loss, 0)(_cdev(zh), to_dev(xh)).get()
    got, want = onp.asarray(got), onp.asarray(want)
    assert got.dtype == want.dtype == onp.complex128
    assert_close(got.real.copy(), want.real.copy(), TOL_REDUCE, "complex grad wrt z (real part)")
    assert_close(got.imag.copy(), want.imag.copy(), TOL_REDUCE, "complex grad wrt z (imaginary part)")
    want = _ref(grad(real_loss, 1), zh, xh)
    got = grad(real_loss, 1)(_cdev(zh), to_dev(xh))
    assert isinstance(got, b200.DeviceArray)                      # a real argument gets a real gradient (match_complex)
    assert_close(got.get(), want, TOL_REDUCE, "complex grad wrt real x")
    z = _cdev(zh)
    for fn in (np.exp, np.abs):
        check_grads(fn, modes=["rev", "fwd"], order=2)(z)
    for fn in (np.log, np.sqrt, np.tanh, np.arcsinh, np.angle, np.real, np.imag, np.conj):   # (every generated kernel is
        check_grads(fn, modes=["rev", "fwd"], order=1)(z)                                      #  an NVRTC compile on the GPU)
    check_grads(lambda a, b: a * b + a / b, modes=["rev", "fwd"], order=1)(z, _cdev(zh[::-1].copy() + 2.0))
    check_grads(lambda a, b: np.dot(a, b.T), modes=["rev", "fwd"], order=2)(z, to_dev(xh))
    check_grads(lambda a: np.sum(a, axis=0) * np.prod(a, axis=1)[:3] + np.mean(a), modes=["rev"], order=1)(z)
    check_grads(lambda a: np.max(a, axis=0), modes=["rev"], order=1)(z)


GRAD_CASES["complex"] = grad_case_complex


def grad_case_linalg_composed():
    """np.linalg.solve / inv / det / slogdet / cholesky composed from device array operations (autograd_b200/linalg.py:
    Gauss-Jordan with on-device partial pivoting, column Cholesky) and scipy.stats.multivariate_normal on top of them:
    values against NumPy / SciPy, gradients against the reference rules (autograd/numpy/linalg.py:57-88,260-279,
    autograd/scipy/stats/multivariate_normal.py)."""
    import scipy.stats as sst
    import autograd.scipy.stats.multivariate_normal as mvn

    import autograd.numpy as np

    r = rs(81)
    for n in (1, 3, 6):
        a = r.randn(n, n) + 0.5 * onp.eye(n)
        bm, bv = r.randn(n, 2), r.randn(n)
        check(lambda a, b: onp.linalg.solve(a, b), a, bm, tol=TOL_REDUCE, what=f"solve {n} matrix rhs")
        check(lambda a, b: onp.linalg.solve(a, b), a, bv, tol=TOL_REDUCE, what=f"solve {n} vector rhs")
        check(lambda a: (onp.linalg.inv(a), onp.linalg.det(a)), a, tol=TOL_REDUCE, what=f"inv/det {n}")
        check(lambda a: tuple(onp.linalg.slogdet(a)), a, tol=TOL_REDUCE, what=f"slogdet {n}")
        spd = a @ a.T + n * onp.eye(n)
        check(lambda s: onp.linalg.cholesky(s), spd, tol=TOL_REDUCE, what=f"cholesky {n}")
    perm = onp.array([[0.0, 2.0, 1.0], [1e-9, 1.0, 3.0], [4.0, 0.5, 0.25]])          # needs the row exchanges
    check(lambda a: (onp.linalg.inv(a), onp.linalg.det(a), onp.linalg.solve(a, a[:, 0] + 1.0)), perm, tol=TOL_REDUCE, what="pivoting")
    stack = r.randn(3, 4, 4) + onp.eye(4)
    check(lambda a, b: (onp.linalg.inv(a), onp.linalg.det(a), onp.linalg.solve(a, b)), stack, r.randn(3, 4, 2), tol=TOL_REDUCE,
          what="batched")
    a = r.randn(4, 4) + onp.eye(4)
    b = r.randn(4, 3)
    check_grad(lambda a, b: np.sum(np.linalg.solve(a, b) ** 2), (a, b), TOL_REDUCE, "solve da", 0)
    check_grad(lambda a, b: np.sum(np.linalg.solve(a, b) ** 2), (a, b), TOL_REDUCE, "solve db", 1)
    check_grad(lambda a: np.sum(np.linalg.inv(a) * np.cos(a)) + np.linalg.det(a) + np.linalg.slogdet(a)[1], (a,), TOL_REDUCE, "inv/det/slogdet")
    spd = a @ a.T + 4 * onp.eye(4)
    check_grad(lambda s: np.sum(np.linalg.cholesky(s) ** 3), (spd,), TOL_REDUCE, "cholesky")

    x, mean = r.randn(7, 4), r.randn(4)
    with onp.errstate(all="ignore"):
        want = sst.multivariate_normal.logpdf(x, mean, spd)
    got = to_host(mvn.logpdf(to_dev(x), to_dev(mean), to_dev(spd)))
    assert_close(got, want, TOL_REDUCE, "mvn.logpdf")
    assert_close(to_host(mvn.logpdf(to_dev(x[0]), to_dev(mean), to_dev(spd))), onp.asarray(sst.multivariate_normal.logpdf(x[0], mean, spd)),
                 TOL_REDUCE, "mvn.logpdf one point")
    assert_close(to_host(mvn.entropy(to_dev(mean), to_dev(spd))), onp.asarray(sst.multivariate_normal.entropy(mean, spd)), TOL_REDUCE, "mvn.entropy")
    for argnum in (0, 1, 2):
        check_grad(lambda x, m, c: np.sum(mvn.logpdf(x, m, c)) + np.sum(mvn.pdf(x, m, c)), (x, mean, spd), TOL_REDUCE, f"mvn d{argnum}", argnum)


GRAD_CASES["linalg_composed"] = grad_case_linalg_composed


def case_inplace_write_semantics():
    """NumPy evaluates eagerly, so a later in-place write never changes an earlier result; the device records ufunc
    calls lazily and must give the same values: pending readers of a buffer are computed before it is written, a copy
    made by astype() is independent, and overlapping source / destination slices go through a temporary."""
    import pytest

    def script(xp, asarr):
        out = []
        x = asarr(onp.arange(8.0))
        y = x * 2.0                      # pending when x is written
        z = xp.exp(x)
        x[0] = 100.0
        x[1:3] = 1.0
        out += [y, z, x * 1.0]
        t = asarr(onp.arange(12.0).reshape(3, 4))
        g = t[asarr(onp.array([2, 0, 1])), asarr(onp.array([1, 3, 0]))]   # lazy gather holds its table
        onp.add.at(t, (slice(None), slice(0, 2)), 10.0)
        out += [g * 1.0]
        a = asarr(onp.arange(10.0))
        c = a.astype(onp.float64)        # NumPy: a copy
        c[0] = 9.0
        out += [a * 1.0, c * 1.0]
        s = asarr(onp.arange(10.0))
        s[1:] = s[:-1]                   # overlapping, same buffer
        out += [s * 1.0]
        s2 = asarr(onp.arange(10.0))
        s2[:-1] = s2[1:]
        out += [s2 * 1.0]
        # a value recorded from a PENDING operand that is given its payload later, and written after that
        p = asarr(onp.arange(12.0).reshape(3, 4)) + 1.0      # pending
        q = p * 2.0                                          # recorded with p pending
        q2 = q * q + p[1:2]                                  # (a slice computes p)
        p[0, 0] = 100.0                                      # q and q2 must keep the old p
        p[2] = -1.0
        out += [q, q2, p * 1.0]
        return out

    want = script(onp, lambda h: h.copy())
    got = script(onp, lambda h: b200.asarray(h))
    for i, (g, w) in enumerate(zip(got, want)):
        assert_close(to_host(g), w, 0, f"in-place semantics #{i}")
    with pytest.raises(IndexError):
        b200.asarray(onp.arange(6.0).reshape(2, 3))[[7], [0]]
    with pytest.raises(IndexError):
        onp.add.at(b200.asarray(onp.zeros(4)), onp.array([0, -5]), 1.0)
    with pytest.raises(IndexError):
        b200.asarray(onp.zeros(4))[onp.array([1, 4])] = 1.0


def case_item_assignment_duplicates_last_write_wins():
    """x[idx] = v with duplicate indices: NumPy's sequential assignment leaves the LAST value (bit-exact index work)."""
    r = rs(11)
    idx = r.randint(0, 50, 4000)
    vals = r.randn(4000)
    want = onp.zeros(50)
    want[idx] = vals
    x = b200.asarray(onp.zeros(50))
    x[b200.asarray(idx)] = b200.asarray(vals)
    assert_close(x.get(), want, 0, "duplicate item assignment")
    rows = r.randint(0, 7, 300)
    v2 = r.randn(300, 5)
    want2 = onp.full((7, 5), -1.0)
    want2[rows] = v2
    x2 = b200.asarray(onp.full((7, 5), -1.0))
    x2[b200.asarray(rows)] = b200.asarray(v2)
    assert_close(x2.get(), want2, 0, "duplicate row assignment")


OP_CASES["inplace_write_semantics"] = case_inplace_write_semantics
OP_CASES["item_assignment_duplicates_last_write_wins"] = case_item_assignment_duplicates_last_write_wins


def grad_case_extra_workloads():
    """the model code of examples/rnn.py, bayesian_neural_net.py (through black_box_svi.py's
    variational objective, einsum over weight samples) and black_box_svi.py's own funnel density (scipy.stats.norm),
    device gradients against the reference on NumPy.  The stochastic objectives draw from a host RandomState: a fresh
    objective per evaluation gives both sides the same draws."""
    import autograd.numpy as np
    from autograd import grad

    from examples import workloads as w

    params, X, Y = w.rnn_data()
    check_grad(w.rnn_objective, (params, X, Y), TOL_REDUCE, "rnn.py gradient")

    rbf = lambda x: np.exp(-(x**2))
    nw, _pred, logprob = w.bnn_make_nn_funs([1, 20, 20, 1], 0.1, 0.01, nonlinearity=rbf)
    inputs, targets = w.bnn_toy_dataset()
    init = onp.concatenate([rs(0).randn(nw), -5 * onp.ones(nw)])
    make = lambda lift: w.svi_objective(lambda weights, t: logprob(weights, lift(inputs), lift(targets)), nw, 20)
    want = _ref(grad(make(lambda a: a)), init, 0)
    got = grad(make(b200.asarray))(b200.asarray(init), 0).get()
    assert_close(got, want, TOL_REDUCE, "bayesian_neural_net.py gradient")

    init = onp.concatenate([-1 * onp.ones(2), -5 * onp.ones(2)])
    want = _ref(grad(w.svi_objective(w.svi_funnel_log_density, 2, 2000)), init, 0)
    got = grad(w.svi_objective(w.svi_funnel_log_density, 2, 2000))(b200.asarray(init), 0).get()
    assert_close(got, want, TOL_REDUCE, "black_box_svi.py gradient")


def grad_case_checkpoint():
    """Gradient checkpointing around every fluid time step.  The reference's ``autograd.checkpoint``
    (differential_operators.py:230-242) runs unchanged on device values and gives the reference's gradient; its v1.9.1
    staging (core.py:60-65 calls the rule's maker when the node is recorded) re-traces during the FORWARD pass, so it
    retains as much as the plain tape - on NumPy and here alike.  ``autograd_b200.checkpoint`` defers the re-trace to the
    backward pass as the docstring promises: same gradient, and the retained memory drops to the state between steps."""
    from autograd import value_and_grad

    from examples import workloads as w

    be = b200.backend.get()
    rows = cols = 64
    steps = 30          # the plain tape grows by ~11 arrays per step, the checkpointed one by the 3 state arrays + 3 results
    p, s0, tg = w.fluid_data(rows, cols)
    vw, gw = _ref(value_and_grad(w.fluid_objective), p, s0, tg, rows, cols, steps)
    args = (b200.asarray(p), b200.asarray(s0), b200.asarray(tg), rows, cols, steps)

    def run(fn, *extra):
        be.synchronize()
        be.mem_stats()                                # reading the statistics resets the peak to the current use
        v, g = value_and_grad(fn)(*args, *extra)
        gh, vh = g.get(), float(v)
        return vh, gh, be.mem_stats()["peak_in_use"]

    v1, g1, plain_peak = run(w.fluid_objective)
    v2, g2, ref_ckpt_peak = run(w.fluid_objective_checkpointed)
    v3, g3, ckpt_peak = run(w.fluid_objective_checkpointed, b200.checkpoint)
    for v, g, what in ((v1, g1, "plain tape"), (v2, g2, "autograd.checkpoint"), (v3, g3, "autograd_b200.checkpoint")):
        assert_close(g, gw, TOL_REDUCE, f"fluid gradient, {what}")
        assert abs(v - float(vw)) <= TOL_REDUCE * abs(float(vw)), what
    # the checkpointed run retains the state between steps (3 arrays per step) plus one step's tape; the plain tape of this
    # workload is itself lean since reshapes of pending values stay pending (4.7 arrays per step, 11 before), so the
    # bound is stated in arrays per step as well as relative to the plain tape
    arr = rows * cols * 8
    assert ckpt_peak < 0.9 * plain_peak, f"checkpoint kept {ckpt_peak} bytes, plain tape {plain_peak}"
    assert ckpt_peak < 4.5 * steps * arr, f"checkpoint kept {ckpt_peak / arr / steps:.2f} arrays per step"
    assert ref_ckpt_peak > 0.8 * plain_peak            # documents the reference's behaviour; not a requirement on us


GRAD_CASES["extra_workloads"] = grad_case_extra_workloads
GRAD_CASES["checkpoint"] = grad_case_checkpoint
