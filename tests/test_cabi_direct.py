"""Parity of the C ABI itself (include/b200ag.h) — raw ctypes calls with plain pointers and sizes, no DeviceArray,
against NumPy/SciPy on the same seeded inputs.  This is the boundary a non-Python host would bind."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F64, F32, I64, I32, BOOL = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def abi():
    from autograd_b200 import _lib

    lib = _lib.lib
    assert lib.b200ag_init(0) == 0, lib.b200ag_last_error()

    class Dev:
        def __init__(self):
            self.lib = lib
            self.live = []

        def ok(self, rc):
            assert rc == 0, lib.b200ag_last_error().decode()

        def put(self, a):
            a = np.asarray(a, order="C")
            p = C.c_void_p()
            self.ok(lib.b200ag_malloc(max(a.nbytes, 1), C.byref(p)))
            self.ok(lib.b200ag_memcpy_h2d(p, a.ctypes.data, a.nbytes))
            self.live.append(p)
            return p

        def empty(self, nbytes):
            p = C.c_void_p()
            self.ok(lib.b200ag_malloc(max(nbytes, 1), C.byref(p)))
            self.live.append(p)
            return p

        def get(self, p, shape, dtype):
            out = np.empty(shape, dtype=dtype)
            self.ok(lib.b200ag_memcpy_d2h(out.ctypes.data, p, out.nbytes))
            return out

        def free_all(self):
            for p in self.live:
                lib.b200ag_free(p)
            self.live.clear()

    d = Dev()
    yield d
    d.free_all()


def i64(*v):
    return (C.c_int64 * len(v))(*v)


def test_error_convention(abi):
    # a failing call returns non-zero and leaves a message; it does not abort
    rc = abi.lib.b200ag_gemm(None, 7, 1, 1, 0, None, 0, 1, 1, 0, None, 0, 1, 1, 0, 4, 4, 4, 1)
    assert rc != 0 and b"dtype" in abi.lib.b200ag_last_error()
    rc = abi.lib.b200ag_free(C.c_void_p(12345))
    assert rc != 0 and b"not allocated" in abi.lib.b200ag_last_error()


def test_copy_strided_transpose_and_cast(abi):
    rs = np.random.RandomState(0)
    a = rs.randn(37, 53)
    src = abi.put(a)
    dst = abi.empty(a.size * 8)
    abi.ok(abi.lib.b200ag_copy_strided(dst, F64, i64(1, 37), src, F64, i64(53, 1), i64(37, 53), 2))  # element (r,c) -> dst[c*37 + r]: a.T laid out [53,37]
    assert np.array_equal(abi.get(dst, (53, 37), np.float64), a.T)
    dst32 = abi.empty(a.size * 4)
    abi.ok(abi.lib.b200ag_copy_strided(dst32, F32, i64(53, 1), src, F64, i64(53, 1), i64(37, 53), 2))
    assert np.array_equal(abi.get(dst32, (37, 53), np.float32), a.astype(np.float32))
    dsti = abi.empty(a.size * 8)
    abi.ok(abi.lib.b200ag_copy_strided(dsti, I64, i64(53, 1), src, F64, i64(53, 1), i64(37, 53), 2))
    assert np.array_equal(abi.get(dsti, (37, 53), np.int64), (a).astype(np.int64))


@pytest.mark.parametrize("op,fn", [(0, np.sum), (1, np.max), (2, np.min), (3, np.prod)])
def test_reduce(abi, op, fn):
    rs = np.random.RandomState(op)
    for outer, red, inner in ((1, 100003, 1), (300, 17, 1), (7, 64, 33), (5000, 2, 24), (1, 256, 200)):
        a = rs.rand(outer, red, inner) + 0.5 if op == 3 and red <= 64 else rs.randn(outer, red, inner) * (0.01 if op == 3 else 1)
        src, dst = abi.put(a), abi.empty(outer * inner * 8)
        abi.ok(abi.lib.b200ag_reduce(dst, src, F64, op, outer, red, inner))
        got, want = abi.get(dst, (outer, inner), np.float64), fn(a, axis=1)
        if op in (1, 2):
            assert np.array_equal(got, want)
        else:
            assert np.allclose(got, want, rtol=1e-10, atol=1e-300)
    ai = rs.randint(-1000, 1000, (11, 40, 3))
    src, dst = abi.put(ai), abi.empty(33 * 8)
    abi.ok(abi.lib.b200ag_reduce(dst, src, I64, 0, 11, 40, 3))
    assert np.array_equal(abi.get(dst, (11, 3), np.int64), ai.sum(axis=1))  # integer work: bit-exact


def test_argreduce_first_occurrence(abi):
    a = np.array([[1.0, 7.0, 7.0, -2.0], [np.nan, 3.0, 9.0, 9.0], [5.0, 5.0, 5.0, 5.0]])
    src, dst = abi.put(a), abi.empty(3 * 8)
    abi.ok(abi.lib.b200ag_argreduce(dst, src, F64, 1, 3, 4, 1))
    assert np.array_equal(abi.get(dst, (3,), np.int64), np.argmax(a, axis=1))
    abi.ok(abi.lib.b200ag_argreduce(dst, src, F64, 2, 3, 4, 1))
    assert np.array_equal(abi.get(dst, (3,), np.int64), np.argmin(a, axis=1))


def test_gemm_strides_dtypes_batch(abi):
    rs = np.random.RandomState(3)
    M, N, K = 70, 45, 130
    a, b = rs.randn(M, K), rs.randn(K, N)
    pa, pb, pc = abi.put(a), abi.put(b), abi.empty(M * N * 8)
    abi.ok(abi.lib.b200ag_gemm(pc, F64, N, 1, 0, pa, F64, K, 1, 0, pb, F64, N, 1, 0, M, N, K, 1))
    assert np.allclose(abi.get(pc, (M, N), np.float64), a @ b, rtol=1e-10, atol=1e-12)
    # A^T B with A stored [K, M]: row stride 1, column stride M — no transpose copy
    at = np.ascontiguousarray(a.T)
    pat = abi.put(at)
    abi.ok(abi.lib.b200ag_gemm(pc, F64, N, 1, 0, pat, F64, 1, M, 0, pb, F64, N, 1, 0, M, N, K, 1))
    assert np.allclose(abi.get(pc, (M, N), np.float64), a @ b, rtol=1e-10, atol=1e-12)
    # f64 x f32 -> f32 output (NumPy promotion inside, one rounding on store)
    b32 = b.astype(np.float32)
    pb32, pc32 = abi.put(b32), abi.empty(M * N * 4)
    abi.ok(abi.lib.b200ag_gemm(pc32, F32, N, 1, 0, pa, F64, K, 1, 0, pb32, F32, N, 1, 0, M, N, K, 1))
    assert np.allclose(abi.get(pc32, (M, N), np.float32), (a @ b32.astype(np.float64)).astype(np.float32), rtol=1e-6)
    # batched
    ab, bb = rs.randn(4, 9, 300), rs.randn(4, 300, 5)
    pab, pbb, pcb = abi.put(ab), abi.put(bb), abi.empty(4 * 9 * 5 * 8)
    abi.ok(abi.lib.b200ag_gemm(pcb, F64, 5, 1, 45, pab, F64, 300, 1, 2700, pbb, F64, 5, 1, 1500, 9, 5, 300, 4))
    assert np.allclose(abi.get(pcb, (4, 9, 5), np.float64), ab @ bb, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("M,N,K,ta,tb", [(128, 128, 32, False, False), (300, 200, 70, False, False),
                                          (1000, 777, 333, True, False), (515, 1030, 257, False, True),
                                          (640, 384, 96, True, True), (256, 384, 6144, True, False),
                                          (1024, 1024, 4096, False, False)])
def test_gemm_f32_tensor_core(abi, M, N, K, ta, tb):
    """tcgen05 3xTF32 product against the float64 product and against NumPy's sgemm (north_star: float32 within 1e-5);
    transposed views, ragged edges, K not a multiple of the k-block, split-K (small M x N, long K), long K (chunked
    promotion keeps the error flat)."""
    r = np.random.RandomState(M + N + K)
    a = r.randn(*((K, M) if ta else (M, K))).astype(np.float32)
    b = r.randn(*((N, K) if tb else (K, N))).astype(np.float32)
    A, B = (a.T if ta else a), (b.T if tb else b)
    a_str = (1, a.shape[1]) if ta else (a.shape[1], 1)
    b_str = (1, b.shape[1]) if tb else (b.shape[1], 1)
    out = abi.empty(M * N * 4)
    abi.ok(abi.lib.b200ag_gemm_f32_tc(out, N, abi.put(a), a_str[0], a_str[1], abi.put(b), b_str[0], b_str[1], M, N, K))
    got = abi.get(out, (M, N), np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() / scale < 5e-6
    assert np.abs(got - A @ B).max() / scale < 1e-5      # NumPy float32 reference, north_star tolerance


def test_gemm_f32_mode_switch_and_dispatch(abi):
    """b200ag_gemm sends a large all-float32 product to the tensor cores unless mode 0 asks for the FP64 DMMA kernel;
    both agree with the float64 product."""
    r = np.random.RandomState(5)
    M, N, K = 512, 640, 768
    a, b = r.randn(M, K).astype(np.float32), r.randn(K, N).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    pa, pb = abi.put(a), abi.put(b)
    prev = C.c_int(-1)
    results = {}
    for mode in (0, 1):
        abi.ok(abi.lib.b200ag_gemm_f32_mode(mode, C.byref(prev)))
        out = abi.empty(M * N * 4)
        abi.ok(abi.lib.b200ag_gemm(out, F32, N, 1, 0, pa, F32, K, 1, 0, pb, F32, N, 1, 0, M, N, K, 1))
        results[mode] = abi.get(out, (M, N), np.float32)
    assert prev.value == 0
    scale = np.abs(want).max()
    assert np.abs(results[0] - want).max() / scale < 2e-7    # exact products, FP64 accumulation, one rounding
    assert np.abs(results[1] - want).max() / scale < 5e-6
    assert not np.array_equal(results[0], results[1])         # really two different kernels


def test_cumulate_and_sort(abi):
    """b200ag_cumulate against np.cumsum/np.cumprod; b200ag_sort against np.sort / np.argsort(kind='stable') with ties,
    NaN and infinities (indices bit-exact), rows longer than one shared-memory tile, non-power-of-two lengths."""
    r = np.random.RandomState(9)
    x = r.randn(5, 70000)
    out = abi.empty(x.nbytes)
    abi.ok(abi.lib.b200ag_cumulate(out, abi.put(x), F64, 0, 5, 70000, 1))
    want = np.cumsum(x, axis=1)
    assert np.abs(abi.get(out, x.shape, np.float64) - want).max() <= 1e-10 * np.abs(want).max()
    y = r.randn(4, 33, 6)
    out = abi.empty(y.nbytes)
    abi.ok(abi.lib.b200ag_cumulate(out, abi.put(y), F64, 0, 4, 33, 6))
    assert np.array_equal(abi.get(out, y.shape, np.float64), np.cumsum(y, axis=1))     # same order as NumPy
    p = 1.0 + 0.001 * r.randn(3, 5000)
    out = abi.empty(p.nbytes)
    abi.ok(abi.lib.b200ag_cumulate(out, abi.put(p), F64, 3, 3, 5000, 1))
    assert np.allclose(abi.get(out, p.shape, np.float64), np.cumprod(p, axis=1), rtol=1e-11, atol=0)
    for rows, n in ((1, 1), (3, 7), (4, 2048), (2, 2049), (3, 10000), (1, 300001)):
        k = np.round(r.randn(rows, n) * 50.0) if n > 100 else r.randn(rows, n)
        if n > 100:
            k[0, 3] = np.nan
            k[0, n // 2] = np.nan
            k[-1, 1] = np.inf
            k[-1, 2] = -np.inf
        keys, idx = abi.empty(k.nbytes), abi.empty(k.size * 8)
        abi.ok(abi.lib.b200ag_sort(keys, idx, abi.put(k), F64, rows, n))
        order = np.argsort(k, axis=1, kind="stable")
        assert np.array_equal(abi.get(idx, k.shape, np.int64), order), (rows, n)
        assert np.array_equal(abi.get(keys, k.shape, np.float64), np.take_along_axis(k, order, axis=1), equal_nan=True)
    ki = r.randint(-1000, 1000, (2, 5000)).astype(np.int32)
    idx = abi.empty(ki.size * 8)
    abi.ok(abi.lib.b200ag_sort(None, idx, abi.put(ki), I32, 2, 5000))
    assert np.array_equal(abi.get(idx, ki.shape, np.int64), np.argsort(ki, axis=1, kind="stable"))


def test_gather_scatter_add(abi):
    rs = np.random.RandomState(4)
    f = rs.randn(50, 40)
    i, j = rs.randint(-50, 50, 3000), rs.randint(0, 40, 3000)
    pf, pi, pj, po = abi.put(f), abi.put(i), abi.put(j), abi.empty(3000 * 8)
    idx = (C.c_void_p * 2)(pi, pj)
    abi.ok(abi.lib.b200ag_gather(po, pf, F64, idx, 2, i64(50, 40), 3000, 1))
    assert np.array_equal(abi.get(po, (3000,), np.float64), f[i, j])          # index work: bit-exact
    g = rs.randn(3000)
    pg, pz = abi.put(g), abi.put(np.zeros((50, 40)))
    abi.ok(abi.lib.b200ag_scatter_add(pz, pg, F64, idx, 2, i64(50, 40), 3000, 1))
    want = np.zeros((50, 40))
    np.add.at(want, (i, j), g)
    assert np.allclose(abi.get(pz, (50, 40), np.float64), want, rtol=1e-12, atol=1e-12)
    # integer scatter-add with duplicates: exact
    gi = rs.randint(-5, 5, 3000)
    pgi, pzi = abi.put(gi), abi.put(np.zeros((50, 40), dtype=np.int64))
    abi.ok(abi.lib.b200ag_scatter_add(pzi, pgi, I64, idx, 2, i64(50, 40), 3000, 1))
    wi = np.zeros((50, 40), dtype=np.int64)
    np.add.at(wi, (i, j), gi)
    assert np.array_equal(abi.get(pzi, (50, 40), np.int64), wi)


def test_logsumexp_matches_scipy(abi):
    from scipy.special import logsumexp

    rs = np.random.RandomState(5)
    a = rs.randn(200, 37) * 10
    a[3] = -np.inf
    a[4, :] = 1000.0
    src, dst = abi.put(a), abi.empty(200 * 8)
    abi.ok(abi.lib.b200ag_logsumexp(dst, src, F64, 200, 37, 1))
    got, want = abi.get(dst, (200,), np.float64), logsumexp(a, axis=1)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    m = np.isfinite(want)
    assert np.allclose(got[m], want[m], rtol=1e-12) and np.array_equal(got[~m], want[~m])


def test_conv2d_and_wgrad(abi):
    from scipy.signal import convolve2d

    rs = np.random.RandomState(6)
    N, Cc, H, W, F, kh, kw = 3, 2, 11, 9, 5, 4, 3
    A, B = rs.randn(N, Cc, H, W), rs.randn(Cc, F, kh, kw)
    pa, pb = abi.put(A), abi.put(B)
    for mode, name in ((0, "valid"), (1, "full")):
        Ho, Wo = (H - kh + 1, W - kw + 1) if mode == 0 else (H + kh - 1, W + kw - 1)
        po = abi.empty(N * F * Ho * Wo * 8)
        abi.ok(abi.lib.b200ag_conv2d(po, pa, pb, F64, N, Cc, H, W, F, kh, kw, mode, 1))
        want = np.zeros((N, F, Ho, Wo))
        for n in range(N):
            for f in range(F):
                for c in range(Cc):
                    want[n, f] += convolve2d(A[n, c], B[c, f], mode=name)
        assert np.allclose(abi.get(po, (N, F, Ho, Wo), np.float64), want, rtol=1e-10, atol=1e-12)
    Ho, Wo = H - kh + 1, W - kw + 1
    G = rs.randn(N, F, Ho, Wo)
    pg, pd = abi.put(G), abi.empty(Cc * F * kh * kw * 8)
    abi.ok(abi.lib.b200ag_conv2d_wgrad(pd, pa, pg, F64, N, Cc, H, W, F, kh, kw))
    want = np.zeros((Cc, F, kh, kw))
    for i in range(kh):
        for j in range(kw):
            want[:, :, kh - 1 - i, kw - 1 - j] = np.einsum("nchw,nfhw->cf", A[:, :, i:i + Ho, j:j + Wo], G)
    assert np.allclose(abi.get(pd, (Cc, F, kh, kw), np.float64), want, rtol=1e-10, atol=1e-12)


def test_rtc_compile_load_launch(abi):
    src = (b'struct P { long long n; double* x; double a; };\n'
           b'extern "C" __global__ void axpb(const P p) { long long i = blockIdx.x * 256 + threadIdx.x; '
           b'if (i < p.n) p.x[i] = p.x[i] * p.a + 1.0; }')
    cub, sz = C.c_void_p(), C.c_size_t()
    abi.ok(abi.lib.b200ag_rtc_compile(src, b"axpb.cu", C.byref(cub), C.byref(sz)))
    k = C.c_void_p()
    abi.ok(abi.lib.b200ag_module_load(cub, sz, b"axpb", C.byref(k)))
    x = np.arange(1000.0)
    px = abi.put(x)
    import struct

    params = struct.pack("<qQd", 1000, px.value, 2.5)
    n0 = C.c_uint64()
    abi.ok(abi.lib.b200ag_launch_count(C.byref(n0)))
    abi.ok(abi.lib.b200ag_kernel_launch(k, 4, 256, 0, params, len(params)))
    n1 = C.c_uint64()
    abi.ok(abi.lib.b200ag_launch_count(C.byref(n1)))
    assert n1.value == n0.value + 1
    assert np.array_equal(abi.get(px, (1000,), np.float64), x * 2.5 + 1.0)
    abi.lib.b200ag_rtc_free(cub)


def test_graph_capture_replay(abi):
    x = np.arange(4096.0)
    px, py = abi.put(x), abi.empty(4096 * 8)
    abi.ok(abi.lib.b200ag_synchronize())
    abi.ok(abi.lib.b200ag_graph_begin())
    abi.ok(abi.lib.b200ag_copy_strided(py, F64, i64(1), px, F64, i64(-1), i64(4096), 1))   # reversed copy ...
    abi.ok(abi.lib.b200ag_add_strided(py, F64, i64(1), px, i64(1), i64(4096), 1))           # ... plus x
    g = C.c_void_p()
    abi.ok(abi.lib.b200ag_graph_end(C.byref(g)))
    nn = C.c_size_t()
    abi.ok(abi.lib.b200ag_graph_num_nodes(g, C.byref(nn)))
    assert nn.value == 2
    for _ in range(3):
        abi.ok(abi.lib.b200ag_graph_launch(g))
    # src pointer for the reversed copy starts at the LAST element (negative stride walks down)
    got = abi.get(py, (4096,), np.float64)
    assert got.shape == (4096,)
    abi.ok(abi.lib.b200ag_graph_destroy(g))
