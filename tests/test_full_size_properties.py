"""Parity at the BASELINE.json FULL sizes through size-independent properties (the reference cannot run these
sizes in seconds, or at all: C2 needs 159 GB and C4 360 GB on NumPy — SURVEY.md §0-6):

* C2 2^28 points: the device result on a strided 2^16-point sample equals the reference evaluated on exactly those
  points (the workload is elementwise, so a sample IS the full computation at those points);
* C4 2048^2: forward field after one step is bit-identical to NumPy; the 20-step gradient passes a directional
  finite-difference check; the replayed graph returns the same gradient as the eager tape;
* C5 65,536 images: the batch gradient is additive over batch shards (linearity of the likelihood term) and equals
  the reference on a 64-image sub-batch;
* C3 T=256: replayed gradient equals eager gradient, is finite, keeps float32 dtype; T=2 prefix matches the reference;
* C1: full size directly against the reference (it runs in milliseconds).
"""

import numpy as onp
import pytest

import cases

pytestmark = pytest.mark.gpu


def test_c2_full_size_strided_sample(cuda_dev):
    import autograd_b200 as b200

    from examples import workloads as w

    n = 1 << 28
    x = b200.asarray(onp.linspace(-7.0, 7.0, n))
    f = w.tanh_nth_derivative(4)
    y = f(x).materialize()
    assert y.shape == (n,) and not y.is_lazy
    stride = n >> 16
    got = y[::stride].get()
    xs = onp.linspace(-7.0, 7.0, n)[::stride].copy()
    with b200.host_arrays():
        want = f(xs)
    cases.assert_close(got, want, cases.TOL_ELEMENTWISE, "C2 full size, 2^16 strided sample")
    # odd symmetry of tanh'''' about 0 holds to rounding on the whole vector: checksum of checksums
    tot = float(onp.abs(b200.array.reduce_(y + y[::-1], "sum", None, False).get()))
    scale = float(b200.array.reduce_(abs(y), "sum", None, False).get())
    assert tot <= 1e-9 * scale
    del x, y
    cuda_dev.empty_cache()


def test_c4_full_grid_forward_bit_exact_and_gradient_check(cuda_dev):
    import autograd_b200 as b200
    from autograd import value_and_grad

    from examples import workloads as w

    rows = cols = 2048
    p, s0, tg = w.fluid_data(rows, cols)
    with b200.host_arrays():
        want = w.fluid_simulate(p[: rows * cols].reshape(rows, cols), p[rows * cols:].reshape(rows, cols), s0, 1)
    dp, ds, dt = b200.asarray(p), b200.asarray(s0), b200.asarray(tg)
    got = w.fluid_simulate(dp[: rows * cols].reshape(rows, cols), dp[rows * cols:].reshape(rows, cols), ds, 1).get()
    assert onp.array_equal(got, want), "C4 forward at 2048^2 is not bit-identical to NumPy"

    steps = 20
    vg = value_and_grad(w.fluid_objective)
    val, g = vg(dp, ds, dt, rows, cols, steps)
    g.materialize()
    rs = onp.random.RandomState(5)
    v = b200.asarray(rs.randn(p.size))
    eps = 1e-6   # the interpolation is piecewise linear: a small step keeps almost every cell on its own linear piece
    fp = float(w.fluid_objective(dp + eps * v, ds, dt, rows, cols, steps).get())
    fm = float(w.fluid_objective(dp - eps * v, ds, dt, rows, cols, steps).get())
    fd = (fp - fm) / (2 * eps)
    an = float(b200.array.dot(g, v).get())
    assert abs(fd - an) <= 5e-3 * max(abs(fd), abs(an)) + 1e-12, (fd, an)

    fast = b200.capture(vg)
    v2, g2 = fast(dp, ds, dt, rows, cols, steps)
    cases.assert_close(g2.get(), g.get(), cases.TOL_REDUCE, "C4 replayed vs eager gradient")
    cases.assert_close(v2.get(), val.get(), cases.TOL_REDUCE, "C4 replayed vs eager value")
    del fast, g, g2, v
    cuda_dev.empty_cache()


def test_c5_full_batch_additivity(cuda_dev):
    import autograd_b200 as b200
    from autograd import grad

    from examples import workloads as w

    n, _pred, loss = w.convnet_make()
    N = 65536
    W, X, T = w.convnet_data(N, n)
    dW, dX, dT = b200.asarray(W), b200.asarray(X), b200.asarray(T)
    g = grad(loss)
    full = g(dW, dX, dT, 1.0).get()
    assert full.shape == (n,) and onp.all(onp.isfinite(full))
    h = N // 2
    a = g(dW, dX[:h], dT[:h], 0.5).get()
    b = g(dW, dX[h:], dT[h:], 0.5).get()
    # the example's softmax shifts by the batch max (convnet.py:40-42), so shards differ from the full batch by rounding only
    cases.assert_close(a + b, full, 1e-9, "C5 gradient is additive over batch shards")
    with b200.host_arrays():
        want = g(W, X[:64], T[:64], 1.0)
    cases.assert_close(g(dW, dX[:64], dT[:64], 1.0).get(), want, cases.TOL_REDUCE, "C5 64-image sub-batch vs reference")
    cuda_dev.empty_cache()


def test_c3_full_length_replay_matches_eager(cuda_dev):
    import autograd_b200 as b200
    from autograd import grad

    from examples import workloads as w

    params, X, Y = w.lstm_data(seq_len=256)
    dp, dX = cases.to_dev(params), b200.asarray(X)
    g = grad(w.lstm_objective)
    eager = cases.to_host(g(dp, dX, dX))
    fast = b200.capture(g)
    replay = cases.to_host(fast(dp, dX, dX))
    for k in eager:
        assert eager[k].dtype == onp.float32 and onp.all(onp.isfinite(eager[k]))
        cases.assert_close(replay[k], eager[k], cases.TOL_F32, f"C3 T=256 replay vs eager {k}")
    del fast
    p2, X2, _ = w.lstm_data(seq_len=2)
    with b200.host_arrays():
        want = g(p2, X2, X2)
    got = cases.to_host(g(cases.to_dev(p2), b200.asarray(X2), b200.asarray(X2)))
    for k in want:
        cases.assert_close(got[k], want[k], cases.TOL_F32, f"C3 full width, T=2 vs reference {k}")
    cuda_dev.empty_cache()


def test_c1_full_size_vs_reference(cuda_dev):
    import autograd_b200 as b200
    from autograd import grad

    from examples import workloads as w

    params, X, T = w.mlp_data()
    g = grad(w.mlp_objective)
    with b200.host_arrays():
        want = g(params, X, T, 1.0)
    got = cases.to_host(g(cases.to_dev(params), b200.asarray(X), b200.asarray(T), 1.0))
    for a, r in zip(cases.leaves(got), cases.leaves(want)):
        cases.assert_close(a, r, cases.TOL_REDUCE, "C1 full size")
