"""CUDA-graph tape replay (autograd_b200.capture) — GPU only."""

import numpy as onp
import pytest

import cases

pytestmark = pytest.mark.gpu


def test_replayed_mlp_grad_matches_eager_and_reference(cuda_dev):
    import autograd_b200 as b200
    from autograd import grad

    from examples import workloads as w

    params, X, T = w.mlp_data(batch=32, layer_sizes=(20, 16, 12, 10))
    g = grad(w.mlp_objective)
    with b200.host_arrays():
        want = g(params, X, T, 1.0)
    fast = b200.capture(g)
    dp, dX, dT = cases.to_dev(params), b200.asarray(X), b200.asarray(T)
    n0 = cuda_dev.launch_count()
    for _ in range(3):
        got = cases.to_host(fast(dp, dX, dT, 1.0))
        for a, r in zip(cases.leaves(got), cases.leaves(want)):
            cases.assert_close(a, r, cases.TOL_REDUCE, "replayed C1 grad")
    assert cuda_dev.launch_count() > n0
    # new input VALUES, same signature: the recording is replayed on refreshed inputs
    params2, X2, T2 = w.mlp_data(batch=32, layer_sizes=(20, 16, 12, 10), seed=7)
    with b200.host_arrays():
        want2 = g(params2, X2, T2, 1.0)
    got2 = cases.to_host(fast(cases.to_dev(params2), b200.asarray(X2), b200.asarray(T2), 1.0))
    for a, r in zip(cases.leaves(got2), cases.leaves(want2)):
        cases.assert_close(a, r, cases.TOL_REDUCE, "replayed C1 grad, new inputs")
    assert len(fast.num_nodes()) == 1
    # a different batch size is a new signature -> second recording
    params3, X3, T3 = w.mlp_data(batch=16, layer_sizes=(20, 16, 12, 10), seed=3)
    with b200.host_arrays():
        want3 = g(params3, X3, T3, 1.0)
    got3 = cases.to_host(fast(cases.to_dev(params3), b200.asarray(X3), b200.asarray(T3), 1.0))
    for a, r in zip(cases.leaves(got3), cases.leaves(want3)):
        cases.assert_close(a, r, cases.TOL_REDUCE, "replayed C1 grad, new signature")
    assert len(fast.num_nodes()) == 2


def test_replayed_lstm_and_fluid(cuda_dev):
    import autograd_b200 as b200
    from autograd import grad, value_and_grad

    from examples import workloads as w

    params, X, Y = w.lstm_data(seq_len=3, batch=8, hidden=16, chars=12)
    g = grad(w.lstm_objective)
    with b200.host_arrays():
        want = g(params, X, Y)
    fast = b200.capture(g)
    for _ in range(2):
        got = cases.to_host(fast(cases.to_dev(params), b200.asarray(X), b200.asarray(Y)))
        for k in want:
            cases.assert_close(got[k], want[k], cases.TOL_F32, f"replayed C3 {k}")

    rows = cols = 16
    p, s0, tg = w.fluid_data(rows, cols)
    vg = value_and_grad(w.fluid_objective)
    with b200.host_arrays():
        vw, gw = vg(p, s0, tg, rows, cols, 2)
    fastf = b200.capture(vg)
    for _ in range(2):
        v, gg = fastf(b200.asarray(p), b200.asarray(s0), b200.asarray(tg), rows, cols, 2)
        cases.assert_close(gg.get(), gw, cases.TOL_REDUCE, "replayed C4 grad")
        cases.assert_close(v.get(), onp.asarray(vw), cases.TOL_REDUCE, "replayed C4 value")


def test_value_dependent_control_flow_is_rejected(cuda_dev):
    import autograd.numpy as np
    import autograd_b200 as b200

    def f(x):
        if float(np.sum(x)) > 0:  # reads a device value on the host
            return x * 2
        return x * 3

    fast = b200.capture(f)
    with pytest.raises(RuntimeError, match="captured"):
        fast(b200.asarray(onp.ones(8)))
    # the backend must be usable again afterwards
    assert float((b200.asarray(onp.ones(4)) * 2).get().sum()) == 8.0


def test_replayed_float32_net_uses_tensor_core_gemm(cuda_dev):
    """A float32 net whose dot VJPs run on the tcgen05 kernels (TMA descriptors, cluster launch, split-K workspace)
    must be capturable and replayable like everything else."""
    import autograd.numpy as np
    import autograd_b200 as b200
    from autograd import grad

    r = onp.random.RandomState(3)
    X = (r.randn(640, 512) * 0.5).astype(onp.float32)
    W1 = (r.randn(512, 384) * 0.05).astype(onp.float32)
    W2 = (r.randn(384, 256) * 0.05).astype(onp.float32)

    def loss(W1, W2, X):
        return np.sum(np.dot(np.tanh(np.dot(X, W1)), W2) ** 2) / X.shape[0]

    g = grad(loss, (0, 1))
    with b200.host_arrays():
        want = g(W1, W2, X)
    fast = b200.capture(g)
    for scale in (1.0, 1.0, 0.5):
        got = fast(b200.asarray(W1 * onp.float32(scale)), b200.asarray(W2), b200.asarray(X))
        if scale != 1.0:
            with b200.host_arrays():
                want = g(W1 * onp.float32(scale), W2, X)
        for a, w_ in zip(got, want):
            assert a.dtype == onp.float32
            cases.assert_close(a.get(), w_, cases.TOL_F32, "replayed float32 net gradient")
    assert len(fast.num_nodes()) == 1


def test_replay_skips_unchanged_inputs_but_sees_every_change(cuda_dev):
    """Inputs are re-copied into the recording only when they changed: the same unmodified array costs nothing, an
    in-place write (buffer version) or a new array (even one recycled at the same address) is picked up."""
    import autograd.numpy as np
    import autograd_b200 as b200
    from autograd import grad

    f = b200.capture(grad(lambda x: np.sum(np.sin(x) * x)))
    ref = lambda v: onp.cos(v) * v + onp.sin(v)
    h = onp.linspace(0.0, 3.0, 4096)
    x = b200.asarray(h)
    cases.assert_close(f(x).get(), ref(h), cases.TOL_REDUCE, "first replay")
    n0 = cuda_dev.launch_count()
    f(x)
    same = cuda_dev.launch_count() - n0
    x[5] = 7.5                                       # in-place write
    h2 = h.copy(); h2[5] = 7.5
    n0 = cuda_dev.launch_count()
    cases.assert_close(f(x).get(), ref(h2), cases.TOL_REDUCE, "after an in-place write")
    assert cuda_dev.launch_count() - n0 >= same      # the refresh copy is back
    for k in range(4):                               # fresh arrays, freed in between: addresses get recycled
        hk = h + k
        xk = b200.asarray(hk)
        cases.assert_close(f(xk).get(), ref(hk), cases.TOL_REDUCE, f"new array {k}")
        del xk
