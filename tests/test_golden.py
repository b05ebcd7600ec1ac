"""Golden vectors generated from the REAL reference examples (oracle/make_golden.py, run where
/root/reference is mounted).

* CPU (`-m "not gpu"`): the restated workloads in examples/workloads.py, run through unmodified
  autograd on NumPy, reproduce the golden outputs BIT FOR BIT — this pins the oracle used everywhere
  else (reference autograd on NumPy + these workload functions) to the reference's own example code.
* GPU (`-m gpu`): the device path reproduces the same golden outputs within the north_star tolerances.
"""

import os

import numpy as onp
import pytest

import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return onp.load(os.path.join(GOLD, name))


def _run_all(asarray, getter, exact):
    """Evaluate every golden case; `asarray` lifts host inputs, `getter` lowers results."""
    import autograd_b200 as b200
    from autograd import grad, value_and_grad

    from examples import workloads as w

    tol_e = 0 if exact else cases.TOL_ELEMENTWISE
    tol_r = 0 if exact else cases.TOL_REDUCE
    tol_f = 0 if exact else cases.TOL_F32
    lift = lambda t: cases.to_dev(t) if asarray else t

    # C1 small + full-size summary
    gold = _load("c1_mlp_small.npz")
    params, X, T = w.mlp_data(batch=32, layer_sizes=(20, 16, 12, 10))
    g = getter(grad(w.mlp_objective)(lift(params), lift(X), lift(T), 1.0))
    for i, wb in enumerate(g):
        for j, a in enumerate(wb):
            cases.assert_close(a, gold[f"g{i}_{j}"], tol_r, f"C1 g{i}_{j}")
    gold = _load("c1_mlp_full_summary.npz")
    params, X, T = w.mlp_data()
    val, g = value_and_grad(w.mlp_objective)(lift(params), lift(X), lift(T), 1.0)
    flat = onp.concatenate([onp.asarray(a).ravel() for wb in getter(g) for a in wb])
    assert flat.size == int(gold["n"])
    cases.assert_close(flat[::997], gold["sample"], tol_r, "C1 full-size gradient sample")
    assert abs(flat.sum() - float(gold["sum"])) <= max(tol_r, 1e-15) * onp.abs(flat).sum()
    assert abs((flat * flat).sum() - float(gold["sumsq"])) <= max(tol_r, 1e-15) * float(gold["sumsq"])
    assert abs(float(getter(val)) - float(gold["value"])) <= max(tol_r, 1e-15) * abs(float(gold["value"]))

    # C2 orders 0..4
    gold = _load("c2_tanh.npz")
    x = gold["x"]
    for order in range(5):
        f = w.tanh_nth_derivative(order)
        cases.assert_close(getter(f(lift(x))), gold[f"d{order}"], tol_e, f"C2 order {order}")

    # C3
    gold = _load("c3_lstm_small.npz")
    params, X, Y = w.lstm_data(seq_len=3, batch=8, hidden=16, chars=12)
    g = getter(grad(w.lstm_objective)(lift(params), lift(X), lift(Y)))
    for k, v in g.items():
        cases.assert_close(v, gold[k.replace(" ", "_")], tol_f, f"C3 {k}")

    # C4: gradient within tolerance, forward smoke field bit-exact even on the device
    gold = _load("c4_fluid_small.npz")
    rows = cols = 16
    p, s0, tg = w.fluid_data(rows, cols)
    val, g = value_and_grad(w.fluid_objective)(lift(p), lift(s0), lift(tg), rows, cols, 2)
    cases.assert_close(getter(g), gold["grad"], tol_r, "C4 grad")
    cases.assert_close(onp.asarray(getter(val)), gold["value"], tol_r, "C4 value")
    dp = lift(p)
    smoke = w.fluid_simulate(dp[: rows * cols].reshape(rows, cols), dp[rows * cols:].reshape(rows, cols), lift(s0), 2)
    cases.assert_close(getter(smoke), gold["smoke"], 0, "C4 forward bit-exact")

    # C5
    gold = _load("c5_convnet_small.npz")
    n, _pred, loss = w.convnet_make()
    assert n == int(gold["n_weights"])
    W, X, T = w.convnet_data(6, n)
    cases.assert_close(getter(grad(loss)(lift(W), lift(X), lift(T))), gold["grad"], tol_r, "C5 grad")


A14_CONFIGS = {"1d": (None, None), "swap": (None, None), "axes": (([1, 2], [1, 0]), None),
               "dot": (([1, 3], [2, 3]), ([0], [0])), "nchw": (([2, 3], [2, 3]), ([1], [0]))}


def _run_a13_a14(lift, getter):
    """Rows a13 / a14 beyond the five configs: weighted logsumexp and general N-d convolve against fixtures produced by
    the reference itself (oracle/make_golden.py)."""
    from autograd import grad
    from autograd.scipy.signal import convolve
    from autograd.scipy.special import logsumexp

    import autograd.numpy as np

    gold = _load("a13_logsumexp_weighted.npz")
    x, b = gold["lse_x"], gold["lse_b"]
    for axis in (None, 0, 1):
        tag = "all" if axis is None else str(axis)
        cases.assert_close(onp.asarray(getter(logsumexp(lift(x), axis=axis, b=lift(b)))), gold[f"lse_val_{tag}"], cases.TOL_REDUCE, f"a13 value {tag}")
        g = grad(lambda x_: np.sum(logsumexp(x_, axis=axis, b=lift(b)) ** 2))(lift(x))
        cases.assert_close(getter(g), gold[f"lse_grad_{tag}"], cases.TOL_REDUCE, f"a13 grad {tag}")
    gold = _load("a14_convolve_general.npz")
    for name, (axes, dot_axes) in A14_CONFIGS.items():
        A, B = gold[f"{name}_A"], gold[f"{name}_B"]
        for mode in ("valid", "full", "same"):
            kw = dict(axes=axes, dot_axes=dot_axes, mode=mode)
            cases.assert_close(getter(convolve(lift(A), lift(B), **kw)), gold[f"{name}_{mode}_val"], cases.TOL_REDUCE, f"a14 {name} {mode}")
            loss = lambda A_, B_: np.sum(np.sin(convolve(A_, B_, **kw)))
            cases.assert_close(getter(grad(loss, 0)(lift(A), lift(B))), gold[f"{name}_{mode}_gA"], cases.TOL_REDUCE, f"a14 {name} {mode} dA")
            cases.assert_close(getter(grad(loss, 1)(lift(A), lift(B))), gold[f"{name}_{mode}_gB"], cases.TOL_REDUCE, f"a14 {name} {mode} dB")


def test_host_logic_matches_golden_a13_a14(fake_dev):
    _run_a13_a14(cases.to_dev, cases.to_host)


@pytest.mark.gpu
def test_device_matches_golden_a13_a14(cuda_dev):
    _run_a13_a14(cases.to_dev, cases.to_host)


def test_oracle_restated_workloads_match_reference_examples_bit_for_bit(fake_dev):
    import autograd_b200 as b200

    with b200.host_arrays(), onp.errstate(all="ignore"):
        _run_all(False, lambda t: t, exact=True)


def test_host_logic_matches_golden(fake_dev):
    _run_all(True, cases.to_host, exact=False)


@pytest.mark.gpu
def test_device_matches_golden(cuda_dev):
    _run_all(True, cases.to_host, exact=False)


def _check_summary(got_flat, gold, prefix, tol, what):
    """Compare a full-size result with the reference's summary of it: strided sample, sum, sum of squares."""
    got_flat = onp.asarray(got_flat, dtype=onp.float64).ravel()
    assert got_flat.size == int(gold[prefix + "n"]), f"{what}: size"
    stride = int(gold[prefix + "stride"])
    cases.assert_close(got_flat[::stride], gold[prefix + "sample"], tol, f"{what} sample")
    assert abs(got_flat.sum() - float(gold[prefix + "sum"])) <= tol * float(gold[prefix + "abssum"]), f"{what}: sum"
    assert abs((got_flat * got_flat).sum() - float(gold[prefix + "sumsq"])) <= 2 * tol * float(gold[prefix + "sumsq"]), f"{what}: sumsq"


def _run_c4_256(lift, getter):
    """C4 at 256 x 256, 20 steps against the reference's own run of examples/fluidsim/fluidsim.py (oracle/make_golden.py):
    gradient at the 1e-10 contract, forward smoke field bit for bit."""
    from autograd import value_and_grad

    from examples import workloads as w

    gold = _load("c4_fluid_256x20_summary.npz")
    rows = cols = 256
    p, s0, tg = w.fluid_data(rows, cols)
    val, g = value_and_grad(w.fluid_objective)(lift(p), lift(s0), lift(tg), rows, cols, 20)
    _check_summary(getter(g), gold, "", cases.TOL_REDUCE, "C4 256x256x20 gradient")
    assert abs(float(getter(val)) - float(gold["value"])) <= cases.TOL_REDUCE * abs(float(gold["value"]))
    dp = lift(p)
    smoke = onp.asarray(getter(w.fluid_simulate(dp[: rows * cols].reshape(rows, cols), dp[rows * cols:].reshape(rows, cols),
                                                lift(s0), 20))).ravel()
    assert onp.array_equal(smoke[:: int(gold["smoke_stride"])], gold["smoke_sample"]), "C4 forward not bit-identical"
    assert smoke.sum() == float(gold["smoke_sum"]) and (smoke * smoke).sum() == float(gold["smoke_sumsq"])


def test_host_logic_matches_golden_c4_256(fake_dev):
    _run_c4_256(cases.to_dev, cases.to_host)


@pytest.mark.gpu
def test_device_matches_full_size_golden(cuda_dev):
    """BASELINE-size (or host-feasible) runs of the reference's example code, summarised by oracle/make_golden.py:
    C3 LSTM at T=256 / B=1024 / H=512 (float32 gradients), C5 convnet at batch 2048, C4 fluid at 256 x 256 x 20."""
    from autograd import value_and_grad

    from examples import workloads as w

    _run_c4_256(cases.to_dev, cases.to_host)

    gold = _load("c5_convnet_2048_summary.npz")
    n, _pred, loss = w.convnet_make()
    W, X, T = w.convnet_data(2048, n)
    val, g = value_and_grad(loss)(cases.to_dev(W), cases.to_dev(X), cases.to_dev(T))
    _check_summary(g.get(), gold, "", cases.TOL_REDUCE, "C5 batch-2048 gradient")
    assert abs(float(val) - float(gold["value"])) <= cases.TOL_REDUCE * abs(float(gold["value"]))
    del val, g, W, X, T

    gold = _load("c3_lstm_full_summary.npz")
    params, X, Y = w.lstm_data()
    dX = cases.to_dev(X)
    val, g = value_and_grad(w.lstm_objective)(cases.to_dev(params), dX, dX)
    for k, v in g.items():
        assert v.dtype == onp.float32, k
        _check_summary(v.get(), gold, k.replace(" ", "_") + "__", cases.TOL_F32, f"C3 full-size d{k}")
    assert abs(float(val) - float(gold["value"])) <= cases.TOL_REDUCE * abs(float(gold["value"]))


def _run_f4(lift, getter, exact):
    """examples/rnn.py, bayesian_neural_net.py and black_box_svi.py against fixtures produced by the example files
    themselves (oracle/make_golden.py)."""
    import autograd.numpy as np
    from autograd import grad, value_and_grad

    from examples import workloads as w

    gold = _load("f4_extra_workloads.npz")
    tol = 0 if exact else cases.TOL_REDUCE
    params, X, Y = w.rnn_data()
    val, g = value_and_grad(w.rnn_objective)(lift(params), lift(X), lift(Y))
    assert abs(float(getter(val)) - float(gold["rnn_value"])) <= max(tol, 0) * abs(float(gold["rnn_value"]))
    for k, v in getter(g).items():
        cases.assert_close(v, gold["rnn_" + k.replace(" ", "_")], tol, f"rnn.py d{k}")
    rbf = lambda x: np.exp(-(x**2))
    nw, _pred, logprob = w.bnn_make_nn_funs([1, 20, 20, 1], 0.1, 0.01, nonlinearity=rbf)
    inputs, targets = w.bnn_toy_dataset()
    obj = w.svi_objective(lambda weights, t: logprob(weights, lift(inputs), lift(targets)), nw, 20)
    cases.assert_close(getter(grad(obj)(lift(gold["bnn_init"]), 0)), gold["bnn_grad"], tol, "bayesian_neural_net.py gradient")
    obj = w.svi_objective(w.svi_funnel_log_density, 2, 2000)
    cases.assert_close(getter(grad(obj)(lift(gold["svi_init"]), 0)), gold["svi_grad"], tol, "black_box_svi.py gradient")


def test_oracle_restated_extra_workloads_match_reference_examples_bit_for_bit(fake_dev):
    import autograd_b200 as b200

    with b200.host_arrays(), onp.errstate(all="ignore"):
        _run_f4(lambda t: t, lambda t: t, exact=True)


def test_host_logic_matches_golden_f4(fake_dev):
    _run_f4(cases.to_dev, cases.to_host, exact=False)


@pytest.mark.gpu
def test_device_matches_golden_f4(cuda_dev):
    _run_f4(cases.to_dev, cases.to_host, exact=False)
