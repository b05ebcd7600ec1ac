"""The NumPy-only port (oracle/numpy_port.py) is pinned against golden vectors generated from the real reference
(CPU), and the device path is checked against the port (GPU)."""

import os
import sys

import numpy as onp
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_port_mlp_gradient_matches_golden():
    import numpy_port

    from examples import workloads as w

    gold = onp.load(os.path.join(GOLD, "c1_mlp_small.npz"))
    params, X, T = w.mlp_data(batch=32, layer_sizes=(20, 16, 12, 10))
    value, grads = numpy_port.mlp_objective_and_grad(params, X, T, 1.0)
    assert abs(value - float(gold["value"])) <= 1e-12 * abs(float(gold["value"]))
    for i, wb in enumerate(grads):
        for j, a in enumerate(wb):
            cases.assert_close(a, gold[f"g{i}_{j}"], 1e-11, f"port vs golden g{i}_{j}")


def test_port_tanh_closed_forms_match_golden():
    import numpy_port

    gold = onp.load(os.path.join(GOLD, "c2_tanh.npz"))
    for k, d in enumerate(numpy_port.tanh_derivatives(gold["x"])):
        cases.assert_close(d, gold[f"d{k}"], 1e-11, f"closed form vs golden, order {k}")


def test_port_convolution_matches_reference_primitive(fake_dev):
    import autograd_b200 as b200
    import numpy_port

    from autograd.scipy.signal import convolve

    r = onp.random.RandomState(0)
    A, B = r.randn(3, 2, 9, 8), r.randn(2, 4, 3, 3)
    with b200.host_arrays():
        want = convolve(A, B, axes=([2, 3], [2, 3]), dot_axes=([1], [0]), mode="valid")
    cases.assert_close(numpy_port.conv_valid_nchw(A, B), want, 1e-12, "port conv vs reference")


@pytest.mark.gpu
def test_device_matches_port(cuda_dev):
    import autograd_b200 as b200
    import numpy_port
    from autograd import grad

    from examples import workloads as w

    params, X, T = w.mlp_data()
    value, grads = numpy_port.mlp_objective_and_grad(params, X, T, 1.0)
    got = cases.to_host(grad(w.mlp_objective)(cases.to_dev(params), b200.asarray(X), b200.asarray(T), 1.0))
    for a, r in zip(cases.leaves(got), cases.leaves(grads)):
        cases.assert_close(a, r, cases.TOL_REDUCE, "device MLP gradient vs hand-written NumPy backprop")
    x = onp.linspace(-7.0, 7.0, 1 << 16)
    for k, d in enumerate(numpy_port.tanh_derivatives(x)):
        cases.assert_close(w.tanh_nth_derivative(k)(b200.asarray(x)).get(), d, 1e-11, f"device vs closed form, order {k}")


def test_port_weighted_logsumexp_matches_golden():
    import numpy_port

    gold = onp.load(os.path.join(GOLD, "a13_logsumexp_weighted.npz"))
    x, b = gold["lse_x"], gold["lse_b"]
    for axis in (None, 0, 1):
        tag = "all" if axis is None else str(axis)
        val, grad = numpy_port.logsumexp_weighted_and_grad(x, b, axis)
        cases.assert_close(onp.asarray(val), gold[f"lse_val_{tag}"], 1e-12, f"port weighted logsumexp {tag}")
        cases.assert_close(grad, gold[f"lse_grad_{tag}"], 1e-11, f"port weighted logsumexp gradient {tag}")


def test_port_direct_convolution_matches_golden():
    """The explicit-sum convolution of the port reproduces the reference's strided-view einsum on every axis layout and
    mode of the a14 fixture — two independent formulations of scipy/signal.py:10-76 agreeing pins both."""
    import numpy_port

    from test_golden import A14_CONFIGS

    gold = onp.load(os.path.join(GOLD, "a14_convolve_general.npz"))
    for name, (axes, dot_axes) in A14_CONFIGS.items():
        A, B = gold[f"{name}_A"], gold[f"{name}_B"]
        for mode in ("valid", "full", "same"):
            got = numpy_port.convolve_direct(A, B, axes=axes, dot_axes=dot_axes or ((), ()), mode=mode)
            cases.assert_close(got, gold[f"{name}_{mode}_val"], 1e-12, f"port convolve {name} {mode}")
