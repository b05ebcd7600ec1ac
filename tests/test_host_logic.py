"""CPU suite: host logic of the backend (graph building, dtype/shape rules, layout bookkeeping, NumPy
protocol dispatch, autograd registration) exercised on tests/fakedev.FakeBackend.  Generated CUDA
sources are still compiled with NVRTC (offline) so code-generation errors fail here."""

import pytest

import cases


@pytest.mark.parametrize("name", cases.UNARY)
def test_unary_f64(name, fake_dev):
    cases.case_unary(name)


@pytest.mark.parametrize("name", ["exp", "tanh", "log", "sqrt", "negative", "abs", "floor"])
def test_unary_f32(name, fake_dev):
    import numpy as np

    cases.case_unary(name, np.float32)


@pytest.mark.parametrize("name", cases.BINARY)
def test_binary_f64(name, fake_dev):
    cases.case_binary(name)


@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_ops(name, fake_dev):
    cases.OP_CASES[name]()


@pytest.mark.parametrize("name", sorted(cases.GRAD_CASES))
def test_grads(name, fake_dev):
    cases.GRAD_CASES[name]()
