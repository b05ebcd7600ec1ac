"""GPU suite (`-m gpu`): the same parity cases through the production path — ctypes → libb200ag.so →
sm_100a kernels — against NumPy / reference autograd on identical inputs."""

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", cases.UNARY)
def test_unary_f64(name, cuda_dev):
    cases.case_unary(name)


@pytest.mark.parametrize("name", ["exp", "tanh", "log", "sqrt", "negative", "abs", "floor"])
def test_unary_f32(name, cuda_dev):
    cases.case_unary(name, np.float32)


@pytest.mark.parametrize("name", cases.BINARY)
def test_binary_f64(name, cuda_dev):
    cases.case_binary(name)


@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_ops(name, cuda_dev):
    cases.OP_CASES[name]()


@pytest.mark.parametrize("name", sorted(cases.GRAD_CASES))
def test_grads(name, cuda_dev):
    cases.GRAD_CASES[name]()


def test_native_library_is_the_one_running(cuda_dev):
    import autograd_b200 as b200

    n0 = cuda_dev.launch_count()
    x = b200.asarray(np.arange(10.0))
    (x * 2 + 1).get()
    assert cuda_dev.launch_count() > n0
    assert cuda_dev.name == "cuda"
