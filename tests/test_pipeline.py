"""Chunked host-buffer streaming (autograd_b200.map_elementwise): same numbers as one monolithic evaluation."""

import numpy as onp
import pytest

import cases


def _check(replay):
    import autograd_b200 as b200

    from examples import workloads as w

    x = onp.linspace(-7.0, 7.0, 100003)
    f = w.tanh_nth_derivative(3)
    with b200.host_arrays():
        want = f(x)
    for chunk in (1 << 14, 30000, 1 << 20):
        got = b200.map_elementwise(f, x, chunk=chunk, replay=replay)
        cases.assert_close(got, want, cases.TOL_ELEMENTWISE, f"map_elementwise chunk={chunk}")
    out = onp.empty_like(x)
    res = b200.map_elementwise(lambda v: v * 2.0 + 1.0, x, out=out, chunk=4096, replay=replay)
    assert res is out and onp.array_equal(out, x * 2.0 + 1.0)


def test_map_elementwise_host_logic(fake_dev):
    _check(False)


@pytest.mark.gpu
@pytest.mark.parametrize("replay", [False, True])
def test_map_elementwise_gpu(cuda_dev, replay):
    _check(replay)
