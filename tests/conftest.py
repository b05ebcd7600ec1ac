"""pytest configuration: paths, the `gpu` marker and the two backend fixtures.

* `fake_dev`  — tests/fakedev.FakeBackend, a host emulation used ONLY to test host logic without a GPU.
* `cuda_dev`  — the production CudaBackend (libb200ag.so on cuda:0); tests using it are marked `gpu`.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(42)


_cuda_backend = None


@pytest.fixture
def fake_dev():
    import autograd_b200 as b200
    from fakedev import FakeBackend

    prev = b200.backend._current
    fb = FakeBackend()
    b200.backend.set_backend(fb)
    b200.install()
    yield fb
    b200.backend.set_backend(prev)


@pytest.fixture
def cuda_dev():
    global _cuda_backend
    import autograd_b200 as b200

    prev = b200.backend._current
    if _cuda_backend is None:
        _cuda_backend = b200.backend.CudaBackend()
    b200.backend.set_backend(_cuda_backend)
    b200.install()
    yield _cuda_backend
    b200.backend.set_backend(prev)
