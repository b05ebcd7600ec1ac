"""Development harness: run the REFERENCE's own test files with device inputs.

The reference's tests draw every input from ``autograd.numpy.random`` (tests/numpy_utils.py:1, tests/test_*.py);
this script copies the test files (never committed — they stay under /tmp) next to a conftest that installs
autograd_b200 and makes those draws return DeviceArrays, then runs pytest on them.  On a machine without a GPU the
host emulation of tests/fakedev.py stands in for the device (host logic + NumPy-protocol dispatch + rule
registration are what these tests exercise); with ``--cuda`` the production backend is used.

    python scripts/run_reference_tests_on_device.py [--cuda] [--ref /root/reference] [pytest args / test files]

Last run (round 1, fake device): 413 passed, 13 skipped, 2 failed (float16 storage; ``isinstance(x, np.ndarray)`` on a
device array).  test_complex.py: all pass.  test_scipy.py: 79 passed, 10 failed (singular-covariance multivariate_normal,
scipy.linalg / scipy.integrate, non-integer Bessel orders, complex erf).  test_linalg.py: 24 passed, 39 failed (eigh/eig/svd/pinv).
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFTEST = '''
import os, sys
import numpy as onp
import pytest

ROOT = {root!r}
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import autograd_b200 as b200

if {cuda!r}:
    b200.backend.set_backend(b200.backend.CudaBackend())
else:
    from fakedev import FakeBackend
    b200.backend.set_backend(FakeBackend(compile_generated=False))
b200.install()

import autograd.numpy.random as npr


_active = [False]     # draws at module import time (parameter tables of test_scipy.py) stay on the host


def _dev(fn):
    def wrapped(*a, **k):
        out = fn(*a, **k)
        if _active[0] and isinstance(out, onp.ndarray) and out.dtype.kind == "f":
            return b200.asarray(out)
        return out
    return wrapped


for _k in ("randn", "rand", "random", "normal", "uniform"):
    setattr(npr, _k, _dev(getattr(npr, _k)))


@pytest.fixture(autouse=True)
def random_seed():
    npr.seed(42)
    _active[0] = True
    yield
    _active[0] = False
'''

DEFAULT_FILES = ["test_systematic.py", "test_numpy.py", "test_graphs.py", "test_wrappers.py", "test_binary_ops.py",
                 "test_scalar_ops.py", "test_jacobian.py", "test_misc.py", "test_logic.py", "test_core.py",
                 "test_vspaces.py", "test_truediv.py", "test_direct.py", "test_dict.py", "test_list.py", "test_tuple.py",
                 "test_builtins.py", "test_ufunc_dispatch.py"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--ref", default="/root/reference")
    args, rest = ap.parse_known_args()
    src = os.path.join(args.ref, "tests")
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found: this harness needs the reference checkout (development container only)")
    tmp = tempfile.mkdtemp(prefix="reftests_")
    try:
        for f in glob.glob(os.path.join(src, "*.py")):
            shutil.copy(f, tmp)
        with open(os.path.join(tmp, "conftest.py"), "w") as fh:
            fh.write(CONFTEST.format(root=ROOT, cuda=args.cuda))
        files = [a for a in rest if a.endswith(".py") or "::" in a] or DEFAULT_FILES
        opts = [a for a in rest if a not in files]
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        return subprocess.call([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *opts, *files], cwd=tmp, env=env)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
