"""Offline install of the UNMODIFIED reference (HIPS/autograd 1.9.1) into baseline/_ref.

TEST / BASELINE INFRASTRUCTURE.  baseline/_ref is git-ignored (no reference source enters the history)
but travels to the GPU box with the repo snapshot, where /root/reference does not exist.

The reference's own build backend (hatchling) is not installed in this image and there is no network, so
`pip install /root/reference` fails while preparing metadata.  The package is pure Python; this script
copies the package directory to /tmp, swaps only the [build-system] table for setuptools (which IS
installed) and runs the documented command on that copy:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/<copy>

The installed files are byte-identical to /root/reference/autograd (checked below).
"""

from __future__ import annotations

import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "autograd"
version = "1.9.1"
requires-python = ">=3.10"
dependencies = ["numpy<3"]

[tool.setuptools.packages.find]
include = ["autograd*"]
"""


def _identical(a: str, b: str) -> bool:
    cmp = filecmp.dircmp(a, b, ignore=["__pycache__"])

    def ok(c):
        if c.left_only or c.right_only or c.diff_files or c.funny_files:
            return False
        return all(ok(s) for s in c.subdirs.values())

    return ok(cmp)


def main() -> int:
    have = os.path.isdir(os.path.join(TARGET, "autograd"))
    if not os.path.isdir(os.path.join(REFERENCE, "autograd")):
        print("install_reference: /root/reference not mounted;", "using existing baseline/_ref" if have else
              "baseline/_ref is ABSENT")
        return 0 if have else 1
    if have and _identical(os.path.join(REFERENCE, "autograd"), os.path.join(TARGET, "autograd")):
        print("install_reference: baseline/_ref is up to date")
        return 0
    tmp = tempfile.mkdtemp(prefix="autograd_ref_")
    try:
        shutil.copytree(os.path.join(REFERENCE, "autograd"), os.path.join(tmp, "autograd"))
        for extra in ("license.txt", "README.md"):
            if os.path.exists(os.path.join(REFERENCE, extra)):
                shutil.copy(os.path.join(REFERENCE, extra), tmp)
        with open(os.path.join(tmp, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT)
        shutil.rmtree(TARGET, ignore_errors=True)
        os.makedirs(TARGET, exist_ok=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", TARGET, tmp], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not _identical(os.path.join(REFERENCE, "autograd"), os.path.join(TARGET, "autograd")):
        print("install_reference: installed copy differs from /root/reference/autograd", file=sys.stderr)
        return 1
    print("install_reference: installed unmodified autograd 1.9.1 into baseline/_ref")
    return 0


if __name__ == "__main__":
    sys.exit(main())
