"""On-disk cache of NVRTC-compiled cubins for generated elementwise kernels.

Keyed by the SHA-256 of the generated source (which already encodes dtypes,
layout classes and baked constants) plus the library build.  Two directories are
consulted: an in-tree one (``autograd_b200/_kcache``; git-ignored ``*.cubin``
files that travel with the repo snapshot) and a per-user one that is writable at
run time.  Set ``B200AG_KERNEL_CACHE=0`` to disable.
"""

from __future__ import annotations

import hashlib
import os

HERE = os.path.dirname(os.path.abspath(__file__))
INTREE_DIR = os.path.join(HERE, "_kcache")
_VERSION = "1"


def _enabled() -> bool:
    return os.environ.get("B200AG_KERNEL_CACHE", "1") != "0"


def _user_dir() -> str:
    d = os.environ.get("B200AG_KERNEL_CACHE_DIR")
    if not d:
        d = os.path.join(os.path.expanduser("~"), ".cache", "autograd_b200")
    return d


def _name(source: str) -> str:
    return hashlib.sha256((_VERSION + source).encode()).hexdigest()[:32] + ".cubin"


def load(source: str):
    if not _enabled():
        return None
    name = _name(source)
    for d in (INTREE_DIR, _user_dir()):
        p = os.path.join(d, name)
        try:
            with open(p, "rb") as f:
                return f.read()
        except OSError:
            continue
    return None


def store(source: str, cubin: bytes, intree: bool = False) -> None:
    if not _enabled():
        return
    d = INTREE_DIR if intree else _user_dir()
    try:
        os.makedirs(d, exist_ok=True)
        tmp = os.path.join(d, _name(source) + f".tmp{os.getpid()}")
        with open(tmp, "wb") as f:
            f.write(cubin)
        os.replace(tmp, os.path.join(d, _name(source)))
    except OSError:
        pass
