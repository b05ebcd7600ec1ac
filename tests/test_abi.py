"""The C-ABI library loads and exports every symbol include/b200ag.h declares (no GPU needed)."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200ag.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ag_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    names = _declared()
    assert len(names) >= 40 and "b200ag_gemm" in names and "b200ag_rtc_compile" in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "autograd_b200", "libb200ag.so"))
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in include/b200ag.h but not exported: {missing}"


def test_ctypes_prototypes_cover_the_header():
    from autograd_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == _declared()


def test_no_cpu_fallback_when_library_is_missing(tmp_path, monkeypatch):
    """Importing the binding with the .so absent must raise, not degrade."""
    import importlib

    from autograd_b200 import _lib

    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib._load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("loading a missing library did not fail")
    importlib.reload(_lib)


def test_nvrtc_compiles_for_sm100a_without_a_gpu():
    from autograd_b200 import _lib

    src = b'extern "C" __global__ void k(double* x, long long n) { long long i = blockIdx.x * 256 + threadIdx.x; if (i < n) x[i] = x[i] * 2.0 + 1.0; }'
    p, sz = ctypes.c_void_p(), ctypes.c_size_t()
    assert _lib.lib.b200ag_rtc_compile(src, b"k.cu", ctypes.byref(p), ctypes.byref(sz)) == 0
    cubin = ctypes.string_at(p.value, sz.value)
    _lib.lib.b200ag_rtc_free(p)
    assert cubin[:4] == b"\x7fELF"
