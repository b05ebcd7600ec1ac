"""autograd_b200 — a B200-native (sm_100a) device array backend for HIPS/autograd.

    import autograd_b200 as b200
    b200.install()                      # register DeviceArray with autograd's Box / VSpace / defvjp tables
    import autograd.numpy as np
    from autograd import grad
    x = b200.asarray(host_array)        # device-resident value; np.* and grad() now run as CUDA kernels
    g = grad(f)(x).get()                # explicit copy back to a numpy.ndarray

There is no CPU fallback: every array operation is a kernel in libb200ag.so (or a
kernel generated and compiled for sm_100a at run time); importing the backend
without the built library raises.
"""

from . import array, backend, lazy
from . import linalg  # noqa: F401  (registers np.linalg.solve/inv/det/slogdet/cholesky for device arrays)
from .array import DeviceArray, asarray, to_device
from .checkpoint import checkpoint
from .complex_array import ComplexDeviceArray
from .graph import capture
from .install import host_arrays, install, is_installed
from .pipeline import map_elementwise


def synchronize() -> None:
    backend.get().synchronize()


def to_host(x):
    return x.get() if isinstance(x, (DeviceArray, ComplexDeviceArray)) else x


__all__ = ["ComplexDeviceArray", "map_elementwise", "capture", "checkpoint", "host_arrays", "DeviceArray", "asarray", "to_device", "to_host", "install", "is_installed", "synchronize", "array",
           "backend", "lazy"]
