"""DeviceArray — the device-resident value type autograd traces.

It is what NumPy hands control to at the reference's raw-call boundary
(autograd/tracer.py:94: ``f_raw(*args)`` on unboxed values) through
``__array_ufunc__`` (NEP 13) and ``__array_function__`` (NEP 18), and what the
VJP closures of autograd/numpy/numpy_vjps.py operate on through plain Python
operators (``-g * x / y**2`` ...).  Shape/dtype/stride metadata lives on the
host; payloads live in HBM and are only produced by libb200ag.so kernels.
"""

from __future__ import annotations

import numbers
import weakref
import operator

import numpy as np

from . import backend, lazy
from .lazy import BOOL, F32, F64, I32, I64, Const, Mat, Node, prod

_Box = None  # autograd.tracer.Box, bound by install() (kept lazy so importing this module never needs autograd)


def _is_box(x) -> bool:
    return _Box is not None and isinstance(x, _Box)


# set by autograd_b200.complex_array (complex values are pairs of real device arrays composed from the ops of this module)
_Complex = None
_complex_ufunc = None
_complex_full = None
_complex_where = None


class _ComplexOperand(TypeError):
    """Raised by ``_operand`` for a complex operand; ``elementwise`` reroutes the call to complex_array.ufunc_call."""


class DeviceArray:
    __slots__ = ("_node", "_weak", "__weakref__")
    __array_priority__ = 1000.0

    def __init__(self, node: Node):
        self._node = node
        # True only for the symbolic result of a REBOUND creation function (``autograd.numpy.zeros/ones/full``,
        # install.py) that no operation has touched yet: such a value carries no evidence of where the caller
        # computes, so combined with host-only operands it behaves as the host array NumPy would have made
        # (``_host_only_operands``) — host-only code keeps host results while the backend is installed.
        self._weak = False
        node.nhandles += 1

    def __del__(self):
        try:
            self._node.nhandles -= 1
        except AttributeError:
            pass

    # ------------------------------------------------------------ metadata (pure host, never syncs)
    @property
    def shape(self):
        return self._node.shape

    @property
    def dtype(self):
        return self._node.dtype

    @property
    def ndim(self):
        return len(self._node.shape)

    @property
    def size(self):
        return prod(self._node.shape)

    @property
    def itemsize(self):
        return self._node.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self._node.dtype.itemsize

    @property
    def T(self):
        return transpose(self)

    @property
    def is_lazy(self) -> bool:
        return self._node.mat is None

    def __len__(self):
        if not self._node.shape:
            raise TypeError("len() of unsized object")
        return self._node.shape[0]

    # ------------------------------------------------------------ materialisation / host transfer
    def _mat(self) -> Mat:
        node = self._node
        return node.mat if node.mat is not None else lazy.materialize(node)

    def _cmat(self) -> Mat:
        """C-contiguous materialised payload (copies a strided / rolled view)."""
        m = self._mat()
        if m.is_contiguous:
            return m
        if m.shifts is not None:
            # a rolled view is made contiguous by the generated identity kernel
            return lazy.materialize(lazy.Node("positive", (lazy.leaf(m),), (m.dtype,), m.shape, m.dtype))
        out = Mat.empty(m.shape, m.dtype, inexact=m.buf.inexact)
        backend.get().copy_strided(out.ptr, out.dtype, out.strides, m.ptr, m.dtype, m.strides, m.shape)
        return out

    def materialize(self):
        """Force evaluation of pending lazy work behind this array (no host sync)."""
        if self._mat().buf.pending:
            lazy.flush_deferred()
        return self

    def get(self, out=None) -> np.ndarray:
        """Copy to host as a numpy.ndarray (synchronises).  ``out`` may be a preallocated (e.g. pinned) array."""
        m = self._cmat()
        if out is None:
            out = np.empty(m.shape, dtype=m.dtype)
        elif out.shape != m.shape or out.dtype != m.dtype or not out.flags.c_contiguous:
            raise ValueError("get(out=...): out must be a C-contiguous array of the same shape and dtype")
        backend.get().d2h(out, m.ptr)
        return out

    to_host = get

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(()).item()

    def tolist(self):
        return self.get().tolist()

    def __array__(self, dtype=None, copy=None):
        raise TypeError(
            "autograd_b200.DeviceArray refuses implicit conversion to a host numpy array "
            "(it would hide a device→host copy); call .get() explicitly.  If this came from a SciPy/NumPy function that "
            "autograd wraps, make sure autograd_b200.install() has run."
        )

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        if self.size != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous. Use a.any() or a.all()")
        return bool(self.item())

    def __index__(self):
        if self.dtype.kind not in "iu" or self.size != 1:
            raise TypeError("only integer scalar arrays can be converted to a scalar index")
        return int(self.item())

    def __repr__(self):
        return "DeviceArray(" + repr(self.get())[len("array("):]

    def __str__(self):
        return str(self.get())

    def __hash__(self):
        return id(self)

    def __iter__(self):
        if not self.shape:
            raise TypeError("iteration over a 0-d array")
        for i in range(self.shape[0]):
            yield self[i]

    def __array_namespace__(self, *, api_version=None):
        import autograd.numpy as anp

        return anp

    @property
    def __cuda_array_interface__(self):
        m = self._cmat()
        if m is not self._node.mat:  # keep the contiguous copy alive for as long as this handle lives
            new = lazy.leaf(m)
            new.nhandles += 1
            self._node.nhandles -= 1
            self._node = new
        return {
            "shape": m.shape,
            "typestr": m.dtype.str,
            "data": (m.ptr, False),
            "version": 3,
            "strides": None,
            "stream": backend.get().stream_handle() or None,
        }

    # ------------------------------------------------------------ NumPy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        for x in inputs:
            if _is_box(x) or (_Complex is not None and type(x) is _Complex):
                return NotImplemented
        if method == "__call__":
            if ufunc is np.matmul:  # a generalized ufunc, not an elementwise one
                return matmul(*inputs)
            op = UFUNC_OPS.get(ufunc)
            if op is None and ufunc in UFUNC_ORDER_OPS and len(inputs) == 2 and not kwargs:
                return order_function(UFUNC_ORDER_OPS[ufunc], inputs[0], inputs[1])
            if op is None:
                raise TypeError(f"autograd_b200: numpy.{ufunc.__name__} has no device implementation")
            out = kwargs.pop("out", None)
            kwargs.pop("casting", None)
            dtype = kwargs.pop("dtype", None)
            if kwargs:
                raise TypeError(f"autograd_b200: numpy.{ufunc.__name__} keyword(s) {sorted(kwargs)} are not supported on device arrays")
            res = elementwise(op, ufunc, *inputs)
            if dtype is not None and np.dtype(dtype) != res.dtype:
                res = res.astype(dtype)
            if out is not None:
                (target,) = out if isinstance(out, tuple) else (out,)
                if isinstance(target, np.ndarray):
                    # `x += y` on a host accumulator (core.py:262-264: the cotangent of a HOST argument, created by
                    # ArrayVSpace.zeros) receiving a device contribution: the value of a host argument's gradient
                    # lives on the host, so this is an explicit device→host read of the contribution.
                    target[...] = res.get()
                    return target
                if not isinstance(target, DeviceArray):
                    raise TypeError("autograd_b200: out= must be a DeviceArray or a host ndarray")
                target._assign(res)
                return target
            return res
        if method == "reduce":
            name = UFUNC_REDUCE.get(ufunc)
            if name is None:
                raise TypeError(f"autograd_b200: numpy.{ufunc.__name__}.reduce has no device implementation")
            kwargs.pop("out", None)
            kwargs.pop("initial", None)
            kwargs.pop("where", None)
            axis = kwargs.pop("axis", 0)
            keepdims = kwargs.pop("keepdims", False)
            dtype = kwargs.pop("dtype", None)
            return reduce_(inputs[0], name, axis, keepdims, dtype)
        if method == "at":
            if ufunc is not np.add:
                raise TypeError(f"autograd_b200: numpy.{ufunc.__name__}.at has no device implementation")
            target, idx, values = inputs
            if isinstance(target, np.ndarray):
                # untake into the cotangent of a HOST argument (numpy_vjps.py:941-950 with vs = ArrayVSpace): the
                # device contribution is read back explicitly, like the `out=` host accumulator above
                idx = tuple(i.get() if isinstance(i, DeviceArray) else i for i in idx) if isinstance(idx, tuple) else \
                    (idx.get() if isinstance(idx, DeviceArray) else idx)
                np.add.at(target, idx, values.get() if isinstance(values, DeviceArray) else values)
                return None
            if not isinstance(target, DeviceArray):
                raise TypeError("autograd_b200: np.add.at target must be a DeviceArray or a host ndarray")
            target._add_at(idx, values)
            return None
        return NotImplemented

    def __array_function__(self, func, types, args, kwargs):
        if _Complex is not None and _Complex in types:
            return NotImplemented              # ComplexDeviceArray.__array_function__ answers
        impl = FUNCTIONS.get(func)
        if impl is None:
            if func in DECOMPOSE:
                return func._implementation(*args, **kwargs)
            raise TypeError(f"autograd_b200: numpy.{func.__name__} has no device implementation")
        return impl(*args, **kwargs)

    # ------------------------------------------------------------ operators used by VJP closures
    def _binop(self, opname, ufunc, other, reflected=False):
        if _is_box(other) or (_Complex is not None and type(other) is _Complex):
            return NotImplemented
        return elementwise(opname, ufunc, other, self) if reflected else elementwise(opname, ufunc, self, other)

    def __add__(self, o): return self._binop("add", np.add, o)
    def __radd__(self, o): return self._binop("add", np.add, o, True)
    def __sub__(self, o): return self._binop("subtract", np.subtract, o)
    def __rsub__(self, o): return self._binop("subtract", np.subtract, o, True)
    def __mul__(self, o): return self._binop("multiply", np.multiply, o)
    def __rmul__(self, o): return self._binop("multiply", np.multiply, o, True)
    def __truediv__(self, o): return self._binop("divide", np.true_divide, o)
    def __rtruediv__(self, o): return self._binop("divide", np.true_divide, o, True)
    def __floordiv__(self, o): return self._binop("floor_divide", np.floor_divide, o)
    def __rfloordiv__(self, o): return self._binop("floor_divide", np.floor_divide, o, True)
    def __mod__(self, o): return self._binop("remainder", np.remainder, o)
    def __rmod__(self, o): return self._binop("remainder", np.remainder, o, True)
    def __pow__(self, o): return self._binop("power", np.power, o)
    def __rpow__(self, o): return self._binop("power", np.power, o, True)
    def __matmul__(self, o): return NotImplemented if _is_box(o) else matmul(self, o)
    def __rmatmul__(self, o): return NotImplemented if _is_box(o) else matmul(o, self)
    def __neg__(self): return elementwise("negative", np.negative, self)
    def __pos__(self): return self
    def __abs__(self): return elementwise("abs", np.absolute, self)
    def __invert__(self): return elementwise("logical_not", np.logical_not, self)
    def __eq__(self, o): return self._binop("equal", np.equal, o)
    def __ne__(self, o): return self._binop("not_equal", np.not_equal, o)
    def __lt__(self, o): return self._binop("less", np.less, o)
    def __le__(self, o): return self._binop("less_equal", np.less_equal, o)
    def __gt__(self, o): return self._binop("greater", np.greater, o)
    def __ge__(self, o): return self._binop("greater_equal", np.greater_equal, o)
    def __and__(self, o): return self._binop("logical_and", np.logical_and, o)
    def __or__(self, o): return self._binop("logical_or", np.logical_or, o)

    # In-place operators rebind this handle to the new expression (VSpace._mut_add: ``x += y``,
    # autograd/core.py:262-264, only ever touches accumulators autograd itself created).
    def _assign(self, value):
        value = asarray(value)
        if value.shape != self.shape:
            value = broadcast_to(value, self.shape)
        if value.dtype != self.dtype:
            value = value.astype(self.dtype)
        old, new = self._node, value._node
        new.nhandles += 1
        old.nhandles -= 1
        self._node = new

    def __iadd__(self, o):
        r = self.__add__(o)
        if r is NotImplemented:
            return r
        self._assign(r)
        return self

    def __isub__(self, o):
        r = self.__sub__(o)
        if r is NotImplemented:
            return r
        self._assign(r)
        return self

    def __imul__(self, o):
        r = self.__mul__(o)
        if r is NotImplemented:
            return r
        self._assign(r)
        return self

    def __itruediv__(self, o):
        r = self.__truediv__(o)
        if r is NotImplemented:
            return r
        self._assign(r)
        return self

    # ------------------------------------------------------------ ndarray-style methods
    def astype(self, dtype, order="K", casting="unsafe", subok=True, copy=True):
        dtype = np.dtype(dtype)
        if dtype == self.dtype:
            if not copy:
                return self
            # NumPy returns an independent copy: share the payload, copy it when either handle is written in place
            _MEMO_NODES.add(self._node)
            return DeviceArray(self._node)
        if dtype not in lazy.SUPPORTED:
            backend.dtype_code(dtype)
        return DeviceArray(lazy.make_cast(self._node, dtype))

    def view(self, *args, **kwargs):
        return self  # `.view(ndarray)` in autograd/numpy/numpy_wrapper.py:59 strips subclasses; nothing to strip here

    def copy(self, order="C"):
        m = self._mat()
        out = Mat.empty(m.shape, m.dtype, inexact=m.buf.inexact)
        if m.shifts is not None:
            return DeviceArray(lazy.leaf(self._cmat()))
        backend.get().copy_strided(out.ptr, out.dtype, out.strides, m.ptr, m.dtype, m.strides, m.shape)
        return DeviceArray(lazy.leaf(out))

    def reshape(self, *shape, order="C"):
        if len(shape) == 1 and not isinstance(shape[0], numbers.Integral):
            shape = shape[0]
        return reshape(self, shape)

    def ravel(self, order="C"): return reshape(self, (-1,))
    def flatten(self, order="C"): return reshape(self, (-1,)).copy()
    def transpose(self, *axes):
        if len(axes) == 1 and (axes[0] is None or not isinstance(axes[0], numbers.Integral)):
            axes = axes[0]
        return transpose(self, axes if axes else None)
    def swapaxes(self, a, b): return swapaxes(self, a, b)
    def squeeze(self, axis=None): return squeeze(self, axis)
    def sum(self, axis=None, dtype=None, out=None, keepdims=False): return reduce_(self, "sum", axis, keepdims, dtype)
    def prod(self, axis=None, dtype=None, out=None, keepdims=False): return reduce_(self, "prod", axis, keepdims, dtype)
    def max(self, axis=None, out=None, keepdims=False): return reduce_(self, "max", axis, keepdims)
    def min(self, axis=None, out=None, keepdims=False): return reduce_(self, "min", axis, keepdims)
    def mean(self, axis=None, dtype=None, out=None, keepdims=False): return mean(self, axis, dtype, None, keepdims)
    def var(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False): return var(self, axis, dtype, None, ddof, keepdims)
    def std(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False): return std(self, axis, dtype, None, ddof, keepdims)
    def all(self, axis=None, out=None, keepdims=False): return reduce_(self, "all", axis, keepdims)
    def any(self, axis=None, out=None, keepdims=False): return reduce_(self, "any", axis, keepdims)
    def argmax(self, axis=None, out=None, keepdims=False): return argreduce(self, "max", axis, keepdims)
    def argmin(self, axis=None, out=None, keepdims=False): return argreduce(self, "min", axis, keepdims)
    def dot(self, other): return dot(self, other)
    def clip(self, a_min=None, a_max=None, out=None): return clip(self, a_min, a_max)
    def repeat(self, repeats, axis=None): return repeat(self, repeats, axis)
    def take(self, indices, axis=None): return take(self, indices, axis)
    def round(self, decimals=0): return round_(self, decimals)
    def conj(self): return self
    conjugate = conj

    @property
    def real(self): return self

    @property
    def imag(self): return zeros(self.shape, self.dtype)

    # ------------------------------------------------------------ indexing
    def __getitem__(self, idx):
        return getitem(self, idx)

    def __setitem__(self, idx, value):
        kind, info = _parse_index(self.shape, idx)
        if kind != "basic":
            return self._assign_at(info, value)
        self._own_buffer()
        view = _basic_view(self._mat(), info)
        value = asarray(value, dtype=self.dtype) if not isinstance(value, DeviceArray) else value
        if value.dtype != self.dtype:
            value = value.astype(self.dtype)
        src = broadcast_to(value, view.shape)._mat()
        if src.shifts is not None:
            src = DeviceArray(lazy.leaf(src))._cmat()
        elif src.buf is view.buf and prod(view.shape) > 0:
            # `x[1:] = x[:-1]`: source and destination may overlap inside one buffer; NumPy copies through a temporary
            tmp = Mat.empty(src.shape, src.dtype, inexact=src.buf.inexact)
            backend.get().copy_strided(tmp.ptr, tmp.dtype, tmp.strides, src.ptr, src.dtype, src.strides, src.shape)
            src = tmp
        backend.get().copy_strided(view.ptr, view.dtype, view.strides, src.ptr, src.dtype, src.strides, view.shape)
        view.buf.inexact = view.buf.inexact or src.buf.inexact

    def _assign_at(self, info, value):
        """x[index arrays] = value (e.g. ``unsort[permutation] = range(n)`` in unpermuter, numpy_vjps.py:803-806)."""
        self._own_buffer()
        m = self._mat()
        if not m.is_contiguous:
            raise NotImplementedError("autograd_b200: item assignment with array indices needs a contiguous target")
        idx_arrays, idx_shape, n_lead = info
        inner_shape = self.shape[n_lead:]
        v = value if isinstance(value, DeviceArray) else asarray(np.asarray(value))
        if v.dtype != self.dtype:
            v = v.astype(self.dtype)
        vm = broadcast_to(v, tuple(idx_shape) + tuple(inner_shape))._cmat()
        ptrs = [a._cmat() for a in idx_arrays]
        backend.get().scatter_assign(m.ptr, vm.ptr, self.dtype, [p.ptr for p in ptrs], list(self.shape[:n_lead]),
                                     prod(idx_shape), prod(inner_shape))
        m.buf.inexact = m.buf.inexact or vm.buf.inexact

    def _own_buffer(self):
        """Make sure this handle is backed by real memory that may be written in place."""
        node = self._node
        if node.mat is None:
            if node.op == "full" and isinstance(node.args[0], Const) and node.args[0].value == 0 and prod(node.shape) > 0 \
                    and str(node.args[0].value)[0] != "-" and node.alias is None:
                # the zero accumulator of a sparse cotangent (vs.zeros() handed to np.add.at, core.py:212-215): a memset at
                # copy-engine speed instead of a generated fill kernel
                zm = Mat.empty(node.shape, node.dtype)
                backend.get().memset(zm.ptr, 0, zm.buf.nbytes)
                node.mat, node.args, node.cost, node.lo, node.leafset = zm, (), 0, node.nid, frozenset((node.nid,))
                lazy.payload_set(node)
            else:
                lazy.materialize(node)
        m = node.mat
        lazy.flush_readers(m.buf)     # pending expressions that read this buffer must see the OLD contents
        m.buf.version += 1            # an in-place write follows: memoised copies of the old contents are stale
        shared = node.nhandles > 1 and node in _MEMO_NODES   # NumPy would have made independent copies: copy on write
        if shared or m.shifts is not None or any(st == 0 and s > 1 for s, st in zip(m.shape, m.strides)):
            c = self._cmat() if m.shifts is not None else None
            if c is None:
                c = Mat.empty(m.shape, m.dtype, inexact=m.buf.inexact)
                backend.get().copy_strided(c.ptr, c.dtype, c.strides, m.ptr, m.dtype, m.strides, m.shape)
            new = lazy.leaf(c)
            new.nhandles += 1
            node.nhandles -= 1
            self._node = new

    def _add_at(self, idx, values):
        """np.add.at(self, idx, values): unbuffered in-place accumulation (numpy_vjps.py:946-948)."""
        self._own_buffer()
        mixed = _mixed_index(self.shape, idx)
        if mixed is not None:
            # scatter into a contiguous stand-in of the (sliced, transposed) view, then one strided in-place add
            v2, n_other = _mixed_view(self, mixed)
            arrays = tuple(mixed[2])
            idx_shape = lazy.broadcast_shapes([tuple(np.shape(i)) for i in arrays])
            probe_shape = tuple(idx_shape) + tuple(v2.shape[len(arrays):])
            nidx, pos = len(probe_shape) - n_other, mixed[3]
            out_shape = probe_shape if not (pos and nidx) else \
                probe_shape[nidx:nidx + pos] + probe_shape[:nidx] + probe_shape[nidx + pos:]
            vals = broadcast_to(asarray(values), out_shape)
            if pos and nidx:
                vals = moveaxis(vals, tuple(range(pos, pos + nidx)), tuple(range(nidx)))
            tmp = zeros(v2.shape, self.dtype).copy()
            tmp._add_at(arrays, vals)
            v2._add_at(Ellipsis, tmp)
            return
        m = self._mat()
        kind, info = _parse_index(self.shape, idx)
        be = backend.get()
        if kind == "basic":
            view = _basic_view(m, info)
            v = asarray(values)
            if v.dtype != self.dtype:
                v = v.astype(self.dtype)
            src = broadcast_to(v, view.shape)
            sm = src._mat()
            if sm.shifts is not None:
                sm = src._cmat()
            be.add_strided(view.ptr, view.dtype, view.strides, sm.ptr, sm.strides, view.shape)
            view.buf.inexact = view.buf.inexact or sm.buf.inexact
            return
        idx_arrays, idx_shape, n_lead = info
        if not m.is_contiguous:
            raise NotImplementedError("autograd_b200: np.add.at with array indices needs a contiguous target")
        inner_shape = self.shape[n_lead:]
        v = asarray(values)
        if v.dtype != self.dtype:
            v = v.astype(self.dtype)
        if not inner_shape and self.dtype.kind in "fi" and self.dtype.itemsize >= 4 and \
                lazy.broadcast_shapes([tuple(idx_shape), v.shape]) == tuple(idx_shape):
            lazy.scatter_add(m, [i._node for i in idx_arrays], v._node)
            return
        v = broadcast_to(v, tuple(idx_shape) + tuple(inner_shape))
        vm = v._cmat()
        ptrs = [a._cmat() for a in idx_arrays]
        be.scatter_add(m.ptr, vm.ptr, self.dtype, [p.ptr for p in ptrs], list(self.shape[:n_lead]), prod(idx_shape),
                       prod(inner_shape))
        m.buf.inexact = True      # atomic accumulation order is not NumPy's sequential order


# =================================================================== construction
def _upload(host: np.ndarray) -> DeviceArray:
    host = np.asarray(host)
    if host.dtype not in lazy.SUPPORTED:
        if host.dtype.kind == "f":
            host = host.astype(np.float64)
        elif host.dtype.kind in "iu":
            host = host.astype(np.int64)
        else:
            backend.dtype_code(host.dtype)
    host = np.asarray(host, order="C")
    m = Mat.empty(host.shape, host.dtype, upload=True)
    if host.size:
        backend.get().h2d(m.ptr, host)
    return DeviceArray(lazy.leaf(m))


def asarray(x, dtype=None) -> DeviceArray:
    """Device version of ``np.asarray``: uploads host data, passes DeviceArrays through."""
    if isinstance(x, DeviceArray):
        return x if dtype is None else x.astype(dtype)
    if _Complex is not None and type(x) is _Complex:
        return x if dtype is None else x.astype(dtype)      # a complex value (pair of device arrays) passes through
    if isinstance(x, (list, tuple)) and any(isinstance(e, DeviceArray) for e in x):
        return stack([asarray(e) for e in x]) if dtype is None else stack([asarray(e) for e in x]).astype(dtype)
    host = np.asarray(x) if dtype is None else np.asarray(x, dtype=dtype)
    return _upload(host)


array = asarray
to_device = asarray


def pinned_empty(shape, dtype=float) -> np.ndarray:
    """A numpy array backed by page-locked host memory (full-rate, truly asynchronous H2D/D2H copies).
    The block lives until ``free_pinned(arr)`` (or process exit): views handed to NumPy must never dangle."""
    import ctypes

    shape = _shape_tuple(shape)
    dtype = np.dtype(dtype)
    nbytes = max(prod(shape) * dtype.itemsize, 1)
    ptr = backend.get().host_alloc(nbytes)
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=prod(shape)).reshape(shape)
    _PINNED[ptr] = buf
    return arr


_PINNED = {}


def free_pinned(arr: np.ndarray) -> None:
    """Release a block obtained from ``pinned_empty`` (the caller guarantees no view of it is used afterwards)."""
    ptr = arr.ctypes.data
    if _PINNED.pop(ptr, None) is not None:
        backend.get().host_free(ptr)


def _shape_tuple(shape):
    if isinstance(shape, (numbers.Integral, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def full(shape, fill_value, dtype=None, **_ignored):
    if isinstance(fill_value, DeviceArray):
        return broadcast_to(fill_value if dtype is None else fill_value.astype(dtype), _shape_tuple(shape)).copy()
    if dtype is None:
        dtype = np.asarray(fill_value).dtype
    dtype = np.dtype(dtype)
    if dtype.kind == "c" and _complex_full is not None:
        return _complex_full(_shape_tuple(shape), fill_value, dtype)
    backend.dtype_code(dtype)
    value = np.asarray(fill_value).astype(dtype).item()
    return DeviceArray(lazy.make_full(value, _shape_tuple(shape), dtype))


def zeros(shape, dtype=float, **_ignored):
    return full(shape, 0, np.dtype(dtype))


def ones(shape, dtype=float, **_ignored):
    return full(shape, 1, np.dtype(dtype))


def empty(shape, dtype=float, **_ignored):
    return full(shape, 0, np.dtype(dtype))


def zeros_like(a, dtype=None, **_ignored):
    return zeros(np.shape(a), dtype or _dtype_of(a))


def ones_like(a, dtype=None, **_ignored):
    return ones(np.shape(a), dtype or _dtype_of(a))


def empty_like(a, dtype=None, **_ignored):
    return zeros(np.shape(a), dtype or _dtype_of(a))


def full_like(a, fill_value, dtype=None, **_ignored):
    return full(np.shape(a), fill_value, dtype or _dtype_of(a))


def _dtype_of(a):
    return a.dtype if hasattr(a, "dtype") else np.asarray(a).dtype


def arange(*args, **kwargs):
    return _upload(np.arange(*args, **kwargs))


def linspace(start, stop, num=50, endpoint=True, retstep=False, dtype=None, axis=0):
    if not isinstance(start, DeviceArray) and not isinstance(stop, DeviceArray):
        return _upload(np.linspace(start, stop, num, endpoint, retstep, dtype, axis))
    # device end points (the JVP ``linspace(g, 0, num)``, numpy_jvps.py:169-173): NumPy's own formula
    # ``arange(num) * (delta / div) + start`` with the last sample pinned to ``stop`` (numpy/_core/function_base.py)
    if retstep or axis != 0 or np.ndim(start) or np.ndim(stop):
        raise NotImplementedError("autograd_b200: linspace with device end points supports scalar ends on axis 0 only")
    num = int(num)
    div = num - 1 if endpoint else num
    start, stop = asarray(start) * 1.0, asarray(stop) * 1.0
    y = _upload(np.arange(0, num, dtype=np.float64))
    if div > 0:
        y = y * ((stop - start) / div) + start
    else:
        y = y * (stop - start) + start
    if endpoint and num > 1:
        y = y.copy()
        y[-1] = stop
    return y if dtype is None else y.astype(dtype)


def eye(*args, **kwargs):
    return _upload(np.eye(*args, **kwargs))


def meshgrid(*xi, copy=True, sparse=False, indexing="xy"):
    """np.meshgrid on device: the outputs are zero-stride broadcast views (no payload is replicated)."""
    arrs = [asarray(x).ravel() for x in xi]
    nd = len(arrs)
    out = []
    for i, a in enumerate(arrs):
        shp = [1] * nd
        shp[i] = a.shape[0]
        out.append(reshape(a, tuple(shp)))
    if indexing == "xy" and nd > 1:
        out[0] = reshape(out[0], (1, -1) + (1,) * (nd - 2))
        out[1] = reshape(out[1], (-1, 1) + (1,) * (nd - 2))
    if not sparse:
        full_shape = lazy.broadcast_shapes([o.shape for o in out])
        out = [broadcast_to(o, full_shape) for o in out]
    return out


# =================================================================== elementwise
def _operand(x):
    """DeviceArray / scalar / host array → lazy.Node or lazy.Const."""
    if isinstance(x, DeviceArray):
        return x._node
    if isinstance(x, np.generic):
        if x.dtype not in lazy.SUPPORTED:
            if x.dtype.kind == "c":
                raise _ComplexOperand("autograd_b200: complex operand")
            return Const(x.item(), None)
        return Const(x.item(), x.dtype)
    if isinstance(x, (bool, int, float)):
        return Const(x, None)
    if isinstance(x, np.ndarray):
        if x.dtype.kind == "c":
            raise _ComplexOperand("autograd_b200: complex operand")
        if x.ndim == 0:
            return _operand(x[()])
        if x.size > 1 and all(s == 0 for s in x.strides):
            # host broadcast of a scalar (e.g. the `+ onp.zeros(shape)` idiom): stay symbolic
            v = x.flat[0]
            return lazy.make_full(v.item(), x.shape, x.dtype if x.dtype in lazy.SUPPORTED else F64)
        return _upload(x)._node
    if isinstance(x, (list, tuple)):
        return asarray(x)._node
    if isinstance(x, complex) or (_Complex is not None and type(x) is _Complex):
        raise _ComplexOperand("autograd_b200: complex operand")
    raise TypeError(f"autograd_b200: cannot use operand of type {type(x).__name__} in a device computation")


def _host_only_operands(inputs):
    """Dispatch on operand type (overrides must keep working for plain ndarray code).  If the only device
    operands are untouched symbolic constants from the rebound creation functions (``anp.zeros(anp.shape(g))`` inside a
    reference rule, numpy_vjps.py:228, numpy_jvps.py:150) and the caller's data is host NumPy, return the operands with
    those constants restated as zero-stride host broadcasts — the computation never concerned the device.  Otherwise
    None.  Nothing is read back from the device: a weak value is (fill value, shape, dtype), all host metadata."""
    weak = host = False
    for x in inputs:
        if isinstance(x, DeviceArray):
            if not x._weak:
                return None
            weak = True
        elif isinstance(x, (np.ndarray, np.generic)):
            host = True
    if not weak:
        return None
    if not host and any(isinstance(x, DeviceArray) and x.ndim for x in inputs):
        return None   # Python scalars only: `anp.zeros(n) + 1.0` is device work; 0-d constants are plain scalars
    out = []
    for x in inputs:
        if isinstance(x, DeviceArray):
            nd_ = x._node
            x = np.broadcast_to(np.asarray(nd_.args[0].value, dtype=nd_.dtype), nd_.shape)
        out.append(x)
    return out


def elementwise(opname, ufunc, *inputs) -> DeviceArray:
    host = _host_only_operands(inputs)
    if host is not None:
        return ufunc(*host)
    try:
        args = [_operand(x) for x in inputs]
    except _ComplexOperand:
        if _complex_ufunc is None:
            raise
        return _complex_ufunc(ufunc, inputs)     # complex values are pairs of real device arrays (complex_array.py)
    return DeviceArray(lazy.make_op(opname, ufunc, args))


def order_function(opname, order, x) -> DeviceArray:
    """Special functions of an integer order and a device argument — ``jn/jv``, ``yn``, ``iv``, ``ive``,
    ``polygamma`` — as they appear in the reference's derivative rules (autograd/scipy/special.py:43-48,76-90:
    ``g * polygamma(1, x)``, ``(j0(x) - jn(2, x)) / 2`` ...).  The order is a host scalar baked into the kernel."""
    if isinstance(order, DeviceArray) or np.ndim(order) != 0:
        raise TypeError(f"autograd_b200: {opname[3:]}(order, x) needs a host scalar order on device arguments")
    n = float(order)
    if n != int(n) or abs(n) > 1e6 or (opname == "sp_polygamma" and n < 0):
        raise TypeError(f"autograd_b200: {opname[3:]} of order {order!r} has no device implementation (integer orders only)")
    x = asarray(x)
    if x.dtype.kind != "f":
        x = x.astype(np.float64)
    return DeviceArray(lazy.make_op(opname, np.copysign, [Const(n, None), x._node]))


def where(cond, x=None, y=None):
    if x is None and y is None:
        raise NotImplementedError("autograd_b200: np.where(cond) (nonzero) is not implemented on device")
    host = _host_only_operands((cond, x, y))
    if host is not None:
        return np.where(*host)
    try:
        c, a, b = _operand(cond), _operand(x), _operand(y)
    except _ComplexOperand:
        if _complex_where is None:
            raise
        return _complex_where(cond, x, y)
    if type(c) is Const:
        return asarray(x if c.value else y)
    dts = []
    for o, raw in ((a, x), (b, y)):
        if type(o) is Node:
            dts.append(o.dtype)
        elif o.dtype is not None:
            dts.append(o.dtype)
        else:
            dts.append(o.value)  # weak Python scalar (NEP 50): pass the value, not its type
    out_dtype = np.result_type(*dts)
    backend.dtype_code(out_dtype)
    return DeviceArray(lazy.make_where(c, a, b, np.dtype(out_dtype)))


def clip(a, a_min=None, a_max=None, out=None, **kw):
    r = asarray(a)
    if a_min is not None:
        r = elementwise("maximum", np.maximum, r, a_min)
    if a_max is not None:
        r = elementwise("minimum", np.minimum, r, a_max)
    return r


def round_(a, decimals=0, out=None):
    a = asarray(a)
    if a.dtype.kind != "f":
        return a
    if decimals == 0:
        return elementwise("rint", np.rint, a)
    s = 10.0 ** decimals
    return elementwise("rint", np.rint, a * s) / s


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    x = asarray(x)
    fi = np.finfo(x.dtype)
    r = where(elementwise("isnan", np.isnan, x), nan, x)
    r = where(r == np.inf, fi.max if posinf is None else posinf, r)
    return where(r == -np.inf, fi.min if neginf is None else neginf, r)


def real(x): return asarray(x)
def imag(x): return zeros_like(x)
def iscomplexobj(x): return False
def isrealobj(x): return True


def isclose(a, b, rtol=1e-05, atol=1e-08, equal_nan=False):
    a, b = asarray(a), asarray(b)
    close = abs(a - b) <= (atol + rtol * abs(b))
    both_inf = (a == b)
    r = (close & elementwise("isfinite", np.isfinite, b)) | both_inf
    if equal_nan:
        r = r | (elementwise("isnan", np.isnan, a) & elementwise("isnan", np.isnan, b))
    return r


def allclose(a, b, rtol=1e-05, atol=1e-08, equal_nan=False):
    return bool(reduce_(isclose(a, b, rtol, atol, equal_nan), "all", None, False))


def array_equal(a1, a2, equal_nan=False):
    a1, a2 = asarray(a1), asarray(a2)
    if a1.shape != a2.shape:
        return False
    return bool(reduce_(a1 == a2, "all", None, False))


# =================================================================== layout (views where possible)
def _normalize_axis(axis, ndim):
    if not -ndim <= axis < ndim:
        raise np.exceptions.AxisError(axis, ndim)
    return axis + ndim if axis < 0 else axis


def _view(x: DeviceArray, shape, strides, offset_delta=0, shifts=None) -> DeviceArray:
    m = x._mat()
    return DeviceArray(lazy.leaf(Mat(m.buf, shape, strides, m.offset + offset_delta, m.dtype, shifts)))


def reshape(a, shape=None, order="C", *, newshape=None, copy=None):
    a = asarray(a)
    if shape is None:
        shape = newshape
    if order not in ("C", None, "A", "K"):
        raise NotImplementedError("autograd_b200: only C-order reshape is supported")
    shape = list(_shape_tuple(shape))
    size = a.size
    if -1 in shape:
        i = shape.index(-1)
        rest = prod(s for j, s in enumerate(shape) if j != i)
        shape[i] = size // rest if rest else 0
    shape = tuple(shape)
    if prod(shape) != size:
        raise ValueError(f"cannot reshape array of size {size} into shape {shape}")
    if shape == a.shape:
        return a
    node = a._node
    if node.mat is None:
        res = None
        if _reshape_transparent(node):
            res = _reshape_lazy(node, shape, node.shape, {})
        if res is None:
            res = _reshape_unit_axes_lazy(node, shape)
        if res is None:
            res = _reshape_strided_lazy(node, shape)
        if res is None and lazy.LAZY_RESHAPE and size >= 4096 and a.dtype.kind != "b" and node.op != "scatter":
            # some operand cannot be re-viewed by itself (a broadcast meshgrid row being flattened): record the reshape
            # and let lazy.materialize pick one iteration space for both sides of it (lazy._plan_reshapes)
            res = Node("reshape", (node,), (a.dtype,), shape, a.dtype)
        if res is not None:
            if res.mat is None:
                lazy.link_alias(res, node)       # NumPy: a view of ``a`` - in-place writes through either must alias
            return DeviceArray(res)
    m = a._mat()
    new_strides = _reshape_strides(m, shape)
    if new_strides is None:
        m = a._cmat()
        new_strides = lazy.c_strides(shape)
    return DeviceArray(lazy.leaf(Mat(m.buf, shape, new_strides, m.offset, m.dtype)))


def _reshape_transparent(node) -> bool:
    """A pending expression can be reshaped symbolically when every operand is either a scalar or a
    contiguous array of exactly the expression's shape (then flat indices line up)."""
    shape = node.shape
    stack, seen = [node], set()
    while stack:
        nd_ = stack.pop()
        if nd_.nid in seen:
            continue
        seen.add(nd_.nid)
        if len(seen) > 64:
            return False
        if nd_.mat is not None:
            if nd_.shape == shape and nd_.mat.is_contiguous:
                continue
            if prod(nd_.shape) == 1:
                continue
            return False
        if nd_.shape != shape and prod(nd_.shape) != 1:
            return False
        for a in nd_.args:
            if type(a) is Node:
                stack.append(a)
    return True


def _reshape_lazy(node, shape, root_shape, memo):
    hit = memo.get(node.nid)
    if hit is not None:
        return hit
    new_shape = shape if node.shape == root_shape else ()  # anything else is a size-1 broadcast operand
    if node.mat is not None:
        m = node.mat
        if new_shape == ():
            res = lazy.leaf(Mat(m.buf, (), (), m.offset, m.dtype))
        else:
            res = lazy.leaf(Mat(m.buf, shape, lazy.c_strides(shape), m.offset, m.dtype))
    else:
        args = tuple(_reshape_lazy(a, shape, root_shape, memo) if type(a) is Node else a for a in node.args)
        res = Node(node.op, args, node.in_dtypes, new_shape, node.dtype)
    memo[node.nid] = res
    return res


def _reshape_strided_lazy(node, new_shape):
    """C-order reshape of a pending elementwise expression whose operands are NOT all contiguous arrays of its own
    shape — a broadcast bias, a strided view — pushed down to the leaves: every operand, seen as a strided (zero-stride
    where it broadcasts) view of the root's shape, is reshaped without a copy (``_reshape_strides``).  This keeps
    ``conv(x, w) + b`` pending through the ``reshape(N, C, H/2, 2, W/2, 2)`` of a max-pool (examples/convnet.py:108-117),
    so the bias add fuses into the pooling kernel instead of costing a pass of its own.  None when some operand cannot
    be reshaped in place (then the caller materialises first, as before)."""
    root_shape = node.shape
    if _pending_size(node, 48) > 48:
        return None
    nd = len(root_shape)
    memo = {}

    def go(n_):
        hit = memo.get(n_.nid, memo)
        if hit is not memo:
            return hit
        res = None
        if prod(n_.shape) == 1 and n_.mat is None and n_.op == "full":
            res = Node("full", n_.args, n_.in_dtypes, (), n_.dtype)
        elif n_.mat is not None:
            m = n_.mat
            if m.shifts is None:
                if prod(m.shape) == 1:
                    res = lazy.leaf(Mat(m.buf, (), (), m.offset, m.dtype))
                else:
                    off = nd - len(m.shape)
                    if off >= 0 and all(s == 1 or s == root_shape[off + i] for i, s in enumerate(m.shape)):
                        bstr = (0,) * off + tuple(0 if s == 1 else st for s, st in zip(m.shape, m.strides))
                        view = Mat(m.buf, root_shape, bstr, m.offset, m.dtype)
                        st_new = _reshape_strides(view, new_shape) if not view.is_contiguous else lazy.c_strides(new_shape)
                        if st_new is not None:
                            res = lazy.leaf(Mat(m.buf, tuple(new_shape), st_new, m.offset, m.dtype))
        elif n_.op == "full":
            res = Node("full", n_.args, n_.in_dtypes, tuple(new_shape), n_.dtype) if n_.shape == root_shape else None
        elif n_.op in ("scatter", "gather"):
            res = None
        elif n_.shape == root_shape or prod(n_.shape) == 1:
            args = []
            for a in n_.args:
                if type(a) is Node:
                    a = go(a)
                    if a is None:
                        break
                args.append(a)
            if len(args) == len(n_.args):
                res = Node(n_.op, tuple(args), n_.in_dtypes, tuple(new_shape) if n_.shape == root_shape else (), n_.dtype)
        memo[n_.nid] = res
        return res

    return go(node)


_SELECTABLE_SPECIAL = ("where", "cast", "full")


def _select_lazy(node, ndim, axis, index, memo):
    """``node[..., index, ...]`` along ``axis`` of the ``ndim``-dimensional root, pushed down to the leaves of a pending
    elementwise expression (elementwise ops commute with slicing).  The axis disappears from every operand; operands
    that broadcast along it (missing leading axis or size 1) keep their data.  Returns None when some operand cannot
    be sliced symbolically (gather tables, rolled views of that axis)."""
    hit = memo.get(node.nid, memo)
    if hit is not memo:
        return hit
    shape = node.shape
    ax = axis - (ndim - len(shape))          # right-aligned broadcasting
    if ax < 0:
        res = node                            # broadcasts along the missing leading axes: unchanged
    elif node.mat is not None:
        m = node.mat
        if m.shifts is not None and m.shifts[ax] != 0:
            res = None
        else:
            i = 0 if shape[ax] == 1 else index
            shifts = None if m.shifts is None else m.shifts[:ax] + m.shifts[ax + 1:]
            if shifts is not None and not any(shifts):
                shifts = None
            res = lazy.leaf(Mat(m.buf, shape[:ax] + shape[ax + 1:], m.strides[:ax] + m.strides[ax + 1:],
                                m.offset + i * m.strides[ax], m.dtype, shifts))
    elif node.op == "full":
        res = Node("full", node.args, node.in_dtypes, shape[:ax] + shape[ax + 1:], node.dtype)
    elif node.op in ("scatter", "reshape"):
        res = None
    else:
        args = []
        for k, a in enumerate(node.args):
            if type(a) is Node and not (node.op == "gather" and k == 0):   # a gather's table is not indexed by position
                a = _select_lazy(a, len(shape), ax, index, memo)
                if a is None:
                    break
            args.append(a)
        res = Node(node.op, tuple(args), node.in_dtypes, shape[:ax] + shape[ax + 1:], node.dtype) \
            if len(args) == len(node.args) else None
    memo[node.nid] = res
    return res


def _insert_axis_lazy(node, ndim, pos, memo):
    """A new size-1 axis at position ``pos`` of the ``ndim``-dimensional root, pushed down to the leaves of a pending
    elementwise expression (the keepdims / expand_dims reshapes of reduction VJPs).  Returns None when impossible."""
    hit = memo.get(node.nid, memo)
    if hit is not memo:
        return hit
    shape = node.shape
    p = pos - (ndim - len(shape))
    if p <= 0:
        res = node                            # the new axis lies left of this operand's axes: still broadcasts as is
    elif node.mat is not None:
        m = node.mat
        shifts = None if m.shifts is None else m.shifts[:p] + (0,) + m.shifts[p:]
        res = lazy.leaf(Mat(m.buf, shape[:p] + (1,) + shape[p:], m.strides[:p] + (0,) + m.strides[p:], m.offset,
                            m.dtype, shifts))
    elif node.op == "full":
        res = Node("full", node.args, node.in_dtypes, shape[:p] + (1,) + shape[p:], node.dtype)
    elif node.op in ("scatter", "reshape"):
        res = None
    else:
        args = []
        for k, a in enumerate(node.args):
            if type(a) is Node and not (node.op == "gather" and k == 0):
                a = _insert_axis_lazy(a, len(shape), p, memo)
                if a is None:
                    break
            args.append(a)
        res = Node(node.op, tuple(args), node.in_dtypes, shape[:p] + (1,) + shape[p:], node.dtype) \
            if len(args) == len(node.args) else None
    memo[node.nid] = res
    return res


def _reshape_unit_axes_lazy(node, new_shape):
    """Reshape of a pending expression that only removes and/or inserts size-1 axes, done symbolically."""
    old_shape = node.shape
    if [d for d in old_shape if d != 1] != [d for d in new_shape if d != 1]:
        return None
    if _pending_size(node, 48) > 48:
        return None
    cur = node
    for ax in range(len(old_shape) - 1, -1, -1):          # drop every size-1 axis, last first
        if old_shape[ax] == 1:
            cur = _select_lazy(cur, len(cur.shape), ax, 0, {})
            if cur is None:
                return None
    for pos, d in enumerate(new_shape):                    # insert the size-1 axes of the target, first to last
        if d == 1:
            if pos == 0 and len(cur.shape) >= 0:
                # a leading axis: every operand keeps broadcasting, only the root's shape changes
                nxt = _insert_leading_axis(cur)
            else:
                nxt = _insert_axis_lazy(cur, len(cur.shape), pos, {})
            if nxt is None:
                return None
            cur = nxt
    return cur if cur.shape == tuple(new_shape) else None


def _insert_leading_axis(node):
    if node.mat is not None:
        m = node.mat
        shifts = None if m.shifts is None else (0,) + m.shifts
        return lazy.leaf(Mat(m.buf, (1,) + node.shape, (0,) + m.strides, m.offset, m.dtype, shifts))
    return Node(node.op, node.args, node.in_dtypes, (1,) + node.shape, node.dtype)


def _pending_size(node, limit):
    """Number of pending (unmaterialised) nodes under ``node``, or ``limit`` + 1 when there are more."""
    stack, seen = [node], set()
    while stack:
        nd_ = stack.pop()
        if nd_.nid in seen or nd_.mat is not None:
            continue
        seen.add(nd_.nid)
        if len(seen) > limit:
            return limit + 1
        stack.extend(a for a in nd_.args if type(a) is Node)
    return len(seen)


_SHORT_REDUCE_ON = __import__("os").environ.get("B200AG_SHORT_REDUCE", "1") != "0"
_SHORT_REDUCE = {"sum": ("add", np.add), "prod": ("multiply", np.multiply), "max": ("maximum", np.maximum),
                 "min": ("minimum", np.minimum)}


def _reduce_short_axes(a, op, axes, work_dtype):
    """Reductions over SHORT axes (pooling windows, `sum(x == ans)` over a window: at most 8 slices in total) as a
    pending elementwise combination of the slices — no reduction kernel, no materialised input; the result fuses into
    its consumers (the max-pool VJP `g * (x == ans) / sum(x == ans)`, numpy_vjps.py:515-530, becomes one kernel).
    Returns None when the fast path does not apply."""
    if op not in _SHORT_REDUCE or not axes or not _SHORT_REDUCE_ON:
        return None
    slices = prod(a.shape[ax] for ax in axes)
    if slices > 8 or slices < 2 or a.size // slices < 4096:
        return None
    node = a._node
    if node.mat is None and _pending_size(node, 24) > 24:
        node = lazy.leaf(a._cmat())            # deep expressions: compute once, then slice the buffer
    opname, ufunc = _SHORT_REDUCE[op]
    cur = [node]
    ndim = a.ndim
    for ax in sorted(axes, reverse=True):       # last axis first keeps the earlier axis numbers valid
        nxt = []
        for piece in cur:
            for i in range(a.shape[ax]):
                sel = _select_lazy(piece, ndim, ax, i, {})
                if sel is None:
                    return None
                nxt.append(sel)
        cur = nxt
        ndim -= 1
    pieces = [DeviceArray(p) for p in cur]
    if pieces[0].dtype != work_dtype:
        pieces = [p.astype(work_dtype) for p in pieces]
    acc = pieces[0]
    for piece in pieces[1:]:
        acc = elementwise(opname, ufunc, acc, piece)
    return acc


def _reshape_strides(m: Mat, new_shape):
    """Strides of a no-copy reshape of view ``m`` or None when a copy is unavoidable."""
    if m.shifts is not None:
        return None
    if m.is_contiguous:
        return lazy.c_strides(new_shape)
    old = [(s, st) for s, st in zip(m.shape, m.strides) if s != 1]
    new_strides = [0] * len(new_shape)
    oi, ni = 0, 0
    nn = len(new_shape)
    while ni < nn and new_shape[ni] == 1:
        ni += 1
    while oi < len(old) and ni < nn:
        # gather a group of old dims and new dims with equal element counts
        op_, np_ = old[oi][0], new_shape[ni]
        oj, nj = oi + 1, ni + 1
        while op_ != np_:
            if op_ < np_:
                if oj >= len(old):
                    return None
                op_ *= old[oj][0]
                oj += 1
            else:
                if nj >= nn:
                    return None
                np_ *= new_shape[nj]
                nj += 1
        for k in range(oi, oj - 1):  # the old group must itself be contiguous
            if old[k][1] != old[k + 1][1] * old[k + 1][0]:
                return None
        st = old[oj - 1][1]
        for k in range(nj - 1, ni - 1, -1):
            new_strides[k] = st
            st *= new_shape[k]
        oi, ni = oj, nj
        while ni < nn and new_shape[ni] == 1:
            ni += 1
    if oi != len(old) or ni != nn:
        return None
    return tuple(new_strides)


def ravel(a, order="C"):
    return reshape(a, (-1,))


def transpose(a, axes=None):
    a = asarray(a)
    nd = a.ndim
    if axes is None:
        axes = tuple(range(nd - 1, -1, -1))
    else:
        axes = tuple(_normalize_axis(int(ax), nd) for ax in axes)
        if sorted(axes) != list(range(nd)):
            raise ValueError("axes don't match array")
    if axes == tuple(range(nd)):
        return a
    m = a._mat()
    shifts = None if m.shifts is None else tuple(m.shifts[ax] for ax in axes)
    return _view(a, tuple(m.shape[ax] for ax in axes), tuple(m.strides[ax] for ax in axes), 0, shifts)


def swapaxes(a, axis1, axis2):
    a = asarray(a)
    nd = a.ndim
    axis1, axis2 = _normalize_axis(axis1, nd), _normalize_axis(axis2, nd)
    axes = list(range(nd))
    axes[axis1], axes[axis2] = axes[axis2], axes[axis1]
    return transpose(a, axes)


def moveaxis(a, source, destination):
    a = asarray(a)
    nd = a.ndim
    src = [_normalize_axis(s, nd) for s in ([source] if isinstance(source, numbers.Integral) else source)]
    dst = [_normalize_axis(d, nd) for d in ([destination] if isinstance(destination, numbers.Integral) else destination)]
    order = [n for n in range(nd) if n not in src]
    for d, s in sorted(zip(dst, src)):
        order.insert(d, s)
    return transpose(a, order)


def expand_dims(a, axis):
    a = asarray(a)
    axes = (axis,) if isinstance(axis, numbers.Integral) else tuple(axis)
    nd = a.ndim + len(axes)
    axes = sorted(_normalize_axis(ax, nd) for ax in axes)
    shape = list(a.shape)
    for ax in axes:
        shape.insert(ax, 1)
    return reshape(a, tuple(shape))


def squeeze(a, axis=None):
    a = asarray(a)
    if axis is None:
        shape = tuple(s for s in a.shape if s != 1)
    else:
        axes = (axis,) if isinstance(axis, numbers.Integral) else tuple(axis)
        axes = [_normalize_axis(ax, a.ndim) for ax in axes]
        for ax in axes:
            if a.shape[ax] != 1:
                raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
        shape = tuple(s for i, s in enumerate(a.shape) if i not in axes)
    return reshape(a, shape)


def atleast_1d(*arys):
    res = [reshape(asarray(a), (1,)) if np.ndim(a) == 0 else asarray(a) for a in arys]
    return res[0] if len(res) == 1 else tuple(res)


def atleast_2d(*arys):
    res = []
    for a in arys:
        a = asarray(a)
        if a.ndim == 0:
            a = reshape(a, (1, 1))
        elif a.ndim == 1:
            a = reshape(a, (1, a.shape[0]))
        res.append(a)
    return res[0] if len(res) == 1 else tuple(res)


def atleast_3d(*arys):
    """np.atleast_3d: () -> (1,1,1), (N,) -> (1,N,1), (M,N) -> (M,N,1) (numpy/_core/shape_base.py)."""
    res = []
    for a in arys:
        a = asarray(a)
        if a.ndim == 0:
            a = reshape(a, (1, 1, 1))
        elif a.ndim == 1:
            a = reshape(a, (1, a.shape[0], 1))
        elif a.ndim == 2:
            a = reshape(a, a.shape + (1,))
        res.append(a)
    return res[0] if len(res) == 1 else tuple(res)


def broadcast_to(a, shape, subok=False):
    a = asarray(a)
    shape = _shape_tuple(shape)
    if a.shape == shape:
        return a
    if lazy.broadcast_shapes([a.shape, shape]) != shape:
        raise ValueError(f"operands could not be broadcast together with remapped shapes [original->remapped]: {a.shape} and requested shape {shape}")
    if a._node.mat is None and prod(a.shape) == 1 and a._node.op == "full":
        c = a._node.args[0]
        return DeviceArray(lazy.make_full(c.value, shape, a.dtype))
    m = a._mat()
    off = len(shape) - len(m.shape)
    strides = [0] * len(shape)
    shifts = [0] * len(shape)
    for i, (s, st) in enumerate(zip(m.shape, m.strides)):
        if s != 1:
            strides[off + i] = st
            if m.shifts is not None:
                shifts[off + i] = m.shifts[i]
    return _view(a, shape, tuple(strides), 0, tuple(shifts) if m.shifts is not None and any(shifts) else None)


def flip(a, axis=None):
    a = asarray(a)
    nd = a.ndim
    axes = range(nd) if axis is None else ([axis] if isinstance(axis, numbers.Integral) else axis)
    idx = [slice(None)] * nd
    for ax in axes:
        idx[_normalize_axis(ax, nd)] = slice(None, None, -1)
    return a[tuple(idx)]


def flipud(a): return flip(a, 0)
def fliplr(a): return flip(a, 1)


def roll(a, shift, axis=None):
    """np.roll as an index transform: the result is a circularly shifted VIEW that fused kernels read
    in place (the VJP is the opposite shift, numpy_vjps.py:193)."""
    a = asarray(a)
    if axis is None:
        return reshape(roll(reshape(a, (-1,)), shift, 0), a.shape)
    axes = (axis,) if isinstance(axis, numbers.Integral) else tuple(axis)
    shifts_in = (shift,) * len(axes) if isinstance(shift, numbers.Integral) else tuple(shift)
    if len(shifts_in) != len(axes):
        raise ValueError("shift and axis must have the same length")
    # (Pushing the shift down through a PENDING expression, so that two Jacobi sweeps share one kernel, was implemented
    # and measured on the B200 in round 2: C4 at 20 steps 30.2 ms without, 30.6 / 36.0 / 35.6 / 38.7 ms with pending
    # expressions of up to 8 / 12 / 24 / 40 nodes pushed - the fused kernels re-derive the inner sweep at five
    # positions per output and become issue-bound.  Removed; a rolled operand is always a materialised view.)
    m = a._mat()
    shifts = list(m.shifts) if m.shifts is not None else [0] * a.ndim
    for s, ax in zip(shifts_in, axes):
        ax = _normalize_axis(ax, a.ndim)
        n = a.shape[ax]
        if n:
            shifts[ax] = (shifts[ax] + int(s)) % n
    return _view(a, m.shape, m.strides, 0, tuple(shifts) if any(shifts) else None)


def concatenate(arrays, axis=0, out=None, dtype=None, casting="same_kind"):
    arrays = [asarray(a) for a in arrays]
    if not arrays:
        raise ValueError("need at least one array to concatenate")
    if axis is None:
        arrays = [ravel(a) for a in arrays]
        axis = 0
    nd = arrays[0].ndim
    if nd == 0:
        raise ValueError("zero-dimensional arrays cannot be concatenated")
    axis = _normalize_axis(axis, nd)
    out_dtype = np.result_type(*[a.dtype for a in arrays]) if dtype is None else np.dtype(dtype)
    shape = list(arrays[0].shape)
    for a in arrays[1:]:
        if a.ndim != nd or any(a.shape[i] != shape[i] for i in range(nd) if i != axis):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    shape[axis] = sum(a.shape[axis] for a in arrays)
    # The same pieces concatenated again (the four gates of an LSTM cell hstack identical (input, hiddens, ones),
    # examples/lstm.py:38-49) return the first copy: the result is immutable unless written in place, which bumps the
    # buffer version and so changes the key.
    for a in arrays:
        if a._node.mat is None and a._node.op != "full" and a.size:
            a._mat()
    # (graph.capture clears the memo when a capture begins and ends, so entries never cross that boundary)
    key = _concat_key(arrays, axis, out_dtype)
    if key is not None:
        hit = _CONCAT_MEMO.get(key)
        if hit is not None:
            node = hit[0]()
            if node is not None and node.mat is not None and node.mat.buf.version == hit[1] and \
                    all(r() is a._node for r, a in zip(hit[2], arrays) if r is not None):
                return DeviceArray(node)
            del _CONCAT_MEMO[key]
    out_m = Mat.empty(shape, out_dtype, inexact=any(a._node.mat is not None and a._node.mat.buf.inexact for a in arrays))
    be = backend.get()
    live = [a for a in arrays if a.size]
    if 3 <= len(live) <= 16 and all(a.dtype == out_dtype and a._node.mat is not None and a._node.mat.is_contiguous
                                    and a._node.mat.shifts is None for a in live):
        # contiguous pieces of one dtype: a single launch instead of one strided copy per piece
        outer = prod(shape[:axis])
        rest = prod(shape[axis + 1:])
        be.concat(out_m.ptr, out_dtype, [a._node.mat.ptr for a in live], [a.shape[axis] * rest for a in live], outer)
        live = []
    if live and 2 <= len(live) <= 16 and out_dtype in (F64, F32) and all(
            (a._node.mat is not None and a._node.mat.is_contiguous and a._node.mat.shifts is None and a.dtype in (F64, F32))
            or (a._node.mat is None and a._node.op == "full" and isinstance(a._node.args[0], lazy.Const)) for a in live):
        # float pieces of mixed dtype and symbolic constants (hstack((input float32, hiddens float64, ones)) of an LSTM
        # cell): still a single launch, the promotion to float64 and the constant fill happen inside it
        outer = prod(shape[:axis])
        rest = prod(shape[axis + 1:])
        pieces = [(a._node.mat.ptr, a.dtype) if a._node.mat is not None else (None, float(a._node.args[0].value))
                  for a in live]
        be.concat_mixed(out_m.ptr, out_dtype, pieces, [a.shape[axis] * rest for a in live], outer)
        live = []
    pos = 0
    for a in (arrays if live else ()):
        if a.size == 0:
            continue
        m = a._mat()
        if m.shifts is not None:
            m = a._cmat()
        dst_off = pos * out_m.strides[axis]
        be.copy_strided(out_m.ptr + dst_off * out_dtype.itemsize, out_dtype, out_m.strides, m.ptr, m.dtype, m.strides,
                        m.shape)
        pos += a.shape[axis]
    out_node = lazy.leaf(out_m)
    if key is not None:
        if len(_CONCAT_MEMO) > 64:
            _CONCAT_MEMO.clear()
        refs = [None if part[0] == "full" else weakref.ref(a._node) for part, a in zip(key[0], arrays)]
        _CONCAT_MEMO[key] = (weakref.ref(out_node), out_m.buf.version, refs)
        _MEMO_NODES.add(out_node)
    return DeviceArray(out_node)


_CONCAT_MEMO = {}
_MEMO_NODES = weakref.WeakSet()   # nodes that several independent results may share: copied before an in-place write


def _concat_key(arrays, axis, out_dtype):
    """Identity of a concatenation: materialised pieces by (node id, buffer version), symbolic constants by value."""
    parts = []
    for a in arrays:
        node = a._node
        if node.mat is not None:
            parts.append((id(node), node.mat.buf.version))
        elif node.op == "full" and isinstance(node.args[0], lazy.Const):
            parts.append(("full", repr(node.args[0].value), node.shape, str(node.dtype)))
        else:
            return None
    return (tuple(parts), axis, str(out_dtype))


def stack(arrays, axis=0, out=None):
    arrays = [asarray(a) for a in arrays]
    nd = arrays[0].ndim + 1
    axis = _normalize_axis(axis, nd)
    return concatenate([expand_dims(a, axis) for a in arrays], axis)


def repeat(a, repeats, axis=None):
    a = asarray(a)
    if not isinstance(repeats, numbers.Integral):
        raise NotImplementedError("autograd_b200: np.repeat supports an integer repeat count only")
    if axis is None:
        a = ravel(a)
        axis = 0
    axis = _normalize_axis(axis, a.ndim)
    shape = a.shape
    expanded = reshape(a, shape[: axis + 1] + (1,) + shape[axis + 1:])
    tiled = broadcast_to(expanded, shape[: axis + 1] + (int(repeats),) + shape[axis + 1:])
    return reshape(tiled, shape[:axis] + (shape[axis] * int(repeats),) + shape[axis + 1:])


def tile(a, reps):
    a = asarray(a)
    reps = (reps,) if isinstance(reps, numbers.Integral) else tuple(reps)
    nd = max(a.ndim, len(reps))
    a = reshape(a, (1,) * (nd - a.ndim) + a.shape)
    reps = (1,) * (nd - len(reps)) + reps
    inter, final, bshape = [], [], []
    for s, r in zip(a.shape, reps):
        inter += [1, s]
        bshape += [r, s]
        final.append(r * s)
    return reshape(broadcast_to(reshape(a, tuple(inter)), tuple(bshape)), tuple(final))


def pad(array, pad_width, mode="constant", **kwargs):
    if mode != "constant":
        raise NotImplementedError("autograd_b200: np.pad supports mode='constant' only")
    value = kwargs.get("constant_values", 0)
    a = asarray(array)
    pw = np.broadcast_to(np.asarray(pad_width, dtype=np.int64), (a.ndim, 2))
    shape = tuple(int(s + l + r) for s, (l, r) in zip(a.shape, pw))
    out = full(shape, value, a.dtype)
    out._own_buffer()
    idx = tuple(slice(int(l), int(l) + s) for s, (l, r) in zip(a.shape, pw))
    out[idx] = a
    return out



# =================================================================== triangular / diagonal helpers (views + masks)
def _tri_mask(shape, k, lower):
    n, m = shape[-2], shape[-1]
    mask = np.tri(n, m, k=k, dtype=bool) if lower else ~np.tri(n, m, k=k - 1, dtype=bool)
    return _upload(mask)


def tril(m, k=0):
    m = asarray(m)
    return where(_tri_mask(m.shape, k, True), m, zeros((), m.dtype))


def triu(m, k=0):
    m = asarray(m)
    return where(_tri_mask(m.shape, k, False), m, zeros((), m.dtype))


def diagonal(a, offset=0, axis1=0, axis2=1):
    a = asarray(a)
    axis1, axis2 = _normalize_axis(axis1, a.ndim), _normalize_axis(axis2, a.ndim)
    mt = a._mat()
    if mt.shifts is not None:
        mt = a._cmat()
    n1, n2 = mt.shape[axis1], mt.shape[axis2]
    s1, s2 = mt.strides[axis1], mt.strides[axis2]
    if offset >= 0:
        length, start = max(0, min(n1, n2 - offset)), offset * s2
    else:
        length, start = max(0, min(n1 + offset, n2)), -offset * s1
    keep = [i for i in range(a.ndim) if i not in (axis1, axis2)]
    shape = tuple(mt.shape[i] for i in keep) + (length,)
    strides = tuple(mt.strides[i] for i in keep) + (s1 + s2,)
    return DeviceArray(lazy.leaf(Mat(mt.buf, shape, strides, mt.offset + start, mt.dtype)))


def diag(v, k=0):
    v = asarray(v)
    if v.ndim == 2:
        return diagonal(v, k)
    if v.ndim != 1:
        raise ValueError("Input must be 1- or 2-d.")
    n = v.shape[0] + abs(k)
    out = zeros((n, n), v.dtype)
    out._own_buffer()
    dview = diagonal(out, k)
    m, src = dview._mat(), v._mat() if v._node.mat is not None and v._mat().shifts is None else v._cmat()
    backend.get().copy_strided(m.ptr, m.dtype, m.strides, src.ptr, src.dtype, src.strides, m.shape)
    return out


def trace(a, offset=0, axis1=0, axis2=1, dtype=None, out=None):
    return reduce_(diagonal(a, offset, axis1, axis2), "sum", -1, False, dtype)


def diff(a, n=1, axis=-1, prepend=None, append=None):
    if prepend is not None or append is not None:
        raise NotImplementedError("autograd_b200: np.diff prepend/append are not supported")
    a = asarray(a)
    axis = _normalize_axis(axis, a.ndim)
    for _ in range(n):
        hi = [slice(None)] * a.ndim
        lo = [slice(None)] * a.ndim
        hi[axis], lo[axis] = slice(1, None), slice(None, -1)
        a = a[tuple(hi)] - a[tuple(lo)]
    return a


def _cumulate(a, op, axis, dtype):
    a = asarray(a)
    if axis is None:
        a, axis = ravel(a), 0
    axis = _normalize_axis(axis, a.ndim) if a.ndim else 0
    if a.ndim == 0:
        a = reshape(a, (1,))
    dt = np.dtype(dtype) if dtype is not None else (np.dtype(np.int64) if a.dtype.kind in "bi" and a.dtype.itemsize < 8
                                                      or a.dtype.kind == "b" else a.dtype)
    if a.dtype != dt:
        a = a.astype(dt)
    if a.size == 0:
        return a
    src = a._cmat()
    out = Mat.empty(a.shape, dt, inexact=dt.kind == "f")      # scan order differs from NumPy's sequential loop
    backend.get().cumulate(out.ptr, src.ptr, dt, op, prod(a.shape[:axis]), a.shape[axis], prod(a.shape[axis + 1:]))
    return DeviceArray(lazy.leaf(out))


def cumsum(a, axis=None, dtype=None, out=None):
    """np.cumsum (VJP: grad_np_cumsum, numpy_vjps.py:539-549) on the device scan kernel."""
    return _cumulate(a, "sum", axis, dtype)


def cumprod(a, axis=None, dtype=None, out=None):
    return _cumulate(a, "prod", axis, dtype)


def _sorted(a, axis, want_keys, want_idx):
    a = asarray(a)
    if a.dtype == np.bool_:
        raise NotImplementedError("autograd_b200: sorting boolean arrays is not supported")
    if axis is None:
        a, axis = ravel(a), 0
    if a.ndim == 0:
        raise np.exceptions.AxisError("axis 0 is out of bounds for array of dimension 0")
    axis = _normalize_axis(axis, a.ndim)
    moved = moveaxis(a, axis, -1) if axis != a.ndim - 1 else a
    shape = moved.shape
    n = shape[-1]
    rows = prod(shape[:-1])
    keys = Mat.empty(shape, a.dtype) if want_keys else None
    idx = Mat.empty(shape, np.dtype(np.int64)) if want_idx else None
    if rows and n:
        src = moved._cmat()
        be = backend.get()
        for r0 in range(0, rows, 65535):      # grid.y limit of the sort kernels
            r1 = min(rows, r0 + 65535)
            be.sort(None if keys is None else keys.ptr + r0 * n * a.dtype.itemsize,
                    None if idx is None else idx.ptr + r0 * n * 8, src.ptr + r0 * n * a.dtype.itemsize, a.dtype, r1 - r0, n)

    def back(m):
        r = DeviceArray(lazy.leaf(m))
        return moveaxis(r, -1, axis) if axis != a.ndim - 1 else r

    return (back(keys) if want_keys else None), (back(idx) if want_idx else None)


def sort(a, axis=-1, kind=None, order=None, **kw):
    """np.sort: stable ascending row sort, NaN last (grad_sort, numpy_vjps.py:774-789)."""
    if order is not None:
        raise NotImplementedError("autograd_b200: structured-array sort order is not supported")
    return _sorted(a, axis, True, False)[0]


def argsort(a, axis=-1, kind=None, order=None, **kw):
    """np.argsort: int64 permutation of the stable sort (bit-exact against kind='stable', and against any kind when
    the keys are distinct)."""
    if order is not None:
        raise NotImplementedError("autograd_b200: structured-array sort order is not supported")
    return _sorted(a, axis, False, True)[1]


def _check_kth(a, kth, axis):
    n = asarray(a).shape[axis] if asarray(a).ndim else 1
    for k in np.atleast_1d(np.asarray(kth)):
        if not -n <= int(k) < n:
            raise ValueError(f"kth(={int(k)}) out of bounds ({n})")


def partition(a, kth, axis=-1, kind="introselect", order=None):
    """np.partition.  Any arrangement with the kth element in its sorted position, nothing larger before it and
    nothing smaller after it is a valid result; the device returns the fully sorted array (one such arrangement:
    the kth element and the two SETS on either side equal NumPy's, the order inside the sets is introselect's
    business and differs).  grad_partition (numpy_vjps.py:787-792) only needs ``partition`` and ``argpartition``
    to describe the same permutation, which they do."""
    if order is not None:
        raise NotImplementedError("autograd_b200: structured-array sort order is not supported")
    if axis is None:
        a, axis = reshape(asarray(a), (-1,)), 0
    _check_kth(a, kth, axis)
    return _sorted(a, axis, True, False)[0]


def argpartition(a, kth, axis=-1, kind="introselect", order=None):
    """np.argpartition: the permutation of :func:`partition` (the stable argsort)."""
    if order is not None:
        raise NotImplementedError("autograd_b200: structured-array sort order is not supported")
    if axis is None:
        a, axis = reshape(asarray(a), (-1,)), 0
    _check_kth(a, kth, axis)
    return _sorted(a, axis, False, True)[1]


def kron(a, b):
    a, b = asarray(a), asarray(b)
    nd = max(a.ndim, b.ndim)
    a = reshape(a, (1,) * (nd - a.ndim) + a.shape)
    b = reshape(b, (1,) * (nd - b.ndim) + b.shape)
    ash = tuple(x for d in a.shape for x in (d, 1))
    bsh = tuple(x for d in b.shape for x in (1, d))
    return reshape(reshape(a, ash) * reshape(b, bsh), tuple(x * y for x, y in zip(a.shape, b.shape)))


def rot90(m, k=1, axes=(0, 1)):
    m = asarray(m)
    a0, a1 = (_normalize_axis(x, m.ndim) for x in axes)
    if a0 == a1:
        raise ValueError("Axes must be different.")
    k %= 4
    if k == 0:
        return m
    if k == 2:
        return flip(flip(m, a0), a1)
    if k == 1:
        return swapaxes(flip(m, a1), a0, a1)
    return flip(swapaxes(m, a0, a1), a1)


def rollaxis(a, axis, start=0):
    a = asarray(a)
    axis = _normalize_axis(axis, a.ndim)
    if start < 0:
        start += a.ndim
    if not 0 <= start <= a.ndim:
        raise np.exceptions.AxisError(f"start {start} is out of bounds")
    if axis < start:
        start -= 1
    return a if axis == start else moveaxis(a, axis, start)


def average(a, axis=None, weights=None, returned=False, keepdims=False):
    a = asarray(a)
    if weights is None:
        avg = mean(a, axis=axis, keepdims=keepdims)
        if returned:
            return avg, full_like(avg, a.size / max(avg.size, 1))
        return avg
    w = asarray(weights)
    if w.shape != a.shape:
        if axis is None or w.ndim != 1 or w.shape[0] != a.shape[axis]:
            raise TypeError("Axis must be specified when shapes of a and weights differ.")
        w = reshape(w, tuple(w.shape[0] if i == _normalize_axis(axis, a.ndim) else 1 for i in range(a.ndim)))
        w = broadcast_to(w, a.shape)
    scl = sum_(w, axis=axis, keepdims=keepdims)
    avg = sum_(a * w, axis=axis, keepdims=keepdims) / scl
    return (avg, scl) if returned else avg


def select(condlist, choicelist, default=0):
    if len(condlist) != len(choicelist):
        raise ValueError("list of cases must be same length as list of conditions")
    out = default
    for c, v in zip(reversed(list(condlist)), reversed(list(choicelist))):
        out = where(c, v, out)
    return out


def sinc(x):
    x = asarray(x)
    if x.dtype.kind != "f":
        x = x.astype(np.float64)
    y = np.pi * where(x == 0, 1.0e-20, x)
    return elementwise("sin", np.sin, y) / y


def cross(a, b, axisa=-1, axisb=-1, axisc=-1, axis=None):
    if axis is not None:
        axisa = axisb = axisc = axis
    a, b = moveaxis(asarray(a), axisa, -1), moveaxis(asarray(b), axisb, -1)
    if a.shape[-1] != 3 or b.shape[-1] != 3:
        raise ValueError("autograd_b200: np.cross supports 3-vectors only")
    a0, a1, a2 = a[..., 0], a[..., 1], a[..., 2]
    b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
    out = stack([a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0], axis=-1)
    return moveaxis(out, -1, axisc)


def gradient(f, *varargs, axis=None, edge_order=1):
    """np.gradient with uniform spacing: central differences inside, one-sided first-order differences at the edges."""
    f = asarray(f)
    if edge_order != 1:
        raise NotImplementedError("autograd_b200: np.gradient supports edge_order=1 only")
    axes = tuple(range(f.ndim)) if axis is None else tuple(_normalize_axis(x, f.ndim) for x in np.atleast_1d(axis))
    if len(varargs) == 0:
        dx = [1.0] * len(axes)
    elif len(varargs) == 1 and np.ndim(varargs[0]) == 0:
        dx = [varargs[0]] * len(axes)
    elif len(varargs) == len(axes) and all(np.ndim(v) == 0 for v in varargs):
        dx = list(varargs)
    else:
        raise NotImplementedError("autograd_b200: np.gradient supports scalar spacings only")
    if f.dtype.kind != "f":
        f = f.astype(np.float64)
    outs = []
    for ax, h in zip(axes, dx):
        if f.shape[ax] < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, at least 2 elements are required.")

        def sl(s):
            idx = [slice(None)] * f.ndim
            idx[ax] = s
            return tuple(idx)

        inner = (f[sl(slice(2, None))] - f[sl(slice(None, -2))]) / (2.0 * h)
        first = (f[sl(slice(1, 2))] - f[sl(slice(0, 1))]) / h
        last = (f[sl(slice(-1, None))] - f[sl(slice(-2, -1))]) / h
        outs.append(concatenate([first, inner, last], axis=ax))
    return outs[0] if axis is not None and np.ndim(axis) == 0 or f.ndim == 1 and len(outs) == 1 else tuple(outs)


def norm(x, ord=None, axis=None, keepdims=False):
    """np.linalg.norm for vector norms and the Frobenius matrix norm (VJP: autograd/numpy/linalg.py:91-144), composed
    from device reductions.  Spectral / nuclear matrix norms need an SVD and are not offered."""
    x = asarray(x)
    if x.dtype.kind != "f":
        x = x.astype(np.float64)
    matrix = isinstance(axis, tuple) and len(axis) == 2 or (axis is None and x.ndim == 2 and ord is not None)
    if matrix:
        if ord not in (None, "fro"):
            raise NotImplementedError(f"autograd_b200: matrix norm ord={ord!r} is not implemented on the device")
        return elementwise("sqrt", np.sqrt, reduce_(x * x, "sum", axis, keepdims))
    if axis is None and ord is not None and x.ndim > 2:
        raise ValueError("Improper number of dimensions to norm.")
    if ord is None or ord == 2 or ord == "fro":
        if ord == "fro" and not (axis is None and x.ndim == 2):
            raise ValueError("Invalid norm order 'fro' for vectors")
        return elementwise("sqrt", np.sqrt, reduce_(x * x, "sum", axis, keepdims))
    ax = None if axis is None else axis
    mag = elementwise("abs", np.absolute, x)
    if ord == np.inf:
        return reduce_(mag, "max", ax, keepdims)
    if ord == -np.inf:
        return reduce_(mag, "min", ax, keepdims)
    if ord == 0:
        return reduce_((x != 0).astype(x.dtype), "sum", ax, keepdims)
    if ord == 1:
        return reduce_(mag, "sum", ax, keepdims)
    if isinstance(ord, str):
        raise ValueError(f"Invalid norm order '{ord}' for vectors")
    return reduce_(mag ** ord, "sum", ax, keepdims) ** (1.0 / ord)


def angle(z, deg=False):
    z = asarray(z)
    zero = zeros_like(z, dtype=z.dtype if z.dtype.kind == "f" else np.float64)
    a = elementwise("arctan2", np.arctan2, zero, z)
    return a * (180.0 / np.pi) if deg else a


def real_if_close(a, tol=100):
    return asarray(a)


# =================================================================== indexing
def _parse_index(shape, idx):
    """Classify ``idx``: ("basic", [(kind, ...)]) or ("adv", (index arrays, index shape, n leading dims))."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    has_array = any(isinstance(i, (DeviceArray, np.ndarray, list)) for i in idx)
    if not has_array:
        n_ell = sum(1 for i in idx if i is Ellipsis)
        if n_ell > 1:
            raise IndexError("an index can only have a single ellipsis ('...')")
        n_real = sum(1 for i in idx if i is not None and i is not Ellipsis)
        if n_real > len(shape):
            raise IndexError(f"too many indices for array: array is {len(shape)}-dimensional, but {n_real} were indexed")
        ops = []
        dim = 0
        for i in idx:
            if i is Ellipsis:
                for _ in range(len(shape) - n_real):
                    ops.append(("slice", dim, slice(None)))
                    dim += 1
            elif i is None:
                ops.append(("new",))
            elif isinstance(i, slice):
                ops.append(("slice", dim, i))
                dim += 1
            elif isinstance(i, (numbers.Integral, np.integer)):
                n = shape[dim]
                j = int(i)
                if not -n <= j < n:
                    raise IndexError(f"index {j} is out of bounds for axis {dim} with size {n}")
                ops.append(("int", dim, j + n if j < 0 else j))
                dim += 1
            else:
                raise IndexError(f"unsupported index {i!r}")
        while dim < len(shape):
            ops.append(("slice", dim, slice(None)))
            dim += 1
        return "basic", ops
    # advanced: integer arrays on the leading axes, optionally followed by full slices
    arrays = []
    for k, i in enumerate(idx):
        if isinstance(i, slice) and i == slice(None):
            if any(not (isinstance(j, slice) and j == slice(None)) and j is not Ellipsis for j in idx[k:]):
                raise NotImplementedError("autograd_b200: index arrays must be on the leading axes")
            break
        if i is Ellipsis:
            break
        if isinstance(i, (numbers.Integral, np.integer)):
            arrays.append(np.asarray(int(i), dtype=np.int64))
        else:
            arrays.append(i)
    if len(arrays) > len(shape):
        raise IndexError("too many indices for array")
    dev = []
    for k, a in enumerate(arrays):
        if not isinstance(a, DeviceArray):
            a = np.asarray(a)
            if a.dtype == np.bool_:
                raise NotImplementedError("autograd_b200: boolean mask indexing is not implemented on device")
            if a.dtype.kind not in "iu":
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            if a.size:
                # host index arrays are validated here like NumPy does; device index arrays are checked by the kernels
                # themselves (out-of-range entries are skipped and reported at the next synchronisation)
                lo, hi, n = int(a.min()), int(a.max()), shape[k]
                if lo < -n or hi >= n:
                    bad = hi if hi >= n else lo
                    raise IndexError(f"index {bad} is out of bounds for axis {k} with size {n}")
            a = _upload(a.astype(np.int64))
        elif a.dtype == BOOL:
            raise NotImplementedError("autograd_b200: boolean mask indexing is not implemented on device")
        elif a.dtype != I64:
            a = a.astype(np.int64)
        dev.append(a)
    idx_shape = lazy.broadcast_shapes([a.shape for a in dev])
    dev = [broadcast_to(a, idx_shape) for a in dev]
    return "adv", (dev, idx_shape, len(dev))


def _mixed_index(shape, idx):
    """Index arrays mixed with slices / newaxis, or not on the leading axes (``x[3, 2:, [1, 3]]``, ``x[:, cols]``).
    NumPy's rule (advanced indexing, "combining advanced and basic indexing"): the basic part is a view; the
    broadcast index dimensions replace the indexed axes in place when the advanced entries (arrays AND integers)
    are adjacent, and come first otherwise.  Returns None for the plain leading form ``_parse_index`` handles, else
    (basic index making the view, view axes carrying an advanced index, those indices, destination of the index
    dims or None for "first")."""
    if not isinstance(idx, tuple):
        idx = (idx,)

    def is_arr(i):
        return isinstance(i, (DeviceArray, np.ndarray, list))

    if not any(is_arr(i) for i in idx):
        return None
    k = 0
    while k < len(idx) and (is_arr(idx[k]) or isinstance(idx[k], (numbers.Integral, np.integer))):
        k += 1
    if all((isinstance(j, slice) and j == slice(None)) or j is Ellipsis for j in idx[k:]):
        return None
    if sum(1 for i in idx if i is Ellipsis) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    n_real = sum(1 for i in idx if i is not None and i is not Ellipsis)
    if n_real > len(shape):
        raise IndexError(f"too many indices for array: array is {len(shape)}-dimensional, but {n_real} were indexed")
    entries = []
    for i in idx:
        if i is Ellipsis:
            entries += [slice(None)] * (len(shape) - n_real)
        else:
            entries.append(i)
    entries += [slice(None)] * (len(shape) - sum(1 for e in entries if e is not None))
    basic, adv_dims, arrays, adv_entries = [], [], [], []
    for e_i, e in enumerate(entries):
        if e is None or isinstance(e, slice):
            basic.append(e)
        elif is_arr(e) or isinstance(e, (numbers.Integral, np.integer)):
            adv_dims.append(len(basic))
            basic.append(slice(None))
            arrays.append(e if is_arr(e) else int(e))
            adv_entries.append(e_i)
        else:
            raise IndexError(f"unsupported index {e!r}")
    adjacent = adv_entries[-1] - adv_entries[0] + 1 == len(adv_entries)
    return tuple(basic), adv_dims, arrays, (adv_dims[0] if adjacent else None)


def _mixed_view(a, mixed):
    """The basic view of ``a`` with the advanced axes moved to the front, ready for the leading-axes gather/scatter."""
    basic, adv_dims, arrays, pos = mixed
    v = getitem(a, basic)
    others = [d for d in range(v.ndim) if d not in adv_dims]
    return transpose(v, adv_dims + others), len(others)


def _basic_view(m: Mat, ops) -> Mat:
    if m.shifts is not None:
        raise AssertionError("rolled views must be made contiguous before slicing")
    shape, strides, offset = [], [], m.offset
    for op in ops:
        if op[0] == "new":
            shape.append(1)
            strides.append(0)
        elif op[0] == "int":
            _, dim, j = op
            offset += j * m.strides[dim]
        else:
            _, dim, sl = op
            start, stop, step = sl.indices(m.shape[dim])
            n = max(0, (stop - start + (step - (1 if step > 0 else -1))) // step)
            shape.append(n)
            strides.append(m.strides[dim] * step)
            offset += start * m.strides[dim]
    view = Mat(m.buf, tuple(shape), tuple(strides), offset, m.dtype, track=False)
    if m.buf.pending:
        lazy.note_view(m, view)        # a window on the result of a queued product: remember which part is wanted
    return view


def getitem(a: DeviceArray, idx):
    if _is_box(idx):
        return NotImplemented
    mixed = _mixed_index(a.shape, idx)
    if mixed is not None:
        v2, n_other = _mixed_view(a, mixed)
        r = getitem(v2, tuple(mixed[2]))
        nidx, pos = r.ndim - n_other, mixed[3]
        if pos and nidx:
            r = moveaxis(r, tuple(range(nidx)), tuple(range(pos, pos + nidx)))
        return r
    kind, info = _parse_index(a.shape, idx)
    if kind == "basic":
        m = a._mat()
        if m.shifts is not None:
            m = a._cmat()
        return DeviceArray(lazy.leaf(_basic_view(m, info)))
    idx_arrays, idx_shape, n_lead = info
    m = a._cmat()
    inner_shape = a.shape[n_lead:]
    if not inner_shape and prod(idx_shape) > 1:
        # every axis is indexed: a pure element lookup, kept symbolic and fused into its consumers
        table = a._node if a._node.mat is m else lazy.leaf(m)
        return DeviceArray(lazy.make_gather(table, [i._node for i in idx_arrays]))
    out = Mat.empty(tuple(idx_shape) + tuple(inner_shape), a.dtype, inexact=m.buf.inexact)
    mats = [i._cmat() for i in idx_arrays]
    backend.get().gather(out.ptr, m.ptr, a.dtype, [p.ptr for p in mats], list(a.shape[:n_lead]), prod(idx_shape),
                         prod(inner_shape))
    return DeviceArray(lazy.leaf(out))


def take(a, indices, axis=None, out=None, mode="raise"):
    a = asarray(a)
    if axis is None:
        return ravel(a)[indices]
    axis = _normalize_axis(axis, a.ndim)
    if axis == 0:
        return a[indices]
    moved = moveaxis(a, axis, 0)
    res = moved[indices]
    nidx = np.ndim(indices)
    return moveaxis(res, tuple(range(nidx)), tuple(range(axis, axis + nidx)))


# =================================================================== reductions
def _axes_tuple(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, (numbers.Integral, np.integer)):
        axis = (axis,)
    axes = tuple(sorted(_normalize_axis(int(ax), ndim) for ax in axis))
    if len(set(axes)) != len(axes):
        raise ValueError("duplicate value in 'axis'")
    return axes


def reduce_(a, op, axis=None, keepdims=False, dtype=None):
    """sum / prod / max / min / all / any over ``axis`` (None, int or tuple of ints)."""
    a = asarray(a)
    axes = _axes_tuple(axis, a.ndim)
    in_dtype = a.dtype
    if op in ("all", "any"):
        a = a.astype(np.bool_).astype(np.int64) if a.dtype != BOOL else a.astype(np.int64)
        kern_op, out_dtype, work_dtype = ("min" if op == "all" else "max"), BOOL, I64
    elif op in ("sum", "prod"):
        if dtype is not None:
            work_dtype = np.dtype(dtype)
        elif in_dtype.kind in "bi":
            work_dtype = I64
        else:
            work_dtype = in_dtype
        if work_dtype != a.dtype:
            a = a.astype(work_dtype)
        kern_op, out_dtype = op, work_dtype
    else:  # max / min keep the dtype
        out_dtype = in_dtype
        work_dtype = in_dtype if in_dtype in (F64, F32, I64) else I64
        if work_dtype != a.dtype:
            a = a.astype(work_dtype)
        kern_op = op
    shape = a.shape
    kept_shape = tuple(1 if i in axes else s for i, s in enumerate(shape))
    final_shape = kept_shape if keepdims else tuple(s for i, s in enumerate(shape) if i not in axes)
    if not axes:
        res = a
    elif a.size == 0 or any(shape[ax] == 0 for ax in axes):
        if op in ("max", "min"):
            raise ValueError("zero-size array to reduction operation which has no identity")
        ident = {"sum": 0, "prod": 1, "all": 1, "any": 0}[op]
        res = full(kept_shape, ident, work_dtype)
    elif op in _SHORT_REDUCE and (fast := _reduce_short_axes(a, op, axes, work_dtype)) is not None:
        res = fast                              # shape without the reduced axes; reshape below re-inserts them lazily
    else:
        be = backend.get()
        cur = a._cmat()
        cur_shape = list(shape)
        # reduce one run of adjacent axes at a time, last run first (keeps earlier axis numbers valid)
        runs, run = [], [axes[-1]]
        for ax in reversed(axes[:-1]):
            if ax == run[0] - 1:
                run.insert(0, ax)
            else:
                runs.append(run)
                run = [ax]
        runs.append(run)
        for run in runs:
            outer = prod(cur_shape[: run[0]])
            red = prod(cur_shape[run[0]: run[-1] + 1])
            inner = prod(cur_shape[run[-1] + 1:])
            out = Mat.empty((outer, inner), work_dtype,
                            inexact=work_dtype.kind == "f" and (kern_op in ("sum", "prod") or cur.buf.inexact))
            be.reduce(out.ptr, cur.ptr, work_dtype, kern_op, outer, red, inner)
            cur = out
            for ax in run:
                cur_shape[ax] = 1
        res = DeviceArray(lazy.leaf(Mat(cur.buf, kept_shape, lazy.c_strides(kept_shape), 0, work_dtype)))
    if res.shape != final_shape:
        res = reshape(res, final_shape)
    if res.dtype != out_dtype:
        res = res.astype(out_dtype)
    return res


def sum_(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    return reduce_(a, "sum", axis, keepdims, dtype)


def prod_(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    return reduce_(a, "prod", axis, keepdims, dtype)


def amax(a, axis=None, out=None, keepdims=False, **kw):
    return reduce_(a, "max", axis, keepdims)


def amin(a, axis=None, out=None, keepdims=False, **kw):
    return reduce_(a, "min", axis, keepdims)


def all_(a, axis=None, out=None, keepdims=False, **kw):
    return reduce_(a, "all", axis, keepdims)


def any_(a, axis=None, out=None, keepdims=False, **kw):
    return reduce_(a, "any", axis, keepdims)


def count_nonzero(a, axis=None, keepdims=False):
    return reduce_(asarray(a) != 0, "sum", axis, keepdims)


def _count(shape, axes):
    return prod(shape[ax] for ax in axes)


def mean(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    a = asarray(a)
    axes = _axes_tuple(axis, a.ndim)
    if a.dtype.kind in "bi":
        a = a.astype(np.float64)
    s = reduce_(a, "sum", axis, keepdims, dtype)
    return elementwise("divide", np.true_divide, s, _count(a.shape, axes))


def var(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw):
    a = asarray(a)
    axes = _axes_tuple(axis, a.ndim)
    if a.dtype.kind in "bi":
        a = a.astype(np.float64)
    mu = mean(a, axis, keepdims=True)
    d = a - mu
    s = reduce_(d * d, "sum", axis, keepdims)
    return elementwise("divide", np.true_divide, s, max(_count(a.shape, axes) - ddof, 0))


def std(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw):
    return elementwise("sqrt", np.sqrt, var(a, axis, dtype, out, ddof, keepdims))


def argreduce(a, op, axis=None, keepdims=False):
    a = asarray(a)
    if axis is None:
        flat = ravel(a)
        res = argreduce(flat, op, 0, False)
        return reshape(res, (1,) * a.ndim) if keepdims else res
    axis = _normalize_axis(axis, a.ndim)
    work = a if a.dtype in (F64, F32, I64) else a.astype(np.int64)
    m = work._cmat()
    shape = a.shape
    outer, red, inner = prod(shape[:axis]), shape[axis], prod(shape[axis + 1:])
    if red == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    out_shape = shape[:axis] + ((1,) if keepdims else ()) + shape[axis + 1:]
    out = Mat.empty(out_shape, I64)
    backend.get().argreduce(out.ptr, m.ptr, work.dtype, op, outer, red, inner)
    return DeviceArray(lazy.leaf(out))


def argmax(a, axis=None, out=None, keepdims=False): return argreduce(a, "max", axis, keepdims)
def argmin(a, axis=None, out=None, keepdims=False): return argreduce(a, "min", axis, keepdims)


def _logsumexp_weighted(a, axis, b, keepdims, return_sign):
    """``logsumexp(a, b=weights)`` / ``return_sign=True``: log|sum(b * exp(a))| with the usual max shift, composed
    from device reductions and fused elementwise kernels.  SciPy's conventions (scipy/special/_logsumexp.py): the
    shift is the max over the entries whose weight is non-zero, a non-finite shift is replaced by 0, the sign of
    the sum is returned separately (and the log of a negative sum is NaN when it is not asked for)."""
    a = asarray(a)
    if a.dtype.kind != "f":
        a = a.astype(np.float64)
    if b is not None:
        b = asarray(b)
        shape = np.broadcast_shapes(a.shape, b.shape)
        a, b = broadcast_to(a, shape), broadcast_to(b, shape)
        shift_src = where(b != 0, a, -np.inf)
    else:
        shift_src = a
    m = reduce_(shift_src, "max", axis, True)
    m = where(elementwise("isfinite", np.isfinite, m), m, 0.0)
    terms = elementwise("exp", np.exp, a - m)
    if b is not None:
        terms = b * terms
    s = reduce_(terms, "sum", axis, keepdims)
    if not keepdims:
        m = reshape(m, s.shape)
    if return_sign:
        sgn = elementwise("sign", np.sign, s)
        return elementwise("log", np.log, abs(s)) + m, sgn
    return elementwise("log", np.log, s) + m


def logsumexp(a, axis=None, b=None, keepdims=False, return_sign=False):
    """scipy.special.logsumexp on device (fused max / exp / sum / log kernel; special.py:120)."""
    if b is not None or return_sign:
        return _logsumexp_weighted(a, axis, b, keepdims, return_sign)
    a = asarray(a)
    if a.dtype.kind != "f":
        a = a.astype(np.float64)
    axes = _axes_tuple(axis, a.ndim)
    shape = a.shape
    kept_shape = tuple(1 if i in axes else s for i, s in enumerate(shape))
    final_shape = kept_shape if keepdims else tuple(s for i, s in enumerate(shape) if i not in axes)
    if a.ndim == 0:
        return a
    # bring the reduced axes together: they must form one adjacent run for the fused kernel
    if list(axes) != list(range(axes[0], axes[-1] + 1)):
        keep = [i for i in range(a.ndim) if i not in axes]
        a = transpose(a, keep + list(axes))
        shape = a.shape
        axes = tuple(range(len(keep), a.ndim))
    m = a._cmat()
    outer, red, inner = prod(shape[: axes[0]]), prod(shape[axes[0]: axes[-1] + 1]), prod(shape[axes[-1] + 1:])
    out = Mat.empty((outer, inner), a.dtype, inexact=True)
    if red == 0:
        return full(final_shape, -np.inf, a.dtype)
    backend.get().logsumexp(out.ptr, m.ptr, a.dtype, outer, red, inner)
    return DeviceArray(lazy.leaf(Mat(out.buf, final_shape, lazy.c_strides(final_shape), 0, a.dtype)))


# =================================================================== contractions
def _float_result(a, b):
    dt = np.result_type(a.dtype, b.dtype)
    if dt.kind != "f":
        dt = F64
    if dt not in (F64, F32):
        backend.dtype_code(dt)
    return np.dtype(dt)


def _as_gemm_operand(x: DeviceArray):
    """A 2-D (or batched 3-D) operand whose payload the GEMM kernel can address by strides."""
    if x.dtype not in (F64, F32):
        x = x.astype(np.float64)
    m = x._mat()
    if m.shifts is not None:
        m = x._cmat()
    return m


def _gemm2d(a: DeviceArray, b: DeviceArray, out_dtype=None) -> DeviceArray:
    """(M,K) @ (K,N) on the DMMA kernel; operands may be arbitrary strided 2-D views."""
    M, K = a.shape
    K2, N = b.shape
    if K != K2:
        raise ValueError(f"shapes {a.shape} and {b.shape} not aligned: {K} (dim 1) != {K2} (dim 0)")
    out_dtype = out_dtype or _float_result(a, b)
    am, bm = _as_gemm_operand(a), _as_gemm_operand(b)
    out = Mat.empty((M, N), out_dtype, inexact=True)     # accumulation order differs from OpenBLAS'
    if M and N:
        return DeviceArray(lazy.defer_gemm(out, am, bm, M, N, K))   # launched, grouped with its neighbours, when first needed
    return DeviceArray(lazy.leaf(out))


def tensordot(a, b, axes=2):
    a, b = asarray(a), asarray(b)
    if isinstance(axes, (numbers.Integral, np.integer)):
        n = int(axes)
        axes_a, axes_b = list(range(a.ndim - n, a.ndim)), list(range(0, n))
    else:
        axes_a, axes_b = axes
        axes_a = [axes_a] if isinstance(axes_a, (numbers.Integral, np.integer)) else list(axes_a)
        axes_b = [axes_b] if isinstance(axes_b, (numbers.Integral, np.integer)) else list(axes_b)
    axes_a = [_normalize_axis(int(x), a.ndim) for x in axes_a]
    axes_b = [_normalize_axis(int(x), b.ndim) for x in axes_b]
    if len(axes_a) != len(axes_b) or any(a.shape[i] != b.shape[j] for i, j in zip(axes_a, axes_b)):
        raise ValueError("shape-mismatch for sum")
    free_a = [i for i in range(a.ndim) if i not in axes_a]
    free_b = [i for i in range(b.ndim) if i not in axes_b]
    M = prod(a.shape[i] for i in free_a)
    K = prod(a.shape[i] for i in axes_a)
    N = prod(b.shape[i] for i in free_b)
    out_shape = tuple(a.shape[i] for i in free_a) + tuple(b.shape[i] for i in free_b)
    a2 = reshape(transpose(a, free_a + axes_a), (M, K))
    b2 = reshape(transpose(b, axes_b + free_b), (K, N))
    if K == 0:
        return zeros(out_shape, _float_result(a, b))
    if K == 1:
        # outer product (e.g. the 1-D . 1-D dot adjoint tensordot(g, B, 0), numpy_vjps.py:621): a lazy broadcast multiply
        # that fuses into its consumer instead of a K=1 GEMM over M x N tiles
        dt = _float_result(a, b)
        prodv = reshape(a2, (M, 1)) * reshape(b2, (1, N))
        return reshape(prodv if prodv.dtype == dt else prodv.astype(dt), out_shape)
    return reshape(_gemm2d(a2, b2), out_shape)


def dot(a, b, out=None):
    a, b = asarray(a), asarray(b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    if a.ndim == 1 and b.ndim == 1:
        if a.shape != b.shape:
            raise ValueError(f"shapes {a.shape} and {b.shape} not aligned")
        return reduce_(a * b, "sum", None, False)
    if b.ndim == 1:
        return tensordot(a, b, ([a.ndim - 1], [0]))
    return tensordot(a, b, ([a.ndim - 1], [b.ndim - 2]))


def matmul(a, b, out=None, **kw):
    a, b = asarray(a), asarray(b)
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: Input operand does not have enough dimensions")
    if a.ndim <= 2 and b.ndim <= 2:
        return dot(a, b)
    a_vec, b_vec = a.ndim == 1, b.ndim == 1
    if a_vec:
        a = reshape(a, (1, a.shape[0]))
    if b_vec:
        b = reshape(b, (b.shape[0], 1))
    batch_shape = lazy.broadcast_shapes([a.shape[:-2], b.shape[:-2]])
    M, K = a.shape[-2:]
    K2, N = b.shape[-2:]
    if K != K2:
        raise ValueError(f"matmul: core dimension mismatch {K} vs {K2}")
    out_dtype = _float_result(a, b)
    nb = prod(batch_shape)
    ab = reshape(broadcast_to(a, batch_shape + (M, K)).copy() if a.shape[:-2] != batch_shape else a, (nb, M, K))
    bb = reshape(broadcast_to(b, batch_shape + (K, N)).copy() if b.shape[:-2] != batch_shape else b, (nb, K, N))
    am, bm = _as_gemm_operand(ab), _as_gemm_operand(bb)
    out = Mat.empty((nb, M, N), out_dtype, inexact=True)
    if nb and M and N:
        backend.get().gemm(out.ptr, out_dtype, (N, 1, M * N), am.ptr, am.dtype, (am.strides[1], am.strides[2], am.strides[0]),
                           bm.ptr, bm.dtype, (bm.strides[1], bm.strides[2], bm.strides[0]), M, N, K, nb)
    res = reshape(DeviceArray(lazy.leaf(out)), batch_shape + (M, N))
    if a_vec:
        res = squeeze(res, -2)
    if b_vec:
        res = squeeze(res, -1)
    return res


def inner(a, b):
    a, b = asarray(a), asarray(b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    return tensordot(a, b, ([a.ndim - 1], [b.ndim - 1]))


def outer(a, b, out=None):
    a, b = ravel(asarray(a)), ravel(asarray(b))
    return reshape(a, (-1, 1)) * reshape(b, (1, -1))




# =================================================================== einsum (pairwise contraction on the GEMM kernel)
def _einsum_parse(subscripts, shapes):
    """Explicit input/output subscript lists for ``subscripts`` (ellipsis expanded to upper-case letters)."""
    subscripts = subscripts.replace(" ", "")
    if "->" in subscripts:
        lhs, out = subscripts.split("->")
    else:
        lhs, out = subscripts, None
    ins = lhs.split(",")
    if len(ins) != len(shapes):
        raise ValueError("more operands provided to einstein sum function than specified in the subscripts string")
    ell = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    max_ell = 0
    ex = []
    for sub, shp in zip(ins, shapes):
        if "..." in sub:
            n = len(shp) - (len(sub) - 3)
            if n < 0:
                raise ValueError("einsum: ellipsis does not match operand rank")
            max_ell = max(max_ell, n)
            sub = sub.replace("...", ell[len(ell) - n:] if n else "")
        if len(sub) != len(shp):
            raise ValueError(f"einsum: operand has {len(shp)} dims but subscripts {sub!r}")
        ex.append(sub)
    if out is None:
        counts = {}
        for sub in ex:
            for ch in sub:
                counts[ch] = counts.get(ch, 0) + 1
        out = "".join(sorted(ch for ch in counts if ch in ell)) + "".join(sorted(ch for ch, c in counts.items() if c == 1 and ch not in ell))
    else:
        out = out.replace("...", ell[len(ell) - max_ell:] if max_ell else "")
    return ex, out


def _einsum_diagonal(sub, o):
    """A repeated index inside one operand ('lli') reads that operand's diagonal over the repeated axes: a strided
    view (stride = sum of the axes' strides) kept at the first occurrence's position — no copy."""
    mt = o._mat()
    if mt.shifts is not None:
        mt = o._cmat()
    new_sub, shape, strides = [], [], []
    for i, ch in enumerate(sub):
        if ch in new_sub:
            j = new_sub.index(ch)
            if mt.shape[i] != shape[j]:
                raise ValueError(f"einsum: dimensions in operand for collapsing index {ch!r} don't match")
            strides[j] += mt.strides[i]
        else:
            new_sub.append(ch)
            shape.append(mt.shape[i])
            strides.append(mt.strides[i])
    return "".join(new_sub), DeviceArray(lazy.leaf(Mat(mt.buf, tuple(shape), tuple(strides), mt.offset, mt.dtype)))


def einsum(*operands, **kwargs):
    """np.einsum for device operands: every pairwise step is transposes + one (batched) DMMA GEMM."""
    if kwargs.get("out") is not None:
        raise NotImplementedError("autograd_b200: einsum(out=...) is not supported")
    if not isinstance(operands[0], str):
        # (op0, sublist0, op1, sublist1, ..., [sublistout]) convention -> subscript string
        letters = "abcdefghijklmnopqrstuvwxyz"
        ops, subs = [], []
        rest = list(operands)
        while len(rest) >= 2:
            ops.append(rest.pop(0))
            sl = rest.pop(0)
            subs.append("".join("..." if x is Ellipsis else letters[x] for x in sl))
        spec = ",".join(subs)
        if rest:
            spec += "->" + "".join("..." if x is Ellipsis else letters[x] for x in rest[0])
        return einsum(spec, *ops)
    spec, ops = operands[0], [asarray(o) for o in operands[1:]]
    ins, out = _einsum_parse(spec, [o.shape for o in ops])
    for k, sub in enumerate(ins):
        if len(set(sub)) != len(sub):
            ins[k], ops[k] = _einsum_diagonal(sub, ops[k])
    sizes = {}
    for sub, o in zip(ins, ops):
        for ch, n in zip(sub, o.shape):
            if n != 1:
                if sizes.get(ch, 1) not in (1, n):    # a size-1 axis broadcasts against any extent (NumPy einsum)
                    raise ValueError(f"einsum: size mismatch for index {ch!r}")
                sizes[ch] = n
            sizes.setdefault(ch, n)
    # broadcast size-1 dims that pair with larger ones
    ops = [broadcast_to(o, tuple(sizes[ch] for ch in sub)) if o.shape != tuple(sizes[ch] for ch in sub) else o
           for sub, o in zip(ins, ops)]

    def reduce_unneeded(sub, o, needed):
        drop = [i for i, ch in enumerate(sub) if ch not in needed]
        if drop:
            o = reduce_(o, "sum", tuple(drop), False)
            sub = "".join(ch for ch in sub if ch in needed)
        return sub, o

    cur_sub, cur = ins[0], ops[0]
    for k in range(1, len(ops)):
        later = set(out).union(*ins[k + 1:]) if k + 1 < len(ops) else set(out)
        cur_sub, cur = reduce_unneeded(cur_sub, cur, later | set(ins[k]))
        b_sub, b = reduce_unneeded(ins[k], ops[k], later | set(cur_sub))
        batch = [ch for ch in cur_sub if ch in b_sub and ch in later]
        contr = [ch for ch in cur_sub if ch in b_sub and ch not in later]
        free_a = [ch for ch in cur_sub if ch not in b_sub]
        free_b = [ch for ch in b_sub if ch not in cur_sub]
        a_t = transpose(cur, [cur_sub.index(ch) for ch in batch + free_a + contr])
        b_t = transpose(b, [b_sub.index(ch) for ch in batch + contr + free_b])
        nb = prod(sizes[ch] for ch in batch)
        M, K, N = prod(sizes[ch] for ch in free_a), prod(sizes[ch] for ch in contr), prod(sizes[ch] for ch in free_b)
        res = matmul(reshape(a_t, (nb, M, K)), reshape(b_t, (nb, K, N))) if nb != 1 or batch else \
            _gemm2d(reshape(a_t, (M, K)), reshape(b_t, (K, N)))
        cur_sub = "".join(batch + free_a + free_b)
        cur = reshape(res, tuple(sizes[ch] for ch in cur_sub))
    cur_sub, cur = reduce_unneeded(cur_sub, cur, set(out))
    if cur_sub != out:
        cur = transpose(cur, [cur_sub.index(ch) for ch in out])
    return cur


# =================================================================== convolution (NCHW, tiny channel counts)
def _conv_operands(x, w):
    x, w = asarray(x), asarray(w)
    dt = _float_result(x, w)
    if x.dtype != dt:
        x = x.astype(dt)
    if w.dtype != dt:
        w = w.astype(dt)
    return x, w, dt


def conv2d(x, w, mode="valid", flip=True):
    """x[N,C,H,W] (*) w[C,F,kh,kw] -> [N,F,Ho,Wo]; ``flip`` = true convolution (signal.py:49-54)."""
    x, w, dt = _conv_operands(x, w)
    N, C, H, W = x.shape
    C2, F, kh, kw = w.shape
    if C != C2:
        raise ValueError(f"convolve: channel mismatch {C} vs {C2}")
    if mode == "valid":
        Ho, Wo = H - kh + 1, W - kw + 1
    elif mode == "full":
        Ho, Wo = H + kh - 1, W + kw - 1
    else:
        raise NotImplementedError(f"autograd_b200: convolve mode {mode!r} is not implemented on device")
    if Ho <= 0 or Wo <= 0:
        raise ValueError("One array must be larger than the other along all convolved dimensions")
    out = Mat.empty((N, F, Ho, Wo), dt, inexact=True)
    xm, wm = x._cmat(), w._cmat()  # keep the (possibly temporary) contiguous payloads alive across the launch
    backend.get().conv2d(out.ptr, xm.ptr, wm.ptr, dt, N, C, H, W, F, kh, kw, 0 if mode == "valid" else 1,
                         1 if flip else 0)
    return DeviceArray(lazy.leaf(out))


def conv2d_input_grad(g, w, x_shape, mode="valid"):
    """VJP of conv2d wrt x (signal.py:153-214, argnum 0): correlate g with the channel-swapped filters."""
    wt = swapaxes(asarray(w), 0, 1)  # [F, C, kh, kw]
    return conv2d(g, wt, mode="full" if mode == "valid" else "valid", flip=False)


def conv2d_filter_grad(g, x, w_shape, mode="valid"):
    """VJP of conv2d wrt w (signal.py:153-214, argnum 1): batch-summed valid correlation of x with g."""
    if mode != "valid":
        raise NotImplementedError("autograd_b200: filter gradient of a 'full' convolution is not implemented on device")
    g, x, dt = _conv_operands(g, x)
    N, C, H, W = x.shape
    _, F, kh, kw = w_shape
    out = Mat.empty((C, F, kh, kw), dt, inexact=True)
    xm, gm = x._cmat(), g._cmat()
    backend.get().conv2d_wgrad(out.ptr, xm.ptr, gm.ptr, dt, N, C, H, W, F, kh, kw)
    return DeviceArray(lazy.leaf(out))


def convolve_nd(a, b, axes=None, dot_axes=None, mode="full"):
    """General N-d ``autograd.scipy.signal.convolve`` on device operands (scipy/signal.py:10-76): any set of
    convolved axes, contracted ("dot") axes and untouched axes, modes 'valid' / 'full' / 'same'.

    Same formulation as the reference, so the output axis order (untouched axes of ``a``, untouched axes of ``b``,
    convolved axes) and the 'same'-mode alignment agree: the larger operand is zero-padded, read through a
    sliding-window view (one extra trailing axis per convolved axis; a pure stride edit — the windows are never
    copied as such), the smaller one is reversed along the convolved axes (a negative-stride view), and the two are
    contracted over window + dot axes by the device einsum (transposes + DMMA GEMM).  The NCHW pattern of
    examples/convnet.py does not come here: it has dedicated kernels (conv2d / conv2d_wgrad)."""
    if mode not in ("valid", "full", "same"):
        raise ValueError(f"Mode {mode} undefined, it can be one of 'valid', 'full', and 'same'")
    a, b = asarray(a), asarray(b)
    dt = _float_result(a, b) if (a.dtype.kind == "f" or b.dtype.kind == "f") else np.result_type(a.dtype, b.dtype)
    a = a if a.dtype == dt else a.astype(dt)
    b = b if b.dtype == dt else b.astype(dt)
    ca, cb = ([int(i) for i in ax] for ax in (axes if axes is not None else (range(a.ndim), range(a.ndim))))
    da, db = ([int(i) for i in ax] for ax in (dot_axes if dot_axes is not None else ((), ())))
    n_free_b = b.ndim - len(db) - len(cb)
    n_free_a = a.ndim - len(da) - len(ca)

    if any(b.shape[j] < a.shape[i] for i, j in zip(ca, cb)):
        # the filter role goes to the operand that is smaller along the convolved axes: compute with swapped roles
        # and move the untouched axes back to (a's, b's, convolved) order
        if mode == "valid" and not all(b.shape[j] <= a.shape[i] for i, j in zip(ca, cb)):
            raise ValueError("One array must be larger than the other along all convolved dimensions")
        if mode == "valid" or (mode == "full" and b.size <= a.size):
            swapped = convolve_nd(b, a, axes=(cb, ca), dot_axes=(db, da), mode=mode)
            order = (list(range(n_free_b, n_free_b + n_free_a)) + list(range(n_free_b))
                     + list(range(n_free_b + n_free_a, swapped.ndim)))
            return transpose(swapped, order)

    # roles: ``filt`` is reversed, ``slid`` is padded and windowed ('same' keeps a's extent, so a is the one slid over)
    filt, slid, cf, cs, df, ds = (b, a, cb, ca, db, da) if mode == "same" else (a, b, ca, cb, da, db)
    if mode != "valid":
        widths = [(0, 0)] * slid.ndim
        for i_f, i_s in zip(cf, cs):
            k = filt.shape[i_f] - 1
            widths[i_s] = (k, k) if mode == "full" else (k - k // 2, k // 2)
        slid = pad(slid, widths)
    sm = slid._cmat()
    win_shape, win_strides = list(sm.shape), list(sm.strides)
    for i_f, i_s in zip(cf, cs):
        win_shape.append(abs(sm.shape[i_s] - filt.shape[i_f]) + 1)     # output positions along this axis
        win_strides.append(sm.strides[i_s])
        win_shape[i_s] = filt.shape[i_f]                                # extent of one window
    windows = DeviceArray(lazy.leaf(Mat(sm.buf, tuple(win_shape), tuple(win_strides), sm.offset, sm.dtype)))
    rev = [slice(None)] * filt.ndim
    for i_f in cf:
        rev[i_f] = slice(None, None, -1)
    filt = filt[tuple(rev)]

    def contract(x, y, x_axes, y_axes):
        # einsum in sublist form: contracted pairs share a label, every other axis keeps its own (x's first)
        lx, ly = list(range(x.ndim)), list(range(x.ndim, x.ndim + y.ndim))
        for k, (i, j) in enumerate(zip(x_axes, y_axes)):
            lx[i] = ly[j] = x.ndim + y.ndim + k
        return einsum(x, lx, y, ly)

    if mode == "same":
        out = contract(windows, filt, list(cs) + list(ds), list(cf) + list(df))
        # einsum order: a's untouched axes, output positions, b's untouched axes -> (a's, b's, positions)
        na, nc = a.ndim - len(da) - len(ca), len(ca)
        return transpose(out, list(range(na)) + list(range(na + nc, out.ndim)) + list(range(na, na + nc)))
    return contract(filt, windows, list(cf) + list(df), list(cs) + list(ds))


# =================================================================== dispatch tables
UFUNC_OPS = {
    np.add: "add", np.subtract: "subtract", np.multiply: "multiply", np.true_divide: "divide",
    np.negative: "negative", np.positive: "positive", np.power: "power", np.float_power: "power",
    np.square: "square", np.reciprocal: "reciprocal", np.absolute: "abs", np.fabs: "abs", np.sign: "sign",
    np.exp: "exp", np.exp2: "exp2", np.expm1: "expm1", np.log: "log", np.log2: "log2", np.log10: "log10",
    np.log1p: "log1p", np.sqrt: "sqrt", np.cbrt: "cbrt",
    np.sin: "sin", np.cos: "cos", np.tan: "tan", np.arcsin: "arcsin", np.arccos: "arccos", np.arctan: "arctan",
    np.sinh: "sinh", np.cosh: "cosh", np.tanh: "tanh", np.arcsinh: "arcsinh", np.arccosh: "arccosh",
    np.arctanh: "arctanh", np.arctan2: "arctan2", np.hypot: "hypot", np.logaddexp: "logaddexp",
    np.floor: "floor", np.ceil: "ceil", np.trunc: "trunc", np.rint: "rint",
    np.remainder: "remainder", np.floor_divide: "floor_divide",
    np.maximum: "maximum", np.minimum: "minimum", np.fmax: "fmax", np.fmin: "fmin",
    np.equal: "equal", np.not_equal: "not_equal", np.less: "less", np.less_equal: "less_equal",
    np.greater: "greater", np.greater_equal: "greater_equal",
    np.logical_and: "logical_and", np.logical_or: "logical_or", np.logical_xor: "logical_xor",
    np.logical_not: "logical_not", np.isfinite: "isfinite", np.isnan: "isnan", np.isinf: "isinf",
    np.conjugate: "conjugate", np.copysign: "copysign",
    np.deg2rad: "deg2rad", np.radians: "deg2rad", np.rad2deg: "rad2deg", np.degrees: "rad2deg",
    np.logaddexp2: "logaddexp2",
}
try:   # scipy.special's ufuncs arrive through __array_ufunc__ like NumPy's (autograd/scipy/special.py wraps them)
    import scipy.special as _sps

    for _name, _op in (("erf", "sp_erf"), ("erfc", "sp_erfc"), ("erfinv", "sp_erfinv"), ("erfcinv", "sp_erfcinv"),
                       ("erfcx", "sp_erfcx"), ("gammaln", "sp_gammaln"), ("gamma", "sp_gamma"), ("psi", "sp_digamma"),
                       ("digamma", "sp_digamma"), ("ndtr", "sp_ndtr"), ("ndtri", "sp_ndtri"), ("log_ndtr", "sp_log_ndtr"),
                       ("expit", "sp_expit"), ("logit", "sp_logit"), ("j0", "sp_j0"), ("j1", "sp_j1"), ("y0", "sp_y0"),
                       ("y1", "sp_y1"), ("i0", "sp_i0"), ("i1", "sp_i1"), ("gammainc", "sp_gammainc"),
                       ("gammaincc", "sp_gammaincc"), ("betainc", "sp_betainc")):
        _uf = getattr(_sps, _name, None)
        if isinstance(_uf, np.ufunc):
            UFUNC_OPS[_uf] = _op
    UFUNC_ORDER_OPS = {getattr(_sps, _n): _o for _n, _o in (("jv", "sp_jn"), ("yn", "sp_yn"), ("iv", "sp_iv"), ("ive", "sp_ive"))
                       if isinstance(getattr(_sps, _n, None), np.ufunc)}
except ImportError:   # SciPy is an optional dependency of the reference (pyproject.toml:55-56)
    UFUNC_ORDER_OPS = {}
UFUNC_REDUCE = {np.add: "sum", np.multiply: "prod", np.maximum: "max", np.minimum: "min",
                np.logical_and: "all", np.logical_or: "any"}


def _result_type(*args):
    return np.result_type(*[a.dtype if isinstance(a, DeviceArray) else a for a in args])


def _np_shape(a): return a.shape
def _np_ndim(a): return a.ndim
def _np_size(a, axis=None): return a.size if axis is None else a.shape[axis]


FUNCTIONS = {
    np.shape: _np_shape, np.ndim: _np_ndim, np.size: _np_size, np.result_type: _result_type,
    np.iscomplexobj: iscomplexobj, np.isrealobj: isrealobj,
    np.reshape: reshape, np.ravel: ravel, np.transpose: transpose, np.swapaxes: swapaxes, np.moveaxis: moveaxis,
    np.expand_dims: expand_dims, np.squeeze: squeeze, np.atleast_1d: atleast_1d, np.atleast_2d: atleast_2d, np.atleast_3d: atleast_3d,
    np.broadcast_to: broadcast_to, np.flip: flip, np.flipud: flipud, np.fliplr: fliplr, np.roll: roll,
    np.concatenate: concatenate, np.stack: stack, np.repeat: repeat, np.tile: tile, np.pad: pad, np.take: take,
    np.sum: sum_, np.prod: prod_, np.max: amax, np.amax: amax, np.min: amin, np.amin: amin, np.mean: mean,
    np.var: var, np.std: std, np.all: all_, np.any: any_, np.count_nonzero: count_nonzero,
    np.argmax: argmax, np.argmin: argmin,
    np.dot: dot, np.matmul: matmul, np.tensordot: tensordot, np.inner: inner, np.outer: outer, np.einsum: einsum,
    np.where: where, np.clip: clip, np.round: round_, np.around: round_, np.nan_to_num: nan_to_num,
    np.real: real, np.imag: imag, np.isclose: isclose, np.allclose: allclose, np.array_equal: array_equal,
    np.zeros_like: zeros_like, np.ones_like: ones_like, np.empty_like: empty_like, np.full_like: full_like,
    np.copy: lambda a, **kw: asarray(a).copy(),
    np.astype: lambda a, dtype, copy=True, **kw: asarray(a).astype(dtype),
    np.meshgrid: meshgrid, np.linspace: linspace,
    np.tril: tril, np.triu: triu, np.diag: diag, np.diagonal: diagonal, np.trace: trace, np.diff: diff,
    np.cumsum: cumsum, np.cumprod: cumprod, np.sort: sort, np.argsort: argsort, np.partition: partition, np.argpartition: argpartition, np.kron: kron, np.rot90: rot90,
    np.rollaxis: rollaxis, np.average: average, np.select: select, np.sinc: sinc, np.cross: cross,
    np.gradient: gradient, np.angle: angle, np.real_if_close: real_if_close, np.linalg.norm: norm,
}
# NumPy's own pure-Python implementations of these only call other dispatched functions.
DECOMPOSE = {np.hstack, np.vstack, np.dstack, np.column_stack, np.atleast_3d,
             np.split, np.array_split, np.hsplit, np.vsplit, np.dsplit}
