"""Lazy elementwise expression graph and its flush into fused kernels.

Every NumPy ufunc call that reaches the device type (forward primitives at
autograd/tracer.py:94 and the operator chains inside VJP closures,
autograd/numpy/numpy_vjps.py:76-191) only records a :class:`Node` here.  Nothing
is computed until a value is needed (a reduction, a contraction, a gather, a host
read); then the whole DAG behind that value is compiled into ONE grid-stride
kernel that reads each distinct input once and writes each needed output once —
the "tape fusion" of BASELINE.json's north_star.  Each arithmetic op keeps
NumPy's loop dtype and is rounded on its own (``--fmad=false``), so a fused chain
gives the same bits as the chain of NumPy ufunc calls it replaces.
"""

from __future__ import annotations

import itertools
import math
import weakref

import numpy as np

from . import backend

F64 = np.dtype(np.float64)
F32 = np.dtype(np.float32)
I64 = np.dtype(np.int64)
I32 = np.dtype(np.int32)
BOOL = np.dtype(np.bool_)
SUPPORTED = (F64, F32, I64, I32, BOOL)

_ids = itertools.count(1)

# Relative ALU cost of one op (used to decide what to keep vs. recompute).
_COST = {
    "add": 1, "subtract": 1, "multiply": 1, "negative": 1, "positive": 0, "square": 1, "abs": 1, "sign": 2,
    "where": 1, "cast": 1, "full": 0, "maximum": 2, "minimum": 2, "fmax": 2, "fmin": 2,
    "equal": 1, "not_equal": 1, "less": 1, "less_equal": 1, "greater": 1, "greater_equal": 1,
    "logical_and": 1, "logical_or": 1, "logical_not": 1, "logical_xor": 1,
    "isfinite": 1, "isnan": 1, "isinf": 1, "floor": 1, "ceil": 1, "trunc": 1, "rint": 1,
    "divide": 6, "reciprocal": 6, "sqrt": 6, "remainder": 4, "floor_divide": 4, "gather": 4,
    "reshape": 0,
}
_DEFAULT_COST = 12  # transcendental functions

# An interior node that Python can still see (a VJP closure holds it) is written
# out as an extra output of the fused kernel once recomputing it would cost more
# than this; cheaper ones are simply recomputed by whoever needs them later.
_HEAVY_VEC = int(__import__("os").environ.get("B200AG_HEAVY_VEC", "2"))
_ROW_TUNE = tuple(int(v) for v in __import__("os").environ.get("B200AG_ROW_TUNE", "").split(",")) \
    if __import__("os").environ.get("B200AG_ROW_TUNE") else None
_GRID_MULT = int(__import__("os").environ.get("B200AG_GRID_MULT", "16"))
_GRID_MULT_HEAVY = int(__import__("os").environ.get("B200AG_GRID_MULT_HEAVY", __import__("os").environ.get("B200AG_GRID_MULT", "64")))
# Repeated op-units per element of the SMALLER space (sub-expression cost x (replication - 1)) below which a broadcast
# sub-expression is recomputed in place: writing it out and reading it back (16 B per element) costs the FP64 pipe time of
# roughly 50 instructions.  Measured on C5 (scripts/gpu_r2_zstr.sh): the max-pool VJP chain (chooser of the outer max, cost
# ~30, replicated 2x into the chooser of the inner max) 10.68 ms fully fused against 11.93 ms cut in two.
EXPAND_MIN_COST = int(__import__("os").environ.get("B200AG_EXPAND_MIN_COST", "64"))
EXPAND_WORK = float(__import__("os").environ.get("B200AG_EXPAND_WORK", "3.0e7"))     # repeated op-units (elements x cost) that justify one extra launch, see materialize()
KEEP_COST = 96
# Re-materialising nodes that keep being recomputed was measured on C4 (fluidsim): +1 % speed for +65 % retained
# memory (the 200-step tape would no longer fit 180 GB), so it is off; the hook stays for experiments.
REUSE_LIMIT = int(__import__("os").environ.get("B200AG_REUSE_LIMIT", str(1 << 30)))
# B200AG_EXACT=1 keeps one IEEE rounding per recorded ufunc call in every generated kernel (no contraction anywhere);
# by default only values that are still bit-reproducible against NumPy are computed that way (Buffer.inexact).
RELAXED = __import__("os").environ.get("B200AG_EXACT", "0") in ("", "0")
UNROLL_WINDOWS = __import__("os").environ.get("B200AG_UNROLL_WINDOWS", "1") != "0"
LAZY_RESHAPE = __import__("os").environ.get("B200AG_LAZY_RESHAPE", "1") != "0"
UNROLL_INNER = __import__("os").environ.get("B200AG_UNROLL_INNER", "1") != "0"
KEEP_WIDE = __import__("os").environ.get("B200AG_KEEP_WIDE", "1") != "0"
KEEP_WIDE_USES = int(__import__("os").environ.get("B200AG_KEEP_WIDE_USES", "1"))
# Hard caps that force a flush while a graph is still being built.
MAX_NODES = 1500
MAX_LEAVES = 48


class Buffer:
    """Owner of one device allocation; freed back to the pool when unreferenced."""

    __slots__ = ("ptr", "nbytes", "version", "inexact", "readers", "aliases", "pending", "__weakref__")

    def __init__(self, nbytes: int, upload: bool = False):
        self.nbytes = int(nbytes)
        self.version = 0          # bumped before every in-place write (DeviceArray._own_buffer): invalidates memoised copies
        # True once the float payload can no longer be bit-identical to what NumPy computes from the same inputs: it is
        # (or derives from) the result of a transcendental function, a float sum / contraction (reduction order), an
        # atomic accumulation.  Generated kernels contract multiply-add pairs and share reciprocals freely on such
        # values (they are held to north_star's 1e-12, not to bit-exactness) and keep one IEEE rounding per recorded
        # ufunc call everywhere else (codegen.taint).
        self.inexact = False
        # pending (not yet computed) expression nodes that read this buffer directly; an in-place write materialises them
        # first so that `y = x * 2; x[0] = 100` leaves y computed from the old x, as NumPy's eager evaluation does
        self.readers = None
        self.aliases = None       # alias groups (lazy._AliasGroup) whose payload lives here
        self.pending = False      # True while a deferred contraction (defer_gemm) still has to write this buffer
        be = backend.get()
        self.ptr = be.malloc_upload(self.nbytes) if upload else be.malloc(self.nbytes)

    def __del__(self):
        ptr, self.ptr = self.ptr, 0
        if ptr:
            try:
                backend.get().free(ptr)
            except Exception:
                pass


# ---------------------------------------------------------------- deferred contractions
# np.dot on device operands does not launch its GEMM at once: the product is queued and launched when something needs
# its result (Mat.ptr), when the queue is full, before an in-place write, a synchronisation or the end of a graph
# capture.  Products queued together go to the GPU as ONE grouped launch (b200ag_gemm_grouped): the four gate products of
# an LSTM cell (examples/lstm.py:38-49) and the G.B^T / A^T.G adjoints of a tape step (numpy_vjps.py:615-638) are each too
# small to fill 148 SMs on their own.  Operands are materialised when the product is queued and held until it runs.
GEMM_QUEUE_MAX = 16
DEFER_GEMMS = __import__("os").environ.get("B200AG_DEFER_GEMM", "1") != "0"
_gemm_queue = []
_gemm_by_buf = {}      # id(output Buffer) -> queued record


class _QueuedGemm:
    """out[M,N] = a[M,K] @ b[K,N], not launched yet.  While it waits, basic slices taken of the result are noted
    (``note_view``): if, when it is finally launched, nobody can reach the full result any more (its array object is
    gone) and every window on it is a column range ``[:, c0:c1]``, only those columns are computed - the VJP of
    ``hstack((input, hiddens, ones)) @ W`` (numpy_vjps.py:615-624, :750-759) is G @ W^T over all 641 columns, of which
    autograd slices out the 512 that belong to the traced ``hiddens``."""

    __slots__ = ("out", "a", "b", "M", "N", "K", "a_ptr", "b_ptr", "node_ref", "lo", "hi", "all")

    def __init__(self, out, a, b, M, N, K):
        self.out, self.a, self.b, self.M, self.N, self.K = out, a, b, M, N, K
        self.a_ptr, self.b_ptr = a.ptr, b.ptr      # taking the operand addresses launches what THEY still wait for
        self.node_ref, self.lo, self.hi, self.all = None, None, None, False

    def problem(self):
        out, a, b = self.out, self.a, self.b
        c0, n = 0, self.N
        if SLICE_GEMMS and not self.all and self.lo is not None and self.node_ref is not None and self.node_ref() is None:
            c0, n = self.lo, self.hi - self.lo
        return (out.buf.ptr + (out.offset + c0 * out.strides[1]) * out.dtype.itemsize, out.dtype,
                (out.strides[0], out.strides[1]), self.a_ptr, a.dtype, (a.strides[0], a.strides[1]),
                self.b_ptr + c0 * b.strides[1] * b.dtype.itemsize, b.dtype, (b.strides[0], b.strides[1]), self.M, n, self.K)


SLICE_GEMMS = __import__("os").environ.get("B200AG_SLICE_GEMM", "1") != "0"


def defer_gemm(out: "Mat", a: "Mat", b: "Mat", M: int, N: int, K: int) -> "Node":
    """Queue out[M,N] = a[M,K] @ b[K,N] (2-D strided operands, float32/float64); returns the leaf node of ``out``."""
    rec = _QueuedGemm(out, a, b, M, N, K)
    node = leaf(out)
    if not DEFER_GEMMS:
        backend.get().gemm_grouped([rec.problem()])
        return node
    rec.node_ref = weakref.ref(node)
    out.buf.pending = True
    _gemm_queue.append(rec)
    _gemm_by_buf[id(out.buf)] = rec
    if len(_gemm_queue) >= GEMM_QUEUE_MAX:
        flush_gemms()
    return node


def _demand_all(buf) -> None:
    rec = _gemm_by_buf.get(id(buf))
    if rec is not None:
        rec.all = True


def note_view(m: "Mat", view: "Mat") -> None:
    """A basic slice ``view`` of ``m`` was taken while ``m``'s buffer is still waiting for a queued product."""
    rec = _gemm_by_buf.get(id(m.buf))
    if rec is None:
        return
    if m is rec.out and len(view.shape) == 2 and view.shape[0] == rec.M and view.strides == m.strides and view.shape[1] > 0:
        c0 = view.offset - m.offset
        if 0 <= c0 and c0 + view.shape[1] <= rec.N:
            rec.lo = c0 if rec.lo is None else min(rec.lo, c0)
            rec.hi = c0 + view.shape[1] if rec.hi is None else max(rec.hi, c0 + view.shape[1])
            return
    rec.all = True


def flush_gemms() -> None:
    """Launch every queued product (one grouped launch per operand class)."""
    if not _gemm_queue:
        return
    queue = _gemm_queue[:]
    del _gemm_queue[:]
    _gemm_by_buf.clear()
    for rec in queue:
        rec.out.buf.pending = False
    backend.get().gemm_grouped([rec.problem() for rec in queue])


def flush_readers(buf) -> None:
    """Compute every pending expression that reads ``buf`` (called before an in-place write to it)."""
    if _gemm_queue:
        flush_gemms()          # a queued product may read the buffer that is about to change
    if _scatter_queue:
        flush_scatters(reading=buf)
    if buf.aliases:
        settle_aliases(buf)    # pending reshapes of an array that lives here become views of it: they must see the write
    readers = buf.readers
    if not readers:
        return
    buf.readers = None
    for node in list(readers):
        if node.mat is None and node.args:
            materialize(node)


def c_strides(shape):
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def prod(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n


class Mat:
    """A materialised array: a strided (optionally circularly shifted) window on a Buffer."""

    __slots__ = ("buf", "shape", "strides", "offset", "dtype", "shifts")

    def __init__(self, buf, shape, strides, offset, dtype, shifts=None, track=True):
        if track and buf.pending:
            _demand_all(buf)      # an untracked new window on the output of a queued product: all of it may be read
        self.buf = buf
        self.shape = tuple(shape)
        self.strides = tuple(strides)
        self.offset = offset
        self.dtype = dtype
        self.shifts = shifts  # None or per-dim circular shift (np.roll as an index transform)

    @property
    def ptr(self) -> int:
        """Device address of the first element.  Every kernel launch takes its operand addresses from here, so this
        is also where a deferred contraction that still has to produce the payload is launched."""
        if self.buf.pending:
            flush_deferred()
        return self.buf.ptr + self.offset * self.dtype.itemsize

    @property
    def is_contiguous(self) -> bool:
        if self.shifts is not None:
            return False
        acc = 1
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != acc:
                return False
            acc *= s
        return True

    @staticmethod
    def empty(shape, dtype, upload=False, inexact=False):
        shape = tuple(int(s) for s in shape)
        buf = Buffer(prod(shape) * dtype.itemsize, upload)
        buf.inexact = inexact
        return Mat(buf, shape, c_strides(shape), 0, dtype)


class Const:
    """A host scalar operand.  ``dtype is None`` marks a weak Python scalar (NEP 50)."""

    __slots__ = ("value", "dtype")

    def __init__(self, value, dtype=None):
        self.value = value
        self.dtype = dtype


_EMPTY = frozenset()


class Node:
    __slots__ = ("op", "args", "in_dtypes", "shape", "dtype", "mat", "nid", "lo", "cost", "leafset", "nhandles",
                 "uses", "alias", "consumers", "__weakref__")

    def __init__(self, op, args, in_dtypes, shape, dtype, mat=None):
        self.alias = None     # the group of arrays that are symbolic RESHAPES of one another (NumPy: views of one buffer)
        self.consumers = None # weakrefs to the nodes recorded with this one as a PENDING operand (see payload_set)
        self.op = op
        self.args = args
        self.in_dtypes = in_dtypes
        self.shape = shape
        self.dtype = dtype
        self.mat = mat
        self.nid = next(_ids)
        self.nhandles = 0
        self.uses = 0  # times this node was recomputed as an interior node of somebody else's kernel
        if mat is None:
            lo, cost, leafset = self.nid, _COST.get(op, _DEFAULT_COST), _EMPTY
            for a in args:
                if type(a) is Node:
                    if a.mat is None:
                        if a.lo < lo:
                            lo = a.lo
                        cost += a.cost
                        leafset = leafset | a.leafset if leafset else a.leafset
                        if a.consumers is None:
                            a.consumers = [weakref.ref(self)]
                        else:
                            a.consumers.append(weakref.ref(self))
                    else:
                        leafset = leafset | a.leafset
                        buf = a.mat.buf
                        if buf.readers is None:
                            buf.readers = weakref.WeakSet()
                        buf.readers.add(self)
            self.lo = lo
            self.cost = cost if cost < 1 << 30 else 1 << 30
            self.leafset = leafset
        else:
            self.lo = self.nid
            self.cost = 0
            self.leafset = frozenset((self.nid,))


def leaf(mat: Mat) -> Node:
    return Node(None, (), (), mat.shape, mat.dtype, mat)


# ---------------------------------------------------------------- dtype resolution (NumPy's own rules)
_resolve_cache = {}


def _type_key(a):
    if type(a) is Node:
        return a.dtype
    if a.dtype is not None:
        return a.dtype
    v = a.value
    if isinstance(v, bool):
        return BOOL
    if isinstance(v, int):
        return int
    if isinstance(v, float):
        return float
    raise TypeError(f"autograd_b200: unsupported scalar operand {v!r} of type {type(v).__name__}")


def resolve(ufunc, args):
    """(input loop dtypes, output dtype) NumPy would use for ``ufunc(*args)``."""
    key = (ufunc,) + tuple(_type_key(a) for a in args)
    hit = _resolve_cache.get(key)
    if hit is None:
        sig = key[1:] + (None,) * ufunc.nout
        try:
            res = ufunc.resolve_dtypes(sig)
        except TypeError as e:
            raise TypeError(f"autograd_b200: {ufunc.__name__} is not defined for operand types {key[1:]}: {e}") from None
        ins, out = tuple(res[: ufunc.nin]), res[ufunc.nin]
        for d in ins + (out,):
            if d not in SUPPORTED:
                raise TypeError(
                    f"autograd_b200: {ufunc.__name__}{key[1:]} needs dtype {d}, which the device backend does not support"
                )
        hit = (ins, out)
        _resolve_cache[key] = hit
    return hit


def broadcast_shapes(shapes):
    first = shapes[0]
    same = True
    for s in shapes:
        if s != first:
            same = False
            break
    if same:
        return first
    nd = max(len(s) for s in shapes)
    out = [1] * nd
    for s in shapes:
        off = nd - len(s)
        for i, d in enumerate(s):
            j = off + i
            if d != 1:
                if out[j] == 1:
                    out[j] = d
                elif out[j] != d:
                    raise ValueError(f"operands could not be broadcast together with shapes {' '.join(map(str, shapes))}")
    return tuple(out)


def make_op(op, ufunc, args):
    """Record ``ufunc(*args)``; args are Nodes and Consts."""
    in_dtypes, out_dtype = resolve(ufunc, args)
    shapes = [a.shape for a in args if type(a) is Node]
    shape = broadcast_shapes(shapes) if shapes else ()
    node = Node(op, tuple(args), in_dtypes, shape, out_dtype)
    _maybe_flush_growing(node)
    return node


def make_where(c, x, y, out_dtype):
    shapes = [a.shape for a in (c, x, y) if type(a) is Node]
    shape = broadcast_shapes(shapes) if shapes else ()
    node = Node("where", (c, x, y), (BOOL, out_dtype, out_dtype), shape, out_dtype)
    _maybe_flush_growing(node)
    return node


def make_cast(x, dtype):
    return Node("cast", (x,), (x.dtype,), x.shape, dtype)


def make_full(value, shape, dtype):
    return Node("full", (Const(value, dtype),), (dtype,), tuple(shape), dtype)


def make_gather(table: Node, idx_nodes):
    """Lazy ``table[idx0, idx1, ...]`` for integer index expressions covering EVERY axis of a materialised,
    contiguous table: the lookup is folded into whatever fused kernel consumes it (and is simply redone by
    the backward kernels that need the value again, instead of being kept in HBM)."""
    shape = broadcast_shapes([i.shape for i in idx_nodes])
    node = Node("gather", (table,) + tuple(idx_nodes), (table.dtype,) + (I64,) * len(idx_nodes), shape, table.dtype)
    _maybe_flush_growing(node)
    return node


# np.add.at calls are QUEUED per target and launched together: the VJP of a bilinear interpolation
# `(1-rw)*((1-bw)*f[l,t] + bw*f[l,b]) + rw*(...)` (examples/fluidsim/fluidsim.py:45-68) is four np.add.at calls into one
# accumulator (core.py:191-203 -> numpy_vjps.py:946-948) whose index and weight expressions share nearly everything: as
# one kernel they read the cotangent and the velocity fields once and issue four atomics per element.
DEFER_SCATTER = __import__("os").environ.get("B200AG_DEFER_SCATTER", "1") != "0"
SCATTER_GROUP_MAX = 8
_scatter_queue = {}     # (id(buffer), offset, shape) -> [target Mat, shared table leaf, [scatter nodes], {id(buffer) read}]


def scatter_add(target: Mat, idx_nodes, value):
    """``np.add.at(target, (idx0, idx1, ...), value)`` with every axis indexed, as a generated kernel: index and value
    expressions are evaluated in registers and accumulated with FP64 atomics (RED.ADD in L2), so neither the index
    arrays nor the cotangent being scattered are ever written to HBM (VJP of integer-array ``__getitem__``,
    autograd/numpy/numpy_vjps.py:941-954).  The launch is deferred until the target is read (or written by something
    else), so consecutive scatters into one target share a kernel."""
    key = (id(target.buf), target.offset, target.shape)
    rec = _scatter_queue.get(key) if DEFER_SCATTER else None
    table = rec[1] if rec is not None else leaf(target)
    args = (table,) + tuple(idx_nodes) + (value,)
    shapes = [a.shape for a in args[1:] if type(a) is Node]
    shape = broadcast_shapes(shapes) if shapes else ()
    root = Node("scatter", args, (target.dtype,) + (I64,) * len(idx_nodes) + (target.dtype,), shape, target.dtype)
    if target.buf.readers is not None:
        target.buf.readers.discard(root)        # it WRITES the target; as a "reader" it would be flushed by its own group
    target.buf.inexact = True                   # atomic accumulation order is not NumPy's sequential order
    if not DEFER_SCATTER:
        materialize(root, store_root=False)
        return
    if rec is not None and (rec[2][0].shape != shape or len(rec[2]) >= SCATTER_GROUP_MAX):
        _flush_scatter_group(key)
        rec = None
        root.args = (leaf(target),) + root.args[1:]
        if target.buf.readers is not None:
            target.buf.readers.discard(root)
    if rec is None:
        rec = _scatter_queue[key] = [target, root.args[0], [], set()]
    rec[2].append(root)
    _, leaves_ = _collect([a for a in args[1:] if type(a) is Node])
    rec[3].update(id(ln.mat.buf) for ln in leaves_)
    target.buf.pending = True


def _flush_scatter_group(key) -> None:
    target, _table, nodes, _reads = _scatter_queue.pop(key)
    if not any(k[0] == key[0] for k in _scatter_queue):
        target.buf.pending = bool(_gemm_by_buf.get(key[0]))
    root = nodes[0]
    for nd_ in nodes[1:]:                        # a value that depends on every scatter of the group; never stored
        root = Node("add", (root, nd_), (target.dtype, target.dtype), root.shape, target.dtype)
    materialize(root, store_root=False)


def flush_scatters(reading=None) -> None:
    """Launch the queued np.add.at groups - all of them, or only those whose index / value expressions read ``reading``
    (a Buffer about to be written in place)."""
    for key in list(_scatter_queue):
        if key in _scatter_queue and (reading is None or id(reading) in _scatter_queue[key][3]):
            _flush_scatter_group(key)


def flush_deferred() -> None:
    """Everything that was queued instead of launched: contractions first (a scatter may accumulate into a product)."""
    flush_gemms()
    flush_scatters()


def _maybe_flush_growing(node):
    if len(node.leafset) > MAX_LEAVES:
        materialize(node)
    elif node.nid - node.lo > MAX_NODES:
        order, _ = _collect([node])
        if len(order) > MAX_NODES:
            materialize(node)
        else:
            node.lo = node.nid - len(order)  # tighten the bound so we do not re-walk every op


# ---------------------------------------------------------------- flush
class Program:
    """Structural description of one fused kernel; ``key`` identifies the generated code."""

    __slots__ = ("instrs", "leaves", "consts", "outputs", "ndim", "idx64", "vec", "key", "cost", "tables", "mode", "bx",
                 "bdims", "share", "relax", "unroll", "zstr")

    def __init__(self, instrs, leaves, consts, outputs, ndim, idx64, vec, cost, tables=(), mode="flat", bx=256,
                 bdims=None, share=None, relax=(), unroll=(), zstr=None):
        self.instrs = instrs      # tuple of (op, in_dtype chars, out dtype char, arg refs)
        self.leaves = leaves      # tuple of (dtype char, layout class 'C'|'S'|'G', has_shift)
        self.consts = consts      # tuple of C scalar kinds: 'd' | 'q' | '?'
        self.outputs = outputs    # tuple of (instr index, dtype char)
        self.ndim = ndim          # collapsed iteration dims (0 when no 'G' leaf needs coordinates)
        self.idx64 = idx64
        self.vec = vec            # elements per thread per step on the vector path
        self.cost = cost
        self.tables = tables      # gather sources: tuple of (dtype char, number of indexed dims)
        self.mode = mode          # "flat": one flattened index; "row": rows x inner loop (general-layout operands)
        self.bx = bx              # row mode: threads of a CTA along the inner dim (256/bx rows per CTA)
        # flat kernels over large general-layout spaces bake the divisor dims (all but the first) as literals, so the
        # per-element coordinate split compiles to multiply-shift instead of integer division
        self.bdims = bdims
        # flat kernels: general-layout leaves with identical strides (the window slices of a pooled tensor) share one
        # element-offset computation; share[i] = index of the leaf whose offset leaf i reuses
        self.share = share
        # indices of the float add / subtract / divide instructions whose result is already inexact with respect to NumPy
        # (downstream of a Buffer.inexact operand or of a transcendental function) and which the generated code may
        # therefore contract into an FMA / compute with a shared reciprocal (codegen.relax_plan)
        self.relax = relax
        # flat kernels over window layouts: iteration dims of tiny extent (the 2 x 2 of a pooling window) that one THREAD
        # walks itself, fully unrolled, instead of one thread per element: operands that broadcast along those dims
        # (the pooled cotangent, the window maximum, the tie count) are then loaded and computed once per window - the
        # compiler merges the identical sub-expressions of the unrolled bodies - and the coordinate split is done once
        self.unroll = unroll
        # flat kernels over large general-layout spaces: per leaf, the bitmask of iteration dims along which the leaf
        # does NOT move (stride 0: a bias, a pooled cotangent inside its window).  Those terms of the element offset are
        # left out of the generated code instead of multiplied by a zero read from the parameter block - the max-pool
        # VJP of C5 spent more instructions on `x * 0` address terms than on its arithmetic.
        self.zstr = zstr
        self.key = (instrs, leaves, consts, outputs, ndim, idx64, vec, tables, mode, bx, bdims, share, relax, unroll, zstr)

    def launch_config(self, sm_count):
        block = 256
        return block, block * self.vec, sm_count * 16

    def grid(self, n, dims, sm_count, resident=None):
        """CTAs to launch.  Grid-stride kernels are capped at a WHOLE number of waves - ``resident`` CTAs per SM (asked
        from the occupancy calculator for the compiled kernel) times about _GRID_MULT / resident: a cap of 16 CTAs per
        SM is 5.33 waves of a kernel that fits 3 per SM, and the last third-full wave costs as much as a full one
        (C2, scripts/gpu_r2_c2_sweep.sh: 2.285 ms at 16 per SM, 2.231 at 32)."""
        if self.unroll:
            per_thread = 1
            for d in self.unroll:
                per_thread *= dims[d]
            n = n // per_thread
        per_block = 256 * self.vec
        # ALU-bound bodies (hundreds of FP64 instructions per element) run faster from MANY short CTAs than from a
        # few long ones (C2, scripts/gpu_r2_waves2.sh: 2.53 ms from one resident wave, 2.28 at 16 CTAs per SM, 2.22 at
        # 64, 2.21 at 128); bandwidth-bound kernels do not (C4: 210 ms at 16, 212 at 32, 228 at 64)
        mult = _GRID_MULT_HEAVY if self.cost > 400 else _GRID_MULT
        if resident:
            max_grid = sm_count * resident * max(1, (mult + resident // 2) // resident)
        else:
            max_grid = sm_count * mult
        if self.mode == "row":
            inner = dims[-1]
            span = self.bx * self.vec
            gx = (inner + span - 1) // span
            by = 256 // self.bx
            nrows = n // inner
            gy = max(1, min((nrows + by - 1) // by, max_grid // gx))
            return gx * gy
        return max(1, min((n + per_block - 1) // per_block, max_grid))


# small exact constants are baked into the source (lets the compiler fold x*1, x**2 ...)
_BAKED = {0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, 4.0, 0.25}


def _collect(roots):
    """Topological order of the un-materialised nodes behind ``roots`` plus the distinct leaves."""
    order, seen, leaves = [], set(), []
    stack = [(r, False) for r in roots]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if node.nid in seen:
            continue
        seen.add(node.nid)
        if node.mat is not None:
            leaves.append(node)
            continue
        stack.append((node, True))
        for a in node.args:
            if type(a) is Node and a.nid not in seen:
                stack.append((a, False))
    return order, leaves


def _aligned_strides(mat, root_shape):
    """Element strides (and shifts) of ``mat`` broadcast against ``root_shape`` (right-aligned)."""
    nd = len(root_shape)
    off = nd - len(mat.shape)
    strides = [0] * nd
    shifts = [0] * nd
    for i, (s, st) in enumerate(zip(mat.shape, mat.strides)):
        if s != 1:
            strides[off + i] = st
            if mat.shifts is not None:
                shifts[off + i] = mat.shifts[i]
    return strides, shifts


def _collapse_dims(root_shape, leaf_layouts):
    """Drop unit dims, then merge adjacent iteration dims that every general-layout leaf walks contiguously."""
    keep = [i for i, d in enumerate(root_shape) if d != 1]
    dims = [root_shape[i] for i in keep]
    lay = [([s[i] for i in keep], [sh[i] for i in keep]) for s, sh in leaf_layouts]
    i = len(dims) - 1
    while i > 0:
        if all(not sh[i] and not sh[i - 1] and s[i - 1] == s[i] * dims[i] for s, sh in lay):
            for s, sh in lay:
                s[i - 1] = s[i]
                del s[i], sh[i]
            dims[i - 1] *= dims[i]
            del dims[i]
        i -= 1
    if not dims:
        dims = [1]
        lay = [([0], [0]) for _ in lay]
    return dims, lay


class _AliasGroup:
    """Arrays that NumPy would hand out as VIEWS of one buffer: an array and its (chained) reshapes.  While they are pending
    they are independent expressions - that is what lets a reshape fuse into its consumers and what keeps the tape from
    holding a buffer for each of them.  The first member that is given a payload provides the buffer of the group; any
    other member is turned into a strided view of it the moment it would be computed (materialize) or the buffer is about
    to be written in place (flush_readers), so a write through one handle is seen through all of them."""

    __slots__ = ("members", "__weakref__")

    def __init__(self):
        self.members = []        # weakrefs to Nodes

    def payload(self):
        for ref in self.members:
            nd_ = ref()
            if nd_ is not None and nd_.mat is not None:
                m = nd_.mat
                if m.shifts is None and m.is_contiguous:
                    return m
        return None


def link_alias(view: Node, src: Node) -> None:
    """Record that ``view`` is a symbolic reshape of the pending ``src`` (see _AliasGroup)."""
    group = src.alias
    if group is None:
        group = src.alias = _AliasGroup()
        group.members.append(weakref.ref(src))
    view.alias = group
    if len(group.members) >= 32:             # an array reshaped again and again in a loop: forget the handles that died
        group.members = [r for r in group.members if r() is not None]
    group.members.append(weakref.ref(view))


def _become_leaf(nd_, m):
    nd_.mat = m
    nd_.args = ()
    nd_.cost = 0
    nd_.lo = nd_.nid
    nd_.leafset = frozenset((nd_.nid,))


def _adopt_group_payload(nd_) -> bool:
    """If an array of ``nd_``'s alias group already has a payload, ``nd_`` (pending) becomes a view of it."""
    m = nd_.alias.payload()
    if m is None or prod(m.shape) != prod(nd_.shape):
        return False
    _become_leaf(nd_, Mat(m.buf, nd_.shape, c_strides(nd_.shape), m.offset, m.dtype))
    payload_set(nd_)
    return True


def payload_set(nd_) -> None:
    """``nd_`` has just been given a payload.  (1) Expressions that were recorded with it as a PENDING operand and are
    themselves still pending now read its buffer: they are registered as readers of that buffer, so that a later
    in-place write to it evaluates them first (flush_readers; NumPy would have computed them when they were written
    down).  (2) Its alias group is noted on the buffer: before an in-place write, the group's pending members become
    views of it."""
    cons, nd_.consumers = nd_.consumers, None
    buf = nd_.mat.buf
    if cons:
        for ref in cons:
            c = ref()
            if c is not None and c.mat is None and c.args:
                if buf.readers is None:
                    buf.readers = weakref.WeakSet()
                buf.readers.add(c)
    if nd_.alias is not None:
        if buf.aliases is None:
            buf.aliases = [nd_.alias]
        elif nd_.alias not in buf.aliases:
            buf.aliases.append(nd_.alias)


def settle_aliases(buf) -> None:
    """Before an in-place write to ``buf``: the pending members of the alias groups living in it become views of it."""
    groups, buf.aliases = buf.aliases, None
    for group in groups or ():
        for ref in list(group.members):
            nd_ = ref()
            if nd_ is not None and nd_.mat is None:
                _adopt_group_payload(nd_)
    if groups:
        buf.aliases = groups           # later reshapes of the same arrays join the same groups


def _settle_reshape(nd_):
    """Turn a pending ``reshape`` node into a leaf: compute what is behind it (contiguous), view it under the new shape."""
    arg = nd_.args[0]
    m = arg.mat if arg.mat is not None else materialize(arg)
    if not m.is_contiguous:
        m = materialize(Node("positive", (leaf(m),), (m.dtype,), m.shape, m.dtype))
    _become_leaf(nd_, Mat(m.buf, nd_.shape, c_strides(nd_.shape), m.offset, m.dtype))
    payload_set(nd_)


def _plan_reshapes(root):
    """A pending expression may contain ``reshape`` nodes (array.reshape of a pending value whose operands cannot be
    re-viewed one by one: ``(cell_xs - vx).ravel()`` with ``cell_xs`` a broadcast meshgrid row, examples/fluidsim.py:37-41).
    Row-major order is the same on both sides of such a node, so ONE kernel can evaluate all of it provided the operands
    that need coordinates (broadcast / strided / rolled ones) all live under the same shape: that shape becomes the
    iteration space, and every full-size contiguous operand - whatever shape it is recorded under - is just the flat
    index.  Returns (iteration shape, {leaf nid: Mat seen in that space}) or None (then the reshapes are computed first)."""
    root_shape = root.shape
    level, leaf_levels, leaf_node = {}, {}, {}
    stack = [(root, root_shape)]
    while stack:
        nd_, lv = stack.pop()
        if nd_.mat is not None:
            leaf_levels.setdefault(nd_.nid, set()).add(lv)
            leaf_node[nd_.nid] = nd_
            continue
        prev = level.get(nd_.nid)
        if prev is not None:
            if prev != lv:
                return None
            continue
        level[nd_.nid] = lv
        if nd_.op == "reshape":
            if nd_.shape != lv:
                return None                      # a reshaped value broadcast into a larger space
            stack.append((nd_.args[0], nd_.args[0].shape))
        else:
            for k, a in enumerate(nd_.args):
                if type(a) is Node and not (nd_.op in ("gather", "scatter") and k == 0):
                    stack.append((a, lv))
    g_levels, views = set(), {}
    for nid, lvs in leaf_levels.items():
        m = leaf_node[nid].mat
        if prod(m.shape) == 1:
            continue
        if m.shifts is None and m.is_contiguous and all(m.shape == lv for lv in lvs):
            views[nid] = None                    # flat order: re-viewed under the iteration shape below
            continue
        if len(lvs) != 1:
            return None
        g_levels.add(next(iter(lvs)))
    if len(g_levels) > 1:
        return None
    iter_shape = next(iter(g_levels)) if g_levels else root_shape
    if prod(iter_shape) != prod(root_shape):
        return None
    out = {}
    for nid in views:
        m = leaf_node[nid].mat
        if m.shape != iter_shape:
            out[nid] = Mat(m.buf, iter_shape, c_strides(iter_shape), m.offset, m.dtype)
    return iter_shape, out


_taint_memo = {}       # (instrs, leaf taint, table taint) -> (codegen.taint, codegen.relax_plan)


def materialize(root: Node, store_root: bool = True) -> Mat:
    """Compute ``root`` (and any still-visible expensive interior node) with one fused kernel."""
    if root.mat is not None:
        return root.mat
    order, leaf_nodes = _collect([root])
    if any(ln.mat.buf.pending for ln in leaf_nodes):
        # an operand is the result of a queued product / the target of queued scatters: launch those NOW, before the
        # kernel is laid out - launching them computes expressions that may share nodes with this one (they become
        # leaves), which must not happen between collecting the graph and numbering its operands (Mat.ptr would do it)
        flush_deferred()
        if root.mat is not None:
            return root.mat
        if any(nd_.mat is not None for nd_ in order):
            order, leaf_nodes = _collect([root])
    adopted = False
    for nd_ in order:                            # a reshape of an array that has been computed meanwhile: a view of it
        if nd_.alias is not None and nd_.mat is None and _adopt_group_payload(nd_):
            adopted = True
    if adopted:
        if root.mat is not None:
            return root.mat
        order, leaf_nodes = _collect([root])
    root_shape = root.shape
    n = prod(root_shape)
    iter_shape, leaf_view = root_shape, {}
    if any(nd_.op == "reshape" for nd_ in order):
        settled = False
        for nd_ in order:                        # what is behind a reshape has been computed meanwhile: plain views
            if nd_.op == "reshape" and nd_.mat is None and nd_.args[0].mat is not None:
                _settle_reshape(nd_)
                settled = True
        if root.mat is not None:
            return root.mat
        if settled:
            order, leaf_nodes = _collect([root])
        if any(nd_.op == "reshape" for nd_ in order):
            plan = _plan_reshapes(root) if n >= 2 else None
            if plan is None:
                for nd_ in order:
                    if nd_.op == "reshape" and nd_.mat is None:
                        _settle_reshape(nd_)
                if root.mat is not None:
                    return root.mat
                order, leaf_nodes = _collect([root])
            else:
                iter_shape, leaf_view = plan

    # A pending sub-expression that lives on a SMALLER iteration space and is broadcast into a larger one would be
    # recomputed once per element of the larger space (the tanh VJP at pooled resolution feeding the 4x larger
    # max-pool VJP of examples/convnet.py).  When the repeated work outweighs a launch, compute it once at its own size.
    if n >= 65536:
        small = []
        for nd_ in order:
            if nd_.mat is not None:
                continue
            pn = prod(nd_.shape)
            for a in nd_.args:
                if type(a) is Node and a.mat is None and a.op not in ("full", "scatter"):
                    an = prod(a.shape)
                    if an and an * 2 <= pn and a.cost * (pn // an - 1) >= EXPAND_MIN_COST and float(pn - an) * a.cost >= EXPAND_WORK:
                        small.append(a)
        if small:
            for a in small:
                materialize(a)
            order, leaf_nodes = _collect([root])

    # interior nodes that Python still holds and that are costly to recompute become extra outputs
    # (b) or that have already been recomputed REUSE_LIMIT times (a loop keeps coming back to them: the
    #     Jacobi right-hand side of fluidsim.py:21-30 is read by all ten sweeps)
    # (c) or that are being recomputed a second time although reading them back would move FEWER bytes than their
    #     inputs do: the Jacobi right-hand side `div` above is two shifted reads each of vx and vy (16 B per cell) where
    #     the stored value is 8 B - kept from the second sweep on, the other eight read one array instead of two.
    #     Values whose inputs are no wider than themselves (the advection weights: one velocity field + a broadcast
    #     arange) stay recomputed, so nothing extra is retained by the tape.
    outs = [root] if store_root else []
    leaf_by_nid = None
    has_tables = any(nd_.op in ("gather", "scatter") for nd_ in order)   # (c) is for stencil loops, not for lookups
    for nd_ in order:
        if nd_ is not root and nd_.nhandles > 0 and nd_.shape == root_shape and nd_.op not in ("full", "scatter"):
            keep = nd_.cost >= KEEP_COST or (nd_.uses >= REUSE_LIMIT and len(nd_.leafset) > 1)
            if not keep and KEEP_WIDE and not has_tables and nd_.uses >= KEEP_WIDE_USES and len(nd_.leafset) > 1 and n >= 65536:
                if leaf_by_nid is None:
                    leaf_by_nid = {ln.nid: ln for ln in leaf_nodes}
                seen_bufs, width = set(), 0.0
                for nid in nd_.leafset:
                    ln = leaf_by_nid.get(nid)
                    if ln is None or id(ln.mat.buf) in seen_bufs:
                        continue
                    seen_bufs.add(id(ln.mat.buf))
                    width += ln.mat.dtype.itemsize * min(1.0, prod(ln.mat.shape) / float(n))
                keep = width > nd_.dtype.itemsize * 1.25
            if keep:
                outs.append(nd_)
            else:
                nd_.uses += 1

    # ---- gather sources ("tables") are addressed by computed indices, not by the iteration index
    table_index, table_descs, table_rt, elementwise_leaf, table_taint = {}, [], [], set(), []
    for nd_ in order:
        for k, a in enumerate(nd_.args):
            if type(a) is Node and a.mat is not None:
                if nd_.op in ("gather", "scatter") and k == 0:
                    if a.nid not in table_index:
                        table_index[a.nid] = len(table_descs)
                        table_descs.append((a.mat.dtype.char, len(a.mat.shape)))
                        table_rt.append((a.mat.ptr, a.mat.shape))
                        table_taint.append(bool(a.mat.buf.inexact))
                else:
                    elementwise_leaf.add(a.nid)

    # ---- leaves: classify layout relative to the iteration space
    leaf_index, leaf_descs, leaf_ptrs, g_layouts, g_slots, leaf_taint = {}, [], [], [], [], []
    dedup = {}
    for ln in leaf_nodes:
        if ln.nid not in elementwise_leaf:
            continue
        m = leaf_view.get(ln.nid) or ln.mat
        strides, shifts = _aligned_strides(m, iter_shape)
        sig = (m.buf.ptr, m.offset, m.dtype.char, tuple(strides), tuple(shifts))
        if sig in dedup:
            leaf_index[ln.nid] = dedup[sig]
            continue
        has_shift = any(shifts)
        if all(st == 0 for st in strides) or n <= 1:
            cls = "S"
        elif not has_shift and m.shape == iter_shape and m.is_contiguous:
            cls = "C"
        else:
            cls = "G"
        idx = len(leaf_descs)
        dedup[sig] = idx
        leaf_index[ln.nid] = idx
        leaf_descs.append((m.dtype.char, cls, has_shift))
        leaf_ptrs.append([m.ptr, None, None])
        leaf_taint.append(bool(m.buf.inexact) and m.dtype.kind == "f")
        if cls == "G":
            g_layouts.append((strides, shifts))
            g_slots.append(idx)

    if g_layouts:
        dims, lay = _collapse_dims(iter_shape, g_layouts)
        for slot, (s, sh) in zip(g_slots, lay):
            leaf_ptrs[slot][1] = s
            leaf_ptrs[slot][2] = sh
            # specialise the common inner-dim patterns: "u" = unit inner stride, "R" = additionally no inner
            # shift and every row start 16-byte aligned, so the row can be read with vector loads
            ch, _cls, has_shift = leaf_descs[slot]
            item = np.dtype(ch).itemsize
            vecw = max(1, 16 // item)
            if s[-1] == 1:
                aligned = leaf_ptrs[slot][0] % 16 == 0 and all(st % vecw == 0 for st in s[:-1])
                if not sh[-1] and aligned:
                    leaf_descs[slot] = (ch, "G", has_shift, "R")
                else:
                    leaf_descs[slot] = (ch, "G", has_shift, "u")
            else:
                leaf_descs[slot] = (ch, "G", has_shift, "")
        ndim = len(dims)
    else:
        dims, ndim = [], 0
    leaf_descs = [d if len(d) == 4 else d + ("",) for d in leaf_descs]

    # ---- instructions
    const_kinds, const_values, const_dedup = [], [], {}
    instr_index, instrs = {}, []
    for nd_ in order:
        refs = []
        for k, (a, in_dt) in enumerate(zip(nd_.args, nd_.in_dtypes)):
            if type(a) is Node:
                if nd_.op in ("gather", "scatter") and k == 0:
                    refs.append(("t", table_index[a.nid]))
                elif a.mat is not None:
                    refs.append(("l", leaf_index[a.nid]))
                else:
                    refs.append(("n", instr_index[a.nid]))
            else:
                v = a.value
                kind = "d" if in_dt.kind == "f" else ("?" if in_dt.kind == "b" else "q")
                if kind == "d":
                    v = float(np.asarray(v).astype(in_dt))  # NumPy's cast of the scalar to the loop dtype
                elif kind == "q":
                    v = int(v)
                else:
                    v = bool(v)
                if (kind == "d" and v in _BAKED and not (v == 0.0 and math.copysign(1.0, v) < 0)) or (
                    kind == "q" and (-4 <= v <= 4 or (k == 1 and nd_.op in ("remainder", "floor_divide") and abs(v) < 1 << 31))
                ) or kind == "?":  # a baked integer divisor lets the compiler strength-reduce `%` and `/`
                    refs.append(("k", v if kind != "d" else float(v)))
                else:
                    ck = (kind, v)
                    ci = const_dedup.get(ck)
                    if ci is None:
                        ci = len(const_kinds)
                        const_dedup[ck] = ci
                        const_kinds.append(kind)
                        const_values.append(v)
                    refs.append(("c", ci))
        instr_index[nd_.nid] = len(instrs)
        instrs.append(("positive" if nd_.op == "reshape" else nd_.op, tuple(d.char for d in nd_.in_dtypes), nd_.dtype.char,
                       tuple(refs)))

    outputs = tuple((instr_index[o.nid], o.dtype.char) for o in outs)

    # ---- vector width: 16-byte accesses on every contiguous leaf and output when alignment allows
    max_item = max([o.dtype.itemsize for o in outs] + [np.dtype(d[0]).itemsize for d in leaf_descs if d[1] == "C" or d[3] == "R"],
                   default=16)   # a scatter of broadcast scalars has neither outputs nor contiguous leaves
    if table_descs:
        max_item = 16  # gathers are latency-bound scalar loads: no vector path
    vec = max(1, 16 // max_item)
    total_cost = sum(_COST.get(nd_.op, _DEFAULT_COST) for nd_ in order)
    if n < 4096:
        vec = 1
    elif total_cost > 400:
        vec = min(vec, _HEAVY_VEC)  # heavy bodies are ALU-bound: keep the code size down
    if vec > 1:
        for (ch, cls, _, _f), lp in zip(leaf_descs, leaf_ptrs):
            if cls == "C" and lp[0] % (np.dtype(ch).itemsize * vec):
                vec = 1
                break
    max_extent = n
    for lp in leaf_ptrs:
        if lp[1] is not None:
            ext = sum(abs(s) * (d - 1) for s, d in zip(lp[1], dims)) + 1
            max_extent = max(max_extent, ext)
    idx64 = max_extent >= (1 << 31) - (1 << 24)
    mode = "row" if g_layouts and dims[-1] >= 128 else "flat"
    if mode == "row" and vec > 1 and dims[-1] % vec:
        vec = 1
    # short rows: several rows per CTA (64 threads along the row) instead of one mostly idle 256-thread row slice
    bx = 256 if (mode != "row" or dims[-1] >= 2048 * vec) else 64
    if mode == "row" and _ROW_TUNE:            # development knobs: "bx,vec" (0 keeps the default)
        tb, tv = _ROW_TUNE
        if tb:
            bx = tb
        if tv and dims[-1] % tv == 0 and tv <= vec:
            vec = tv

    bdims = tuple(int(d) for d in dims[1:]) if (mode == "flat" and g_layouts and n >= (1 << 20)) else None
    zstr = None
    if bdims is not None:
        zstr = [0] * len(leaf_descs)
        for slot in g_slots:
            zstr[slot] = sum(1 << d for d, st in enumerate(leaf_ptrs[slot][1]) if st == 0)
        zstr = tuple(zstr) if any(zstr) else None
    share = None
    if mode == "flat" and len(g_slots) > 1:
        rep, first = list(range(len(leaf_descs))), {}
        for slot in g_slots:
            st_, sh_ = leaf_ptrs[slot][1], leaf_ptrs[slot][2]
            if any(sh_):
                continue
            rep[slot] = first.setdefault(tuple(st_), slot)
        if any(r != i for i, r in enumerate(rep)):
            share = tuple(rep)
    from . import codegen

    instrs = tuple(instrs)
    memo_key = (instrs, tuple(leaf_taint), tuple(table_taint))
    hit = _taint_memo.get(memo_key)
    if hit is None:                              # (a training loop records the same bodies over and over)
        inexact = codegen.taint(instrs, leaf_taint, table_taint)
        relax = codegen.relax_plan(instrs, inexact) if RELAXED else ()
        if len(_taint_memo) >= 4096:
            _taint_memo.clear()
        _taint_memo[memo_key] = (inexact, relax)
    else:
        inexact, relax = hit
    unroll = ()
    if UNROLL_WINDOWS and mode == "flat" and g_layouts and not table_descs and n >= (1 << 20) and len(dims) >= 2:
        # dims of extent <= 4 - neither the outermost nor the INNERMOST one, whose consecutive elements belong to consecutive
        # threads - along which a real array operand (one that varies along at least two other dims: the pooled cotangent,
        # the window maximum, not a bias) broadcasts: that operand's loads and everything computed from it alone is what
        # the unrolling shares.  Measured on the B200 (scripts/gpu_r2_unroll.sh): the 226 M-element max-pool VJP of C5
        # 1.59 -> 1.28 ms with the window rows unrolled; unrolling the innermost dim instead breaks up the coalesced /
        # vector accesses and loses (pooled-size kernels 0.14 -> 0.26 ms), so it is never picked.
        picked, per_thread = [], 1
        shifted = any(any(leaf_ptrs[slot][2]) for slot in g_slots)
        for d in range(len(dims) - 2, 0, -1):
            if shifted:
                break
            if dims[d] <= 4 and per_thread * dims[d] <= 8 and any(
                    leaf_ptrs[slot][1][d] == 0 and sum(1 for st in leaf_ptrs[slot][1] if st) >= 2 for slot in g_slots):
                picked.append(d)
                per_thread *= dims[d]
        # ... except when the innermost dim IS a window dim of extent 2 over 8-byte elements and every operand either
        # does not move along it or has unit stride with 16-byte aligned rows: then one thread reads / writes the pair
        # as a 16-byte vector (consecutive threads still touch consecutive 16-byte pieces) and the whole 2 x 2 window
        # is one thread's work - half the index arithmetic and window-level arithmetic per element again
        # (scripts/gpu_r2_window.sh)
        last = len(dims) - 1
        if (UNROLL_INNER and not shifted and dims[last] == 2 and per_thread * 2 <= 8
                and all(o.dtype.itemsize == 8 for o in outs)
                and all(np.dtype(ch).itemsize == 8 for ch, cls, _s, _f in leaf_descs if cls != "S")
                and all(lp[0] % 16 == 0 for (ch, cls, _s, _f), lp in zip(leaf_descs, leaf_ptrs) if cls == "C")
                and all(leaf_ptrs[slot][1][last] == 0 or leaf_descs[slot][3] == "R" for slot in g_slots)
                and (picked or any(leaf_ptrs[slot][1][last] == 0 and sum(1 for st in leaf_ptrs[slot][1] if st) >= 2
                                   for slot in g_slots))):
            picked.append(last)
            per_thread *= 2
        if per_thread >= 2 and total_cost >= 8:
            unroll = tuple(sorted(picked))
            vec = 1
    program = Program(instrs, tuple(leaf_descs), tuple(const_kinds), outputs, ndim, idx64, vec, total_cost,
                      tuple(table_descs), mode, bx, bdims, share, relax, unroll, zstr)
    out_mats = [Mat.empty(root_shape, o.dtype, inexact=inexact[j] and o.dtype.kind == "f") for o, (j, _) in zip(outs, outputs)]
    backend.get().run_program(program, (leaf_ptrs, dims, table_rt), const_values, [m.ptr for m in out_mats], n)
    for o, m in zip(outs, out_mats):
        o.mat = m
        if o.consumers or o.alias is not None:
            payload_set(o)
    if not store_root:
        root.args = ()
    # drop the graph behind everything that is now materialised or no longer needed
    for nd_ in order:
        if nd_.mat is not None:
            nd_.args = ()
            nd_.cost = 0
            nd_.lo = nd_.nid
            nd_.leafset = frozenset((nd_.nid,))
    return root.mat
