"""Tape replay as a CUDA graph.

autograd re-traces the user function on every ``grad(f)(x)`` call: a Python loop over the tape
(``backward_pass``, autograd/core.py:26-33; 64 / 10,252 / 35,003 nodes for BASELINE configs C1 / C3 / C4).
Once the kernels are on the GPU that loop *is* the run time.  ``capture(fn)`` runs the traced call once
under CUDA stream capture — every kernel the tape launches, fused elementwise programs, GEMMs, reductions,
gathers and scatter-adds alike, becomes a graph node — and afterwards replays the recorded forward+backward
tape with one ``cudaGraphLaunch``: no Python per op, no per-kernel launch latency.

Validity is the usual one for traced programs: the recorded tape must not depend on array *values*
(a device→host read during capture raises) and is keyed on the argument structure, shapes and dtypes;
a call with a new signature is captured afresh.  Python scalars are constants of the recording and are
part of the key.
"""

from __future__ import annotations

import weakref

import numpy as np

from . import backend, lazy
from .array import DeviceArray, asarray


def _flatten(tree, leaves, spec):
    if isinstance(tree, DeviceArray) or isinstance(tree, np.ndarray):
        spec.append(("a", tuple(tree.shape), np.dtype(tree.dtype).char))
        leaves.append(tree)
    elif isinstance(tree, (list, tuple)):
        spec.append(("seq", type(tree).__name__, len(tree)))
        for t in tree:
            _flatten(t, leaves, spec)
    elif isinstance(tree, dict):
        keys = sorted(tree)
        spec.append(("dict", tuple(keys)))
        for k in keys:
            _flatten(tree[k], leaves, spec)
    else:
        spec.append(("const", type(tree).__name__, tree if isinstance(tree, (int, float, bool, str, type(None))) else id(tree)))


def _rebuild(tree, it):
    if isinstance(tree, DeviceArray) or isinstance(tree, np.ndarray):
        return next(it)
    if isinstance(tree, (list, tuple)):
        return type(tree)(_rebuild(t, it) for t in tree)
    if isinstance(tree, dict):
        return {k: _rebuild(tree[k], it) for k in sorted(tree)}
    return tree


def _out_leaves(tree, acc):
    if isinstance(tree, DeviceArray):
        acc.append(tree)
    elif isinstance(tree, (list, tuple)):
        for t in tree:
            _out_leaves(t, acc)
    elif isinstance(tree, dict):
        for k in sorted(tree):
            _out_leaves(tree[k], acc)


class _Recording:
    __slots__ = ("graph", "static_inputs", "outputs", "nodes", "seen", "out_bufs", "__weakref__")

    def __init__(self, graph, static_inputs, outputs, nodes):
        self.graph, self.static_inputs, self.outputs, self.nodes = graph, static_inputs, outputs, nodes
        self.seen = [None] * len(static_inputs)   # per input: what was copied in last (buffer identity + version)
        acc = []
        _out_leaves(outputs, acc)
        self.out_bufs = [a._node.mat.buf for a in acc if a._node.mat is not None]   # rewritten by every replay

    def __del__(self):
        try:
            backend.get().graph_destroy(self.graph)
        except Exception:
            pass


class CapturedFunction:
    """``fn`` replayed as a CUDA graph.  Returned arrays are owned by the recording and are overwritten by
    the next call with the same signature (copy them if they must survive)."""

    def __init__(self, fn, copy_inputs: bool = True):
        self.fn = fn
        self.copy_inputs = copy_inputs   # False: bake the caller's own input buffers into the graph (pipelines)
        self._recordings = {}
        self.max_recordings = 16

    def __call__(self, *args):
        be = backend.get()
        leaves, spec = [], []
        _flatten(args, leaves, spec)
        key = tuple(spec)
        rec = self._recordings.pop(key, None)
        if rec is None:
            rec = self._record(be, args, leaves)
            # least-recently-used bound: Python scalars are constants of a recording and part of its key, so a call
            # site that varies one (a step counter) would otherwise keep every graph and its arena alive for ever
            while len(self._recordings) >= self.max_recordings:
                self._recordings.pop(next(iter(self._recordings)))
        else:
            for i, (dst, src) in enumerate(zip(rec.static_inputs, leaves)):
                self._refresh(be, dst, src, rec.seen, i)
        self._recordings[key] = rec      # most recently used last
        # The replay rewrites the recording's output buffers in place: expressions still pending on the previous
        # results are computed first, and memoised copies (array._CONCAT_MEMO keys on buffer versions) become stale.
        for buf in rec.out_bufs:
            lazy.flush_readers(buf)
            buf.version += 1
        be.graph_launch(rec.graph)
        return rec.outputs

    @staticmethod
    def _refresh(be, dst: DeviceArray, src, seen, i):
        dm = dst._node.mat
        if isinstance(src, DeviceArray):
            sm = src._cmat()
            if sm.ptr != dm.ptr:
                # the same, unmodified source buffer as last time (same Buffer OBJECT, not merely the same address, and
                # no in-place write since): the static copy is still current
                tag = seen[i]
                if tag is not None and tag[0]() is sm.buf and tag[1] == sm.buf.version and tag[2] == sm.offset:
                    return
                lazy.flush_readers(dm.buf)
                dm.buf.version += 1
                be.d2d(dm.ptr, sm.ptr, src.nbytes)
                seen[i] = (weakref.ref(sm.buf), sm.buf.version, sm.offset)
        else:
            lazy.flush_readers(dm.buf)
            dm.buf.version += 1
            be.h2d(dm.ptr, np.asarray(src, dtype=dm.dtype, order="C"))

    def _record(self, be, args, leaves):
        # 1. eager warm-up: compiles/loads every kernel the tape needs, sizes the caches
        out = self.fn(*args)
        acc = []
        _out_leaves(out, acc)
        for a in acc:
            a.materialize()
        be.synchronize()
        del out, acc
        be.empty_cache()  # the recording allocates from its own arena: give the warm-up's blocks back first
        # 2. private copies of the inputs: their addresses are what the graph bakes in
        static = []
        for leaf in leaves:
            d = asarray(leaf)
            static.append(d.copy() if isinstance(leaf, DeviceArray) and self.copy_inputs else d)
        be.synchronize()
        static_args = _rebuild(args, iter(static))
        # 3. capture the same call.  Memoised results (array._CONCAT_MEMO) must not cross the capture boundary in either
        #    direction: a buffer computed outside would be baked in stale, one recorded inside only exists after a replay.
        from . import array as _array

        _array._CONCAT_MEMO.clear()
        be.graph_begin()
        try:
            out = self.fn(*static_args)
            acc = []
            _out_leaves(out, acc)
            for a in acc:
                a.materialize()
        except BaseException:
            _array._CONCAT_MEMO.clear()
            try:
                g = be.graph_end()
                be.graph_destroy(g)
            except Exception:
                pass
            raise
        graph = be.graph_end()
        _array._CONCAT_MEMO.clear()
        return _Recording(graph, static, out, be.graph_num_nodes(graph))

    def num_nodes(self):
        return {k: r.nodes for k, r in self._recordings.items()}


def capture(fn, copy_inputs: bool = True):
    """Wrap ``fn`` (typically ``grad(f)`` / ``value_and_grad(f)``) so that repeated calls replay a CUDA graph."""
    return CapturedFunction(fn, copy_inputs)
