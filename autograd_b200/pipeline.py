"""Host-buffer streaming of elementwise workloads.

``map_elementwise(fn, x_host, out=...)`` evaluates ``fn`` (typically an ``elementwise_grad`` tower) over a host
array that is processed in chunks: chunk i+1 is uploaded while chunk i is computed and chunk i-1 is downloaded
(two copy streams + the compute stream, PCIe in both directions at once).  Only valid when every output element
depends on the matching input element alone — exactly the property ``elementwise_grad`` assumes
(autograd/differential_operators.py:40-50) and the same one the multi-GPU sharding of this workload relies on.
"""

from __future__ import annotations

import numpy as np

from . import backend, lazy
from .array import DeviceArray, pinned_empty
from .lazy import Mat


_STATE = {}


def map_elementwise(fn, x_host: np.ndarray, out: np.ndarray | None = None, chunk: int = 1 << 23,
                    replay: bool = True) -> np.ndarray:
    """``replay=True`` records ``fn`` once per buffer slot as a CUDA graph (the trace must not depend on values),
    so a chunk costs one graph launch instead of a re-trace through autograd and small chunks stay efficient."""
    from .graph import CapturedFunction

    be = backend.get()
    x_flat = np.asarray(x_host, order="C").reshape(-1)
    n = x_flat.size
    chunk = min(chunk, n)
    nchunks = (n + chunk - 1) // chunk
    up, down = be.stream_create(), be.stream_create()
    # staging buffers and recorded graphs are kept per (function, chunk, dtype): a second call replays immediately
    key = (id(fn), chunk, x_flat.dtype.char, replay)
    state = _STATE.get(key)
    if state is None or state[0] is not fn:
        state = (fn, [Mat.empty((chunk,), x_flat.dtype) for _ in range(2)], {},
                 [CapturedFunction(fn, copy_inputs=False) if replay else fn for _ in range(2)])
        if len(_STATE) >= 8:
            _STATE.pop(next(iter(_STATE)))
        _STATE[key] = state
    _, in_bufs, in_arrays, fast = state   # in_arrays: (slot, m) -> DeviceArray over the slot's first m elements
    ev_in = [be.event_create() for _ in range(2)]      # upload of slot finished
    ev_free = [be.event_create() for _ in range(2)]    # compute no longer reads slot
    ev_done = [be.event_create() for _ in range(2)]    # result of slot computed
    ev_out = [be.event_create() for _ in range(2)]     # download of slot finished
    results = [None, None]
    busy = [False, False]
    out_flat = None
    try:
        for i in range(nchunks):
            s = i & 1
            lo, hi = i * chunk, min(n, (i + 1) * chunk)
            m = hi - lo
            if i >= 2:
                be.stream_wait_event(up, ev_free[s])            # slot's previous contents have been consumed
            in_bufs[s].buf.version += 1                         # rewritten in place: memoised copies are stale
            be.h2d_on(in_bufs[s].ptr, x_flat[lo:hi], up)
            be.event_record_on(ev_in[s], up)
            be.stream_wait_event(None, ev_in[s])
            xin = in_arrays.get((s, m))
            if xin is None:
                xin = DeviceArray(lazy.leaf(Mat(in_bufs[s].buf, (m,), (1,), 0, x_flat.dtype)))
                in_arrays[(s, m)] = xin
            if busy[s]:
                if replay:
                    be.stream_wait_event(None, ev_out[s])       # the replay overwrites the slot's output buffer
                else:
                    be.event_synchronize(ev_out[s])             # its download is done: the buffer may be recycled
            y = fast[s](xin)
            ym = y._cmat()
            be.event_record_on(ev_free[s], None)
            be.event_record_on(ev_done[s], None)
            if out_flat is None:
                if out is None:
                    out = pinned_empty(x_host.shape, ym.dtype)
                out_flat = out.reshape(-1)
                if out_flat.size != n or not out.flags.c_contiguous:
                    raise ValueError("map_elementwise: out must be a C-contiguous array with as many elements as x")
            results[s] = ym
            busy[s] = True
            be.stream_wait_event(down, ev_done[s])
            be.d2h_on(out_flat[lo:hi], ym.ptr, down)
            be.event_record_on(ev_out[s], down)
            del y
        be.stream_synchronize(down)
        be.stream_synchronize(up)
        be.synchronize()
    finally:
        results[:] = [None, None]
        for evs in (ev_in, ev_free, ev_done, ev_out):
            for e in evs:
                be.event_destroy(e)
        be.stream_destroy(up)
        be.stream_destroy(down)
    return out
