"""Device backend: the only place that talks to libb200ag.so.

Everything above (``DeviceArray``, the lazy expression graph, the NumPy
dispatch tables, the autograd registrations) calls ``backend.get()`` and works
on raw device pointers (Python ints) plus shape/stride metadata kept on the
host.  There is one production implementation, :class:`CudaBackend`.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

DTYPE_CODE = {
    np.dtype(np.float64): 0,
    np.dtype(np.float32): 1,
    np.dtype(np.int64): 2,
    np.dtype(np.int32): 3,
    np.dtype(np.bool_): 4,
}
REDUCE_CODE = {"sum": 0, "max": 1, "min": 2, "prod": 3}


def dtype_code(dt) -> int:
    try:
        return DTYPE_CODE[np.dtype(dt)]
    except KeyError:
        raise TypeError(
            f"autograd_b200: dtype {np.dtype(dt)} is not supported on the device "
            "(supported: float64, float32, int64, int32, bool)"
        ) from None


def _i64_array(seq):
    n = len(seq)
    return (C.c_int64 * max(n, 1))(*seq)


class CudaBackend:
    """libb200ag.so on one B200 (one process per GPU)."""

    name = "cuda"

    def __init__(self, device: int | None = None):
        from . import _lib

        self._lib = _lib.lib
        self._check = _lib.check
        if device is None:
            device = int(os.environ.get("B200AG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device = device
        self._check(self._lib.b200ag_init(device))
        self._kernels = {}  # Program.key -> compiled kernel record
        self.capturing = False  # True between graph_begin and graph_end (results computed now are baked into a graph)
        self.profile = None  # set to a list to collect (program, n, start_event, stop_event) per fused launch
        sms, major, minor, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self._lib.b200ag_device_info(C.byref(sms), C.byref(major), C.byref(minor), C.byref(mem)))
        self.sm_count = sms.value
        self.total_mem = mem.value
        flag = C.c_void_p()
        self._check(self._lib.b200ag_index_flag(C.byref(flag)))
        self._index_flag = flag.value

    # ------------------------------------------------------------ memory
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_malloc(nbytes, C.byref(p)))
        return p.value or 0

    def malloc_upload(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_malloc_upload(nbytes, C.byref(p)))
        return p.value or 0

    def free(self, ptr: int) -> None:
        self._lib.b200ag_free(ptr)

    def h2d(self, ptr: int, host: np.ndarray) -> None:
        host = np.asarray(host, order="C")
        self._check(self._lib.b200ag_memcpy_h2d(ptr, host.ctypes.data, host.nbytes))

    def d2h(self, host: np.ndarray, ptr: int) -> None:
        assert host.flags.c_contiguous
        self._check(self._lib.b200ag_memcpy_d2h(host.ctypes.data, ptr, host.nbytes))
        self._check_index_error()

    def _check_index_error(self) -> None:
        """Device index arrays are bounds-checked by the kernels that use them; what NumPy raises at the indexing
        call surfaces here, at the first synchronisation after it."""
        raised = C.c_int(0)
        self._lib.b200ag_index_error(C.byref(raised))
        if raised.value & 2:
            raise RuntimeError("autograd_b200: a peer-memory allreduce gave up waiting for another rank (bounded spin timed "
                               "out); the gradient on this rank is not the sum over ranks and replicas may have diverged - "
                               "abort the job")
        if raised.value & 1:
            raise IndexError("autograd_b200: an index array on the device held an index that is out of bounds for the "
                             "indexed axis (detected by a gather / np.add.at / item-assignment kernel since the last "
                             "synchronisation; the offending accesses were skipped)")

    def d2d(self, dst: int, src: int, nbytes: int) -> None:
        self._check(self._lib.b200ag_memcpy_d2d(dst, src, nbytes))

    def memset(self, ptr: int, byte: int, nbytes: int) -> None:
        self._check(self._lib.b200ag_memset(ptr, byte, nbytes))

    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_host_alloc(nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int) -> None:
        self._check(self._lib.b200ag_host_free(ptr))

    def mem_stats(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._check(self._lib.b200ag_mem_stats(C.byref(a), C.byref(b), C.byref(c)))
        return {"in_use": a.value, "cached": b.value, "peak_in_use": c.value}

    def empty_cache(self) -> None:
        self._check(self._lib.b200ag_empty_cache())

    # ------------------------------------------------------------ runtime
    def synchronize(self) -> None:
        from . import lazy

        lazy.flush_deferred()       # queued contractions are work the caller expects to be done after a synchronisation
        self._check(self._lib.b200ag_synchronize())
        self._check_index_error()

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.b200ag_launch_count(C.byref(n)))
        return n.value

    def stream_handle(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_stream(C.byref(p)))
        return p.value or 0

    def event_create(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_event_create(C.byref(p)))
        return p.value

    def event_record(self, ev: int) -> None:
        self._check(self._lib.b200ag_event_record(ev))

    def event_synchronize(self, ev: int) -> None:
        self._check(self._lib.b200ag_event_synchronize(ev))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self._lib.b200ag_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    def event_destroy(self, ev: int) -> None:
        self._check(self._lib.b200ag_event_destroy(ev))

    # ------------------------------------------------------------ copy streams
    def stream_create(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.b200ag_stream_create(C.byref(p)))
        return p.value

    def stream_destroy(self, s: int) -> None:
        self._check(self._lib.b200ag_stream_destroy(s))

    def stream_synchronize(self, s) -> None:
        self._check(self._lib.b200ag_stream_synchronize(s))

    def event_record_on(self, ev: int, s) -> None:
        self._check(self._lib.b200ag_event_record_on(ev, s))

    def stream_wait_event(self, s, ev: int) -> None:
        self._check(self._lib.b200ag_stream_wait_event(s, ev))

    def h2d_on(self, ptr: int, host: np.ndarray, s) -> None:
        assert host.flags.c_contiguous
        self._check(self._lib.b200ag_memcpy_h2d_on(ptr, host.ctypes.data, host.nbytes, s))

    def d2h_on(self, host: np.ndarray, ptr: int, s) -> None:
        assert host.flags.c_contiguous
        self._check(self._lib.b200ag_memcpy_d2h_on(host.ctypes.data, ptr, host.nbytes, s))

    # ------------------------------------------------------------ peer-memory allreduce
    def comm_create(self, rank: int, world: int, slot_bytes: int) -> bytes:
        h = C.create_string_buffer(64)
        self._check(self._lib.b200ag_comm_create(rank, world, slot_bytes, h))
        return h.raw

    def comm_connect(self, handles: bytes) -> None:
        buf = C.create_string_buffer(handles, len(handles))
        self._check(self._lib.b200ag_comm_connect(buf))

    def comm_block_bytes(self, world: int, slot_bytes: int) -> int:
        n = C.c_size_t()
        self._check(self._lib.b200ag_comm_block_bytes(world, slot_bytes, C.byref(n)))
        return n.value

    def comm_attach(self, rank: int, world: int, slot_bytes: int, peer_ptrs, multicast_ptr: int) -> None:
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        self._check(self._lib.b200ag_comm_attach(rank, world, slot_bytes, arr, multicast_ptr or None))

    def comm_allreduce_sum_f64(self, out: int, src: int, n: int) -> None:
        self._check(self._lib.b200ag_comm_allreduce_sum_f64(out, src, n))

    def comm_status(self) -> int:
        v = C.c_int()
        self._check(self._lib.b200ag_comm_status(C.byref(v)))
        return v.value

    def comm_destroy(self) -> None:
        self._lib.b200ag_comm_destroy()

    # ------------------------------------------------------------ CUDA graphs
    def graph_begin(self) -> None:
        from . import lazy

        lazy.flush_deferred()       # work queued before the capture must not be recorded into it
        self._check(self._lib.b200ag_graph_begin())
        self.capturing = True

    def graph_end(self) -> int:
        from . import lazy

        lazy.flush_deferred()       # everything recorded so far belongs to this graph
        p = C.c_void_p()
        self.capturing = False
        self._check(self._lib.b200ag_graph_end(C.byref(p)))
        return p.value

    def graph_launch(self, g: int) -> None:
        self._check(self._lib.b200ag_graph_launch(g))

    def graph_destroy(self, g: int) -> None:
        self._lib.b200ag_graph_destroy(g)

    def graph_num_nodes(self, g: int) -> int:
        n = C.c_size_t()
        self._check(self._lib.b200ag_graph_num_nodes(g, C.byref(n)))
        return n.value

    # ------------------------------------------------------------ generated elementwise kernels
    def compile(self, source: str, name: str):
        """NVRTC-compile ``source`` for sm_100a and return a launchable kernel handle."""
        from . import kernel_cache

        cubin = kernel_cache.load(source)
        if cubin is None:
            p, n = C.c_void_p(), C.c_size_t()
            self._check(self._lib.b200ag_rtc_compile(source.encode(), (name + ".cu").encode(), C.byref(p), C.byref(n)))
            cubin = C.string_at(p.value, n.value)
            self._lib.b200ag_rtc_free(p)
            kernel_cache.store(source, cubin)
        k = C.c_void_p()
        buf = C.create_string_buffer(cubin, len(cubin))
        self._check(self._lib.b200ag_module_load(buf, len(cubin), name.encode(), C.byref(k)))
        return k.value, buf  # keep the image alive alongside the handle

    def run_program(self, program, leaf_ptrs, const_values, out_ptrs, n: int) -> None:
        """Launch the fused elementwise kernel generated for ``program`` (see codegen.py)."""
        rec = self._kernels.get(program.key)
        if rec is None:
            from . import codegen

            source, name, packer = codegen.generate(program)
            handle, image = self.compile(source, name)
            occ = C.c_int(0)
            self._check(self._lib.b200ag_kernel_max_active_blocks(handle, 256, 0, C.byref(occ)))
            rec = (handle, image, packer, program.launch_config(self.sm_count), max(1, occ.value))
            self._kernels[program.key] = rec
        handle, _image, packer, (block, _elems_per_block, _max_grid), occ = rec
        if n == 0:
            return
        params = packer(n, tuple(leaf_ptrs) + (self._index_flag,) if program.tables else leaf_ptrs, const_values, out_ptrs,
                        program)
        grid = program.grid(n, leaf_ptrs[1], self.sm_count, occ)
        if self.profile is None:
            self._check(self._lib.b200ag_kernel_launch(handle, grid, block, 0, params, len(params)))
        else:
            e0, e1 = self.event_create(), self.event_create()
            self.event_record(e0)
            self._check(self._lib.b200ag_kernel_launch(handle, grid, block, 0, params, len(params)))
            self.event_record(e1)
            self.profile.append((program, n, e0, e1))

    # ------------------------------------------------------------ static kernels
    def copy_strided(self, dst, dst_dt, dst_strides, src, src_dt, src_strides, shape) -> None:
        nd = len(shape)
        self._check(self._lib.b200ag_copy_strided(
            dst, dtype_code(dst_dt), _i64_array(dst_strides), src, dtype_code(src_dt), _i64_array(src_strides),
            _i64_array(shape), nd))

    def add_strided(self, dst, dt, dst_strides, src, src_strides, shape) -> None:
        nd = len(shape)
        self._check(self._lib.b200ag_add_strided(
            dst, dtype_code(dt), _i64_array(dst_strides), src, _i64_array(src_strides), _i64_array(shape), nd))

    def concat(self, dst, dt, src_ptrs, inner, outer: int) -> None:
        """dst[outer, sum(inner)] = hstack of contiguous pieces [outer, inner[k]] of dtype ``dt`` (<= 16 pieces)."""
        arr = (C.c_void_p * len(src_ptrs))(*src_ptrs)
        self._check(self._lib.b200ag_concat(dst, dtype_code(dt), arr, _i64_array(inner), len(src_ptrs), outer))

    def concat_mixed(self, dst, dt, pieces, inner, outer: int) -> None:
        """hstack of float pieces of mixed dtype / constants: pieces = [(ptr, dtype) | (None, fill value)] (<= 16)."""
        n = len(pieces)
        ptrs = (C.c_void_p * n)(*[p if p is not None else None for p, _ in pieces])
        dts = (C.c_int * n)(*[dtype_code(d) if p is not None else 0 for p, d in pieces])
        fills = (C.c_double * n)(*[0.0 if p is not None else float(d) for p, d in pieces])
        self._check(self._lib.b200ag_concat_mixed(dst, dtype_code(dt), ptrs, dts, fills, _i64_array(inner), n, outer))

    def reduce(self, dst, src, dt, op: str, outer: int, red: int, inner: int) -> None:
        self._check(self._lib.b200ag_reduce(dst, src, dtype_code(dt), REDUCE_CODE[op], outer, red, inner))

    def argreduce(self, dst, src, dt, op: str, outer: int, red: int, inner: int) -> None:
        self._check(self._lib.b200ag_argreduce(dst, src, dtype_code(dt), REDUCE_CODE[op], outer, red, inner))

    def logsumexp(self, dst, src, dt, outer: int, red: int, inner: int) -> None:
        self._check(self._lib.b200ag_logsumexp(dst, src, dtype_code(dt), outer, red, inner))

    def gemm(self, c, c_dt, c_str, a, a_dt, a_str, b, b_dt, b_str, M, N, K, batch=1) -> None:
        """C[b] = A[b] @ B[b]; *_str = (row_stride, col_stride, batch_stride) in elements."""
        self._check(self._lib.b200ag_gemm(
            c, dtype_code(c_dt), *c_str, a, dtype_code(a_dt), *a_str, b, dtype_code(b_dt), *b_str, M, N, K, batch))

    def gemm_grouped(self, problems) -> None:
        """Independent products in one launch per operand class; ``problems`` = sequence of
        (c, c_dt, (c_rs, c_cs), a, a_dt, (a_rs, a_cs), b, b_dt, (b_rs, b_cs), M, N, K)."""
        from ._lib import GemmProblem

        n = len(problems)
        arr = (GemmProblem * n)()
        for q, (c, c_dt, c_str, a, a_dt, a_str, b, b_dt, b_str, M, N, K) in zip(arr, problems):
            q.C, q.A, q.B = c, a, b
            q.c_dtype, q.a_dtype, q.b_dtype = dtype_code(c_dt), dtype_code(a_dt), dtype_code(b_dt)
            q.c_rs, q.c_cs = c_str
            q.a_rs, q.a_cs = a_str
            q.b_rs, q.b_cs = b_str
            q.M, q.N, q.K = M, N, K
        self._check(self._lib.b200ag_gemm_grouped(arr, n))

    def gemm_tune(self, tile: int = 0, splits: int = 0) -> None:
        """Measurement hook: force tile (1 = 64x64, 2 = 128x128) and K split of the following GEMM launches; 0 = automatic."""
        self._check(self._lib.b200ag_gemm_tune(tile, splits))

    def gemm_f32_tc(self, c, c_rs, a, a_str, b, b_str, M, N, K) -> None:
        """float32 C = A @ B on the tcgen05 tensor cores (3xTF32); *_str = (row_stride, col_stride) in elements."""
        self._check(self._lib.b200ag_gemm_f32_tc(c, c_rs, a, a_str[0], a_str[1], b, b_str[0], b_str[1], M, N, K))

    def gemm_f32_mode(self, mode: int) -> int:
        """0 = float32 products on the FP64 DMMA kernel, 1 = tcgen05 3xTF32 for large ones; returns the old mode."""
        prev = C.c_int(0)
        self._check(self._lib.b200ag_gemm_f32_mode(mode, C.byref(prev)))
        return prev.value

    def gather(self, dst, src, dt, idx_ptrs, dims, n_index: int, inner: int) -> None:
        arr = (C.c_void_p * len(idx_ptrs))(*idx_ptrs)
        self._check(self._lib.b200ag_gather(dst, src, dtype_code(dt), arr, len(idx_ptrs), _i64_array(dims),
                                            n_index, inner))

    def scatter_add(self, dst, src, dt, idx_ptrs, dims, n_index: int, inner: int) -> None:
        arr = (C.c_void_p * len(idx_ptrs))(*idx_ptrs)
        self._check(self._lib.b200ag_scatter_add(dst, src, dtype_code(dt), arr, len(idx_ptrs), _i64_array(dims),
                                                 n_index, inner))

    def scatter_assign(self, dst, src, dt, idx_ptrs, dims, n_index: int, inner: int) -> None:
        arr = (C.c_void_p * len(idx_ptrs))(*idx_ptrs)
        self._check(self._lib.b200ag_scatter_assign(dst, src, dtype_code(dt), arr, len(idx_ptrs), _i64_array(dims),
                                                    n_index, inner))

    def cumulate(self, dst, src, dt, op: str, outer: int, n: int, inner: int) -> None:
        self._check(self._lib.b200ag_cumulate(dst, src, dtype_code(dt), REDUCE_CODE[op], outer, n, inner))

    def sort(self, keys_out, idx_out, src, dt, rows: int, n: int) -> None:
        """Row-wise stable ascending sort; keys_out / idx_out (int64 argsort) may be None."""
        self._check(self._lib.b200ag_sort(keys_out, idx_out, src, dtype_code(dt), rows, n))

    def conv2d(self, out, a, b, dt, N, Cc, H, W, F, kh, kw, mode: int, flip: int) -> None:
        self._check(self._lib.b200ag_conv2d(out, a, b, dtype_code(dt), N, Cc, H, W, F, kh, kw, mode, flip))

    def comm_mode(self, mode: int) -> int:
        """Peer allreduce algorithm: 0 automatic, 1 one-shot pull, 2 two-shot, 3 push, 4 NVLS multimem; returns the previous setting."""
        prev = C.c_int(0)
        self._check(self._lib.b200ag_comm_mode(mode, C.byref(prev)))
        return prev.value

    def conv_mode(self, direct: int) -> int:
        """1 = direct CUDA-core kernels for few-channel float64 layers (default), 0 = DMMA implicit GEMM only."""
        prev = C.c_int(0)
        self._check(self._lib.b200ag_conv_mode(direct, C.byref(prev)))
        return prev.value

    def conv2d_wgrad(self, db, a, g, dt, N, Cc, H, W, F, kh, kw) -> None:
        self._check(self._lib.b200ag_conv2d_wgrad(db, a, g, dtype_code(dt), N, Cc, H, W, F, kh, kw))


_current = None


def get():
    """The active backend; created on first use.  Raises if the CUDA library or a GPU is missing."""
    global _current
    if _current is None:
        _current = CudaBackend()
    return _current


def set_backend(b) -> None:
    """Install a backend object (used by the CPU test-suite to exercise host logic without a GPU)."""
    global _current
    if _current is not None and _current is not b:
        from . import lazy

        try:
            lazy.flush_deferred()           # queued work belongs to the backend that is being replaced
        except Exception:
            del lazy._gemm_queue[:]
            lazy._gemm_by_buf.clear()
            lazy._scatter_queue.clear()
    _current = b
