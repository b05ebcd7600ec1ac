"""ctypes binding of libb200ag.so (ABI: include/b200ag.h).

There is deliberately no fallback: if the shared library is missing or does not
export a declared symbol, importing this module raises.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200ag.so")

c_i64 = C.c_int64
c_i64_p = C.POINTER(C.c_int64)
c_void_pp = C.POINTER(C.c_void_p)
c_size_p = C.POINTER(C.c_size_t)

# name -> argtypes; every function returns int (0 = ok) unless listed in _RESTYPE.
PROTOTYPES = {
    "b200ag_init": [C.c_int],
    "b200ag_shutdown": [],
    "b200ag_last_error": [],
    "b200ag_device_count": [C.POINTER(C.c_int)],
    "b200ag_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_size_p],
    "b200ag_stream": [c_void_pp],
    "b200ag_synchronize": [],
    "b200ag_launch_count": [C.POINTER(C.c_uint64)],
    "b200ag_malloc": [C.c_size_t, c_void_pp],
    "b200ag_malloc_upload": [C.c_size_t, c_void_pp],
    "b200ag_free": [C.c_void_p],
    "b200ag_empty_cache": [],
    "b200ag_mem_stats": [c_size_p, c_size_p, c_size_p],
    "b200ag_host_alloc": [C.c_size_t, c_void_pp],
    "b200ag_host_free": [C.c_void_p],
    "b200ag_memcpy_h2d": [C.c_void_p, C.c_void_p, C.c_size_t],
    "b200ag_memcpy_d2h": [C.c_void_p, C.c_void_p, C.c_size_t],
    "b200ag_memcpy_d2d": [C.c_void_p, C.c_void_p, C.c_size_t],
    "b200ag_memset": [C.c_void_p, C.c_int, C.c_size_t],
    "b200ag_event_create": [c_void_pp],
    "b200ag_event_record": [C.c_void_p],
    "b200ag_event_synchronize": [C.c_void_p],
    "b200ag_event_elapsed_ms": [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)],
    "b200ag_event_destroy": [C.c_void_p],
    "b200ag_stream_create": [c_void_pp],
    "b200ag_stream_destroy": [C.c_void_p],
    "b200ag_stream_synchronize": [C.c_void_p],
    "b200ag_event_record_on": [C.c_void_p, C.c_void_p],
    "b200ag_stream_wait_event": [C.c_void_p, C.c_void_p],
    "b200ag_memcpy_h2d_on": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    "b200ag_memcpy_d2h_on": [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    "b200ag_rtc_compile": [C.c_char_p, C.c_char_p, c_void_pp, c_size_p],
    "b200ag_rtc_free": [C.c_void_p],
    "b200ag_rtc_log": [C.POINTER(C.c_char_p)],
    "b200ag_module_load": [C.c_void_p, C.c_size_t, C.c_char_p, c_void_pp],
    "b200ag_kernel_launch": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t],
    "b200ag_kernel_max_active_blocks": [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int)],
    "b200ag_copy_strided": [C.c_void_p, C.c_int, c_i64_p, C.c_void_p, C.c_int, c_i64_p, c_i64_p, C.c_int],
    "b200ag_add_strided": [C.c_void_p, C.c_int, c_i64_p, C.c_void_p, c_i64_p, c_i64_p, C.c_int],
    "b200ag_concat": [C.c_void_p, C.c_int, c_void_pp, c_i64_p, C.c_int, c_i64],
    "b200ag_concat_mixed": [C.c_void_p, C.c_int, c_void_pp, C.POINTER(C.c_int), C.POINTER(C.c_double), c_i64_p, C.c_int, c_i64],
    "b200ag_reduce": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_i64, c_i64, c_i64],
    "b200ag_argreduce": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_i64, c_i64, c_i64],
    "b200ag_gemm": [C.c_void_p, C.c_int, c_i64, c_i64, c_i64,
                    C.c_void_p, C.c_int, c_i64, c_i64, c_i64,
                    C.c_void_p, C.c_int, c_i64, c_i64, c_i64,
                    c_i64, c_i64, c_i64, c_i64],
    "b200ag_gemm_grouped": [C.c_void_p, C.c_int],
    "b200ag_gemm_tune": [C.c_int, C.c_int],
    "b200ag_gemm_f32_tc": [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64],
    "b200ag_gemm_f32_mode": [C.c_int, C.POINTER(C.c_int)],
    "b200ag_index_flag": [c_void_pp],
    "b200ag_index_error": [C.POINTER(C.c_int)],
    "b200ag_gather": [C.c_void_p, C.c_void_p, C.c_int, c_void_pp, C.c_int, c_i64_p, c_i64, c_i64],
    "b200ag_scatter_add": [C.c_void_p, C.c_void_p, C.c_int, c_void_pp, C.c_int, c_i64_p, c_i64, c_i64],
    "b200ag_scatter_assign": [C.c_void_p, C.c_void_p, C.c_int, c_void_pp, C.c_int, c_i64_p, c_i64, c_i64],
    "b200ag_cumulate": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_i64, c_i64, c_i64],
    "b200ag_sort": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_i64, c_i64],
    "b200ag_logsumexp": [C.c_void_p, C.c_void_p, C.c_int, c_i64, c_i64, c_i64],
    "b200ag_conv2d": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                      C.c_int, C.c_int],
    "b200ag_conv2d_wgrad": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                            c_i64],
    "b200ag_conv_mode": [C.c_int, C.POINTER(C.c_int)],
    "b200ag_comm_create": [C.c_int, C.c_int, C.c_size_t, C.c_void_p],
    "b200ag_comm_connect": [C.c_void_p],
    "b200ag_comm_block_bytes": [C.c_int, C.c_size_t, c_size_p],
    "b200ag_comm_attach": [C.c_int, C.c_int, C.c_size_t, c_void_pp, C.c_void_p],
    "b200ag_comm_allreduce_sum_f64": [C.c_void_p, C.c_void_p, c_i64],
    "b200ag_comm_mode": [C.c_int, C.POINTER(C.c_int)],
    "b200ag_comm_status": [C.POINTER(C.c_int)],
    "b200ag_comm_destroy": [],
    "b200ag_graph_begin": [],
    "b200ag_graph_end": [c_void_pp],
    "b200ag_graph_launch": [C.c_void_p],
    "b200ag_graph_destroy": [C.c_void_p],
    "b200ag_graph_num_nodes": [C.c_void_p, c_size_p],
}
_RESTYPE = {"b200ag_last_error": C.c_char_p}


class GemmProblem(C.Structure):
    """b200ag_gemm_problem (include/b200ag.h)."""

    _fields_ = [("C", C.c_void_p), ("A", C.c_void_p), ("B", C.c_void_p),
                ("c_dtype", C.c_int), ("a_dtype", C.c_int), ("b_dtype", C.c_int), ("reserved", C.c_int),
                ("c_rs", c_i64), ("c_cs", c_i64), ("a_rs", c_i64), ("a_cs", c_i64), ("b_rs", c_i64), ("b_cs", c_i64),
                ("M", c_i64), ("N", c_i64), ("K", c_i64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m autograd_b200.build` "
            "(or __graft_entry__.build()); autograd_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.b200ag_last_error()
        raise RuntimeError("libb200ag: " + (msg.decode(errors="replace") if msg else "unknown error"))
