/*
 * b200ag.h — C ABI of libb200ag.so, the sm_100a kernel library behind the
 * device-resident array type that plugs into HIPS/autograd.
 *
 * The reference (autograd 1.9.1) is pure Python and has no FFI of its own: its
 * "kernel boundary" is the raw NumPy call at autograd/tracer.py:94 (forward)
 * and the NumPy operators used inside VJP closures (autograd/core.py:30,
 * autograd/numpy/numpy_vjps.py).  Every entry point below names the NumPy /
 * SciPy behaviour (and the reference call site that reaches it) it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, int64 shapes/strides (in ELEMENTS), enums as
 *     int.  No torch / Python types cross this boundary.
 *   - every function returns 0 on success; non-zero means failure and
 *     b200ag_last_error() holds a message (the ctypes layer raises RuntimeError).
 *   - kernels never allocate: outputs are preallocated by the caller
 *     (autograd's lifetime model: intermediates live as long as VJP closures
 *     hold them, docs/updateguide.md:71).
 *   - all work is enqueued on ONE library-owned stream (b200ag_stream); the
 *     reference is single-threaded (tracer.py:154 global trace_stack).
 */
#ifndef B200AG_H
#define B200AG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AG_MAX_DIMS 8

/* dtype codes (NumPy dtypes the five BASELINE configs produce) */
enum {
  B200AG_F64 = 0,  /* numpy.float64 — autograd's default */
  B200AG_F32 = 1,  /* numpy.float32 */
  B200AG_I64 = 2,  /* numpy.int64   (astype(int), np.mod indices; fluidsim.py:56-62) */
  B200AG_I32 = 3,  /* numpy.int32 */
  B200AG_BOOL = 4  /* numpy.bool_   (comparison masks; numpy_vjps.py:521,911) */
};

/* reduction codes */
enum { B200AG_SUM = 0, B200AG_MAX = 1, B200AG_MIN = 2, B200AG_PROD = 3 };

/* ---------------------------------------------------------------- runtime */
int b200ag_init(int device);                /* select GPU, create stream; idempotent */
int b200ag_shutdown(void);
const char* b200ag_last_error(void);
int b200ag_device_count(int* n);
int b200ag_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
int b200ag_stream(void** stream);           /* cudaStream_t every kernel is launched on */
int b200ag_synchronize(void);
int b200ag_launch_count(uint64_t* n);       /* kernels launched by this library so far */

/* ---------------------------------------------------------------- memory
 * Replaces numpy's allocator for array payloads (np.zeros/np.empty results,
 * ufunc outputs).  Caching, stream-ordered on the library stream. */
int b200ag_malloc(size_t nbytes, void** ptr);
int b200ag_malloc_upload(size_t nbytes, void** ptr);  /* destination of a host upload (graph-constant safe) */
int b200ag_free(void* ptr);
int b200ag_empty_cache(void);
int b200ag_mem_stats(size_t* in_use, size_t* cached, size_t* peak_in_use);
int b200ag_host_alloc(size_t nbytes, void** ptr);   /* pinned host staging */
int b200ag_host_free(void* ptr);
int b200ag_memcpy_h2d(void* dst, const void* src, size_t nbytes);   /* async on the stream */
int b200ag_memcpy_d2h(void* dst, const void* src, size_t nbytes);   /* returns after the copy */
int b200ag_memcpy_d2d(void* dst, const void* src, size_t nbytes);
int b200ag_memset(void* dst, int byte, size_t nbytes);

/* ---------------------------------------------------------------- events (bench timing on the library stream) */
int b200ag_event_create(void** ev);
int b200ag_event_record(void* ev);
int b200ag_event_synchronize(void* ev);
int b200ag_event_elapsed_ms(void* start, void* stop, float* ms);
int b200ag_event_destroy(void* ev);

/* ---------------------------------------------------------------- copy streams
 * Host<->device copies that overlap the compute stream (np.asarray uploads / .get() downloads of chunked
 * elementwise workloads).  A NULL stream argument means the library's compute stream. */
int b200ag_stream_create(void** stream);
int b200ag_stream_destroy(void* stream);
int b200ag_stream_synchronize(void* stream);
int b200ag_event_record_on(void* ev, void* stream);
int b200ag_stream_wait_event(void* stream, void* ev);
int b200ag_memcpy_h2d_on(void* dst, const void* src, size_t nbytes, void* stream);  /* async; src should be pinned */
int b200ag_memcpy_d2h_on(void* dst, const void* src, size_t nbytes, void* stream);  /* async; dst should be pinned */

/* ---------------------------------------------------------------- generated elementwise kernels
 * Replaces NumPy ufunc inner loops (tracer.py:94 → numpy add/multiply/exp/...)
 * and the operator chains inside VJP closures (numpy_vjps.py:76-191).  The
 * Python layer emits CUDA C for a fused expression DAG; these compile it for
 * sm_100a (NVRTC, --fmad=false so every op is individually rounded like a
 * NumPy ufunc) and launch it. */
int b200ag_rtc_compile(const char* src, const char* name, void** cubin, size_t* cubin_size);
int b200ag_rtc_free(void* cubin);
int b200ag_rtc_log(const char** log);       /* log of the last compile */
int b200ag_module_load(const void* cubin, size_t size, const char* kernel_name, void** kernel);
int b200ag_kernel_launch(void* kernel, uint32_t grid, uint32_t block, uint32_t smem,
                         const void* params, size_t params_size);   /* one by-value struct param */
int b200ag_kernel_max_active_blocks(void* kernel, int block, size_t smem, int* nblocks);

/* ---------------------------------------------------------------- layout
 * Strided copy with dtype conversion: ndarray.astype / np.ascontiguousarray /
 * transposed & sliced views made contiguous / np.concatenate pieces / np.pad
 * interior (numpy_wrapper.py:57-59, numpy_vjps.py:246-251,750-759,957-975).
 * dst and src are described by shape + element strides (may be 0 or negative). */
int b200ag_copy_strided(void* dst, int dst_dtype, const int64_t* dst_strides,
                        const void* src, int src_dtype, const int64_t* src_strides,
                        const int64_t* shape, int ndim);
/* dst[...] += src[...] over a strided destination window (VJP of basic slicing:
 * np.add.at(A, slice_tuple, g), numpy_vjps.py:941-950) */
int b200ag_add_strided(void* dst, int dtype, const int64_t* dst_strides,
                       const void* src, const int64_t* src_strides,
                       const int64_t* shape, int ndim);
/* np.concatenate of up to 16 contiguous pieces of one dtype in a single launch: piece k is
 * [outer, inner[k]], dst is [outer, sum(inner)] (np.concatenate behind concatenate_args and
 * misc.flatten, numpy_wrapper.py:57-59, misc/flatten.py:11-27). */
int b200ag_concat(void* dst, int dtype, const void* const* srcs, const int64_t* inner,
                  int npieces, int64_t outer);
/* The same for float pieces of mixed dtype and symbolic constants: piece k is float32 or float64
 * (src_dtypes[k]), or the constant fills[k] when srcs[k] is NULL; dst is float64 or float32.
 * hstack((input float32, hiddens float64, ones)) of an LSTM / RNN cell (examples/rnn.py:21,
 * examples/lstm.py:33-36: NumPy promotes the row to float64) in one launch. */
int b200ag_concat_mixed(void* dst, int dtype, const void* const* srcs, const int* src_dtypes,
                        const double* fills, const int64_t* inner, int npieces, int64_t outer);

/* ---------------------------------------------------------------- reductions
 * np.sum / np.max / np.min / np.prod over one contiguous group of axes of a
 * C-contiguous array viewed as [outer, reduce, inner] (numpy_vjps.py:443-530,
 * unbroadcast :883-892). */
int b200ag_reduce(void* dst, const void* src, int dtype, int op,
                  int64_t outer, int64_t reduce, int64_t inner);
/* argmax/argmin over the reduce axis of [outer, reduce, inner]; first occurrence
 * wins (numpy semantics); dst is int64. */
int b200ag_argreduce(int64_t* dst, const void* src, int dtype, int op,
                     int64_t outer, int64_t reduce, int64_t inner);

/* ---------------------------------------------------------------- contractions
 * C[M,N] = A[M,K] · B[K,N] with arbitrary element strides on every operand, so
 * swapaxes/transpose views feed the kernel without a copy.  Accumulates in
 * FP64 on the DMMA tensor pipe (mma.sync m8n8k4.f64); inputs may be f32 or f64
 * (NumPy promotes f32·f64 → f64, examples/rnn.py:21) and the result is
 * rounded to c_dtype on store (onp.asarray(out, dtype=A_dtype),
 * numpy_vjps.py:624,638).  batch > 1 = stacked matmul with batch strides.
 * Replaces np.dot / np.matmul / np.tensordot (numpy_vjps.py:567-742). */
int b200ag_gemm(void* C, int c_dtype, int64_t c_rs, int64_t c_cs, int64_t c_bs,
                const void* A, int a_dtype, int64_t a_rs, int64_t a_cs, int64_t a_bs,
                const void* B, int b_dtype, int64_t b_rs, int64_t b_cs, int64_t b_bs,
                int64_t M, int64_t N, int64_t K, int64_t batch);

/* A group of independent products in ONE launch (per operand type / layout class):
 * the four gate products of an LSTM cell that share their left operand
 * (examples/lstm.py:38-49), the G·Bᵀ / Aᵀ·G adjoints of one tape step
 * (numpy_vjps.py:615-638), the layer gradients of an MLP.  Each problem is what
 * b200ag_gemm computes for batch = 1; together they fill the SMs without split-K.
 * Results are deterministic (K-split partials are added in split order). */
typedef struct {
  void* C; const void* A; const void* B;
  int c_dtype, a_dtype, b_dtype, reserved;
  int64_t c_rs, c_cs, a_rs, a_cs, b_rs, b_cs;
  int64_t M, N, K;
} b200ag_gemm_problem;
int b200ag_gemm_grouped(const b200ag_gemm_problem* problems, int count);
/* Measurement hook: force the K split (0 = automatic) of the following launches; scripts/gemm_sweep.py uses it to
 * calibrate the automatic choice.  `tile` is accepted for ABI stability and ignored (one tile shape, 64x64). */
int b200ag_gemm_tune(int tile, int splits);

/* float32 x float32 -> float32 products on the tcgen05 tensor cores (accumulator in
 * TMEM, operands staged by TMA): each operand is split into two TF32 terms and
 * Ahi*Bhi + Ahi*Blo + Alo*Bhi is accumulated in FP32 ("3xTF32"), which keeps the
 * result within the float32 tolerance of onp.tensordot's sgemm (numpy_vjps.py:621-637,
 * 677-719).  C has unit column stride and row stride c_rs; A and B take arbitrary
 * element strides.  b200ag_gemm routes large all-float32 products here unless
 * b200ag_gemm_f32_mode(0, ...) selected the FP64 DMMA kernel for them. */
int b200ag_gemm_f32_tc(void* C, int64_t c_rs, const void* A, int64_t a_rs, int64_t a_cs,
                       const void* B, int64_t b_rs, int64_t b_cs, int64_t M, int64_t N, int64_t K);
int b200ag_gemm_f32_mode(int mode, int* previous);

/* ---------------------------------------------------------------- gather / scatter-add
 * out[i] = src[idx0[i], idx1[i], ..., :] (integer-array indexing on the leading
 * nidx axes, ArrayBox.__getitem__, numpy_boxes.py:16-18) and its VJP
 * np.add.at(A, idx, g) with duplicate indices accumulated (numpy_vjps.py:941-950).
 * Negative indices wrap like NumPy.  inner = product of the trailing,
 * un-indexed dimensions (contiguous). */
/* Index entries outside [-dim, dim) (NumPy: IndexError, raised by the index machinery
 * under numpy_boxes.py:16-18 / np.add.at, numpy_vjps.py:947) are never dereferenced:
 * gather yields 0 for them, scatter skips them, and a library-wide flag in mapped host
 * memory is raised.  b200ag_index_error reports and clears it (call after a
 * synchronisation); b200ag_index_flag gives generated kernels the flag's device address. */
int b200ag_index_flag(void** device_ptr);
int b200ag_index_error(int* raised);   /* bit 0: index out of bounds (cleared by the call); bit 1: collective timed out */
int b200ag_gather(void* dst, const void* src, int dtype,
                  const int64_t* const* idx, int nidx, const int64_t* src_dims,
                  int64_t n_index, int64_t inner);
int b200ag_scatter_add(void* dst, const void* src, int dtype,
                       const int64_t* const* idx, int nidx, const int64_t* dst_dims,
                       int64_t n_index, int64_t inner);
/* x[idx] = values for integer index arrays (item assignment on a raw array, e.g.
 * `unsort[permutation] = range(n)` in unpermuter, numpy_vjps.py:803-806).  Same
 * index convention as b200ag_scatter_add; among duplicate indices the LAST one in
 * the index list wins, as in NumPy's sequential assignment (two passes: atomicMax
 * of the writer position per row, then only that writer stores). */
int b200ag_scatter_assign(void* dst, const void* src, int dtype,
                          const int64_t* const* idx, int nidx, const int64_t* dst_dims,
                          int64_t n_index, int64_t inner);

/* ---------------------------------------------------------------- scans and sorts
 * b200ag_cumulate: np.cumsum (op = B200AG_SUM) / np.cumprod (B200AG_PROD) along the
 * middle axis of a contiguous [outer, n, inner] view (numpy_vjps.py:539-549).
 * b200ag_sort: every row of a contiguous [rows, n] array sorted ascending in NumPy's
 * order (NaN last); keys_out receives the sorted values, idx_out the int64 argsort
 * (either may be NULL).  The sort is stable, so indices are bit-exact against
 * np.argsort(kind="stable") and, without ties, against any kind
 * (np.sort / np.argsort behind grad_sort, numpy_vjps.py:774-812). */
int b200ag_cumulate(void* dst, const void* src, int dtype, int op,
                    int64_t outer, int64_t n, int64_t inner);
int b200ag_sort(void* keys_out, int64_t* idx_out, const void* src, int dtype,
                int64_t rows, int64_t n);


/* ---------------------------------------------------------------- logsumexp
 * scipy.special.logsumexp(a, axis, keepdims=True) over the middle axis of
 * [outer, reduce, inner], following SciPy 1.18.1 _logsumexp (max; m = #max;
 * s = sum exp(a - max) over the non-max entries; log1p(s/m) + log(m) + max);
 * reference primitive: autograd/scipy/special.py:120. */
int b200ag_logsumexp(void* dst, const void* src, int dtype,
                     int64_t outer, int64_t reduce, int64_t inner);

/* ---------------------------------------------------------------- convolution
 * autograd.scipy.signal.convolve(A, B, axes=([2,3],[2,3]), dot_axes=([1],[0]),
 * mode) for A[N,C,H,W], B[C,F,kh,kw] → out[N,F,Ho,Wo] (true convolution: the
 * kernel is flipped), mode 0 = valid, 1 = full (autograd/scipy/signal.py:10-65),
 * plus the weight-gradient form used by its VJP (signal.py:153-214):
 * dB[C,F,kh,kw] = sum_n valid-correlate(A[n,c], G[n,f]). */
int b200ag_conv2d(void* out, const void* A, const void* B, int dtype,
                  int64_t N, int64_t C, int64_t H, int64_t W,
                  int64_t F, int64_t kh, int64_t kw, int mode, int flip);
int b200ag_conv2d_wgrad(void* dB, const void* A, const void* G, int dtype,
                        int64_t N, int64_t C, int64_t H, int64_t W,
                        int64_t F, int64_t kh, int64_t kw);
/* Kernel family for float64 convolutions: 1 (default) = layers with few output channels
 * (F <= 8, C <= 4, 3- or 5-wide filters) run on register-tiled CUDA-core kernels, everything
 * else on the implicit-GEMM DMMA kernels; 0 = DMMA kernels only. */
int b200ag_conv_mode(int direct, int* previous);


/* ---------------------------------------------------------------- gradient allreduce over NVLink peer memory
 * Not in the reference (single process).  Sum of batch-sharded gradient vectors, one process per GPU: every rank
 * publishes its vector in a CUDA-IPC buffer and one kernel per rank reads all peers through NVLink and adds them
 * in rank order.  handles = world x 64-byte cudaIpcMemHandle_t, exchanged by the caller (torch.distributed). */
int b200ag_comm_create(int rank, int world, size_t slot_bytes, void* ipc_handle_out);
int b200ag_comm_connect(const void* handles);
/* The same communicator over comm blocks the caller allocated symmetrically on every rank (zero-filled,
 * b200ag_comm_block_bytes each) and mapped into this process — peer_bases[p] = rank p's block as addressable from
 * here, multicast_base = the NVLS multicast mapping of those blocks, or NULL (torch.distributed's symmetric memory
 * hands out both).  With a multicast mapping the sum is taken INSIDE the NVSwitch: rank r reduces chunk r with one
 * multimem.ld_reduce.add.f64 per element and broadcasts it with one multimem.st (mode 4). */
int b200ag_comm_block_bytes(int world, size_t slot_bytes, size_t* total);
int b200ag_comm_attach(int rank, int world, size_t slot_bytes, void* const* peer_bases, void* multicast_base);
int b200ag_comm_allreduce_sum_f64(void* out, const void* src, int64_t n);
/* 1 when a collective gave up waiting for a peer (bounded spin): the flag lives in mapped host memory, is also
 * reported by b200ag_index_error (bit 1) after any synchronisation, and is never cleared — replicas may have diverged. */
int b200ag_comm_status(int* timed_out);
/* 0 = choose automatically (NVLS above 4 ranks when a multicast mapping is attached, else one-shot; two-shot
 * reduce-scatter + all-gather from 4 ranks and 4 MB up), 1 = one-shot (pull), 2 = two-shot, 3 = one-shot push (every
 * rank writes its vector into the peers' inboxes, the sum reads local memory only), 4 = NVLS multimem. */
int b200ag_comm_mode(int mode, int* previous);
int b200ag_comm_destroy(void);

/* ---------------------------------------------------------------- CUDA graphs
 * Capture everything the library enqueues between begin and end (a recorded
 * forward+backward tape) and replay it with one launch (core.py:26-33 is the
 * loop being replayed).  Memory allocated while capturing stays owned by the
 * graph until it is destroyed. */
int b200ag_graph_begin(void);
int b200ag_graph_end(void** graph);
int b200ag_graph_launch(void* graph);
int b200ag_graph_destroy(void* graph);
int b200ag_graph_num_nodes(void* graph, size_t* n);

#ifdef __cplusplus
}
#endif
#endif /* B200AG_H */
