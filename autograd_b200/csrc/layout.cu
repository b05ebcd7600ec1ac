// Layout kernels: strided copy with dtype conversion, strided in-place add,
// gather and scatter-add.  HBM-bound byte movers: coalesced on the side that is
// contiguous, 64-bit elements, grids sized in multiples of the SM count.
#include "common.cuh"

namespace b200ag {

struct StridedDesc {
  int ndim;
  int64_t shape[B200AG_MAX_DIMS];
  int64_t dst_strides[B200AG_MAX_DIMS];
  int64_t src_strides[B200AG_MAX_DIMS];
};

// Drop unit dims and merge neighbours that are contiguous in BOTH operands so
// the per-element index math is as short as possible.
static int64_t collapse(StridedDesc& d, const int64_t* shape, const int64_t* ds, const int64_t* ss, int ndim) {
  int64_t n = 1;
  int k = 0;
  for (int i = 0; i < ndim; ++i) {
    n *= shape[i];
    if (shape[i] == 1) continue;
    if (k > 0 && d.dst_strides[k - 1] == ds[i] * shape[i] && d.src_strides[k - 1] == ss[i] * shape[i]) {
      d.shape[k - 1] *= shape[i];
      d.dst_strides[k - 1] = ds[i];
      d.src_strides[k - 1] = ss[i];
    } else {
      d.shape[k] = shape[i];
      d.dst_strides[k] = ds[i];
      d.src_strides[k] = ss[i];
      ++k;
    }
  }
  if (k == 0) {
    d.shape[0] = 1;
    d.dst_strides[0] = 1;
    d.src_strides[0] = 1;
    k = 1;
  }
  d.ndim = k;
  return n;
}

template <typename T>
struct Conv {
  template <typename S>
  __device__ static T from(S v) { return (T)v; }
};
template <>
struct Conv<bool> {
  template <typename S>
  __device__ static bool from(S v) { return v != (S)0; }
};

template <typename D, typename S, bool ACCUM>
__global__ void __launch_bounds__(256) strided_kernel(D* __restrict__ dst, const S* __restrict__ src, StridedDesc d,
                                                       int64_t n) {
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    int64_t rem = i, so = 0, dof = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k > 0; --k) {
      int64_t q = rem / d.shape[k];
      int64_t c = rem - q * d.shape[k];
      rem = q;
      so += c * d.src_strides[k];
      dof += c * d.dst_strides[k];
    }
    so += rem * d.src_strides[0];
    dof += rem * d.dst_strides[0];
    if (ACCUM)
      dst[dof] += Conv<D>::from(src[so]);
    else
      dst[dof] = Conv<D>::from(src[so]);
  }
}

// 2-D transpose-like copies (inner dim contiguous in src, strided in dst or the
// reverse) go through a padded shared-memory tile so both sides stay coalesced.
template <typename T>
__global__ void __launch_bounds__(256) transpose_tile_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                              int64_t rows, int64_t cols, int64_t dst_rs,
                                                              int64_t dst_cs, int64_t src_rs, int64_t src_cs,
                                                              int64_t batch, int64_t dst_bs, int64_t src_bs) {
  // src is contiguous along cols (src_cs == 1); dst is contiguous along rows (dst_rs == 1)
  __shared__ T tile[32][33];
  int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  int64_t ntiles = tiles_c * tiles_r * batch;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int64_t b = t / (tiles_c * tiles_r);
    int64_t tr = (t / tiles_c) % tiles_r, tc = t % tiles_c;
    const T* s = src + b * src_bs;
    T* dd = dst + b * dst_bs;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t r = tr * 32 + ty + j, c = tc * 32 + tx;
      if (r < rows && c < cols) tile[ty + j][tx] = s[r * src_rs + c * src_cs];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t c = tc * 32 + ty + j, r = tr * 32 + tx;
      if (r < rows && c < cols) dd[r * dst_rs + c * dst_cs] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

template <typename D, typename S, bool ACCUM>
static int launch_strided(void* dst, const void* src, const StridedDesc& d, int64_t n) {
  if (n == 0) return 0;
  unsigned grid = grid_for(n, 256, 16);
  strided_kernel<D, S, ACCUM><<<grid, 256, 0, stream()>>>((D*)dst, (const S*)src, d, n);
  B200AG_LAUNCH_CHECK();
  return 0;
}

template <typename D, bool ACCUM>
static int dispatch_src(void* dst, const void* src, int src_dtype, const StridedDesc& d, int64_t n) {
  switch (src_dtype) {
    case B200AG_F64: return launch_strided<D, double, ACCUM>(dst, src, d, n);
    case B200AG_F32: return launch_strided<D, float, ACCUM>(dst, src, d, n);
    case B200AG_I64: return launch_strided<D, int64_t, ACCUM>(dst, src, d, n);
    case B200AG_I32: return launch_strided<D, int32_t, ACCUM>(dst, src, d, n);
    case B200AG_BOOL: return launch_strided<D, uint8_t, ACCUM>(dst, src, d, n);
  }
  B200AG_FAIL("copy_strided: unsupported source dtype");
}

template <typename T>
static bool try_transpose(void* dst, const void* src, const StridedDesc& d) {
  // pattern: 2 or 3 collapsed dims, one of the two inner dims unit-stride in src and the other in dst
  if (d.ndim != 2 && d.ndim != 3) return false;
  int o = d.ndim - 2;
  int r = o, c = o + 1;  // the kernel wants src contiguous along "cols" and dst contiguous along "rows"
  if (d.src_strides[o] == 1 && d.dst_strides[o + 1] == 1) { r = o + 1; c = o; }
  else if (!(d.src_strides[o + 1] == 1 && d.dst_strides[o] == 1)) return false;
  int64_t rows = d.shape[r], cols = d.shape[c];
  if (rows < 16 || cols < 16) return false;
  int64_t batch = d.ndim == 3 ? d.shape[0] : 1;
  int64_t ntiles = ((rows + 31) / 32) * ((cols + 31) / 32) * batch;
  unsigned grid = (unsigned)(ntiles < (int64_t)sm_count() * 16 ? ntiles : (int64_t)sm_count() * 16);
  transpose_tile_kernel<T><<<grid, 256, 0, stream()>>>(
      (T*)dst, (const T*)src, rows, cols, d.dst_strides[r], d.dst_strides[c], d.src_strides[r],
      d.src_strides[c], batch, d.ndim == 3 ? d.dst_strides[0] : 0, d.ndim == 3 ? d.src_strides[0] : 0);
  return true;
}

// ---------------------------------------------------------------- gather / scatter-add
struct IndexDesc {
  int nidx;
  const int64_t* idx[B200AG_MAX_DIMS];
  int64_t dims[B200AG_MAX_DIMS];
  int* err;   // library-wide "index out of bounds" flag (mapped host memory, read at the next synchronisation)
};

// Row addressed by the i-th index tuple, or -1 when an entry is outside [-dim, dim): NumPy raises IndexError there;
// the kernels skip the access (never touch memory outside the array) and raise the flag the host reads back.
__device__ __forceinline__ int64_t linear_row(const IndexDesc& d, int64_t i) {
  int64_t row = 0;
#pragma unroll 1
  for (int k = 0; k < d.nidx; ++k) {
    int64_t v = d.idx[k][i];
    if (v < 0) v += d.dims[k];  // NumPy negative-index wrap
    if ((uint64_t)v >= (uint64_t)d.dims[k]) {
      *d.err = 1;
      return -1;
    }
    row = row * d.dims[k] + v;
  }
  return row;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(T* __restrict__ dst, const T* __restrict__ src, IndexDesc d,
                                                      int64_t n_index, int64_t inner) {
  int64_t n = n_index * inner;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (inner == 1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
      const int64_t row = linear_row(d, i);
      dst[i] = row < 0 ? T(0) : src[row];
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
      int64_t r = i / inner, c = i - r * inner;
      const int64_t row = linear_row(d, r);
      dst[i] = row < 0 ? T(0) : src[row * inner + c];
    }
  }
}

__device__ __forceinline__ void atomic_add_t(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(int64_t* p, int64_t v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}
__device__ __forceinline__ void atomic_add_t(int32_t* p, int32_t v) { atomicAdd(p, v); }

// np.add.at semantics: every (possibly duplicated) index contributes; RED.ADD.F64
// in L2 makes the accumulation order free, which only perturbs the last bits.
template <typename T>
__global__ void __launch_bounds__(256) scatter_add_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                           IndexDesc d, int64_t n_index, int64_t inner) {
  int64_t n = n_index * inner;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (inner == 1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
      const int64_t row = linear_row(d, i);
      if (row >= 0) atomic_add_t(dst + row, src[i]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
      int64_t r = i / inner, c = i - r * inner;
      const int64_t row = linear_row(d, r);
      if (row >= 0) atomic_add_t(dst + row * inner + c, src[i]);
    }
  }
}

// np.concatenate of contiguous pieces of one dtype in ONE launch (flatten of a gradient tree: misc/flatten.py:11-27,
// the pieces handed to the allreduce).  Piece k is viewed as [outer, inner[k]], the result as [outer, sum inner].
struct ConcatDesc {
  const void* src[16];
  int64_t inner[16];
  int64_t start[17];
  int npieces;
  int64_t outer, total_inner;
};

template <typename T>
__global__ void __launch_bounds__(256) concat_kernel(T* __restrict__ dst, ConcatDesc d) {
  const int64_t n = d.outer * d.total_inner;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += step) {
    const int64_t row = idx / d.total_inner, col = idx - row * d.total_inner;
    int k = 0;
#pragma unroll 1
    while (k + 1 < d.npieces && col >= d.start[k + 1]) ++k;
    dst[idx] = reinterpret_cast<const T*>(d.src[k])[row * d.inner[k] + (col - d.start[k])];
  }
}

// x[idx] = values (integer-array item assignment, e.g. unpermuter in numpy_vjps.py:803-806).  With duplicate indices
// NumPy's sequential loop leaves the LAST value in place; two passes reproduce that deterministically: pass 1 records,
// per destination row, the largest position in the index list that addresses it (atomicMax), pass 2 lets exactly that
// position write.
__global__ void __launch_bounds__(256) scatter_winner_kernel(long long* __restrict__ winner, IndexDesc d, int64_t n_index) {
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_index; i += step) {
    const int64_t row = linear_row(d, i);
    if (row >= 0) atomicMax(winner + row, (long long)i);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_assign_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                              const long long* __restrict__ winner, IndexDesc d,
                                                              int64_t n_index, int64_t inner) {
  int64_t n = n_index * inner;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    int64_t r = i / inner, c = i - r * inner;
    const int64_t row = linear_row(d, r);
    if (row >= 0 && winner[row] == (long long)r) dst[row * inner + c] = src[i];
  }
}

// np.hstack of float pieces of MIXED dtype and of symbolic constants in ONE launch: the cell input of an LSTM step is
// hstack(input float32, hiddens float64, ones) (examples/rnn.py:21 via examples/lstm.py:33-36), which NumPy promotes to
// float64.  Piece k is [outer, inner[k]] of float32 / float64, or a constant fill when its source is NULL.
struct ConcatMixedDesc {
  const void* src[16];
  double fill[16];
  int src_f32[16];
  int64_t inner[16];
  int64_t start[17];
  int npieces;
  int64_t outer, total_inner;
};

template <typename T>
__global__ void __launch_bounds__(256) concat_mixed_kernel(T* __restrict__ dst, ConcatMixedDesc d) {
  const int64_t n = d.outer * d.total_inner;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += step) {
    const int64_t row = idx / d.total_inner, col = idx - row * d.total_inner;
    int k = 0;
#pragma unroll 1
    while (k + 1 < d.npieces && col >= d.start[k + 1]) ++k;
    const int64_t off = row * d.inner[k] + (col - d.start[k]);
    double v;
    if (!d.src[k]) v = d.fill[k];
    else if (d.src_f32[k]) v = (double)reinterpret_cast<const float*>(d.src[k])[off];
    else v = reinterpret_cast<const double*>(d.src[k])[off];
    dst[idx] = (T)v;
  }
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

int b200ag_copy_strided(void* dst, int dst_dtype, const int64_t* dst_strides, const void* src, int src_dtype,
                        const int64_t* src_strides, const int64_t* shape, int ndim) {
  B200AG_CHECK_INIT();
  if (ndim < 0 || ndim > B200AG_MAX_DIMS) B200AG_FAIL("copy_strided: ndim out of range");
  StridedDesc d;
  int64_t n = collapse(d, shape, dst_strides, src_strides, ndim);
  if (n == 0) return 0;
  if (dst_dtype == src_dtype) {
    size_t es = dtype_size(dst_dtype);
    if (d.ndim == 1 && d.dst_strides[0] == 1 && d.src_strides[0] == 1) {
      B200AG_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * es, cudaMemcpyDeviceToDevice, stream()));
      count_launch();
      return 0;
    }
    bool done = false;
    if (es == 8) done = try_transpose<double>(dst, src, d);
    if (es == 4) done = try_transpose<float>(dst, src, d);
    if (done) {
      B200AG_LAUNCH_CHECK();
      return 0;
    }
  }
  switch (dst_dtype) {
    case B200AG_F64: return dispatch_src<double, false>(dst, src, src_dtype, d, n);
    case B200AG_F32: return dispatch_src<float, false>(dst, src, src_dtype, d, n);
    case B200AG_I64: return dispatch_src<int64_t, false>(dst, src, src_dtype, d, n);
    case B200AG_I32: return dispatch_src<int32_t, false>(dst, src, src_dtype, d, n);
    case B200AG_BOOL: return dispatch_src<bool, false>(dst, src, src_dtype, d, n);
  }
  B200AG_FAIL("copy_strided: unsupported destination dtype");
}

int b200ag_add_strided(void* dst, int dtype, const int64_t* dst_strides, const void* src, const int64_t* src_strides,
                       const int64_t* shape, int ndim) {
  B200AG_CHECK_INIT();
  if (ndim < 0 || ndim > B200AG_MAX_DIMS) B200AG_FAIL("add_strided: ndim out of range");
  StridedDesc d;
  int64_t n = collapse(d, shape, dst_strides, src_strides, ndim);
  switch (dtype) {
    case B200AG_F64: return launch_strided<double, double, true>(dst, src, d, n);
    case B200AG_F32: return launch_strided<float, float, true>(dst, src, d, n);
    case B200AG_I64: return launch_strided<int64_t, int64_t, true>(dst, src, d, n);
    case B200AG_I32: return launch_strided<int32_t, int32_t, true>(dst, src, d, n);
  }
  B200AG_FAIL("add_strided: unsupported dtype");
}

static int fill_index_desc(IndexDesc& d, const int64_t* const* idx, int nidx, const int64_t* dims) {
  if (nidx < 1 || nidx > B200AG_MAX_DIMS) B200AG_FAIL("gather/scatter: nidx out of range");
  d.nidx = nidx;
  for (int k = 0; k < nidx; ++k) {
    d.idx[k] = idx[k];
    d.dims[k] = dims[k];
  }
  d.err = index_flag();
  if (!d.err) return 1;
  return 0;
}

int b200ag_gather(void* dst, const void* src, int dtype, const int64_t* const* idx, int nidx, const int64_t* src_dims,
                  int64_t n_index, int64_t inner) {
  B200AG_CHECK_INIT();
  IndexDesc d;
  if (fill_index_desc(d, idx, nidx, src_dims)) return 1;
  int64_t n = n_index * inner;
  if (n == 0) return 0;
  unsigned grid = grid_for(n, 256, 16);
  switch (dtype_size(dtype)) {
    case 8: gather_kernel<int64_t><<<grid, 256, 0, stream()>>>((int64_t*)dst, (const int64_t*)src, d, n_index, inner); break;
    case 4: gather_kernel<int32_t><<<grid, 256, 0, stream()>>>((int32_t*)dst, (const int32_t*)src, d, n_index, inner); break;
    case 1: gather_kernel<uint8_t><<<grid, 256, 0, stream()>>>((uint8_t*)dst, (const uint8_t*)src, d, n_index, inner); break;
    default: B200AG_FAIL("gather: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_scatter_add(void* dst, const void* src, int dtype, const int64_t* const* idx, int nidx,
                       const int64_t* dst_dims, int64_t n_index, int64_t inner) {
  B200AG_CHECK_INIT();
  IndexDesc d;
  if (fill_index_desc(d, idx, nidx, dst_dims)) return 1;
  int64_t n = n_index * inner;
  if (n == 0) return 0;
  unsigned grid = grid_for(n, 256, 16);
  switch (dtype) {
    case B200AG_F64: scatter_add_kernel<double><<<grid, 256, 0, stream()>>>((double*)dst, (const double*)src, d, n_index, inner); break;
    case B200AG_F32: scatter_add_kernel<float><<<grid, 256, 0, stream()>>>((float*)dst, (const float*)src, d, n_index, inner); break;
    case B200AG_I64: scatter_add_kernel<int64_t><<<grid, 256, 0, stream()>>>((int64_t*)dst, (const int64_t*)src, d, n_index, inner); break;
    case B200AG_I32: scatter_add_kernel<int32_t><<<grid, 256, 0, stream()>>>((int32_t*)dst, (const int32_t*)src, d, n_index, inner); break;
    default: B200AG_FAIL("scatter_add: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_scatter_assign(void* dst, const void* src, int dtype, const int64_t* const* idx, int nidx,
                          const int64_t* dst_dims, int64_t n_index, int64_t inner) {
  B200AG_CHECK_INIT();
  IndexDesc d;
  if (fill_index_desc(d, idx, nidx, dst_dims)) return 1;
  int64_t n = n_index * inner;
  if (n == 0) return 0;
  int64_t rows = 1;
  for (int k = 0; k < nidx; ++k) rows *= dst_dims[k];
  void* winner = nullptr;
  if (b200ag_malloc((size_t)rows * sizeof(long long), &winner)) return 1;
  if (b200ag_memset(winner, 0xFF, (size_t)rows * sizeof(long long))) return 1;   // every row: no writer (-1)
  scatter_winner_kernel<<<grid_for(n_index, 256, 16), 256, 0, stream()>>>((long long*)winner, d, n_index);
  B200AG_LAUNCH_CHECK();
  unsigned grid = grid_for(n, 256, 16);
  const long long* w = (const long long*)winner;
  switch (dtype_size(dtype)) {
    case 8: scatter_assign_kernel<int64_t><<<grid, 256, 0, stream()>>>((int64_t*)dst, (const int64_t*)src, w, d, n_index, inner); break;
    case 4: scatter_assign_kernel<int32_t><<<grid, 256, 0, stream()>>>((int32_t*)dst, (const int32_t*)src, w, d, n_index, inner); break;
    case 1: scatter_assign_kernel<uint8_t><<<grid, 256, 0, stream()>>>((uint8_t*)dst, (const uint8_t*)src, w, d, n_index, inner); break;
    default: b200ag_free(winner); B200AG_FAIL("scatter_assign: unsupported dtype");
  }
  b200ag_free(winner);   // stream-ordered allocator: the block is reused only by work enqueued after these kernels
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_concat(void* dst, int dtype, const void* const* srcs, const int64_t* inner, int npieces, int64_t outer) {
  B200AG_CHECK_INIT();
  if (npieces < 1 || npieces > 16) B200AG_FAIL("concat: between 1 and 16 pieces per call");
  ConcatDesc d;
  d.npieces = npieces;
  d.outer = outer;
  int64_t pos = 0;
  for (int k = 0; k < npieces; ++k) {
    d.src[k] = srcs[k];
    d.inner[k] = inner[k];
    d.start[k] = pos;
    pos += inner[k];
  }
  d.start[npieces] = pos;
  d.total_inner = pos;
  int64_t n = outer * pos;
  if (n == 0) return 0;
  unsigned grid = grid_for(n, 256, 8);
  switch (dtype_size(dtype)) {
    case 8: concat_kernel<int64_t><<<grid, 256, 0, stream()>>>((int64_t*)dst, d); break;
    case 4: concat_kernel<int32_t><<<grid, 256, 0, stream()>>>((int32_t*)dst, d); break;
    case 1: concat_kernel<uint8_t><<<grid, 256, 0, stream()>>>((uint8_t*)dst, d); break;
    default: B200AG_FAIL("concat: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_concat_mixed(void* dst, int dtype, const void* const* srcs, const int* src_dtypes, const double* fills,
                        const int64_t* inner, int npieces, int64_t outer) {
  B200AG_CHECK_INIT();
  if (npieces < 1 || npieces > 16) B200AG_FAIL("concat_mixed: between 1 and 16 pieces per call");
  if (dtype != B200AG_F64 && dtype != B200AG_F32) B200AG_FAIL("concat_mixed: output dtype must be float64 or float32");
  ConcatMixedDesc d;
  d.npieces = npieces;
  d.outer = outer;
  int64_t pos = 0;
  for (int k = 0; k < npieces; ++k) {
    if (srcs[k] && src_dtypes[k] != B200AG_F64 && src_dtypes[k] != B200AG_F32)
      B200AG_FAIL("concat_mixed: pieces must be float64 or float32");
    d.src[k] = srcs[k];
    d.src_f32[k] = src_dtypes[k] == B200AG_F32;
    d.fill[k] = fills[k];
    d.inner[k] = inner[k];
    d.start[k] = pos;
    pos += inner[k];
  }
  d.start[npieces] = pos;
  d.total_inner = pos;
  int64_t n = outer * pos;
  if (n == 0) return 0;
  unsigned grid = grid_for(n, 256, 8);
  if (dtype == B200AG_F64) concat_mixed_kernel<double><<<grid, 256, 0, stream()>>>((double*)dst, d);
  else concat_mixed_kernel<float><<<grid, 256, 0, stream()>>>((float*)dst, d);
  B200AG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
