// Reductions over the middle axis of a C-contiguous [outer, reduce, inner] view:
// sum / max / min / prod, argmax / argmin, and SciPy-style logsumexp.
// Warp-shuffle + shared-memory tree inside a CTA; when the reduced axis is long
// and there are few outputs, the axis is split across CTAs into partials and
// reduced again (two deterministic passes, no atomics).
#include <math_constants.h>

#include "common.cuh"

namespace b200ag {

template <int OP, typename A>
struct Red;
template <typename A>
struct Red<B200AG_SUM, A> {
  __device__ static A identity() { return (A)0; }
  __device__ static A apply(A a, A b) { return a + b; }
};
template <typename A>
struct Red<B200AG_PROD, A> {
  __device__ static A identity() { return (A)1; }
  __device__ static A apply(A a, A b) { return a * b; }
};
// NumPy max/min propagate NaN.
__device__ __forceinline__ double max_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double min_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ int64_t max_nan(int64_t a, int64_t b) { return a > b ? a : b; }
__device__ __forceinline__ int64_t min_nan(int64_t a, int64_t b) { return a < b ? a : b; }
template <>
struct Red<B200AG_MAX, double> {
  __device__ static double identity() { return -CUDART_INF; }
  __device__ static double apply(double a, double b) { return max_nan(a, b); }
};
template <>
struct Red<B200AG_MIN, double> {
  __device__ static double identity() { return CUDART_INF; }
  __device__ static double apply(double a, double b) { return min_nan(a, b); }
};
template <>
struct Red<B200AG_MAX, int64_t> {
  __device__ static int64_t identity() { return INT64_MIN; }
  __device__ static int64_t apply(int64_t a, int64_t b) { return a > b ? a : b; }
};
template <>
struct Red<B200AG_MIN, int64_t> {
  __device__ static int64_t identity() { return INT64_MAX; }
  __device__ static int64_t apply(int64_t a, int64_t b) { return a < b ? a : b; }
};

template <typename A>
__device__ __forceinline__ A shfl_down_t(A v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// ---- inner == 1: each output is the reduction of a contiguous run of `reduce` elements.
// One CTA (256 threads) per (row, split) pair; lanes stride the run so loads coalesce.
template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) row_reduce_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                          int64_t outer, int64_t reduce, int64_t split,
                                                          int64_t chunk) {
  __shared__ A warp_part[8];
  int64_t nwork = outer * split;
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    int64_t row = w / split, s = w - row * split;
    int64_t lo = s * chunk, hi = lo + chunk < reduce ? lo + chunk : reduce;
    const T* p = src + row * reduce;
    A acc = Red<OP, A>::identity();
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc = Red<OP, A>::apply(acc, (A)p[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = Red<OP, A>::apply(acc, shfl_down_t(acc, d));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      A v = threadIdx.x < 8 ? warp_part[threadIdx.x] : Red<OP, A>::identity();
#pragma unroll
      for (int d = 4; d > 0; d >>= 1) v = Red<OP, A>::apply(v, shfl_down_t(v, d));
      if (threadIdx.x == 0) dst[w] = (T)v;
    }
    __syncthreads();
  }
}

// ---- inner == 1 and short rows: one warp per row (rows of 2..~1024 elements, e.g. the
// class axis of a softmax or a 2-wide pooling window).
template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) row_reduce_warp_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                               int64_t outer, int64_t reduce) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < outer; row += nwarps) {
    const T* p = src + row * reduce;
    A acc = Red<OP, A>::identity();
    for (int64_t i = lane; i < reduce; i += 32) acc = Red<OP, A>::apply(acc, (A)p[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = Red<OP, A>::apply(acc, shfl_down_t(acc, d));
    if (lane == 0) dst[row] = (T)acc;
  }
}

// ---- very short rows (reduce <= 8): one thread per row, the warp still reads a
// contiguous 32*reduce-element span.
template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) row_reduce_thread_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                                 int64_t outer, int64_t reduce) {
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < outer; row += step) {
    const T* p = src + row * reduce;
    A acc = Red<OP, A>::identity();
    for (int64_t i = 0; i < reduce; ++i) acc = Red<OP, A>::apply(acc, (A)p[i]);
    dst[row] = (T)acc;
  }
}

// ---- inner > 1: column-style reduction.  Threads run along `inner` (coalesced), 8 thread
// rows of the CTA take interleaved slices of the reduced axis and combine through smem;
// gridDim.y splits the reduced axis further when there are few output columns.
template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) col_reduce_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                          int64_t outer, int64_t reduce, int64_t inner,
                                                          int64_t split, int64_t chunk) {
  __shared__ A part[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int64_t col_tiles = (inner + 31) / 32;
  int64_t nwork = outer * split * col_tiles;
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    int64_t ct = w % col_tiles;
    int64_t s = (w / col_tiles) % split;
    int64_t o = w / (col_tiles * split);
    int64_t col = ct * 32 + tx;
    int64_t lo = s * chunk, hi = lo + chunk < reduce ? lo + chunk : reduce;
    A acc = Red<OP, A>::identity();
    if (col < inner) {
      const T* p = src + o * reduce * inner + col;
      for (int64_t r = lo + ty; r < hi; r += 8) acc = Red<OP, A>::apply(acc, (A)p[r * inner]);
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int j = 1; j < 8; ++j) acc = Red<OP, A>::apply(acc, part[j][tx]);
      if (col < inner) dst[(o * split + s) * inner + col] = (T)acc;
    }
    __syncthreads();
  }
}

// ---- inner > 1 and a SHORT reduced axis (pooling windows, small channel counts): one thread per output,
// consecutive threads on consecutive `inner` positions, so every 32-byte sector that is fetched is used.
template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) col_reduce_small_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                                int64_t outer, int64_t reduce, int64_t inner) {
  int64_t total = outer * inner;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    int64_t o = idx / inner, c = idx - o * inner;
    const T* p = src + o * reduce * inner + c;
    A acc = Red<OP, A>::identity();
    for (int64_t r = 0; r < reduce; ++r) acc = Red<OP, A>::apply(acc, (A)p[r * inner]);
    dst[idx] = (T)acc;
  }
}

template <typename T, typename A, int OP>
static int reduce_impl(T* dst, const T* src, int64_t outer, int64_t reduce, int64_t inner) {
  const int64_t target_ctas = (int64_t)sm_count() * 4;
  if (inner == 1) {
    if (reduce <= 8 && outer >= 1024) {
      row_reduce_thread_kernel<T, A, OP><<<grid_for(outer, 256, 8), 256, 0, stream()>>>(dst, src, outer, reduce);
      B200AG_LAUNCH_CHECK();
      return 0;
    }
    if (reduce <= 2048 && outer >= 64) {
      row_reduce_warp_kernel<T, A, OP><<<grid_for(outer * 32, 256, 8), 256, 0, stream()>>>(dst, src, outer, reduce);
      B200AG_LAUNCH_CHECK();
      return 0;
    }
    int64_t split = 1;
    if (outer < target_ctas) {
      split = target_ctas / outer;
      int64_t max_split = (reduce + 4095) / 4096;  // at least 4096 elements per CTA
      if (split > max_split) split = max_split;
      if (split < 1) split = 1;
    }
    if (outer * reduce <= 65536) split = 1;   // tiny inputs: a second launch costs more than the parallelism gains
    int64_t chunk = (reduce + split - 1) / split;
    split = (reduce + chunk - 1) / chunk;
    if (split == 1) {
      int64_t g = outer < target_ctas * 2 ? outer : target_ctas * 2;
      row_reduce_kernel<T, A, OP><<<(unsigned)g, 256, 0, stream()>>>(dst, src, outer, reduce, 1, reduce);
      B200AG_LAUNCH_CHECK();
      return 0;
    }
    void* scratch = nullptr;
    if (b200ag_malloc((size_t)(outer * split) * sizeof(T), &scratch)) return 1;
    row_reduce_kernel<T, A, OP><<<(unsigned)(outer * split), 256, 0, stream()>>>((T*)scratch, src, outer, reduce, split, chunk);
    B200AG_LAUNCH_CHECK();
    int rc = reduce_impl<T, A, OP>(dst, (const T*)scratch, outer, split, 1);
    b200ag_free(scratch);
    return rc;
  }
  if (reduce <= 16 && outer * inner >= 32768) {
    col_reduce_small_kernel<T, A, OP><<<grid_for(outer * inner, 256, 8), 256, 0, stream()>>>(dst, src, outer, reduce, inner);
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  int64_t col_tiles = (inner + 31) / 32;
  int64_t base = outer * col_tiles;
  int64_t split = 1;
  if (base < target_ctas) {
    split = target_ctas / base;
    int64_t max_split = (reduce + 63) / 64;  // at least 64 rows per CTA
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
  }
  if (outer * reduce * inner <= 131072) split = 1;   // tiny inputs (bias gradients): one launch instead of two
  int64_t chunk = (reduce + split - 1) / split;
  split = (reduce + chunk - 1) / chunk;
  int64_t nwork = base * split;
  unsigned grid = (unsigned)(nwork < target_ctas * 4 ? nwork : target_ctas * 4);
  if (split == 1) {
    col_reduce_kernel<T, A, OP><<<grid, 256, 0, stream()>>>(dst, src, outer, reduce, inner, 1, reduce);
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  void* scratch = nullptr;
  if (b200ag_malloc((size_t)(outer * split * inner) * sizeof(T), &scratch)) return 1;
  col_reduce_kernel<T, A, OP><<<grid, 256, 0, stream()>>>((T*)scratch, src, outer, reduce, inner, split, chunk);
  B200AG_LAUNCH_CHECK();
  int rc = reduce_impl<T, A, OP>(dst, (const T*)scratch, outer, split, inner);
  b200ag_free(scratch);
  return rc;
}

template <typename T, typename A>
static int reduce_op(void* dst, const void* src, int op, int64_t outer, int64_t reduce, int64_t inner) {
  switch (op) {
    case B200AG_SUM: return reduce_impl<T, A, B200AG_SUM>((T*)dst, (const T*)src, outer, reduce, inner);
    case B200AG_MAX: return reduce_impl<T, A, B200AG_MAX>((T*)dst, (const T*)src, outer, reduce, inner);
    case B200AG_MIN: return reduce_impl<T, A, B200AG_MIN>((T*)dst, (const T*)src, outer, reduce, inner);
    case B200AG_PROD: return reduce_impl<T, A, B200AG_PROD>((T*)dst, (const T*)src, outer, reduce, inner);
  }
  B200AG_FAIL("reduce: unknown op");
}

// ---------------------------------------------------------------- argmax / argmin
// One warp per output; first occurrence wins, NaN wins over everything (NumPy).
template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256) argreduce_kernel(int64_t* __restrict__ dst, const T* __restrict__ src,
                                                         int64_t outer, int64_t reduce, int64_t inner) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t nout = outer * inner;
  for (int64_t w = warp; w < nout; w += nwarps) {
    int64_t o = w / inner, c = w - o * inner;
    const T* p = src + o * reduce * inner + c;
    bool have = false, is_nan = false;
    T best = 0;
    int64_t besti = 0;
    for (int64_t r = lane; r < reduce; r += 32) {
      T v = p[r * inner];
      bool vnan = v != v;
      if (is_nan) continue;
      if (!have || vnan || (IS_MAX ? v > best : v < best)) {
        best = v; besti = r; have = true; is_nan = vnan;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      T ob = __shfl_down_sync(0xffffffffu, best, d);
      int64_t oi = __shfl_down_sync(0xffffffffu, besti, d);
      int oh = __shfl_down_sync(0xffffffffu, (int)have, d);
      int on = __shfl_down_sync(0xffffffffu, (int)is_nan, d);
      if (!oh) continue;
      bool take;
      if (!have) take = true;
      else if (is_nan && on) take = oi < besti;
      else if (is_nan) take = false;
      else if (on) take = true;
      else if (ob == best) take = oi < besti;
      else take = IS_MAX ? ob > best : ob < best;
      if (take) { best = ob; besti = oi; have = true; is_nan = on; }
    }
    if (lane == 0) dst[w] = besti;
  }
}

// ---------------------------------------------------------------- logsumexp
// SciPy 1.18.1 scipy/special/_logsumexp.py:_logsumexp for real input, b=None:
//   a_max = max(a); m = count(a == a_max); s = sum(exp(a - a_max) over a != a_max)
//   s = s == 0 ? s : s / m;  out = log1p(s) + log(m) + a_max
// and, where that is not finite, the direct log(sum(exp(a))) (logsumexp(), same file).
template <typename T>
__global__ void __launch_bounds__(256) logsumexp_kernel(T* __restrict__ dst, const T* __restrict__ src,
                                                         int64_t outer, int64_t reduce, int64_t inner) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t nout = outer * inner;
  for (int64_t w = warp; w < nout; w += nwarps) {
    int64_t o = w / inner, c = w - o * inner;
    const T* p = src + o * reduce * inner + c;
    double mx = -CUDART_INF;
    for (int64_t r = lane; r < reduce; r += 32) mx = max_nan(mx, (double)p[r * inner]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    double m = 0.0, s = 0.0;
    for (int64_t r = lane; r < reduce; r += 32) {
      double a = (double)p[r * inner];
      if (a == mx) m += 1.0;
      else s += exp(a - mx);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      m += __shfl_xor_sync(0xffffffffu, m, d);
      s += __shfl_xor_sync(0xffffffffu, s, d);
    }
    double sm = (s == 0.0) ? s : s / m;
    double out = log1p(sm) + log(m) + mx;
    if (!isfinite(out)) {  // rare (inf / nan / empty-ish rows): SciPy falls back to the direct formula
      double direct = 0.0;
      for (int64_t r = lane; r < reduce; r += 32) direct += exp((double)p[r * inner]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) direct += __shfl_xor_sync(0xffffffffu, direct, d);
      out = log(direct);
    }
    if (lane == 0) dst[w] = (T)out;
  }
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

int b200ag_reduce(void* dst, const void* src, int dtype, int op, int64_t outer, int64_t reduce, int64_t inner) {
  B200AG_CHECK_INIT();
  if (outer * inner == 0) return 0;
  if (reduce == 0) B200AG_FAIL("reduce: zero-size reduction axis must be handled by the caller");
  switch (dtype) {
    case B200AG_F64: return reduce_op<double, double>(dst, src, op, outer, reduce, inner);
    case B200AG_F32: return reduce_op<float, double>(dst, src, op, outer, reduce, inner);  // accumulate f32 in f64
    case B200AG_I64: return reduce_op<int64_t, int64_t>(dst, src, op, outer, reduce, inner);
  }
  B200AG_FAIL("reduce: unsupported dtype (cast to float64/float32/int64 first)");
}

int b200ag_argreduce(int64_t* dst, const void* src, int dtype, int op, int64_t outer, int64_t reduce, int64_t inner) {
  B200AG_CHECK_INIT();
  int64_t nout = outer * inner;
  if (nout == 0) return 0;
  if (reduce == 0) B200AG_FAIL("argreduce: attempt to get argmax of an empty sequence");
  if (op != B200AG_MAX && op != B200AG_MIN) B200AG_FAIL("argreduce: op must be MAX or MIN");
  unsigned grid = grid_for(nout * 32, 256, 8);
  bool mx = op == B200AG_MAX;
  switch (dtype) {
    case B200AG_F64:
      if (mx) argreduce_kernel<double, true><<<grid, 256, 0, stream()>>>(dst, (const double*)src, outer, reduce, inner);
      else argreduce_kernel<double, false><<<grid, 256, 0, stream()>>>(dst, (const double*)src, outer, reduce, inner);
      break;
    case B200AG_F32:
      if (mx) argreduce_kernel<float, true><<<grid, 256, 0, stream()>>>(dst, (const float*)src, outer, reduce, inner);
      else argreduce_kernel<float, false><<<grid, 256, 0, stream()>>>(dst, (const float*)src, outer, reduce, inner);
      break;
    case B200AG_I64:
      if (mx) argreduce_kernel<int64_t, true><<<grid, 256, 0, stream()>>>(dst, (const int64_t*)src, outer, reduce, inner);
      else argreduce_kernel<int64_t, false><<<grid, 256, 0, stream()>>>(dst, (const int64_t*)src, outer, reduce, inner);
      break;
    default: B200AG_FAIL("argreduce: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_logsumexp(void* dst, const void* src, int dtype, int64_t outer, int64_t reduce, int64_t inner) {
  B200AG_CHECK_INIT();
  int64_t nout = outer * inner;
  if (nout == 0) return 0;
  unsigned grid = grid_for(nout * 32, 256, 8);
  switch (dtype) {
    case B200AG_F64: logsumexp_kernel<double><<<grid, 256, 0, stream()>>>((double*)dst, (const double*)src, outer, reduce, inner); break;
    case B200AG_F32: logsumexp_kernel<float><<<grid, 256, 0, stream()>>>((float*)dst, (const float*)src, outer, reduce, inner); break;
    default: B200AG_FAIL("logsumexp: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
