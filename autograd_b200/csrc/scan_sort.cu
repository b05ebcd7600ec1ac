// Prefix scans (np.cumsum / np.cumprod) and row sorts (np.sort / np.argsort) — the remaining array primitives the
// reference's VJP table reaches (numpy_vjps.py:539-549 grad_np_cumsum, :774-812 grad_sort / unpermuter).
//
// cumulate: input viewed as [outer, n, inner], scan along n.
//   inner > 1 : one thread per (outer, inner) column walks n sequentially — coalesced across inner and the same
//               addition order as NumPy's add.accumulate (bit-identical).
//   inner == 1: rows are cut into 16 Ki-element segments; pass 1 reduces every segment, pass 2 scans the segment
//               totals, pass 3 scans inside each segment starting from its carry (2 reads + 1 write of the data).
// sort: every row is sorted ascending with NumPy's order (NaN last) by a bitonic network on (key, original index)
//   pairs — a strict total order, so the result is the STABLE sort and argsort is bit-exact for kind="stable" and
//   for any input without ties.  Rows up to 2048 elements are sorted entirely in shared memory; longer rows use
//   global compare-exchange passes for strides >= 2048 and the shared-memory kernel below that.
#include "common.cuh"

namespace b200ag {

// ------------------------------------------------------------------------------------------------ cumulate
template <typename T, bool PROD>
__device__ __forceinline__ T comb(T a, T b) { return PROD ? a * b : a + b; }
template <typename T, bool PROD>
__device__ __forceinline__ T ident() { return PROD ? (T)1 : (T)0; }

template <typename T, bool PROD>
__global__ void __launch_bounds__(256) cumulate_cols_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t outer,
                                                             int64_t n, int64_t inner) {
  const int64_t total = outer * inner;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
    const int64_t o = t / inner, c = t - o * inner;
    const T* s = src + o * n * inner + c;
    T* d = dst + o * n * inner + c;
    T acc = s[0];
    d[0] = acc;
    for (int64_t k = 1; k < n; ++k) {
      acc = comb<T, PROD>(acc, s[k * inner]);
      d[k * inner] = acc;
    }
  }
}

constexpr int SEG = 16384;        // elements per segment
constexpr int PER_THREAD = 8;     // 256 threads x 8 = 2048 elements per block step

template <typename T, bool PROD>
__device__ __forceinline__ T block_exclusive_scan(T v, T* warp_tot, T& block_total) {
  // exclusive scan of one value per thread over a 256-thread block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = comb<T, PROD>(up, inc);
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  T wprefix = ident<T, PROD>();
  T tot = ident<T, PROD>();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    T x = warp_tot[k];
    if (k < w) wprefix = comb<T, PROD>(wprefix, x);
    tot = comb<T, PROD>(tot, x);
  }
  block_total = tot;
  T excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = ident<T, PROD>();
  __syncthreads();
  return comb<T, PROD>(wprefix, excl);
}

// pass 1: totals[row, seg] = combine of the segment
template <typename T, bool PROD>
__global__ void __launch_bounds__(256) segment_totals_kernel(T* __restrict__ totals, const T* __restrict__ src, int64_t n,
                                                              int64_t nseg) {
  __shared__ T warp_tot[8];
  const int64_t row = blockIdx.y, seg = blockIdx.x;
  const T* s = src + row * n;
  const int64_t lo = seg * SEG, hi = min(lo + (int64_t)SEG, n);
  T acc = ident<T, PROD>();
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc = comb<T, PROD>(acc, s[i]);
  T tot;
  block_exclusive_scan<T, PROD>(acc, warp_tot, tot);
  if (threadIdx.x == 0) totals[row * nseg + seg] = tot;
}

// pass 2: in-place exclusive scan of each row of totals (few thousand entries at most: one thread per row)
template <typename T, bool PROD>
__global__ void scan_totals_kernel(T* __restrict__ totals, int64_t rows, int64_t nseg) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  T* t = totals + row * nseg;
  T acc = ident<T, PROD>();
  for (int64_t k = 0; k < nseg; ++k) {
    T v = t[k];
    t[k] = acc;
    acc = comb<T, PROD>(acc, v);
  }
}

// pass 3: inclusive scan inside each segment, seeded with the segment's carry
template <typename T, bool PROD>
__global__ void __launch_bounds__(256) segment_scan_kernel(T* __restrict__ dst, const T* __restrict__ src, const T* __restrict__ carries,
                                                            int64_t n, int64_t nseg) {
  __shared__ T warp_tot[8];
  const int64_t row = blockIdx.y, seg = blockIdx.x;
  const T* s = src + row * n;
  T* d = dst + row * n;
  const int64_t lo = seg * SEG, hi = min(lo + (int64_t)SEG, n);
  T carry = carries ? carries[row * nseg + seg] : ident<T, PROD>();
  for (int64_t base = lo; base < hi; base += 256 * PER_THREAD) {
    const int64_t i0 = base + (int64_t)threadIdx.x * PER_THREAD;
    T v[PER_THREAD];
    T run = ident<T, PROD>();
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
      v[k] = (i0 + k < hi) ? s[i0 + k] : ident<T, PROD>();
      run = comb<T, PROD>(run, v[k]);
      v[k] = run;
    }
    T tot;
    T excl = block_exclusive_scan<T, PROD>(run, warp_tot, tot);
    const T off = comb<T, PROD>(carry, excl);
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k)
      if (i0 + k < hi) d[i0 + k] = comb<T, PROD>(off, v[k]);
    carry = comb<T, PROD>(carry, tot);
  }
}

template <typename T, bool PROD>
static int cumulate_t(void* dst, const void* src, int64_t outer, int64_t n, int64_t inner) {
  if (inner > 1 || n <= 32) {
    cumulate_cols_kernel<T, PROD><<<grid_for(outer * inner, 256, 8), 256, 0, stream()>>>((T*)dst, (const T*)src, outer, n, inner);
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  if (outer > 65535) B200AG_FAIL("cumulate: too many rows for a last-axis scan (65535 max)");
  const int64_t nseg = (n + SEG - 1) / SEG;
  void* totals = nullptr;
  if (nseg > 1) {
    if (b200ag_malloc((size_t)(outer * nseg) * sizeof(T), &totals)) return 1;
    dim3 g((unsigned)nseg, (unsigned)outer);
    segment_totals_kernel<T, PROD><<<g, 256, 0, stream()>>>((T*)totals, (const T*)src, n, nseg);
    count_launch();
    scan_totals_kernel<T, PROD><<<(unsigned)((outer + 63) / 64), 64, 0, stream()>>>((T*)totals, outer, nseg);
    count_launch();
  }
  dim3 g((unsigned)nseg, (unsigned)outer);
  segment_scan_kernel<T, PROD><<<g, 256, 0, stream()>>>((T*)dst, (const T*)src, (const T*)totals, n, nseg);
  cudaError_t e = cudaPeekAtLastError();
  if (totals) b200ag_free(totals);
  if (e != cudaSuccess) {
    set_error(std::string("cumulate launch: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------ sort
template <typename T>
__device__ __forceinline__ bool is_nan_t(T) { return false; }
template <>
__device__ __forceinline__ bool is_nan_t<double>(double v) { return v != v; }
template <>
__device__ __forceinline__ bool is_nan_t<float>(float v) { return v != v; }

// strict total order on (key, original index); indices >= n are padding and sort after everything
template <typename T>
__device__ __forceinline__ bool pair_less(T a, uint32_t ia, T b, uint32_t ib, uint32_t n) {
  const bool pa = ia >= n, pb = ib >= n;
  if (pa | pb) return (pa == pb) ? ia < ib : pb;
  const bool na = is_nan_t(a), nb = is_nan_t(b);
  if (na | nb) return (na == nb) ? ia < ib : nb;
  if (a < b) return true;
  if (b < a) return false;
  return ia < ib;
}

constexpr int TILE = 2048;   // elements sorted in shared memory by one 1024-thread block

template <typename T>
__device__ __forceinline__ void cmp_swap(T* k, uint32_t* x, uint32_t i, uint32_t p, bool ascending, uint32_t n) {
  T a = k[i], b = k[p];
  uint32_t ia = x[i], ib = x[p];
  if (pair_less(b, ib, a, ia, n) == ascending) {
    k[i] = b; k[p] = a; x[i] = ib; x[p] = ia;
  }
}

// Runs the bitonic stages (k, j) with j < TILE for k in [k_lo, k_hi] on one TILE-sized slice held in shared memory.
// For k <= TILE every j is local; for k > TILE only the tail j = TILE/2 .. 1 of that k is done here.
template <typename T>
__global__ void __launch_bounds__(1024) bitonic_local_kernel(T* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t P,
                                                              uint32_t n, uint32_t k_lo, uint32_t k_hi, int load_input,
                                                              const T* __restrict__ src, int64_t src_pitch) {
  __shared__ T sk[TILE];
  __shared__ uint32_t sx[TILE];
  const uint32_t row = blockIdx.y;
  const uint32_t base = blockIdx.x * TILE;
  T* krow = keys + (size_t)row * P;
  uint32_t* xrow = idx + (size_t)row * P;
  const uint32_t tile_n = min((uint32_t)TILE, P);
  for (uint32_t t = threadIdx.x; t < tile_n; t += blockDim.x) {
    const uint32_t i = base + t;
    if (load_input) {
      sk[t] = (i < n) ? src[(int64_t)row * src_pitch + i] : (T)0;
      sx[t] = i;
    } else {
      sk[t] = krow[i];
      sx[t] = xrow[i];
    }
  }
  __syncthreads();
  for (uint32_t k = k_lo; k <= k_hi; k <<= 1) {
    for (uint32_t j = min(k >> 1, (uint32_t)TILE >> 1); j >= 1; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < tile_n / 2; t += blockDim.x) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // local index with bit j clear
        const bool ascending = (((base + i) & k) == 0);
        cmp_swap(sk, sx, i, i | j, ascending, n);
      }
      __syncthreads();
    }
  }
  for (uint32_t t = threadIdx.x; t < tile_n; t += blockDim.x) {
    krow[base + t] = sk[t];
    xrow[base + t] = sx[t];
  }
}

// one global compare-exchange stage (k, j) with j >= TILE
template <typename T>
__global__ void __launch_bounds__(256) bitonic_global_kernel(T* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t P, uint32_t n,
                                                              uint32_t k, uint32_t j) {
  const uint32_t row = blockIdx.y;
  T* krow = keys + (size_t)row * P;
  uint32_t* xrow = idx + (size_t)row * P;
  const uint32_t half = P >> 1;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < half; t += gridDim.x * blockDim.x) {
    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    cmp_swap(krow, xrow, i, i | j, (i & k) == 0, n);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) sort_emit_kernel(T* __restrict__ keys_out, int64_t* __restrict__ idx_out, const T* __restrict__ keys,
                                                         const uint32_t* __restrict__ idx, int64_t rows, uint32_t n, uint32_t P) {
  const int64_t total = rows * (int64_t)n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / n, c = t - r * n;
    if (keys_out) keys_out[t] = keys[r * P + c];
    if (idx_out) idx_out[t] = (int64_t)idx[r * P + c];
  }
}

template <typename T>
static int sort_t(void* keys_out, int64_t* idx_out, const void* src, int64_t rows, int64_t n) {
  uint32_t P = 1;
  while (P < (uint32_t)n) P <<= 1;
  void *keys = nullptr, *idx = nullptr;
  if (b200ag_malloc((size_t)rows * P * sizeof(T), &keys)) return 1;
  if (b200ag_malloc((size_t)rows * P * 4, &idx)) { b200ag_free(keys); return 1; }
  const uint32_t tiles = (P + TILE - 1) / TILE;
  dim3 gl(tiles, (unsigned)rows);
  // stage 1: every tile fully sorted (alternating direction), reading the caller's rows and padding to P
  bitonic_local_kernel<T><<<gl, 1024, 0, stream()>>>((T*)keys, (uint32_t*)idx, P, (uint32_t)n, 2u, min(P, (uint32_t)TILE), 1,
                                                      (const T*)src, n);
  count_launch();
  for (uint32_t k = 2 * TILE; k <= P && k != 0; k <<= 1) {
    for (uint32_t j = k >> 1; j >= TILE; j >>= 1) {
      dim3 gg(min((P / 2 + 255) / 256, 4u * 148u), (unsigned)rows);
      bitonic_global_kernel<T><<<gg, 256, 0, stream()>>>((T*)keys, (uint32_t*)idx, P, (uint32_t)n, k, j);
      count_launch();
    }
    bitonic_local_kernel<T><<<gl, 1024, 0, stream()>>>((T*)keys, (uint32_t*)idx, P, (uint32_t)n, k, k, 0, nullptr, 0);
    count_launch();
  }
  sort_emit_kernel<T><<<grid_for(rows * n, 256, 8), 256, 0, stream()>>>((T*)keys_out, idx_out, (const T*)keys, (const uint32_t*)idx,
                                                                        rows, (uint32_t)n, P);
  cudaError_t e = cudaPeekAtLastError();
  b200ag_free(keys);
  b200ag_free(idx);
  if (e != cudaSuccess) {
    set_error(std::string("sort launch: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  count_launch();
  return 0;
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

int b200ag_cumulate(void* dst, const void* src, int dtype, int op, int64_t outer, int64_t n, int64_t inner) {
  B200AG_CHECK_INIT();
  if (outer <= 0 || n <= 0 || inner <= 0) return 0;
  if (op != B200AG_SUM && op != B200AG_PROD) B200AG_FAIL("cumulate: op must be SUM or PROD");
  const bool prod = op == B200AG_PROD;
  switch (dtype) {
    case B200AG_F64: return prod ? cumulate_t<double, true>(dst, src, outer, n, inner) : cumulate_t<double, false>(dst, src, outer, n, inner);
    case B200AG_F32: return prod ? cumulate_t<float, true>(dst, src, outer, n, inner) : cumulate_t<float, false>(dst, src, outer, n, inner);
    case B200AG_I64: return prod ? cumulate_t<int64_t, true>(dst, src, outer, n, inner) : cumulate_t<int64_t, false>(dst, src, outer, n, inner);
    case B200AG_I32: return prod ? cumulate_t<int32_t, true>(dst, src, outer, n, inner) : cumulate_t<int32_t, false>(dst, src, outer, n, inner);
  }
  B200AG_FAIL("cumulate: unsupported dtype");
}

int b200ag_sort(void* keys_out, int64_t* idx_out, const void* src, int dtype, int64_t rows, int64_t n) {
  B200AG_CHECK_INIT();
  if (rows <= 0 || n <= 0) return 0;
  if (n > (1ll << 30)) B200AG_FAIL("sort: rows longer than 2^30 elements are not supported");
  if (rows > 65535) B200AG_FAIL("sort: more than 65535 rows are not supported");
  switch (dtype) {
    case B200AG_F64: return sort_t<double>(keys_out, idx_out, src, rows, n);
    case B200AG_F32: return sort_t<float>(keys_out, idx_out, src, rows, n);
    case B200AG_I64: return sort_t<int64_t>(keys_out, idx_out, src, rows, n);
    case B200AG_I32: return sort_t<int32_t>(keys_out, idx_out, src, rows, n);
  }
  B200AG_FAIL("sort: unsupported dtype");
}

}  // extern "C"
