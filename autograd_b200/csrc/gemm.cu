// FP64-accumulating GEMM on the DMMA tensor pipe (mma.sync.m8n8k4.f64) for
// np.dot / np.matmul / np.tensordot and their adjoints (numpy_vjps.py:567-742).
// tcgen05 has no f64 kind, so the legacy warp-level DMMA is the tensor path for
// autograd's default dtype.  Operands are addressed through (row, col, batch)
// element strides: G·Bᵀ and Aᵀ·G read the swapped views directly, no transpose
// copy.  f32 operands are widened on load (NumPy promotion f32·f64 → f64) and
// the accumulator is rounded once to the output dtype on store.
//
// Tiling: CTA tile BM×BN×16, double-buffered in shared memory as [row][k] with a
// 4-double pad (conflict-free 64-bit fragment loads), register-staged global
// prefetch of the next k-slab while the current one is multiplied.
#include "common.cuh"

namespace b200ag {

constexpr int BK = 16;
constexpr int LDS = BK + 4;  // row pitch in doubles: (r*20 + k) mod 16 is a bijection on r<4,k<4

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  void* C; const void* A; const void* B;
  int64_t c_rs, c_cs, c_bs;
  int64_t a_rs, a_cs, a_bs;
  int64_t b_rs, b_cs, b_bs;
  int64_t M, N, K;
  int c_is_f32;
  int splits;        // split-K factor: blockIdx.z = batch * splits + split
  int64_t k_chunk;   // K range per split (multiple of BK)
};

template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_dmma_kernel(GemmArgs g) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LA = BM * BK / NT;
  constexpr int LB = BN * BK / NT;
  constexpr int TM = WM / 8, TN = WN / 8;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");

  extern __shared__ double smem[];
  double* As = smem;                    // [2][BM][LDS]
  double* Bs = smem + 2 * BM * LDS;     // [2][BN][LDS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t bz = blockIdx.z / g.splits, kz = blockIdx.z % g.splits;
  const int64_t k_lo = kz * g.k_chunk;
  const int64_t k_hi = (k_lo + g.k_chunk < g.K) ? k_lo + g.k_chunk : g.K;
  const TA* A = (const TA*)g.A + bz * g.a_bs;
  const TB* B = (const TB*)g.B + bz * g.b_bs;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[LA], rb[LB];

  auto load_global = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT;
      int m = A_KMAJOR ? e / BK : e % BM;
      int k = A_KMAJOR ? e % BK : e / BM;
      int64_t gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < k_hi) ? (double)A[gm * g.a_rs + gk * g.a_cs] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT;
      int n = B_KMAJOR ? e / BK : e % BN;
      int k = B_KMAJOR ? e % BK : e / BN;
      int64_t gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < g.N && gk < k_hi) ? (double)B[gk * g.b_rs + gn * g.b_cs] : 0.0;
    }
  };
  auto store_shared = [&](int buf) {
    double* as = As + buf * BM * LDS;
    double* bs = Bs + buf * BN * LDS;
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT;
      int m = A_KMAJOR ? e / BK : e % BM;
      int k = A_KMAJOR ? e % BK : e / BM;
      as[m * LDS + k] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT;
      int n = B_KMAJOR ? e / BK : e % BN;
      int k = B_KMAJOR ? e % BK : e / BN;
      bs[n * LDS + k] = rb[i];
    }
  };

  const int64_t ktiles = (k_hi - k_lo + BK - 1) / BK;
  load_global(k_lo);
  store_shared(0);
  __syncthreads();

  for (int64_t kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_global(k_lo + (kt + 1) * BK);
    const double* as = As + buf * BM * LDS + (wm * WM + (lane >> 2)) * LDS + (lane & 3);
    const double* bs = Bs + buf * BN * LDS + (wn * WN + (lane >> 2)) * LDS + (lane & 3);
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = as[i * 8 * LDS + k4 * 4];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = bs[j * 8 * LDS + k4 * 4];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt + 1 < ktiles) {
      store_shared(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: one rounding to the output dtype (split-K partials are accumulated with FP64 atomics instead)
  const int64_t cb = bz * g.c_bs;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t row = m0 + wm * WM + i * 8 + (lane >> 2);
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t col = n0 + wn * WN + j * 8 + (lane & 3) * 2;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (col + t < g.N) {
          int64_t off = cb + row * g.c_rs + (col + t) * g.c_cs;
          if (g.splits > 1) atomicAdd((double*)g.C + off, acc[i][j][t]);
          else if (g.c_is_f32) ((float*)g.C)[off] = (float)acc[i][j][t];
          else ((double*)g.C)[off] = acc[i][j][t];
        }
      }
    }
  }
}

template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool AK, bool BKM>
static int launch_gemm(const GemmArgs& g, int64_t batch) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  size_t smem = (size_t)2 * (BM + BN) * LDS * sizeof(double);
  auto kern = gemm_dmma_kernel<TA, TB, BM, BN, WM, WN, AK, BKM>;
  static bool configured = false;
  if (!configured) {
    B200AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + BM - 1) / BM), (unsigned)(batch * g.splits));
  kern<<<grid, NT, smem, stream()>>>(g);
  B200AG_LAUNCH_CHECK();
  return 0;
}

template <typename TA, typename TB, bool AK, bool BKM>
static int pick_tile(const GemmArgs& g, int64_t batch) {
  // Large tiles only when they still give every SM a CTA; otherwise 64x64 tiles for more parallelism.
  int64_t big = ((g.M + 127) / 128) * ((g.N + 127) / 128) * batch;
  if (big >= 2 * (int64_t)sm_count()) return launch_gemm<TA, TB, 128, 128, 64, 32, AK, BKM>(g, batch);
  return launch_gemm<TA, TB, 64, 64, 32, 32, AK, BKM>(g, batch);
}

template <typename TA, typename TB>
static int pick_layout(const GemmArgs& g, int64_t batch) {
  bool ak = g.a_cs == 1 || g.K == 1;   // A contiguous along k
  bool am = g.a_rs == 1 || g.M == 1;   // A contiguous along m
  bool bk = g.b_rs == 1 || g.K == 1;   // B contiguous along k
  bool bn = g.b_cs == 1 || g.N == 1;   // B contiguous along n
  bool a_kmajor = ak || !am;           // fall back to k-major mapping for odd strides (correct, just slower)
  bool b_kmajor = bk && !bn ? true : (bn ? false : true);
  if (a_kmajor && b_kmajor) return pick_tile<TA, TB, true, true>(g, batch);
  if (a_kmajor && !b_kmajor) return pick_tile<TA, TB, true, false>(g, batch);
  if (!a_kmajor && b_kmajor) return pick_tile<TA, TB, false, true>(g, batch);
  return pick_tile<TA, TB, false, false>(g, batch);
}

}  // namespace b200ag

using namespace b200ag;

extern "C" int b200ag_gemm(void* C, int c_dtype, int64_t c_rs, int64_t c_cs, int64_t c_bs, const void* A, int a_dtype,
                           int64_t a_rs, int64_t a_cs, int64_t a_bs, const void* B, int b_dtype, int64_t b_rs,
                           int64_t b_cs, int64_t b_bs, int64_t M, int64_t N, int64_t K, int64_t batch) {
  B200AG_CHECK_INIT();
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (c_dtype != B200AG_F64 && c_dtype != B200AG_F32) B200AG_FAIL("gemm: output dtype must be float64 or float32");
  if (batch > 65535) B200AG_FAIL("gemm: batch too large");
  GemmArgs g{C, A, B, c_rs, c_cs, c_bs, a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, M, N, K, c_dtype == B200AG_F32, 1, K};
  // Split-K: when the output has too few tiles to occupy 148 SMs (weight gradients: small M x N, long K), slices of
  // K go to different CTAs and are combined with FP64 atomics into a zeroed float64 accumulator.
  int64_t tiles = ((M + 63) / 64) * ((N + 63) / 64) * batch;
  int64_t want = (3 * (int64_t)sm_count() + tiles - 1) / tiles;
  int64_t max_by_k = K / 128;
  int64_t splits = want < max_by_k ? want : max_by_k;
  if (splits > 64) splits = 64;
  void* scratch = nullptr;
  void* real_c = C;
  bool dense_c = (c_cs == 1 && c_rs == N && (batch == 1 || c_bs == M * N));
  if (tiles < 2 * (int64_t)sm_count() && splits >= 2 && dense_c && batch * splits <= 65535) {
    int64_t chunk = (K + splits - 1) / splits;
    chunk = (chunk + BK - 1) / BK * BK;
    splits = (K + chunk - 1) / chunk;
    g.splits = (int)splits;
    g.k_chunk = chunk;
    size_t bytes = (size_t)(M * N * batch) * sizeof(double);
    if (g.c_is_f32) {
      if (b200ag_malloc(bytes, &scratch)) return 1;
      g.C = scratch;
      g.c_is_f32 = 0;
    }
    if (b200ag_memset(g.C, 0, bytes)) return 1;
  }
  int rc;
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F64) rc = pick_layout<double, double>(g, batch);
  else if (a_dtype == B200AG_F64 && b_dtype == B200AG_F32) rc = pick_layout<double, float>(g, batch);
  else if (a_dtype == B200AG_F32 && b_dtype == B200AG_F64) rc = pick_layout<float, double>(g, batch);
  else if (a_dtype == B200AG_F32 && b_dtype == B200AG_F32) rc = pick_layout<float, float>(g, batch);
  else B200AG_FAIL("gemm: operand dtypes must be float64 or float32");
  if (scratch) {
    if (rc == 0) {
      int64_t shape[1] = {M * N * batch}, one[1] = {1};
      rc = b200ag_copy_strided(real_c, B200AG_F32, one, scratch, B200AG_F64, one, shape, 1);
    }
    b200ag_free(scratch);
  }
  return rc;
}
