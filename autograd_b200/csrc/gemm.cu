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
#include <stdlib.h>

#include "common.cuh"

namespace b200ag {

constexpr int BK = 16;

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

// Shared-memory layouts (doubles).  An operand that is contiguous along k in global memory is staged as
// [row][k] with pitch BK+4; one that is contiguous along m/n is staged as [k][row] with pitch ROWS+4.  In both,
// consecutive threads of the loader write consecutive addresses, and the 64-bit fragment reads of a half-warp
// (row = lane/4 in 0..3, k = lane%4) hit 16 distinct bank pairs because the pitch is 4 mod 16.
template <int ROWS, bool KMAJOR>
struct Tile {
  static constexpr int PITCH = KMAJOR ? BK + 4 : ROWS + 4;
  static constexpr int SIZE = KMAJOR ? ROWS * PITCH : BK * PITCH;
  __device__ static int at(int row, int k) { return KMAJOR ? row * PITCH + k : k * PITCH + row; }
};

template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_dmma_kernel(GemmArgs g) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LA = BM * BK / NT;
  constexpr int LB = BN * BK / NT;
  constexpr int TM = WM / 8, TN = WN / 8;
  using TileA = Tile<BM, A_KMAJOR>;
  using TileB = Tile<BN, B_KMAJOR>;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
  static_assert(NT % BK == 0 && NT % BM == 0 && NT % BN == 0 || NT >= BM, "loader mapping");

  extern __shared__ double smem[];
  double* As = smem;                       // [2][TileA::SIZE]
  double* Bs = smem + 2 * TileA::SIZE;     // [2][TileB::SIZE]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t bz = blockIdx.z / g.splits, kz = blockIdx.z % g.splits;
  const int64_t k_lo = kz * g.k_chunk;
  const int64_t k_hi = (k_lo + g.k_chunk < g.K) ? k_lo + g.k_chunk : g.K;

  // ---- loader mapping: element e = tid + i*NT of a ROWS x BK slab.
  //   k-major:  row = e / BK, k = e % BK   -> per thread: k fixed, row = row0 + i * (NT / BK)
  //   row-major: row = e % ROWS, k = e / ROWS -> per thread: row fixed, k = k0 + i * (NT / ROWS)
  constexpr int A_RSTEP = A_KMAJOR ? NT / BK : 0, A_KSTEP = A_KMAJOR ? 0 : NT / BM;
  constexpr int B_RSTEP = B_KMAJOR ? NT / BK : 0, B_KSTEP = B_KMAJOR ? 0 : NT / BN;
  const int a_row0 = A_KMAJOR ? tid / BK : tid % BM, a_k0 = A_KMAJOR ? tid % BK : tid / BM;
  const int b_row0 = B_KMAJOR ? tid / BK : tid % BN, b_k0 = B_KMAJOR ? tid % BK : tid / BN;
  const TA* pa = (const TA*)g.A + bz * g.a_bs + (m0 + a_row0) * g.a_rs + (k_lo + a_k0) * g.a_cs;
  const TB* pb = (const TB*)g.B + bz * g.b_bs + (n0 + b_row0) * g.b_cs + (k_lo + b_k0) * g.b_rs;
  const int64_t a_istep = (int64_t)A_RSTEP * g.a_rs + (int64_t)A_KSTEP * g.a_cs;
  const int64_t b_istep = (int64_t)B_RSTEP * g.b_cs + (int64_t)B_KSTEP * g.b_rs;
  const int64_t a_kadv = (int64_t)BK * g.a_cs, b_kadv = (int64_t)BK * g.b_rs;
  unsigned a_rowok = 0, b_rowok = 0;  // bit i: row of element i is inside the matrix
#pragma unroll
  for (int i = 0; i < LA; ++i) a_rowok |= (m0 + a_row0 + i * A_RSTEP < g.M) ? (1u << i) : 0u;
#pragma unroll
  for (int i = 0; i < LB; ++i) b_rowok |= (n0 + b_row0 + i * B_RSTEP < g.N) ? (1u << i) : 0u;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // staging registers keep the operand's own type: the f32 -> f64 widening happens when the slab is written to
  // shared memory, i.e. AFTER the multiply of the previous slab, so the global loads stay in flight behind it
  TA ra[LA];
  TB rb[LB];

  auto load_global = [&](int64_t kbase) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      bool ok = ((a_rowok >> i) & 1u) && (kbase + a_k0 + i * A_KSTEP < k_hi);
      ra[i] = ok ? pa[i * a_istep] : (TA)0;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      bool ok = ((b_rowok >> i) & 1u) && (kbase + b_k0 + i * B_KSTEP < k_hi);
      rb[i] = ok ? pb[i * b_istep] : (TB)0;
    }
    pa += a_kadv;
    pb += b_kadv;
  };
  auto store_shared = [&](int buf) {
    double* as = As + buf * TileA::SIZE;
    double* bs = Bs + buf * TileB::SIZE;
#pragma unroll
    for (int i = 0; i < LA; ++i) as[TileA::at(a_row0 + i * A_RSTEP, a_k0 + i * A_KSTEP)] = (double)ra[i];
#pragma unroll
    for (int i = 0; i < LB; ++i) bs[TileB::at(b_row0 + i * B_RSTEP, b_k0 + i * B_KSTEP)] = (double)rb[i];
  };

  const int64_t ktiles = (k_hi - k_lo + BK - 1) / BK;
  load_global(k_lo);
  store_shared(0);
  __syncthreads();

  const int a_frag = TileA::at(wm * WM + (lane >> 2), lane & 3);
  const int b_frag = TileB::at(wn * WN + (lane >> 2), lane & 3);
  constexpr int A_ROW8 = A_KMAJOR ? 8 * TileA::PITCH : 8, A_K4 = A_KMAJOR ? 4 : 4 * TileA::PITCH;
  constexpr int B_ROW8 = B_KMAJOR ? 8 * TileB::PITCH : 8, B_K4 = B_KMAJOR ? 4 : 4 * TileB::PITCH;

  for (int64_t kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_global(k_lo + (kt + 1) * BK);
    const double* as = As + buf * TileA::SIZE + a_frag;
    const double* bs = Bs + buf * TileB::SIZE + b_frag;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = as[i * A_ROW8 + k4 * A_K4];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = bs[j * B_ROW8 + k4 * B_K4];
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
  size_t smem = (size_t)2 * (Tile<BM, AK>::SIZE + Tile<BN, BKM>::SIZE) * sizeof(double);
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

// development knobs (read once): B200AG_GEMM_TILE = 1 (64x64) / 2 (128x128) / 3 (128x64), B200AG_GEMM_SPLITS = n
static int tune_tile() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200AG_GEMM_TILE"); v = e ? atoi(e) : 0; }
  return v;
}
static int tune_splits() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200AG_GEMM_SPLITS"); v = e ? atoi(e) : 0; }
  return v;
}

template <typename TA, typename TB, bool AK, bool BKM>
static int pick_tile(const GemmArgs& g, int64_t batch) {
  // Large tiles only when they still give every SM a CTA; otherwise 64x64 tiles for more parallelism.
  int64_t big = ((g.M + 127) / 128) * ((g.N + 127) / 128) * batch;
  const int forced = tune_tile();
  if (forced == 2 || (forced == 0 && big >= 2 * (int64_t)sm_count())) return launch_gemm<TA, TB, 128, 128, 64, 32, AK, BKM>(g, batch);
  if (forced == 3) return launch_gemm<TA, TB, 128, 64, 32, 32, AK, BKM>(g, batch);
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
  // large float32 x float32 -> float32 products run on tcgen05 tensor cores (3xTF32, gemm_tc.cu)
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F32 && c_dtype == B200AG_F32 && f32_tc_eligible(M, N, K, c_cs, batch))
    return b200ag_gemm_f32_tc(C, c_rs, A, a_rs, a_cs, B, b_rs, b_cs, M, N, K);
  GemmArgs g{C, A, B, c_rs, c_cs, c_bs, a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, M, N, K, c_dtype == B200AG_F32, 1, K};
  // Split-K: when the output has too few tiles to occupy 148 SMs (weight gradients: small M x N, long K), slices of
  // K go to different CTAs and are combined with FP64 atomics into a zeroed float64 accumulator.
  int64_t tiles = ((M + 63) / 64) * ((N + 63) / 64) * batch;
  // aim for ONE full wave: 3 CTAs of the 64x64 kernel are resident per SM (register-limited)
  int64_t slots = 3 * (int64_t)sm_count();
  int64_t want = slots / tiles;
  // tiny outputs (a handful of tiles) are pure load-latency chains of K/16 dependent slabs: cut K finer there
  int64_t max_by_k = (tiles <= (int64_t)sm_count() / 4) ? K / 48 : K / 128;
  int64_t splits = want < max_by_k ? want : max_by_k;
  if (splits > 64) splits = 64;
  if (tune_splits() > 0) splits = tune_splits() <= K / 16 ? tune_splits() : K / 16;
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
