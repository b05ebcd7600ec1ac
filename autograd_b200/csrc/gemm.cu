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
//
// Launch shape: ONE grid covers a GROUP of up to 16 independent products (the four
// gate products of an LSTM cell, examples/lstm.py:38-49, or the eight G·Bᵀ / Aᵀ·G
// adjoints of one time step, numpy_vjps.py:615-638): a linear CTA index is mapped to
// (problem, k-split, tile).  Products that would leave SMs idle on their own fill the
// machine together, without split-K.  Where the whole group still has too few tiles,
// K is split; the partial tiles go to a workspace and the LAST CTA to arrive at a tile
// (a per-tile counter) adds them in split order 0..S-1 and rounds once to the output
// dtype — no atomics on the result, so it is bit-identical from run to run.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace b200ag {

constexpr int BK = 16;
constexpr int MAX_GROUP = 16;
constexpr int MAX_SPLIT_TILES = 8192;   // output tiles (of all problems of one launch) that may be K-split

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
  int splits;        // split-K factor
  int64_t k_chunk;   // K range per split (multiple of BK)
  // grouped launches only
  int tiles_n;       // output tiles along N
  int tiles_mn;      // output tiles of this problem
  int cta_start;     // first linear CTA index of this problem (CTAs = tiles_mn * splits)
  int ctr_start;     // first per-tile arrival counter (split problems)
  double* ws;        // partial tiles [splits][tiles_mn][BM*BN] (split problems)
};

struct GroupArgs {
  GemmArgs g[MAX_GROUP];
  int count;
  int* counters;
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

template <int BM, int BN, int WM, int WN>
struct Frag {
  static constexpr int TM = WM / 8, TN = WN / 8;
  double acc[TM][TN][2];
};

// One BM x BN output tile over k in [k_lo, k_hi): the accumulators stay in registers.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, int64_t m0, int64_t n0, int64_t bz, int64_t k_lo,
                                          int64_t k_hi, double (&acc)[WM / 8][WN / 8][2], double* smem) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LA = BM * BK / NT;
  constexpr int LB = BN * BK / NT;
  constexpr int TM = WM / 8, TN = WN / 8;
  using TileA = Tile<BM, A_KMAJOR>;
  using TileB = Tile<BN, B_KMAJOR>;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
  static_assert(NT % BK == 0 && NT % BM == 0 && NT % BN == 0 || NT >= BM, "loader mapping");

  double* As = smem;                       // [2][TileA::SIZE]
  double* Bs = smem + 2 * TileA::SIZE;     // [2][TileB::SIZE]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);

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
}

// epilogue: one rounding to the output dtype
template <int BM, int BN, int WM, int WN>
__device__ __forceinline__ void store_tile(const GemmArgs& g, int64_t m0, int64_t n0, int64_t bz,
                                           const double (&acc)[WM / 8][WN / 8][2]) {
  constexpr int TM = WM / 8, TN = WN / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);
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
          if (g.c_is_f32) ((float*)g.C)[off] = (float)acc[i][j][t];
          else ((double*)g.C)[off] = acc[i][j][t];
        }
      }
    }
  }
}

// Batched products (np.matmul on stacks): one problem, blockIdx.z = batch index, no split-K.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_dmma_kernel(GemmArgs g) {
  extern __shared__ double smem[];
  double acc[WM / 8][WN / 8][2];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN, bz = blockIdx.z;
  gemm_tile<TA, TB, BM, BN, WM, WN, A_KMAJOR, B_KMAJOR>(g, m0, n0, bz, 0, g.K, acc, smem);
  store_tile<BM, BN, WM, WN>(g, m0, n0, bz, acc);
}

// Grouped launch: CTA -> (problem, k-split, tile); see the header comment.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_grouped_kernel(const __grid_constant__ GroupArgs ga) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int TM = WM / 8, TN = WN / 8;
  extern __shared__ double smem[];
  __shared__ int s_last;
  int p = 0;
#pragma unroll 1
  while (p + 1 < ga.count && (int)blockIdx.x >= ga.g[p + 1].cta_start) ++p;
  const GemmArgs& g = ga.g[p];
  const int local = (int)blockIdx.x - g.cta_start;
  const int kz = local / g.tiles_mn, t = local - kz * g.tiles_mn;
  const int tm = t / g.tiles_n, tn = t - tm * g.tiles_n;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int64_t k_lo = (int64_t)kz * g.k_chunk;
  const int64_t k_hi = (k_lo + g.k_chunk < g.K) ? k_lo + g.k_chunk : g.K;
  double acc[TM][TN][2];
  gemm_tile<TA, TB, BM, BN, WM, WN, A_KMAJOR, B_KMAJOR>(g, m0, n0, 0, k_lo, k_hi, acc, smem);
  if (g.splits > 1) {
    // partial tile -> workspace, element (i, j, t) of thread tid at [(i*TN + j)*2 + t][tid]: coalesced, and the
    // thread that owns an output element in one split owns it in every split
    double* mine = g.ws + ((int64_t)kz * g.tiles_mn + t) * (BM * BN) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        mine[((i * TN + j) * 2 + 0) * NT] = acc[i][j][0];
        mine[((i * TN + j) * 2 + 1) * NT] = acc[i][j][1];
      }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int prev = atomicAdd(ga.counters + g.ctr_start + t, 1);
      s_last = prev == g.splits - 1;
      if (s_last) ga.counters[g.ctr_start + t] = 0;   // ready for the next launch / graph replay
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int s = 0; s < g.splits; ++s) {            // fixed order: the sum does not depend on which CTA came last
      const double* part = g.ws + ((int64_t)s * g.tiles_mn + t) * (BM * BN) + threadIdx.x;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j][0] += __ldcg(part + ((i * TN + j) * 2 + 0) * NT);
          acc[i][j][1] += __ldcg(part + ((i * TN + j) * 2 + 1) * NT);
        }
    }
  }
  store_tile<BM, BN, WM, WN>(g, m0, n0, 0, acc);
}

static int* g_counters = nullptr;

// per-tile arrival counters of the K-split epilogue: allocated and zeroed once, at library initialisation (never inside
// a stream capture); every launch leaves them at zero again
int gemm_init() {
  if (g_counters) return 0;
  B200AG_CUDA(cudaMalloc((void**)&g_counters, MAX_SPLIT_TILES * sizeof(int)));
  B200AG_CUDA(cudaMemset(g_counters, 0, MAX_SPLIT_TILES * sizeof(int)));
  B200AG_CUDA(cudaDeviceSynchronize());
  return 0;
}
static int ensure_counters() { return gemm_init(); }

// development knobs (read once): B200AG_GEMM_TILE = 1 (64x64) / 2 (128x128), B200AG_GEMM_SPLITS = n
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

template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool AK, bool BKM>
static int launch_batched(const GemmArgs& g, int64_t batch) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  size_t smem = (size_t)2 * (Tile<BM, AK>::SIZE + Tile<BN, BKM>::SIZE) * sizeof(double);
  auto kern = gemm_dmma_kernel<TA, TB, BM, BN, WM, WN, AK, BKM>;
  static bool configured = false;
  if (!configured) {
    B200AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + BM - 1) / BM), (unsigned)batch);
  kern<<<grid, NT, smem, stream()>>>(g);
  B200AG_LAUNCH_CHECK();
  return 0;
}

// One launch for a group of problems that share operand types and layouts.  Chooses the tile and the K split.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool AK, bool BKM>
static int launch_group_tiles(GroupArgs& ga) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  const int64_t sms = sm_count();
  int64_t tiles = 0;
  for (int p = 0; p < ga.count; ++p) {
    GemmArgs& g = ga.g[p];
    g.tiles_n = (int)((g.N + BN - 1) / BN);
    g.tiles_mn = (int)(((g.M + BM - 1) / BM) * g.tiles_n);
    tiles += g.tiles_mn;
  }
  // Split K only when the whole group leaves SMs idle: aim for one full wave of resident CTAs (3 per SM for the
  // 64x64 kernel, register-limited; 1 for 128x128).  Tiny outputs (a handful of tiles) are chains of K/16 dependent
  // slabs, pure load latency: K is cut finer there.
  const int64_t slots = (BM == 64 ? 3 : 1) * sms;
  int64_t want = tiles ? slots / tiles : 1;
  int ctr = 0, cta = 0;
  size_t ws_doubles = 0;
  for (int p = 0; p < ga.count; ++p) {
    GemmArgs& g = ga.g[p];
    int64_t max_by_k = (tiles <= sms / 4) ? g.K / 48 : g.K / 128;
    int64_t splits = want < max_by_k ? want : max_by_k;
    if (splits > 64) splits = 64;
    if (tune_splits() > 0) splits = tune_splits() <= g.K / 16 ? tune_splits() : g.K / 16;
    g.splits = 1;
    g.k_chunk = g.K;
    g.ws = nullptr;
    g.ctr_start = 0;
    if (tiles < 2 * sms && splits >= 2 && ctr + g.tiles_mn <= MAX_SPLIT_TILES) {
      int64_t chunk = (g.K + splits - 1) / splits;
      chunk = (chunk + BK - 1) / BK * BK;
      splits = (g.K + chunk - 1) / chunk;
      if (splits >= 2) {
        g.splits = (int)splits;
        g.k_chunk = chunk;
        g.ctr_start = ctr;
        ctr += g.tiles_mn;
        g.ws = (double*)(uintptr_t)(ws_doubles + 1);   // offset + 1 for now (0 = none); rebased below
        ws_doubles += (size_t)splits * g.tiles_mn * BM * BN;
      }
    }
    g.cta_start = cta;
    cta += g.tiles_mn * g.splits;
  }
  void* ws = nullptr;
  if (ws_doubles) {
    if (ensure_counters()) return 1;
    if (b200ag_malloc(ws_doubles * sizeof(double), &ws)) return 1;
    for (int p = 0; p < ga.count; ++p)
      if (ga.g[p].ws) ga.g[p].ws = (double*)ws + ((uintptr_t)ga.g[p].ws - 1);
  }
  ga.counters = g_counters;
  size_t smem = (size_t)2 * (Tile<BM, AK>::SIZE + Tile<BN, BKM>::SIZE) * sizeof(double);
  auto kern = gemm_grouped_kernel<TA, TB, BM, BN, WM, WN, AK, BKM>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      if (ws) b200ag_free(ws);
      B200AG_FAIL(std::string("gemm: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    }
    configured = true;
  }
  kern<<<(unsigned)cta, NT, smem, stream()>>>(ga);
  if (ws) b200ag_free(ws);   // stream-ordered allocator: reused only by work enqueued after this kernel
  B200AG_LAUNCH_CHECK();
  return 0;
}

template <typename TA, typename TB, bool AK, bool BKM>
static int launch_group(GroupArgs& ga) {
  // Large tiles only when they still give (nearly) every SM a CTA and do not pad much more than the small ones
  const int64_t sms = sm_count();
  int64_t big = 0, small = 0, area = 0;
  for (int p = 0; p < ga.count; ++p) {
    const GemmArgs& g = ga.g[p];
    big += ((g.M + 127) / 128) * ((g.N + 127) / 128);
    small += ((g.M + 63) / 64) * ((g.N + 63) / 64);
    area += g.M * g.N;
  }
  const double pad_big = (double)big * 128 * 128 / (double)(area ? area : 1);
  const double pad_small = (double)small * 64 * 64 / (double)(area ? area : 1);
  // time in units of "one SM busy for one 64x64 tile": the 128x128 kernel runs 1 CTA per SM (4 units each), the 64x64
  // kernel 3 CTAs per SM that share its DMMA pipe
  const double t_big = 4.0 * (double)((big + sms - 1) / sms);
  const double waves_small = (double)small / (double)(3 * sms);
  const double t_small = 3.0 * (waves_small < 1.0 ? 1.0 : waves_small + 0.35);
  const int forced = tune_tile();
  bool use_big = forced == 2 || (forced == 0 && big >= sms * 3 / 4 && pad_big <= pad_small * 1.12 && t_big <= t_small * 1.05);
  if (use_big) return launch_group_tiles<TA, TB, 128, 128, 64, 32, AK, BKM>(ga);
  return launch_group_tiles<TA, TB, 64, 64, 32, 32, AK, BKM>(ga);
}

static void layout_of(const GemmArgs& g, bool& a_kmajor, bool& b_kmajor) {
  bool ak = g.a_cs == 1 || g.K == 1;   // A contiguous along k
  bool am = g.a_rs == 1 || g.M == 1;   // A contiguous along m
  bool bk = g.b_rs == 1 || g.K == 1;   // B contiguous along k
  bool bn = g.b_cs == 1 || g.N == 1;   // B contiguous along n
  a_kmajor = ak || !am;                // fall back to k-major mapping for odd strides (correct, just slower)
  b_kmajor = bk && !bn ? true : (bn ? false : true);
}

template <typename TA, typename TB>
static int launch_group_layout(GroupArgs& ga, bool a_kmajor, bool b_kmajor) {
  if (a_kmajor && b_kmajor) return launch_group<TA, TB, true, true>(ga);
  if (a_kmajor && !b_kmajor) return launch_group<TA, TB, true, false>(ga);
  if (!a_kmajor && b_kmajor) return launch_group<TA, TB, false, true>(ga);
  return launch_group<TA, TB, false, false>(ga);
}

static int launch_group_typed(GroupArgs& ga, int a_dtype, int b_dtype, bool a_kmajor, bool b_kmajor) {
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F64) return launch_group_layout<double, double>(ga, a_kmajor, b_kmajor);
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F32) return launch_group_layout<double, float>(ga, a_kmajor, b_kmajor);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F64) return launch_group_layout<float, double>(ga, a_kmajor, b_kmajor);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F32) return launch_group_layout<float, float>(ga, a_kmajor, b_kmajor);
  B200AG_FAIL("gemm: operand dtypes must be float64 or float32");
}

template <typename TA, typename TB>
static int batched_layout(const GemmArgs& g, int64_t batch) {
  bool ak, bk;
  layout_of(g, ak, bk);
  if (ak && bk) return launch_batched<TA, TB, 64, 64, 32, 32, true, true>(g, batch);
  if (ak && !bk) return launch_batched<TA, TB, 64, 64, 32, 32, true, false>(g, batch);
  if (!ak && bk) return launch_batched<TA, TB, 64, 64, 32, 32, false, true>(g, batch);
  return launch_batched<TA, TB, 64, 64, 32, 32, false, false>(g, batch);
}

}  // namespace b200ag

using namespace b200ag;

extern "C" int b200ag_gemm_grouped(const b200ag_gemm_problem* problems, int count) {
  B200AG_CHECK_INIT();
  if (count < 0) B200AG_FAIL("gemm_grouped: negative count");
  std::vector<char> done((size_t)count, 0);
  for (int i = 0; i < count; ++i) {
    const b200ag_gemm_problem& q = problems[i];
    if (done[i]) continue;
    if (q.M == 0 || q.N == 0) { done[i] = 1; continue; }
    if (q.c_dtype != B200AG_F64 && q.c_dtype != B200AG_F32) B200AG_FAIL("gemm: output dtype must be float64 or float32");
    // large float32 x float32 -> float32 products run on tcgen05 tensor cores (3xTF32, gemm_tc.cu), one by one
    if (q.a_dtype == B200AG_F32 && q.b_dtype == B200AG_F32 && q.c_dtype == B200AG_F32 &&
        f32_tc_eligible(q.M, q.N, q.K, q.c_cs, 1)) {
      if (b200ag_gemm_f32_tc(q.C, q.c_rs, q.A, q.a_rs, q.a_cs, q.B, q.b_rs, q.b_cs, q.M, q.N, q.K)) return 1;
      done[i] = 1;
      continue;
    }
    // gather every remaining problem with the same operand types and layouts into launches of up to MAX_GROUP
    GroupArgs ga;
    ga.count = 0;
    bool ak0 = false, bk0 = false;
    for (int j = i; j < count; ++j) {
      const b200ag_gemm_problem& r = problems[j];
      if (done[j] || r.M == 0 || r.N == 0) continue;
      if (r.a_dtype != q.a_dtype || r.b_dtype != q.b_dtype) continue;
      if (r.a_dtype == B200AG_F32 && r.b_dtype == B200AG_F32 && r.c_dtype == B200AG_F32 &&
          f32_tc_eligible(r.M, r.N, r.K, r.c_cs, 1))
        continue;
      GemmArgs g{r.C, r.A, r.B, r.c_rs, r.c_cs, 0, r.a_rs, r.a_cs, 0, r.b_rs, r.b_cs, 0, r.M, r.N, r.K,
                 r.c_dtype == B200AG_F32, 1, r.K, 0, 0, 0, 0, nullptr};
      bool ak, bk;
      layout_of(g, ak, bk);
      if (ga.count == 0) { ak0 = ak; bk0 = bk; }
      else if (ak != ak0 || bk != bk0) continue;
      if (r.c_dtype != B200AG_F64 && r.c_dtype != B200AG_F32) B200AG_FAIL("gemm: output dtype must be float64 or float32");
      ga.g[ga.count++] = g;
      done[j] = 1;
      if (ga.count == MAX_GROUP) {
        if (launch_group_typed(ga, q.a_dtype, q.b_dtype, ak0, bk0)) return 1;
        ga.count = 0;
      }
    }
    if (ga.count && launch_group_typed(ga, q.a_dtype, q.b_dtype, ak0, bk0)) return 1;
  }
  return 0;
}

extern "C" int b200ag_gemm(void* C, int c_dtype, int64_t c_rs, int64_t c_cs, int64_t c_bs, const void* A, int a_dtype,
                           int64_t a_rs, int64_t a_cs, int64_t a_bs, const void* B, int b_dtype, int64_t b_rs,
                           int64_t b_cs, int64_t b_bs, int64_t M, int64_t N, int64_t K, int64_t batch) {
  B200AG_CHECK_INIT();
  if (M == 0 || N == 0 || batch == 0) return 0;
  if (c_dtype != B200AG_F64 && c_dtype != B200AG_F32) B200AG_FAIL("gemm: output dtype must be float64 or float32");
  if (batch == 1) {
    b200ag_gemm_problem q{C, A, B, c_dtype, a_dtype, b_dtype, 0, c_rs, c_cs, a_rs, a_cs, b_rs, b_cs, M, N, K};
    return b200ag_gemm_grouped(&q, 1);
  }
  if (batch > 65535) B200AG_FAIL("gemm: batch too large");
  GemmArgs g{C, A, B, c_rs, c_cs, c_bs, a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, M, N, K, c_dtype == B200AG_F32, 1, K,
             0, 0, 0, 0, nullptr};
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F64) return batched_layout<double, double>(g, batch);
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F32) return batched_layout<double, float>(g, batch);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F64) return batched_layout<float, double>(g, batch);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F32) return batched_layout<float, float>(g, batch);
  B200AG_FAIL("gemm: operand dtypes must be float64 or float32");
}
