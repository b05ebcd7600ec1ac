// FP64-accumulating GEMM on the DMMA tensor pipe (mma.sync.m8n8k4.f64) for
// np.dot / np.matmul / np.tensordot and their adjoints (numpy_vjps.py:567-742).
// tcgen05 has no f64 kind, so the legacy warp-level DMMA is the tensor path for
// autograd's default dtype.  Operands are addressed through (row, col, batch)
// element strides: G·Bᵀ and Aᵀ·G read the swapped views directly, no transpose
// copy.  f32 operands are widened on load (NumPy promotion f32·f64 → f64) and
// the accumulator is rounded once to the output dtype on store.
//
// Tiling: CTA tile 64×64×16, operand slabs staged in shared memory in their own dtype
// ([row][k] with a 4-element pad: conflict-free fragment loads) through a cp.async
// ring (the next k-slab is in flight while the current one is multiplied).
//
// Launch shape: ONE grid covers a GROUP of up to 16 independent products (the four
// gate products of an LSTM cell, examples/lstm.py:38-49, or the eight G·Bᵀ / Aᵀ·G
// adjoints of one time step, numpy_vjps.py:615-638): a linear CTA index is mapped to
// (problem, k-split, tile).  Products that would leave SMs idle on their own fill the
// machine together.  Where the tiles of the whole group do not balance over the SMs
// (too few, or a ragged last wave), K is split S ways and the S CTAs of a tile form a
// THREAD-BLOCK CLUSTER: every CTA parks its partial tile in its own shared memory, and
// after one cluster barrier each CTA adds 1/S of the tile over all S partials through
// distributed shared memory, in split order 0..S-1, and rounds once to the output dtype.
// No workspace in HBM, no atomics, no counters: the result is bit-identical from run to
// run and the reduction is spread over the S SMs instead of serialised on one.
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace b200ag {

constexpr int BK = 16;
constexpr int MAX_GROUP = 16;
constexpr int MAX_SPLITS = 8;            // = portable cluster size
constexpr int MAX_WS_SPLITS = 64;        // deepest split through the HBM workspace

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
  int cta_start;     // first linear CTA index of this problem (CTAs = tiles_mn * splits, a tile's splits adjacent)
};

struct GroupArgs {
  GemmArgs g[MAX_GROUP];
  int count;
  int splits;        // split-K factor of the launch (1: plain launch; 2..MAX_SPLITS: one cluster per tile)
  double* ws;        // splits > MAX_SPLITS only: partial tiles [tile][split][BM*BN] in HBM, added up by
                     // gemm_splitk_reduce_kernel (huge-K, tiny-output products: FC weight gradients over a 65536 batch)
};

// Shared-memory layouts, in elements of the operand's OWN type (float32 operands are staged as float32 and widened
// when a fragment is read: NumPy promotion f32·f64 -> f64 costs no extra pass and half the staging bytes).  An operand
// that is contiguous along k in global memory is staged as [row][k] with pitch BK+4; one that is contiguous along m/n
// as [k][row] with pitch ROWS+4.  In both, consecutive threads of the loader write consecutive addresses, and the
// fragment reads of a warp (row = lane/4, k = lane%4) are conflict-free because the pitch is 4 mod 16.
template <int ROWS, bool KMAJOR>
struct Tile {
  static constexpr int PITCH = KMAJOR ? BK + 4 : ROWS + 4;
  static constexpr int SIZE = KMAJOR ? ROWS * PITCH : BK * PITCH;
  __device__ static int at(int row, int k) { return KMAJOR ? row * PITCH + k : k * PITCH + row; }
};

#ifndef B200AG_GEMM_STAGES
#define B200AG_GEMM_STAGES 2
#endif
// k-slabs per CTA in the cp.async ring.  Measured on the B200 (scripts/gpu_r2_stage_ab.sh): 2 stages (40 KB of shared
// memory, more CTAs per SM) beat 3 stages (61 KB) on the LSTM config (104.6 vs 107.4 ms) and tie on C1 / C5.
constexpr int NSTAGE = B200AG_GEMM_STAGES;

template <typename TA, typename TB, int BM, int BN, bool AK, bool BKM>
__host__ __device__ constexpr size_t stage_bytes() {
  return ((size_t)Tile<BM, AK>::SIZE * sizeof(TA) + (size_t)Tile<BN, BKM>::SIZE * sizeof(TB) + 15) / 16 * 16;
}

// One element global -> shared, asynchronously; `ok == false` writes a zero without touching global memory.
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* dst, const T* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = ok ? (int)sizeof(T) : 0;
  if constexpr (sizeof(T) == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One BM x BN output tile over k in [k_lo, k_hi): the accumulators stay in registers.  The operand slabs travel
// global -> shared through a ring of NSTAGE cp.async stages: while slab kt is multiplied the following slab(s) are in
// flight, and nothing is staged through registers (30 registers per thread less than a register-staged prefetch, so
// more CTAs fit an SM).
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, int64_t m0, int64_t n0, int64_t bz, int64_t k_lo,
                                          int64_t k_hi, double (&acc)[WM / 8][WN / 8][2], unsigned char* smem) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LA = BM * BK / NT;
  constexpr int LB = BN * BK / NT;
  constexpr int TM = WM / 8, TN = WN / 8;
  using TileA = Tile<BM, A_KMAJOR>;
  using TileB = Tile<BN, B_KMAJOR>;
  constexpr size_t STAGE = stage_bytes<TA, TB, BM, BN, A_KMAJOR, B_KMAJOR>();
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
  static_assert(NT % BK == 0 && NT % BM == 0 && NT % BN == 0 || NT >= BM, "loader mapping");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);

  // ---- loader mapping: element e = tid + i*NT of a ROWS x BK slab.
  //   k-major:  row = e / BK, k = e % BK   -> per thread: k fixed, row = row0 + i * (NT / BK)
  //   row-major: row = e % ROWS, k = e / ROWS -> per thread: row fixed, k = k0 + i * (NT / ROWS)
  constexpr int A_RSTEP = A_KMAJOR ? NT / BK : 0, A_KSTEP = A_KMAJOR ? 0 : NT / BM;
  constexpr int B_RSTEP = B_KMAJOR ? NT / BK : 0, B_KSTEP = B_KMAJOR ? 0 : NT / BN;
  const int a_row0 = A_KMAJOR ? tid / BK : tid % BM, a_k0 = A_KMAJOR ? tid % BK : tid / BM;
  const int b_row0 = B_KMAJOR ? tid / BK : tid % BN, b_k0 = B_KMAJOR ? tid % BK : tid / BN;
  const TA* const a_base = (const TA*)g.A;
  const TB* const b_base = (const TB*)g.B;
  const TA* pa = a_base + bz * g.a_bs + (m0 + a_row0) * g.a_rs + (k_lo + a_k0) * g.a_cs;
  const TB* pb = b_base + bz * g.b_bs + (n0 + b_row0) * g.b_cs + (k_lo + b_k0) * g.b_rs;
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

  // slab `kt` (k = k_lo + kt*BK ..) -> ring stage `stage`
  auto issue = [&](int stage, int64_t kbase) {
    TA* as = reinterpret_cast<TA*>(smem + (size_t)stage * STAGE);
    TB* bs = reinterpret_cast<TB*>(smem + (size_t)stage * STAGE + (size_t)TileA::SIZE * sizeof(TA));
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const bool ok = ((a_rowok >> i) & 1u) && (kbase + a_k0 + i * A_KSTEP < k_hi);
      cp_async_elem(as + TileA::at(a_row0 + i * A_RSTEP, a_k0 + i * A_KSTEP), ok ? pa + i * a_istep : a_base, ok);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const bool ok = ((b_rowok >> i) & 1u) && (kbase + b_k0 + i * B_KSTEP < k_hi);
      cp_async_elem(bs + TileB::at(b_row0 + i * B_RSTEP, b_k0 + i * B_KSTEP), ok ? pb + i * b_istep : b_base, ok);
    }
    pa += a_kadv;
    pb += b_kadv;
  };

  const int64_t ktiles = (k_hi - k_lo + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < ktiles) issue(s, k_lo + (int64_t)s * BK);
    cp_async_commit();                     // one group per slab, empty ones included: the wait counts stay uniform
  }

  const int a_frag = TileA::at(wm * WM + (lane >> 2), lane & 3);
  const int b_frag = TileB::at(wn * WN + (lane >> 2), lane & 3);
  constexpr int A_ROW8 = A_KMAJOR ? 8 * TileA::PITCH : 8, A_K4 = A_KMAJOR ? 4 : 4 * TileA::PITCH;
  constexpr int B_ROW8 = B_KMAJOR ? 8 * TileB::PITCH : 8, B_K4 = B_KMAJOR ? 4 : 4 * TileB::PITCH;

  int stage = 0;
  for (int64_t kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<NSTAGE - 2>();           // this thread's copies of slab kt have landed
    __syncthreads();                       // ... and everyone's; everyone is also done reading the stage refilled below
    {
      const int64_t nxt = kt + NSTAGE - 1;
      int nstage = stage + NSTAGE - 1;
      if (nstage >= NSTAGE) nstage -= NSTAGE;
      if (nxt < ktiles) issue(nstage, k_lo + nxt * BK);
      cp_async_commit();
    }
    const TA* as = reinterpret_cast<const TA*>(smem + (size_t)stage * STAGE) + a_frag;
    const TB* bs = reinterpret_cast<const TB*>(smem + (size_t)stage * STAGE + (size_t)TileA::SIZE * sizeof(TA)) + b_frag;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = (double)as[i * A_ROW8 + k4 * A_K4];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = (double)bs[j * B_ROW8 + k4 * B_K4];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (++stage == NSTAGE) stage = 0;
  }
  cp_async_wait<0>();
  __syncthreads();                         // the ring may be reused by the caller (split-K partial tile)
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
  extern __shared__ __align__(16) unsigned char smem[];
  double acc[WM / 8][WN / 8][2];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN, bz = blockIdx.z;
  gemm_tile<TA, TB, BM, BN, WM, WN, A_KMAJOR, B_KMAJOR>(g, m0, n0, bz, 0, g.K, acc, smem);
  store_tile<BM, BN, WM, WN>(g, m0, n0, bz, acc);
}

// Grouped launch: CTA -> (problem, tile, k-split); see the header comment.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_grouped_kernel(const __grid_constant__ GroupArgs ga) {
  namespace cg = cooperative_groups;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int TM = WM / 8, TN = WN / 8;
  extern __shared__ __align__(16) unsigned char smem[];
  double* const tile_smem = reinterpret_cast<double*>(smem);
  int p = 0;
#pragma unroll 1
  while (p + 1 < ga.count && (int)blockIdx.x >= ga.g[p + 1].cta_start) ++p;
  const GemmArgs& g = ga.g[p];
  const int S = ga.splits;
  const int local = (int)blockIdx.x - g.cta_start;
  const int t = local / S, kz = local - t * S;          // the S splits of a tile are adjacent: one cluster
  const int tm = t / g.tiles_n, tn = t - tm * g.tiles_n;
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
  const int64_t k_lo = (int64_t)kz * g.k_chunk;
  int64_t k_hi = k_lo + g.k_chunk;
  if (k_hi > g.K) k_hi = g.K;
  if (k_hi < k_lo) k_hi = k_lo;                         // a split past the end of K contributes zeros
  double acc[TM][TN][2];
  gemm_tile<TA, TB, BM, BN, WM, WN, A_KMAJOR, B_KMAJOR>(g, m0, n0, 0, k_lo, k_hi, acc, smem);
  if (S == 1) {
    store_tile<BM, BN, WM, WN>(g, m0, n0, 0, acc);
    return;
  }
  if (ga.ws) {
    // deep split (more ways than a cluster holds): partial tile -> HBM workspace, summed by gemm_splitk_reduce_kernel
    double* mine = ga.ws + ((int64_t)(g.cta_start + local)) * (BM * BN) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        mine[((i * TN + j) * 2 + 0) * NT] = acc[i][j][0];
        mine[((i * TN + j) * 2 + 1) * NT] = acc[i][j][1];
      }
    return;
  }
  // ---- split-K: partial tile -> own shared memory, element (i, j, t) of thread tid at [(i*TN + j)*2 + t][tid]
  // (conflict-free; the operand slabs are dead by now), one cluster barrier, then CTA r of the cluster owns the slots
  // f = r*chunk .. of the BM*BN-element tile: it adds them over the S partials in split order through distributed
  // shared memory and stores the rounded result.
  cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      tile_smem[((i * TN + j) * 2 + 0) * NT + threadIdx.x] = acc[i][j][0];
      tile_smem[((i * TN + j) * 2 + 1) * NT + threadIdx.x] = acc[i][j][1];
    }
  cluster.sync();
  const int r = (int)cluster.block_rank();
  constexpr int TILE = BM * BN;
  const int chunk = (TILE / NT + S - 1) / S * NT;         // slots per CTA, a multiple of the thread count
  const int f_hi = (r + 1) * chunk < TILE ? (r + 1) * chunk : TILE;
  for (int f = r * chunk + threadIdx.x; f < f_hi; f += NT) {
    double sum = 0.0;
    for (int sidx = 0; sidx < S; ++sidx) sum += *cluster.map_shared_rank(tile_smem + f, sidx);   // fixed order
    const int tid = f % NT, w = f / NT;
    const int tt = w & 1, j = (w >> 1) % TN, i = (w >> 1) / TN;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / (BN / WN), wn = warp % (BN / WN);
    const int64_t row = m0 + wm * WM + i * 8 + (lane >> 2);
    const int64_t col = n0 + wn * WN + j * 8 + (lane & 3) * 2 + tt;
    if (row < g.M && col < g.N) {
      const int64_t off = row * g.c_rs + col * g.c_cs;
      if (g.c_is_f32) ((float*)g.C)[off] = (float)sum;
      else ((double*)g.C)[off] = sum;
    }
  }
  cluster.sync();   // nobody leaves while its shared memory may still be read
}

// Second pass of a deep split: CTA (tile, w) adds slot w of the tile's S partial tiles in split order - one output
// element per thread, its S loads independent of each other - and rounds once to the output dtype.  TM*TN*2 CTAs per
// tile: the pass is spread over the machine instead of serialised on a handful of CTAs.
template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) gemm_splitk_reduce_kernel(const __grid_constant__ GroupArgs ga) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int TM = WM / 8, TN = WN / 8;
  constexpr int SLOTS = TM * TN * 2;
  const int S = ga.splits;
  const int tile_global = (int)blockIdx.x / SLOTS, w = (int)blockIdx.x - tile_global * SLOTS;
  int p = 0;
#pragma unroll 1
  while (p + 1 < ga.count && tile_global * S >= ga.g[p + 1].cta_start) ++p;
  const GemmArgs& g = ga.g[p];
  const int t = tile_global - g.cta_start / S;
  const int tm = t / g.tiles_n, tn = t - tm * g.tiles_n;
  const double* part = ga.ws + ((int64_t)g.cta_start + (int64_t)t * S) * (BM * BN) + (int64_t)w * NT + threadIdx.x;
  double sum = 0.0;
  int sidx = 0;
  for (; sidx + 8 <= S; sidx += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (int64_t)(sidx + u) * (BM * BN));
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += v[u];
  }
  for (; sidx < S; ++sidx) sum += __ldcg(part + (int64_t)sidx * (BM * BN));
  const int tt = w & 1, j = (w >> 1) % TN, i = (w >> 1) / TN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / (BN / WN), wn = warp % (BN / WN);
  const int64_t row = (int64_t)tm * BM + wm * WM + i * 8 + (lane >> 2);
  const int64_t col = (int64_t)tn * BN + wn * WN + j * 8 + (lane & 3) * 2 + tt;
  if (row < g.M && col < g.N) {
    const int64_t off = row * g.c_rs + col * g.c_cs;
    if (g.c_is_f32) ((float*)g.C)[off] = (float)sum;
    else ((double*)g.C)[off] = sum;
  }
}

int gemm_init() { return 0; }

// development knob: B200AG_GEMM_SPLITS = n (environment, read once), or b200ag_gemm_tune at run time
// (scripts/gemm_sweep.py measures the split choice with it)
static int g_tune_splits = -1;
static int tune_splits() {
  if (g_tune_splits < 0) { const char* e = getenv("B200AG_GEMM_SPLITS"); g_tune_splits = e ? atoi(e) : 0; }
  return g_tune_splits;
}

template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool AK, bool BKM>
static int launch_batched(const GemmArgs& g, int64_t batch) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  size_t smem = NSTAGE * stage_bytes<TA, TB, BM, BN, AK, BKM>();
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

// Split factor S of a launch, from measurements on a B200 (scripts/gemm_sweep.py, profiles/r02_gemm_sweep.txt; every
// shape class of the BASELINE configs, forced S = 1..64 inside a replayed CUDA graph):
//  * a CTA walks its K range one 16-wide slab per global-load round trip (~0.8 us when nothing else runs on the SM), so
//    small outputs are latency chains: 256x200x784 takes 52 us unsplit and 14 us with a cluster of 8; the best S keeps
//    ~2.6 CTAs per SM in flight as long as every split still has >= 24 columns of K;
//  * from one wave of tiles upwards only the balance of the last wave matters: all CTAs of an SM share its DMMA pipe, so
//    the launch lasts as long as the busiest SM; S in {1, 2, 3} is chosen to even that out (4 x 641x512x1024: 128 -> 111 us
//    at S = 2), each extra split costing ~4 %;
//  * huge-K, few-tile products (FC weight gradients over a 65536 batch: 256x120x65536) want far more than the 8 CTAs a
//    cluster holds: S up to 64 through the HBM workspace and a second, tiny reduction launch (3.8 ms -> 0.25 ms).
static int plan_split(int64_t tiles, int64_t min_k, int64_t sms) {
  if (tiles <= 0) return 1;
  const int64_t by_k = min_k / 24;
  if (tiles >= sms) {
    if (tiles >= 4 * sms) return 1;
    int best = 1;
    double best_t = 1e30;
    for (int S = 1; S <= 3 && (S == 1 || S <= min_k / 128); ++S) {
      const double t = (double)((tiles * S + sms - 1) / sms) / (double)S * (1.0 + 0.04 * (S - 1));
      if (t < best_t - 1e-9) { best_t = t; best = S; }
    }
    return best;
  }
  if (tiles * MAX_SPLITS < sms && min_k >= 8192) {           // deep split through the workspace
    int64_t S = (2 * sms + tiles / 2) / tiles;
    if (S > MAX_WS_SPLITS) S = MAX_WS_SPLITS;
    if (S > min_k / 256) S = min_k / 256;
    return (int)(S < 1 ? 1 : S);
  }
  int64_t S = (385 * sms / 148 + tiles / 2) / tiles;           // ~2.6 CTAs per SM
  if (S > MAX_SPLITS) S = MAX_SPLITS;
  if (S > by_k) S = by_k;
  return (int)(S < 1 ? 1 : S);
}

// One launch for a group of problems that share operand types and layouts.  Chooses the K split.
template <typename TA, typename TB, int BM, int BN, int WM, int WN, bool AK, bool BKM>
static int launch_group_tiles(GroupArgs& ga) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  const int64_t sms = sm_count();
  int64_t tiles = 0, min_k = INT64_MAX;
  for (int p = 0; p < ga.count; ++p) {
    GemmArgs& g = ga.g[p];
    g.tiles_n = (int)((g.N + BN - 1) / BN);
    g.tiles_mn = (int)(((g.M + BM - 1) / BM) * g.tiles_n);
    tiles += g.tiles_mn;
    if (g.K < min_k) min_k = g.K;
  }
  int best = plan_split(tiles, min_k, sms);
  if (tune_splits() > 0) best = tune_splits() <= MAX_WS_SPLITS ? tune_splits() : MAX_WS_SPLITS;
  for (int p = 0; p < ga.count; ++p)
    if (best > 1 && ga.g[p].K / best < BK) best = (int)(ga.g[p].K / BK > 1 ? ga.g[p].K / BK : 1);
  ga.splits = best;
  ga.ws = nullptr;
  int cta = 0;
  for (int p = 0; p < ga.count; ++p) {
    GemmArgs& g = ga.g[p];
    int64_t chunk = (g.K + best - 1) / best;
    chunk = (chunk + BK - 1) / BK * BK;
    g.splits = best;
    g.k_chunk = best > 1 ? chunk : g.K;
    g.cta_start = cta;
    cta += g.tiles_mn * best;
  }
  size_t smem = NSTAGE * stage_bytes<TA, TB, BM, BN, AK, BKM>();
  if (best > 1 && best <= MAX_SPLITS && smem < (size_t)BM * BN * sizeof(double)) smem = (size_t)BM * BN * sizeof(double);
  auto kern = gemm_grouped_kernel<TA, TB, BM, BN, WM, WN, AK, BKM>;
  static size_t configured = 0;
  if (smem > configured) {
    B200AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (best > MAX_SPLITS) {
    void* ws = nullptr;
    if (b200ag_malloc((size_t)cta * BM * BN * sizeof(double), &ws)) return 1;
    ga.ws = (double*)ws;
    kern<<<(unsigned)cta, NT, smem, stream()>>>(ga);
    gemm_splitk_reduce_kernel<BM, BN, WM, WN><<<(unsigned)(cta / best) * (WM / 8) * (WN / 8) * 2, NT, 0, stream()>>>(ga);
    b200ag_free(ws);   // stream-ordered allocator: reused only by work enqueued after these kernels
    count_launch();
  } else if (best == 1) {
    kern<<<(unsigned)cta, NT, smem, stream()>>>(ga);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cta);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)best;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200AG_CUDA(cudaLaunchKernelEx(&cfg, kern, ga));
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

template <typename TA, typename TB, bool AK, bool BKM>
static int launch_group(GroupArgs& ga) {
  // One tile shape: in the sweep the 64x64 tile (4 warps, 3 CTAs per SM) was never more than 3 % behind a 128x128 tile
  // (8 warps, 1 CTA per SM) - not even at 4096^3 (30.5 TFLOP/s both) - and far ahead on ragged shapes (N = 641, N = 120).
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

extern "C" int b200ag_gemm_tune(int tile, int splits) {
  (void)tile;                      // one tile shape since the round-2 sweep (see launch_group)
  g_tune_splits = splits < 0 ? 0 : splits;
  return 0;
}

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
                 r.c_dtype == B200AG_F32, 1, r.K, 0, 0, 0};
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
  GemmArgs g{C, A, B, c_rs, c_cs, c_bs, a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, M, N, K, c_dtype == B200AG_F32, 1, K, 0, 0, 0};
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F64) return batched_layout<double, double>(g, batch);
  if (a_dtype == B200AG_F64 && b_dtype == B200AG_F32) return batched_layout<double, float>(g, batch);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F64) return batched_layout<float, double>(g, batch);
  if (a_dtype == B200AG_F32 && b_dtype == B200AG_F32) return batched_layout<float, float>(g, batch);
  B200AG_FAIL("gemm: operand dtypes must be float64 or float32");
}
