// float32 contractions on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) with the accumulator in TMEM,
// operands staged by TMA into 128B-swizzled shared memory, one elected thread issuing the MMAs.
//
// Replaces the float32 case of onp.tensordot / onp.dot behind the dot/tensordot/matmul VJPs
// (reference numpy_vjps.py:615-651, 666-742).  north_star asks for float32 within 1e-5 of NumPy's sgemm, which a
// single TF32 product (10-bit mantissa) cannot give, so every operand is split once into two TF32 terms
//     x = hi + lo,  hi = rna_tf32(x),  lo = rna_tf32(x - hi)
// and the kernel accumulates  Ahi*Bhi + Ahi*Blo + Alo*Bhi  in the FP32 TMEM accumulator ("3xTF32"; the dropped
// lo*lo term is ~2^-22 relative).  The split pre-pass also brings both operands into K-major packed form
// [rows, 2*Kp] (hi block, then lo block, Kp = K rounded up to 32 and zero-filled), whatever strides the caller's
// views have, so the GEMM itself has one layout and the transposed operands of the VJPs cost nothing extra.
//
// Tile: 128 x BN (BN = 256 or 128) per CTA, UMMA 128xBNx8, K-block = 32 floats (one 128-byte swizzle row).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
// (tcgen05.ld 32 lanes x 32 columns per warp -> FP32 register accumulators -> global).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace b200ag {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;          // floats per K-block = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;       // kind::tf32
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int KCB = 8;          // k-blocks per accumulation chunk (256 values of k) before promotion to registers
constexpr uint32_t SPIN_LIMIT = 400000000u;   // bounded waits: a protocol bug traps instead of hanging the GPU

template <int BN>
struct Cfg {
  static constexpr int A_TILE = BM * BK * 4;                 // 16 KB
  static constexpr int B_TILE = BN * BK * 4;                 // 32 KB (BN=256)
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;      // hi and lo of both operands
  static constexpr int STAGES = (BN == 256) ? 2 : 3;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4)   // start address, 16-byte units
         | (1ull << 16)                            // leading byte offset (unused for swizzled K-major)
         | ((uint64_t)(1024 >> 4) << 32)           // stride byte offset between 8-row groups
         | (1ull << 46)                            // descriptor version (Blackwell)
         | (2ull << 61);                           // SWIZZLE_128B
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// C[M,N] (float32, row pitch ldc) = A'[M, 2Kp] x B'[N, 2Kp]^T in 3xTF32.
//
// The tensor core adds into its FP32 accumulator with truncation, so a single long accumulation drifts linearly in K
// (measured 1.5e-5 relative at K=2048, 5e-5 at K=8192).  K is therefore cut into chunks of KCB k-blocks: the MMA warp
// alternates between two TMEM accumulators, and the epilogue warps drain each finished chunk into FP32 registers
// with round-to-nearest adds while the next chunk is being multiplied (the drain is ~2% of a chunk's MMA time).
template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   float* __restrict__ C, int M, int N, int Kp, int64_t ldc, int kb_per_split,
                   int64_t split_stride, int mma_only) {
  using cfg = Cfg<BN>;
  constexpr int CPT = BN / 2;   // accumulator columns per epilogue thread (two warps share a 32-lane quadrant)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle-128B tiles need 1024-byte alignment
  const uint32_t bars = base + cfg::STAGES * cfg::STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (cfg::STAGES + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * cfg::STAGES + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * cfg::STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // split-K: blockIdx.z owns k-blocks [kb0, kb0 + num_kb) and writes its partial product to C + z * split_stride
  const int kb0 = blockIdx.z * kb_per_split;
  const int num_kb = min(kb_per_split, Kp / BK - kb0);
  const int num_chunks = (num_kb + KCB - 1) / KCB;
  C += (int64_t)blockIdx.z * split_stride;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {   // one warp allocates both accumulators and owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      const int load_kb = mma_only ? min(num_kb, cfg::STAGES) : num_kb;   // diagnostics: tensor-pipe ceiling without loads
      for (int kb = 0; kb < load_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t s0 = base + stage * cfg::STAGE;
        mbar_expect_tx(full_bar(stage), (uint32_t)cfg::STAGE);
        const int k = (kb0 + kb) * BK;
        tma_load_2d(s0, &map_a, k, m0, full_bar(stage));                                        // A hi
        tma_load_2d(s0 + cfg::A_TILE, &map_a, Kp + k, m0, full_bar(stage));                      // A lo
        tma_load_2d(s0 + 2 * cfg::A_TILE, &map_b, k, n0, full_bar(stage));                       // B hi
        tma_load_2d(s0 + 2 * cfg::A_TILE + cfg::B_TILE, &map_b, Kp + k, n0, full_bar(stage));    // B lo
        if (++stage == cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
      constexpr uint32_t idesc = (1u << 4)                      // D format: F32
                                 | (2u << 7) | (2u << 10)        // A, B format: TF32; both K-major (bits 15, 16 = 0)
                                 | ((uint32_t)(BN >> 3) << 17)   // N
                                 | ((uint32_t)(BM >> 4) << 24);  // M
      int stage = 0; uint32_t phase = 0;
      int kb = 0;
      for (int chunk = 0; chunk < num_chunks; ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(tempty_bar(buf), (((uint32_t)chunk >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        const int kb_end = min(kb + KCB, num_kb);
        for (; kb < kb_end; ++kb) {
          if (!mma_only || kb < cfg::STAGES) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t s0 = base + stage * cfg::STAGE;
          const uint64_t a_hi = umma_desc(s0), a_lo = umma_desc(s0 + cfg::A_TILE);
          const uint64_t b_hi = umma_desc(s0 + 2 * cfg::A_TILE), b_lo = umma_desc(s0 + 2 * cfg::A_TILE + cfg::B_TILE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // 32 bytes per UMMA_K step, in 16-byte units
            umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, acc);    // small terms first
            acc = 1;
            umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          if (!mma_only) umma_commit(empty_bar(stage));   // frees this smem stage once the MMAs above have read it
          if (++stage == cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(buf));            // this chunk's accumulator is complete
      }
    }
  } else {
    // ===== epilogue: drain every chunk TMEM -> registers (FP32 round-to-nearest sum), then registers -> global =====
    const int e = warp - 2;
    const int q = warp & 3;                     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int half = e >> 2;                    // which CPT-column half of the tile this warp owns
    float acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < num_chunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(tfull_bar(buf), ((uint32_t)chunk >> 1) & 1u);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * CPT);
#pragma unroll
      for (int c = 0; c < CPT / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t0 + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    const int row = m0 + q * 32 + lane;
    float* crow = C + (int64_t)row * ldc;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const int col0 = n0 + half * CPT;
    if (row < M) {
      if (vec_ok && col0 + CPT <= N) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4)
          *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (col0 + j < N) crow[col0 + j] = acc[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster, on the two SMs of a TPC, compute a 256 x 256 tile
// with ONE 256x256x8 UMMA issued by the leader.  Each CTA stages its own 128 rows of A and only HALF of the B tile,
// so a k-block costs 64 KB of L2->SM traffic per SM instead of 96 KB, and three stages fit in shared memory.  (The
// single-CTA kernel above is bound by exactly that traffic: ncu shows 10.8 TB/s of L2->SM reads at 65% of the tensor
// rate, profiles/r01_tc_gemm_ncu.csv.)
//   full[s]   (leader)   count 2: the leader's producer arrives with expect_tx for BOTH CTAs' bytes, the peer's
//                        producer arrives remotely; both CTAs' TMA loads complete_tx on the leader's barrier
//   empty[s]  (each CTA) count 1: tcgen05.commit multicast from the leader's MMA thread
//   tfull[b]  (each CTA) count 1: tcgen05.commit multicast; tempty[b] (leader) count 2*EPI_WARPS, peer arrives remotely
namespace pair {

constexpr int BN = 256;
constexpr int A_TILE = BM * BK * 4;            // 16 KB: this CTA's 128 rows of A
constexpr int B_TILE = (BN / 2) * BK * 4;      // 16 KB: this CTA's half of the B tile
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE; // 64 KB
constexpr int STAGES = 3;
constexpr int SMEM = STAGES * STAGE + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        float* __restrict__ C, int M, int N, int Kp, int64_t ldc, int kb_per_split, int64_t split_stride) {
  constexpr int CPT = BN / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();          // 0 = leader (issues the MMAs), 1 = peer
  const int m0 = blockIdx.x * BM;                // this CTA's 128 rows (blockIdx.x = 2 * pair + rank)
  const int n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * kb_per_split;
  const int num_kb = min(kb_per_split, Kp / BK - kb0);
  const int num_chunks = (num_kb + KCB - 1) / KCB;
  C += (int64_t)blockIdx.z * split_stride;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {   // the same warp of both CTAs allocates the pair's accumulators (2 x 256 columns in each SM)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                            // barriers of BOTH CTAs are initialised before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own A rows, own half of B; bytes are counted on the LEADER's full barrier =====
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t s0 = base + stage * STAGE;
        const uint32_t lead_full = map_to_cta(full_bar(stage), 0);
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)STAGE);
        else mbar_arrive_cluster(lead_full);
        const int k = (kb0 + kb) * BK;
        const int nb = n0 + (int)rank * (BN / 2);
        tma_load_2d_pair(s0, &map_a, k, m0, lead_full);
        tma_load_2d_pair(s0 + A_TILE, &map_a, Kp + k, m0, lead_full);
        tma_load_2d_pair(s0 + 2 * A_TILE, &map_b, k, nb, lead_full);
        tma_load_2d_pair(s0 + 2 * A_TILE + B_TILE, &map_b, Kp + k, nb, lead_full);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer: one thread of the leader CTA drives the tensor cores of both SMs =====
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int kb = 0;
      for (int chunk = 0; chunk < num_chunks; ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(tempty_bar(buf), (((uint32_t)chunk >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        const int kb_end = min(kb + KCB, num_kb);
        for (; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t s0 = base + stage * STAGE;
          const uint64_t a_hi = umma_desc(s0), a_lo = umma_desc(s0 + A_TILE);
          const uint64_t b_hi = umma_desc(s0 + 2 * A_TILE), b_lo = umma_desc(s0 + 2 * A_TILE + B_TILE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            umma_tf32_pair(tmem_d, a_lo + adv, b_hi + adv, idesc, acc);
            acc = 1;
            umma_tf32_pair(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32_pair(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit_pair(empty_bar(stage));    // frees this stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(tfull_bar(buf));        // wakes the epilogue warps of both CTAs
      }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the pair's accumulator =====
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    float acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
    const uint32_t lead_tempty0 = map_to_cta(tempty_bar(0), 0), lead_tempty1 = map_to_cta(tempty_bar(1), 0);
    for (int chunk = 0; chunk < num_chunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(tfull_bar(buf), ((uint32_t)chunk >> 1) & 1u);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * CPT);
#pragma unroll
      for (int c = 0; c < CPT / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t0 + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(buf ? lead_tempty1 : lead_tempty0);
    }
    const int row = m0 + q * 32 + lane;
    float* crow = C + (int64_t)row * ldc;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const int col0 = n0 + half * CPT;
    if (row < M) {
      if (vec_ok && col0 + CPT <= N) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4)
          *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (col0 + j < N) crow[col0 + j] = acc[j];
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                            // nobody leaves while the peer may still touch its smem / barriers / TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

}  // namespace pair

// Split pre-pass: dst[r, k] = hi(src[r,k]), dst[r, Kp + k] = lo(src[r,k]); zero for K <= k < Kp.
// src is any strided float32 view (element strides rs, cs); reads are coalesced along whichever axis is contiguous.
__global__ void __launch_bounds__(256) split_tf32_kernel(float* __restrict__ dst, const float* __restrict__ src, int R, int K,
                                                          int Kp, int64_t rs, int64_t cs, int along_rows) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  if (along_rows) {   // source contiguous along r: read with tx -> r, transpose through shared memory
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, r = r0 + tx;
      tile[ty + 8 * i][tx] = (k < K && r < R) ? src[(int64_t)r * rs + (int64_t)k * cs] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, k = k0 + tx;
      tile[tx][ty + 8 * i] = (k < K && r < R) ? src[(int64_t)r * rs + (int64_t)k * cs] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, k = k0 + tx;
    if (r < R) {
      const float x = tile[tx][ty + 8 * i];
      uint32_t h, l;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
      const float hi = __uint_as_float(h);
      const float rest = (fabsf(hi) <= 3.0e38f) ? x - hi : 0.f;   // inf/nan stay in the hi term only
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
      float* d = dst + (int64_t)r * (2 * (int64_t)Kp) + k;
      d[0] = hi;
      d[Kp] = __uint_as_float(l);
    }
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static bool load_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    set_error("gemm_f32_tc: cuTensorMapEncodeTiled is not available from the driver");
    return false;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return true;
}

// 2-D map over a packed [rows, 2*Kp] float32 matrix; box = 32 floats (128 B, swizzled) x box_rows.
static bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t Kp, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)(2 * Kp), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(2 * Kp) * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_f32_tc: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return false;
  }
  return true;
}

static int g_mma_only = 0;

template <int BN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, float* C, int M, int N, int Kp, int64_t ldc, int splits,
                  int kb_per_split, int64_t split_stride) {
  static bool configured = false;
  if (!configured) {
    B200AG_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    configured = true;
  }
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
  gemm_tf32x3_kernel<BN><<<grid, THREADS, Cfg<BN>::SMEM, stream()>>>(ma, mb, C, M, N, Kp, ldc, kb_per_split, split_stride,
                                                                   g_mma_only);
  B200AG_LAUNCH_CHECK();
  return 0;
}

static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, float* C, int M, int N, int Kp, int64_t ldc, int splits,
                       int kb_per_split, int64_t split_stride) {
  static bool configured = false;
  if (!configured) {
    B200AG_CUDA(cudaFuncSetAttribute(pair::gemm_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM));
    configured = true;
  }
  dim3 grid(2 * ((M + 2 * BM - 1) / (2 * BM)), (N + pair::BN - 1) / pair::BN, splits);
  pair::gemm_tf32x3_pair_kernel<<<grid, THREADS, pair::SMEM, stream()>>>(ma, mb, C, M, N, Kp, ldc, kb_per_split, split_stride);
  B200AG_LAUNCH_CHECK();
  return 0;
}

static int g_variant = 0;   // 0 = choose automatically, 1 = single-CTA kernels only, 2 = CTA-pair kernel whenever legal

}  // namespace tc

static int g_f32_mode = 1;

// Used by b200ag_gemm (gemm.cu) to decide whether a float32 x float32 -> float32 product goes to the tensor cores.
bool f32_tc_eligible(int64_t M, int64_t N, int64_t K, int64_t c_cs, int64_t batch) {
  return g_f32_mode == 1 && batch == 1 && c_cs == 1 && M >= 128 && N >= 128 && K >= 32 &&
         M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 29) && (double)M * (double)N * (double)K >= 1.0e8;
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

// 0 = float32 products on the FP64 DMMA kernel (exact products, FP64 accumulation); 1 = tcgen05 3xTF32 for large
// products (default).  Returns the previous mode through *previous (may be null).
int b200ag_gemm_f32_mode(int mode, int* previous) {
  if (previous) *previous = g_f32_mode;
  if (mode == 0 || mode == 1) g_f32_mode = mode;
  if (mode == 100 || mode == 101) tc::g_mma_only = mode - 100;   // diagnostics: 101 = issue MMAs without reloading operands
  if (mode >= 200 && mode <= 202) tc::g_variant = mode - 200;    // diagnostics: kernel variant (see tc::g_variant)
  return 0;
}

// C[M,N] (float32, unit column stride, row stride c_rs) = A[M,K] * B[K,N]; A and B are float32 views with arbitrary
// element strides.  Three launches: split(A), split(B), tcgen05 GEMM.  Workspace comes from the caching allocator.
int b200ag_gemm_f32_tc(void* C, int64_t c_rs, const void* A, int64_t a_rs, int64_t a_cs, const void* B, int64_t b_rs,
                       int64_t b_cs, int64_t M, int64_t N, int64_t K) {
  B200AG_CHECK_INIT();
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) B200AG_FAIL("gemm_f32_tc: K must be positive");
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 29)) B200AG_FAIL("gemm_f32_tc: dimension too large");
  if (!tc::load_encode()) return 1;
  const int64_t Kp = (K + tc::BK - 1) / tc::BK * tc::BK;
  void *a2 = nullptr, *b2 = nullptr;
  if (b200ag_malloc((size_t)(M * 2 * Kp) * 4, &a2)) return 1;
  if (b200ag_malloc((size_t)(N * 2 * Kp) * 4, &b2)) { b200ag_free(a2); return 1; }
  int rc = 0;
  {
    dim3 ga((unsigned)(Kp / 32), (unsigned)((M + 31) / 32));
    tc::split_tf32_kernel<<<ga, 256, 0, stream()>>>((float*)a2, (const float*)A, (int)M, (int)K, (int)Kp, a_rs, a_cs,
                                                     (a_cs != 1 && a_rs == 1) ? 1 : 0);
    count_launch();
    // B is [K, N]: its packed form is [N, 2Kp], i.e. "rows" are n (stride b_cs) and "columns" are k (stride b_rs)
    dim3 gb((unsigned)(Kp / 32), (unsigned)((N + 31) / 32));
    tc::split_tf32_kernel<<<gb, 256, 0, stream()>>>((float*)b2, (const float*)B, (int)N, (int)K, (int)Kp, b_cs, b_rs,
                                                     (b_rs != 1 && b_cs == 1) ? 1 : 0);
    count_launch();
  }
  const int BN = (N > 128 && ((M + 127) / 128) * ((N + 255) / 256) >= sm_count() / 2) ? 256 : 128;
  const bool use_pair = tc::g_variant == 2 || (tc::g_variant == 0 && BN == 256 && M >= 256);
  // Split-K when the output has too few tiles for 148 SMs (weight gradients: small M x N, K = batch): every slice
  // writes its own partial product, summed afterwards in a fixed order (deterministic, round-to-nearest).
  const int64_t tiles = ((M + 127) / 128) * ((N + BN - 1) / BN);
  const int64_t num_kb = Kp / tc::BK;
  int64_t splits = sm_count() / tiles;
  if (splits > num_kb / 16) splits = num_kb / 16;      // at least 512 values of k per slice
  if (splits > 16) splits = 16;
  if (splits < 2 || c_rs != N) splits = 1;
  int64_t kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per_split - 1) / kb_per_split;
  void* part = nullptr;
  if (splits > 1 && b200ag_malloc((size_t)(splits * M * N) * 4, &part)) splits = 1, kb_per_split = num_kb;
  float* out = splits > 1 ? (float*)part : (float*)C;
  const int64_t ldc = splits > 1 ? N : c_rs;
  CUtensorMap ma, mb;
  if (!tc::make_map(&ma, (const float*)a2, M, Kp, tc::BM) || !tc::make_map(&mb, (const float*)b2, N, Kp, use_pair ? 128 : BN)) rc = 1;
  if (rc == 0 && use_pair) rc = tc::launch_pair(ma, mb, out, (int)M, (int)N, (int)Kp, ldc, (int)splits, (int)kb_per_split, M * N);
  else if (rc == 0) rc = (BN == 256) ? tc::launch<256>(ma, mb, out, (int)M, (int)N, (int)Kp, ldc, (int)splits, (int)kb_per_split, M * N)
                                : tc::launch<128>(ma, mb, out, (int)M, (int)N, (int)Kp, ldc, (int)splits, (int)kb_per_split, M * N);
  if (rc == 0 && splits > 1) rc = b200ag_reduce(C, part, B200AG_F32, B200AG_SUM, 1, splits, M * N);
  if (part) b200ag_free(part);
  b200ag_free(a2);
  b200ag_free(b2);
  return rc;
}

}  // extern "C"
