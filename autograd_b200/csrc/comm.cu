// One-shot sum-allreduce of small gradient vectors over NVLink peer memory (one process per GPU).
//
// The reference has no communication layer; BASELINE.json's north_star adds exactly one collective: the sum of
// the batch-sharded gradient leaves (1.42 MB for the MLP, 355 KB for the convnet).  At that size an allreduce is
// pure latency, so instead of a ring it is done in ONE kernel per rank: every rank publishes its vector in a
// CUDA-IPC buffer, raises a flag in every peer, waits for the peers' flags, then READS all peer vectors through
// NVLink (NVSwitch gives full bandwidth to every peer) and adds them in rank order — identical bits on all ranks.
// The kernel carries its own epoch counter in device memory, so it can sit inside a replayed CUDA graph right
// behind the last weight-gradient kernel.
#include "common.cuh"

namespace b200ag {

struct PeerComm {
  int rank = 0, world = 1;
  size_t slot_bytes = 0;
  unsigned char* local = nullptr;      // [2 slots][slot_bytes] data, [slot_bytes] reduced chunks (two-shot), then flags:
                                       // A [2][world] u32 at +0, B [2][world] u32 at +1024, epoch/status/arrival at +2048
  unsigned char* peers[16] = {nullptr};
  uint32_t* flags_of[16] = {nullptr};  // flags array inside each peer's block
  size_t flags_off = 0;
  unsigned char* mc = nullptr;         // multicast (NVLS) mapping of the same block on all ranks, when attached
  bool external = false;               // block allocated by the caller (b200ag_comm_attach): never freed here
};
static PeerComm g_comm;
static int g_comm_mode = 0;   // 0 = choose by rank count and size, 1 = one-shot pull, 2 = two-shot, 3 = push, 4 = NVLS multimem (b200ag_comm_mode)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Flag rounds are done by `world` threads in parallel - thread p signals / polls rank p - instead of one thread walking
// all ranks: a release store to a remote GPU costs a NVLink round trip, eight of them in sequence were a third of
// the 46 us an 8-rank allreduce took.
__device__ __forceinline__ void signal_all(uint32_t* const* peer_flags, int offset, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();     // everything this GPU wrote before (earlier kernels, the CTA barrier before this call)
    st_release_sys(peer_flags[threadIdx.x] + offset, epoch);
  }
}
// Every thread of the CTA gets the verdict; a timed-out wait sets *status (mapped host memory).
__device__ __forceinline__ int wait_all(const uint32_t* my_flags, int world, uint32_t epoch, uint32_t* status, int good_in) {
  int good = good_in;
  if ((int)threadIdx.x < world && good) {
    const uint32_t* f = my_flags + threadIdx.x;
    long long spins = 0;
    while (ld_acquire_sys(f) != epoch) {
      if (++spins > 20000000LL) { good = 0; *status = 1; break; }   // a peer died or never arrived; do not hang the GPU
      __nanosleep(32);
    }
  }
  return __syncthreads_and(good);
}

struct AllreduceArgs {
  double* out;
  const double* peer_data[16];   // slot base of every rank's buffer (index = rank), as mapped in THIS process
  uint32_t* peer_flags[16];      // flags array of every rank
  uint32_t* my_flags;
  uint32_t* epoch;               // device-resident call counter of this rank
  uint32_t* status;              // set to 1 when a wait timed out
  int64_t n;
  int64_t slot_elems;
  int rank, world;
};

// Copy this rank's vector into the slot the CURRENT epoch selects (epoch lives on the device: graph-replay safe).
__global__ void __launch_bounds__(256) peer_publish_kernel(double* slots, const double* __restrict__ src, int64_t n,
                                                            int64_t slot_elems, const uint32_t* epoch) {
  double* dst = slots + (int64_t)(*epoch & 1) * slot_elems;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dst[i] = src[i];
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n] = 0.0;   // pad to a whole double2 (two-shot path)
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(AllreduceArgs a) {
  const uint32_t epoch = *a.epoch;           // same value on every rank: each call advances every counter once
  const int slot = epoch & 1;
  if (blockIdx.x == 0) signal_all(a.peer_flags, slot * a.world + a.rank, a.world, epoch);
  const int ok = wait_all(a.my_flags + slot * a.world, a.world, epoch, a.status, 1);
  if (ok) {
    // every thread owns ONE pair of doubles (grid sized to the vector): a single NVLink round trip, all peer loads
    // of a thread issued back to back, summed in rank order so every rank produces the same bits
    const int64_t off = (int64_t)slot * a.slot_elems;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t npairs = a.n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += step) {
      double2 v[16];
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if (p < a.world) v[p] = __ldcv(reinterpret_cast<const double2*>(a.peer_data[p] + off) + i);
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if (p < a.world) { acc.x += v[p].x; acc.y += v[p].y; }
      reinterpret_cast<double2*>(a.out)[i] = acc;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double acc = 0.0;
      for (int p = 0; p < a.world; ++p) acc += __ldcv(a.peer_data[p] + off + a.n - 1);
      a.out[a.n - 1] = acc;
    }
  }
  // last block to finish advances the epoch (device-resident, so a replayed graph keeps counting)
  __shared__ unsigned last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned done = atomicAdd(a.epoch + 2, 1u);   // epoch[2] = arrival counter
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.epoch[2] = 0;
    a.epoch[0] = epoch + 1;
  }
}


// Two-shot variant for larger rank counts: rank r first reduces chunk r of every peer's vector (reduce-scatter by
// pulling over NVLink), publishes the reduced chunk, then gathers the other ranks' reduced chunks.  Per rank this
// moves 2 x (world-1)/world of the vector instead of (world-1) x the vector, and every rank receives the same bits.
//   flags A: "my input is published"  flags B: "my reduced chunk is ready"  (both carry the epoch, double-slotted)
__global__ void __launch_bounds__(256) peer_allreduce2_kernel(AllreduceArgs a, double* my_result) {
  __shared__ unsigned last;
  const uint32_t epoch = *a.epoch;
  const int slot = epoch & 1;
  uint32_t* flagsB_mine = a.my_flags + 256;          // B flags live 1024 bytes after the A flags
  if (blockIdx.x == 0) signal_all(a.peer_flags, slot * a.world + a.rank, a.world, epoch);
  int ok = wait_all(a.my_flags + slot * a.world, a.world, epoch, a.status, 1);
  const int64_t off = (int64_t)slot * a.slot_elems;
  const int64_t npairs = (a.n + 1) >> 1;                         // vectors are padded to an even length in the slots
  const int64_t chunk = (npairs + a.world - 1) / a.world;        // double2 pairs per rank
  const int64_t c_lo = (int64_t)a.rank * chunk, c_hi = min(c_lo + chunk, npairs);
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ok) {
    for (int64_t i = c_lo + tid; i < c_hi; i += step) {
      double2 v[16];
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if (p < a.world) v[p] = __ldcv(reinterpret_cast<const double2*>(a.peer_data[p] + off) + i);
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if (p < a.world) { acc.x += v[p].x; acc.y += v[p].y; }
      reinterpret_cast<double2*>(my_result)[i] = acc;            // result area is indexed like the full vector
    }
  }
  // every CTA of this rank has written its part of the chunk: the last one to arrive announces it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned done = atomicAdd(a.epoch + 2, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) a.epoch[2] = 0;
    signal_all(a.peer_flags, 256 + slot * a.world + a.rank, a.world, epoch);
  }
  // gather: wait for every rank's reduced chunk, copy it out
  ok = wait_all(flagsB_mine + slot * a.world, a.world, epoch, a.status, ok);
  if (ok) {
    for (int64_t i = tid; i < npairs; i += step) {
      const int p = (int)(i / chunk);
      // peer p's result area sits 2 slots after its data base
      const double2 v = __ldcv(reinterpret_cast<const double2*>(a.peer_data[p] + 2 * a.slot_elems) + i);
      if (2 * i + 1 < a.n) reinterpret_cast<double2*>(a.out)[i] = v;
      else a.out[2 * i] = v.x;
    }
  }
  // second arrival round: the last CTA advances the epoch (arrival counter epoch[3])
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned done = atomicAdd(a.epoch + 3, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.epoch[3] = 0;
    a.epoch[0] = epoch + 1;
  }
}

// NVLS variant (b200ag_comm_attach with a multicast mapping): the reduction happens INSIDE the NVSwitch.  Rank r owns
// chunk r: one `multimem.ld_reduce.add.f64` per element makes the switch fetch that element from every rank's published
// slot and return the sum (one NVLink round trip instead of world-1 peer loads), and one `multimem.st` writes the sum
// into the result area of EVERY rank.  Each element is reduced exactly once, by one rank, so all ranks hold the same
// bits.  Same flag protocol as the two-shot kernel: A = "my input is published", B = "my chunk is broadcast".
__device__ __forceinline__ double multimem_ld_reduce_add(const double* mc) {
  double v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(double* mc, double v) {
  asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}

__global__ void __launch_bounds__(256) nvls_allreduce_kernel(AllreduceArgs a, const double* mc_slots, double* mc_result,
                                                             const double* my_result) {
  __shared__ unsigned last;
  const uint32_t epoch = *a.epoch;
  const int slot = epoch & 1;
  uint32_t* flagsB_mine = a.my_flags + 256;
  if (blockIdx.x == 0) signal_all(a.peer_flags, slot * a.world + a.rank, a.world, epoch);
  int ok = wait_all(a.my_flags + slot * a.world, a.world, epoch, a.status, 1);
  const int64_t off = (int64_t)slot * a.slot_elems;
  const int64_t chunk = (a.n + a.world - 1) / a.world;
  const int64_t c_lo = (int64_t)a.rank * chunk, c_hi = min(c_lo + chunk, a.n);
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ok) {
    for (int64_t i = c_lo + tid; i < c_hi; i += step) multimem_st(mc_result + i, multimem_ld_reduce_add(mc_slots + off + i));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned done = atomicAdd(a.epoch + 2, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) a.epoch[2] = 0;
    signal_all(a.peer_flags, 256 + slot * a.world + a.rank, a.world, epoch);
  }
  ok = wait_all(flagsB_mine + slot * a.world, a.world, epoch, a.status, ok);
  if (ok) {
    for (int64_t i = tid; i < a.n; i += step) a.out[i] = __ldcv(my_result + i);     // local HBM: every chunk was broadcast here
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned done = atomicAdd(a.epoch + 3, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.epoch[3] = 0;
    a.epoch[0] = epoch + 1;
  }
}

// Push variant: instead of every rank pulling all peers' vectors (round-trip loads over NVLink), every rank WRITES its
// vector into an inbox slot of every peer (posted stores pipeline much better than loads), raises the peers' flags once
// all of its CTAs have finished, and the sum then runs over local HBM only.
//   inbox layout in each rank's block: [parity][source rank][slot]; flags C at +1536 bytes: [parity][source rank]
struct PushArgs {
  unsigned char* peer_base[16];
  uint32_t* peer_flags[16];
  uint32_t* my_flags;
  uint32_t* epoch;
  uint32_t* status;
  const double* src;
  double* out;
  const double* my_inbox;        // this rank's inbox region (base of parity 0, source 0)
  int64_t n, slot_elems, inbox_off;
  int rank, world;
};

__global__ void __launch_bounds__(256) peer_push_kernel(PushArgs a) {
  __shared__ unsigned last;
  const uint32_t epoch = *a.epoch;
  const int par = epoch & 1;
  const int64_t npairs = a.n >> 1;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t dst_off = a.inbox_off + ((int64_t)(par * a.world + a.rank) * a.slot_elems) * 8;
  const double2* src2 = reinterpret_cast<const double2*>(a.src);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += step) {
    const double2 v = src2[i];
#pragma unroll
    for (int p = 0; p < 16; ++p)
      if (p < a.world) reinterpret_cast<double2*>(a.peer_base[p] + dst_off)[i] = v;
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double v = a.src[a.n - 1];
    for (int p = 0; p < a.world; ++p) reinterpret_cast<double*>(a.peer_base[p] + dst_off)[a.n - 1] = v;
  }
  // CTA barrier, then ONE thread orders the CTA's stores before its arrival (a system fence by every thread costs tens
  // of microseconds); the last arriver's system fence + release store publishes everything by cumulativity
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned done = atomicAdd(a.epoch + 4, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) a.epoch[4] = 0;
    signal_all(a.peer_flags, 384 + par * a.world + a.rank, a.world, epoch);
  }
}

__global__ void __launch_bounds__(256) peer_push_reduce_kernel(PushArgs a) {
  __shared__ unsigned last;
  const uint32_t epoch = *a.epoch;
  const int par = epoch & 1;
  const int ok = wait_all(a.my_flags + 384 + par * a.world, a.world, epoch, a.status, 1);
  if (ok) {
    const double* base = a.my_inbox + (int64_t)par * a.world * a.slot_elems;
    const int64_t npairs = a.n >> 1;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += step) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if (p < a.world) {
          const double2 v = __ldcv(reinterpret_cast<const double2*>(base + (int64_t)p * a.slot_elems) + i);
          acc.x += v.x; acc.y += v.y;
        }
      reinterpret_cast<double2*>(a.out)[i] = acc;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double acc = 0.0;
      for (int p = 0; p < a.world; ++p) acc += __ldcv(base + (int64_t)p * a.slot_elems + a.n - 1);
      a.out[a.n - 1] = acc;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned done = atomicAdd(a.epoch + 2, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.epoch[2] = 0;
    a.epoch[0] = epoch + 1;
  }
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

// Allocate this rank's comm block (2 data slots + flags + counters) and return its 64-byte CUDA IPC handle.
int b200ag_comm_create(int rank, int world, size_t slot_bytes, void* ipc_handle_out) {
  B200AG_CHECK_INIT();
  if (world < 1 || world > 16 || rank < 0 || rank >= world) B200AG_FAIL("comm_create: bad rank/world (max 16 ranks)");
  if (g_comm.local) B200AG_FAIL("comm_create: already created");
  slot_bytes = (slot_bytes + 255) & ~size_t(255);
  // [2 data slots][1 result slot (two-shot)][inbox: 2 parities x world slots (push)][flags 4 KB]
  size_t flags_off = (3 + 2 * (size_t)world) * slot_bytes;
  size_t total = flags_off + 4096;
  void* p = nullptr;
  B200AG_CUDA(cudaMalloc(&p, total));
  B200AG_CUDA(cudaMemset(p, 0, total));
  g_comm.rank = rank; g_comm.world = world; g_comm.slot_bytes = slot_bytes; g_comm.flags_off = flags_off;
  g_comm.local = (unsigned char*)p;
  uint32_t one = 1;   // epoch starts at 1 so that zeroed flags never match
  B200AG_CUDA(cudaMemcpy(g_comm.local + flags_off + 2048, &one, 4, cudaMemcpyHostToDevice));
  cudaIpcMemHandle_t h;
  B200AG_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_out, &h, sizeof(h));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  return 0;
}

// Size of one rank's comm block for b200ag_comm_attach.
int b200ag_comm_block_bytes(int world, size_t slot_bytes, size_t* total) {
  slot_bytes = (slot_bytes + 255) & ~size_t(255);
  *total = (3 + 2 * (size_t)world) * slot_bytes + 4096;
  return 0;
}

// Use comm blocks the CALLER allocated symmetrically on every rank (zero-filled, b200ag_comm_block_bytes each) and mapped
// into this process: peer_bases[p] = rank p's block as addressable from here, multicast_base = the NVLS multicast
// mapping of the same blocks (NULL when the platform has none).  torch.distributed's symmetric memory provides both.
int b200ag_comm_attach(int rank, int world, size_t slot_bytes, void* const* peer_bases, void* multicast_base) {
  B200AG_CHECK_INIT();
  if (world < 1 || world > 16 || rank < 0 || rank >= world) B200AG_FAIL("comm_attach: bad rank/world (max 16 ranks)");
  if (g_comm.local) B200AG_FAIL("comm_attach: a communicator already exists");
  slot_bytes = (slot_bytes + 255) & ~size_t(255);
  g_comm.rank = rank; g_comm.world = world; g_comm.slot_bytes = slot_bytes;
  g_comm.flags_off = (3 + 2 * (size_t)world) * slot_bytes;
  g_comm.external = true;
  g_comm.mc = (unsigned char*)multicast_base;
  for (int p = 0; p < world; ++p) {
    if (!peer_bases[p]) { g_comm = PeerComm(); B200AG_FAIL("comm_attach: missing peer mapping"); }
    g_comm.peers[p] = (unsigned char*)peer_bases[p];
    g_comm.flags_of[p] = (uint32_t*)(g_comm.peers[p] + g_comm.flags_off);
  }
  g_comm.local = g_comm.peers[rank];
  uint32_t one = 1;   // epoch starts at 1 so that zeroed flags never match
  B200AG_CUDA(cudaMemcpy(g_comm.local + g_comm.flags_off + 2048, &one, 4, cudaMemcpyHostToDevice));
  return 0;
}

// Map every rank's comm block (handles = world x 64 bytes, in rank order).
int b200ag_comm_connect(const void* handles) {
  B200AG_CHECK_INIT();
  if (!g_comm.local) B200AG_FAIL("comm_connect: call comm_create first");
  for (int p = 0; p < g_comm.world; ++p) {
    if (p == g_comm.rank) {
      g_comm.peers[p] = g_comm.local;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const unsigned char*)handles + 64 * p, 64);
      void* q = nullptr;
      B200AG_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      g_comm.peers[p] = (unsigned char*)q;
    }
    g_comm.flags_of[p] = (uint32_t*)(g_comm.peers[p] + g_comm.flags_off);
  }
  return 0;
}

// out[0:n] = sum over ranks of src[0:n] (float64).  src is first copied into this rank's published slot.
int b200ag_comm_allreduce_sum_f64(void* out, const void* src, int64_t n) {
  B200AG_CHECK_INIT();
  if (!g_comm.local || !g_comm.peers[g_comm.world - 1]) B200AG_FAIL("comm_allreduce: communicator is not connected");
  if ((size_t)(n + 1) * 8 > g_comm.slot_bytes) B200AG_FAIL("comm_allreduce: message larger than the comm slot");
  if (n == 0) return 0;
  if (((uintptr_t)out & 15) != 0 || ((uintptr_t)src & 15) != 0) B200AG_FAIL("comm_allreduce: buffers must be 16-byte aligned");
  // measured at 1.42 MB and 8 ranks: pull 46.7 us, push 49.3 us (both are bound by the all-to-all NVLink traffic): pull stays the default
  if (g_comm_mode == 3) {
    PushArgs q;
    for (int p = 0; p < g_comm.world; ++p) { q.peer_base[p] = g_comm.peers[p]; q.peer_flags[p] = g_comm.flags_of[p]; }
    q.my_flags = g_comm.flags_of[g_comm.rank];
    q.epoch = (uint32_t*)(g_comm.local + g_comm.flags_off + 2048);
    q.status = (uint32_t*)(index_flag() + 1);    // mapped host memory: the host sees a timeout at its next synchronisation
    q.src = (const double*)src; q.out = (double*)out;
    q.n = n; q.slot_elems = (int64_t)(g_comm.slot_bytes / 8);
    q.inbox_off = 3 * (int64_t)g_comm.slot_bytes;
    q.my_inbox = (const double*)(g_comm.local + q.inbox_off);
    q.rank = g_comm.rank; q.world = g_comm.world;
    unsigned grid = (unsigned)((n / 2 + 255) / 256);
    unsigned cap = (unsigned)sm_count() * 2;   // both kernels hold a grid-wide arrival: every CTA must be resident
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    peer_push_kernel<<<grid, 256, 0, stream()>>>(q);
    B200AG_LAUNCH_CHECK();
    peer_push_reduce_kernel<<<grid, 256, 0, stream()>>>(q);
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  AllreduceArgs a;
  a.out = (double*)out;
  for (int p = 0; p < g_comm.world; ++p) {
    a.peer_data[p] = (const double*)g_comm.peers[p];
    a.peer_flags[p] = g_comm.flags_of[p];
  }
  a.my_flags = g_comm.flags_of[g_comm.rank];
  a.epoch = (uint32_t*)(g_comm.local + g_comm.flags_off + 2048);
  a.status = (uint32_t*)(index_flag() + 1);      // mapped host memory: the host sees a timeout at its next synchronisation
  a.n = n;
  a.slot_elems = (int64_t)(g_comm.slot_bytes / 8);
  a.rank = g_comm.rank;
  a.world = g_comm.world;
  unsigned pgrid = (unsigned)((n + 1023) / 1024);
  if (pgrid > 148) pgrid = 148;
  peer_publish_kernel<<<pgrid, 256, 0, stream()>>>((double*)g_comm.local, (const double*)src, n, a.slot_elems, a.epoch);
  B200AG_LAUNCH_CHECK();
  if (((uintptr_t)out & 15) != 0) B200AG_FAIL("comm_allreduce: output must be 16-byte aligned");
  // measured at 1.42 MB: one-shot 21 / 43 us at 2 / 8 ranks, two-shot 31 / 66 us (two flag rounds + a grid-wide arrival):
  // the two-shot form only pays once the (world-1)x larger one-shot traffic dominates, i.e. for multi-megabyte vectors
  // in-switch reduction pays from 8 ranks on (1.42 MB: 25-26 us against 34 pull / 33 push); at 4 ranks the one-shot kernels
  // are faster (21 us against 26), at 2 ranks too (17-19 against 27): profiles/r02_peer_allreduce_latency.txt
  const bool nvls = g_comm.mc && (g_comm_mode == 4 || (g_comm_mode == 0 && g_comm.world > 4));
  if (nvls) {
    int64_t per_rank = (n + g_comm.world - 1) / g_comm.world;
    unsigned grid = (unsigned)((per_rank + 255) / 256);
    unsigned cap = (unsigned)sm_count();          // grid-wide arrival rounds: every CTA must be resident
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    nvls_allreduce_kernel<<<grid, 256, 0, stream()>>>(a, (const double*)g_comm.mc, (double*)(g_comm.mc + 2 * g_comm.slot_bytes),
                                                      (const double*)(g_comm.local + 2 * g_comm.slot_bytes));
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  const bool two_shot = g_comm_mode == 2 || (g_comm_mode == 0 && g_comm.world >= 4 && n >= (1 << 19));
  if (two_shot) {
    // every CTA takes part in a grid-wide arrival round, so all of them must be resident at once
    int64_t pairs_per_rank = ((n + 1) / 2 + g_comm.world - 1) / g_comm.world;
    unsigned grid = (unsigned)((pairs_per_rank + 255) / 256);
    unsigned cap = (unsigned)sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    double* my_result = (double*)(g_comm.local + 2 * g_comm.slot_bytes);
    peer_allreduce2_kernel<<<grid, 256, 0, stream()>>>(a, my_result);
    B200AG_LAUNCH_CHECK();
    return 0;
  }
  unsigned grid = (unsigned)((n / 2 + 255) / 256);
  unsigned cap = (unsigned)sm_count() * 4;   // all CTAs are resident: they only wait on remote flags, never on each other
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  peer_allreduce_kernel<<<grid, 256, 0, stream()>>>(a);
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_comm_status(int* timed_out) {
  B200AG_CHECK_INIT();
  int flags = 0;
  b200ag_index_error(&flags);
  *timed_out = (flags & 2) ? 1 : 0;
  return 0;
}

int b200ag_comm_destroy(void) {
  if (!g_comm.local) return 0;
  cudaStreamSynchronize(stream());
  if (!g_comm.external) {
    for (int p = 0; p < g_comm.world; ++p)
      if (p != g_comm.rank && g_comm.peers[p]) cudaIpcCloseMemHandle(g_comm.peers[p]);
    cudaFree(g_comm.local);
  }
  g_comm = PeerComm();
  return 0;
}

// Algorithm of b200ag_comm_allreduce_sum_f64: 0 = automatic (two-shot from 4 ranks and 4 MB up), 1 = one-shot,
// 2 = two-shot (reduce-scatter + all-gather over peer memory).
int b200ag_comm_mode(int mode, int* previous) {
  if (previous) *previous = g_comm_mode;
  if (mode >= 0 && mode <= 4) g_comm_mode = mode;
  return 0;
}

}  // extern "C"
