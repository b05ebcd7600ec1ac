// Runtime core of libb200ag.so: device/stream ownership, caching allocator,
// copies, events, NVRTC compilation + launch of generated kernels, CUDA-graph
// capture.  ABI documented in include/b200ag.h.
#include <nvrtc.h>

#include <atomic>
#include <map>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace b200ag {

static thread_local std::string g_error;
static std::string g_rtc_log;
static cudaStream_t g_stream = nullptr;
static cudaStream_t g_side_stream = nullptr;  // uploads issued while g_stream is being captured
static bool g_capturing = false;
static int g_device = -1;
static int g_sm_count = 148;
static std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }
cudaStream_t stream() { return g_stream; }
int sm_count() { return g_sm_count; }
void count_launch(uint64_t n) { g_launches += n; }

bool ensure_init() {
  if (g_device >= 0) return true;
  return b200ag_init(0) == 0;
}

// "index out of bounds" flag: one int in mapped pinned host memory.  Gather / scatter kernels (static and generated)
// set it instead of touching memory outside the array; the host reads it after a synchronisation without any copy.
static volatile int* g_index_flag_host = nullptr;
static int* g_index_flag_dev = nullptr;
int* index_flag() {
  if (g_index_flag_dev) return g_index_flag_dev;
  if (!ensure_init()) return nullptr;
  int* h = nullptr;
  if (cudaHostAlloc((void**)&h, 2 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
    set_error("index flag: cudaHostAlloc failed");
    cudaGetLastError();
    return nullptr;
  }
  h[0] = h[1] = 0;     // [0] index out of bounds, [1] peer collective timed out (comm.cu)
  int* dptr = nullptr;
  if (cudaHostGetDevicePointer((void**)&dptr, h, 0) != cudaSuccess) {
    set_error("index flag: cudaHostGetDevicePointer failed");
    cudaGetLastError();
    return nullptr;
  }
  g_index_flag_host = h;
  g_index_flag_dev = dptr;
  return dptr;
}

// ---------------------------------------------------------------- allocator
// Size-binned caching allocator.  All work runs on one stream, so a block
// handed back by free() can be reused by the next malloc() at once: any kernel
// still reading it is ordered before the kernel that will overwrite it.
// While a CUDA graph is being captured, allocations come from a private arena
// that stays with the graph (addresses baked into the graph must not be handed
// to anybody else while the graph can still be replayed).
struct Arena {
  std::unordered_map<size_t, std::vector<void*>> free_bins;
  std::vector<void*> owned;  // every block cudaMalloc'ed for this arena
};

static std::mutex g_mem_mutex;
static Arena g_main_arena;
static Arena* g_capture_arena = nullptr;
static std::unordered_map<void*, std::pair<size_t, Arena*>> g_live;  // ptr -> (binned size, arena)
static size_t g_in_use = 0, g_cached = 0, g_peak = 0;
static std::unordered_set<void*> g_norecycle;  // graph constants: never handed out again while the graph lives

static size_t bin_size(size_t n) {
  if (n == 0) n = 1;
  if (n <= (1u << 20)) return (n + 511) & ~size_t(511);
  const size_t big = size_t(2) << 20;
  return (n + big - 1) / big * big;
}

static void release_bins(Arena& a) {
  for (auto& kv : a.free_bins)
    for (void* p : kv.second) {
      cudaFree(p);
      g_cached -= kv.first;
    }
  a.free_bins.clear();
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

const char* b200ag_last_error(void) { return g_error.c_str(); }

int b200ag_init(int device) {
  if (g_device >= 0) {
    if (device != g_device) B200AG_FAIL("b200ag_init: already initialised on another device");
    return 0;
  }
  int n = 0;
  B200AG_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) B200AG_FAIL("b200ag_init: no such CUDA device");
  B200AG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200AG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) B200AG_FAIL("b200ag_init: this library is built for sm_100a (B200) only");
  g_sm_count = prop.multiProcessorCount;
  B200AG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  B200AG_CUDA(cudaStreamCreateWithFlags(&g_side_stream, cudaStreamNonBlocking));
  g_device = device;
  if (!index_flag()) return 1;   // allocated up front: never inside a stream capture
  if (gemm_init()) return 1;
  return 0;
}

int b200ag_shutdown(void) {
  if (g_device < 0) return 0;
  cudaStreamSynchronize(g_stream);
  {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    release_bins(g_main_arena);
  }
  return 0;
}

int b200ag_device_count(int* n) {
  B200AG_CUDA(cudaGetDeviceCount(n));
  return 0;
}

int b200ag_device_info(int* sms, int* major, int* minor, size_t* total_mem) {
  B200AG_CHECK_INIT();
  cudaDeviceProp prop;
  B200AG_CUDA(cudaGetDeviceProperties(&prop, g_device));
  *sms = prop.multiProcessorCount;
  *major = prop.major;
  *minor = prop.minor;
  *total_mem = prop.totalGlobalMem;
  return 0;
}

int b200ag_stream(void** s) {
  B200AG_CHECK_INIT();
  *s = (void*)g_stream;
  return 0;
}

int b200ag_synchronize(void) {
  B200AG_CHECK_INIT();
  if (g_capturing) B200AG_FAIL("synchronize() while a CUDA graph is being captured");
  B200AG_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int b200ag_index_flag(void** dev_ptr) {
  B200AG_CHECK_INIT();
  *dev_ptr = (void*)index_flag();
  return *dev_ptr ? 0 : 1;
}

int b200ag_index_error(int* raised) {
  *raised = 0;
  if (g_index_flag_host && g_index_flag_host[0]) {
    *raised = 1;
    g_index_flag_host[0] = 0;
  }
  if (g_index_flag_host && g_index_flag_host[1]) {   // reported in bit 1, left set: a timed-out collective is fatal
    *raised |= 2;
  }
  return 0;
}

int b200ag_launch_count(uint64_t* n) {
  *n = g_launches.load();
  return 0;
}

// ---------------------------------------------------------------- memory
int b200ag_malloc(size_t nbytes, void** ptr) {
  B200AG_CHECK_INIT();
  size_t sz = bin_size(nbytes);
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  Arena* arena = g_capture_arena ? g_capture_arena : &g_main_arena;
  void* p = nullptr;
  auto it = arena->free_bins.find(sz);
  if (it != arena->free_bins.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
    g_cached -= sz;
  } else {
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // give cached blocks back to the driver and retry once
      if (!g_capturing) cudaStreamSynchronize(g_stream);
      release_bins(g_main_arena);
      e = cudaMalloc(&p, sz);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("b200ag_malloc: out of device memory (" + std::to_string(sz) + " bytes requested, " +
                  std::to_string(g_in_use) + " in use)");
        return 1;
      }
    }
    if (arena != &g_main_arena) arena->owned.push_back(p);
  }
  g_live[p] = {sz, arena};
  g_in_use += sz;
  if (g_in_use > g_peak) g_peak = g_in_use;
  *ptr = p;
  return 0;
}

// Destination of a host upload.  Outside a capture this is a plain malloc.  While a graph is being
// captured the upload happens once, at capture time (b200ag_memcpy_h2d), so the block holds a CONSTANT of
// the graph: it must be a block no captured kernel has written or will ever write, i.e. it is allocated
// fresh and never goes back to the arena's free bins.
int b200ag_malloc_upload(size_t nbytes, void** ptr) {
  B200AG_CHECK_INIT();
  if (!g_capturing) return b200ag_malloc(nbytes, ptr);
  size_t sz = bin_size(nbytes);
  void* p = nullptr;
  B200AG_CUDA(cudaMalloc(&p, sz));
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  g_capture_arena->owned.push_back(p);
  g_live[p] = {sz, g_capture_arena};
  g_norecycle.insert(p);
  g_in_use += sz;
  if (g_in_use > g_peak) g_peak = g_in_use;
  *ptr = p;
  return 0;
}

int b200ag_free(void* ptr) {
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  auto it = g_live.find(ptr);
  if (it == g_live.end()) B200AG_FAIL("b200ag_free: pointer was not allocated by b200ag_malloc");
  size_t sz = it->second.first;
  Arena* arena = it->second.second;
  g_live.erase(it);
  g_in_use -= sz;
  if (g_norecycle.count(ptr)) {
    if (arena == nullptr) {  // its graph is gone: release now
      g_norecycle.erase(ptr);
      cudaFree(ptr);
    }
    return 0;  // otherwise the block stays reserved until the owning graph is destroyed
  }
  if (arena == nullptr) {  // block outlived the graph whose arena owned it
    cudaFree(ptr);
    return 0;
  }
  arena->free_bins[sz].push_back(ptr);
  g_cached += sz;
  return 0;
}

int b200ag_empty_cache(void) {
  B200AG_CHECK_INIT();
  B200AG_CUDA(cudaStreamSynchronize(g_stream));
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  release_bins(g_main_arena);
  return 0;
}

int b200ag_mem_stats(size_t* in_use, size_t* cached, size_t* peak) {
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  *in_use = g_in_use;
  *cached = g_cached;
  *peak = g_peak;
  g_peak = g_in_use;
  return 0;
}

int b200ag_host_alloc(size_t nbytes, void** ptr) {
  B200AG_CHECK_INIT();
  B200AG_CUDA(cudaMallocHost(ptr, nbytes ? nbytes : 1));
  return 0;
}

int b200ag_host_free(void* ptr) {
  B200AG_CUDA(cudaFreeHost(ptr));
  return 0;
}

int b200ag_memcpy_h2d(void* dst, const void* src, size_t nbytes) {
  B200AG_CHECK_INIT();
  if (!nbytes) return 0;
  if (g_capturing) {
    // A host pointer must not be baked into a graph (the ndarray is gone at replay time): upload now,
    // on a side stream, into the graph-owned block; the data is then a constant of the graph.
    B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, g_side_stream));
    B200AG_CUDA(cudaStreamSynchronize(g_side_stream));
    return 0;
  }
  B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, g_stream));
  return 0;
}

int b200ag_memcpy_d2h(void* dst, const void* src, size_t nbytes) {
  B200AG_CHECK_INIT();
  if (g_capturing)
    B200AG_FAIL("device->host read while a CUDA graph is being captured: the traced function depends on array "
                "VALUES (float(x), bool(x), .get()); such functions cannot be replayed as a graph");
  if (nbytes) B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, g_stream));
  B200AG_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int b200ag_memcpy_d2d(void* dst, const void* src, size_t nbytes) {
  B200AG_CHECK_INIT();
  if (!nbytes) return 0;
  B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}

int b200ag_memset(void* dst, int byte, size_t nbytes) {
  B200AG_CHECK_INIT();
  if (!nbytes) return 0;
  B200AG_CUDA(cudaMemsetAsync(dst, byte, nbytes, g_stream));
  return 0;
}

// ---------------------------------------------------------------- events
int b200ag_event_create(void** ev) {
  B200AG_CHECK_INIT();
  cudaEvent_t e;
  B200AG_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
int b200ag_event_record(void* ev) {
  B200AG_CUDA(cudaEventRecord((cudaEvent_t)ev, g_stream));
  return 0;
}
int b200ag_event_synchronize(void* ev) {
  B200AG_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}
int b200ag_event_elapsed_ms(void* a, void* b, float* ms) {
  B200AG_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}
int b200ag_event_destroy(void* ev) {
  B200AG_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}

// ---------------------------------------------------------------- auxiliary copy streams (host<->device pipelining)
int b200ag_stream_create(void** s) {
  B200AG_CHECK_INIT();
  cudaStream_t st;
  B200AG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = (void*)st;
  return 0;
}
int b200ag_stream_destroy(void* s) {
  B200AG_CUDA(cudaStreamDestroy((cudaStream_t)s));
  return 0;
}
int b200ag_stream_synchronize(void* s) {
  B200AG_CUDA(cudaStreamSynchronize(s ? (cudaStream_t)s : g_stream));
  return 0;
}
// s == NULL means the library's compute stream
int b200ag_event_record_on(void* ev, void* s) {
  B200AG_CUDA(cudaEventRecord((cudaEvent_t)ev, s ? (cudaStream_t)s : g_stream));
  return 0;
}
int b200ag_stream_wait_event(void* s, void* ev) {
  B200AG_CUDA(cudaStreamWaitEvent(s ? (cudaStream_t)s : g_stream, (cudaEvent_t)ev, 0));
  return 0;
}
int b200ag_memcpy_h2d_on(void* dst, const void* src, size_t nbytes, void* s) {
  B200AG_CHECK_INIT();
  if (!nbytes) return 0;
  B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, s ? (cudaStream_t)s : g_stream));
  return 0;
}
int b200ag_memcpy_d2h_on(void* dst, const void* src, size_t nbytes, void* s) {
  B200AG_CHECK_INIT();
  if (!nbytes) return 0;
  B200AG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, s ? (cudaStream_t)s : g_stream));
  return 0;
}

// ---------------------------------------------------------------- NVRTC
int b200ag_rtc_compile(const char* src, const char* name, void** cubin, size_t* cubin_size) {
  nvrtcProgram prog;
  nvrtcResult r = nvrtcCreateProgram(&prog, src, name, 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) B200AG_FAIL(std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r));
  // --fmad=false: a*b+c must round twice, like two NumPy ufunc calls.
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo",
                        "-default-device"};
  r = nvrtcCompileProgram(prog, 5, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  g_rtc_log.assign(log_size, '\0');
  if (log_size) nvrtcGetProgramLog(prog, &g_rtc_log[0]);
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    B200AG_FAIL(std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(r) + "\n" + g_rtc_log);
  }
  size_t sz = 0;
  nvrtcGetCUBINSize(prog, &sz);
  void* buf = malloc(sz);
  nvrtcGetCUBIN(prog, (char*)buf);
  nvrtcDestroyProgram(&prog);
  *cubin = buf;
  *cubin_size = sz;
  return 0;
}

int b200ag_rtc_free(void* cubin) {
  free(cubin);
  return 0;
}

int b200ag_rtc_log(const char** log) {
  *log = g_rtc_log.c_str();
  return 0;
}

int b200ag_module_load(const void* cubin, size_t size, const char* kernel_name, void** kernel) {
  B200AG_CHECK_INIT();
  (void)size;
  cudaLibrary_t lib;
  B200AG_CUDA(cudaLibraryLoadData(&lib, cubin, nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t k;
  B200AG_CUDA(cudaLibraryGetKernel(&k, lib, kernel_name));
  *kernel = (void*)k;
  return 0;
}

int b200ag_kernel_launch(void* kernel, uint32_t grid, uint32_t block, uint32_t smem, const void* params,
                         size_t params_size) {
  (void)params_size;
  void* args[1] = {const_cast<void*>(params)};
  B200AG_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(block), args, smem, g_stream));
  count_launch();
  return 0;
}

int b200ag_kernel_max_active_blocks(void* kernel, int block, size_t smem, int* nblocks) {
  B200AG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(nblocks, kernel, block, smem));
  return 0;
}

// ---------------------------------------------------------------- CUDA graphs
struct Graph {
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  Arena* arena = nullptr;
};

int b200ag_graph_begin(void) {
  B200AG_CHECK_INIT();
  {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    if (g_capture_arena) B200AG_FAIL("b200ag_graph_begin: a capture is already in progress");
    g_capture_arena = new Arena();
  }
  cudaError_t e = cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    delete g_capture_arena;
    g_capture_arena = nullptr;
    B200AG_FAIL(std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
  }
  g_capturing = true;
  return 0;
}

int b200ag_graph_end(void** out) {
  if (!g_capturing) B200AG_FAIL("b200ag_graph_end: no capture in progress");
  g_capturing = false;
  Graph* g = new Graph();
  cudaError_t e = cudaStreamEndCapture(g_stream, &g->graph);
  {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    g->arena = g_capture_arena;
    g_capture_arena = nullptr;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    b200ag_graph_destroy(g);
    B200AG_FAIL(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    b200ag_graph_destroy(g);
    B200AG_FAIL(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  }
  *out = g;
  return 0;
}

int b200ag_graph_launch(void* graph) {
  Graph* g = (Graph*)graph;
  B200AG_CUDA(cudaGraphLaunch(g->exec, g_stream));
  size_t n = 0;
  cudaGraphGetNodes(g->graph, nullptr, &n);
  count_launch(n);
  return 0;
}

int b200ag_graph_num_nodes(void* graph, size_t* n) {
  Graph* g = (Graph*)graph;
  B200AG_CUDA(cudaGraphGetNodes(g->graph, nullptr, n));
  return 0;
}

int b200ag_graph_destroy(void* graph) {
  Graph* g = (Graph*)graph;
  if (!g) return 0;
  cudaStreamSynchronize(g_stream);
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->arena) {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    // Blocks still referenced by live arrays keep their storage until the arrays
    // die; mark them arena-less so free() releases them, then release the rest now.
    std::unordered_map<void*, bool> live_in_arena;
    for (auto& kv : g_live)
      if (kv.second.second == g->arena) {
        kv.second.second = nullptr;
        live_in_arena[kv.first] = true;
      }
    for (auto& kv : g->arena->free_bins) g_cached -= kv.first * kv.second.size();
    for (void* p : g->arena->owned)
      if (!live_in_arena.count(p)) {
        g_norecycle.erase(p);
        cudaFree(p);
      }
    delete g->arena;
  }
  delete g;
  return 0;
}

}  // extern "C"
