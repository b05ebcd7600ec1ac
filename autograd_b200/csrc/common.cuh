// Shared helpers for libb200ag.so (sm_100a).  See include/b200ag.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/b200ag.h"

namespace b200ag {

void set_error(const std::string& msg);
cudaStream_t stream();
int sm_count();
void count_launch(uint64_t n = 1);
bool ensure_init();
int* index_flag();   // device address of the "index out of bounds" flag (mapped host memory), NULL on failure
int gemm_init();        // gemm.cu: allocates the K-split arrival counters
bool f32_tc_eligible(int64_t M, int64_t N, int64_t K, int64_t c_cs, int64_t batch);   // gemm_tc.cu

#define B200AG_CHECK_INIT()                      \
  do {                                           \
    if (!::b200ag::ensure_init()) return 1;      \
  } while (0)

#define B200AG_CUDA(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      ::b200ag::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));   \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

#define B200AG_LAUNCH_CHECK()                                                    \
  do {                                                                           \
    cudaError_t _e = cudaPeekAtLastError();                                      \
    if (_e != cudaSuccess) {                                                     \
      ::b200ag::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
      cudaGetLastError();                                                        \
      return 1;                                                                  \
    }                                                                            \
    ::b200ag::count_launch();                                                    \
  } while (0)

#define B200AG_FAIL(msg)            \
  do {                              \
    ::b200ag::set_error(msg);       \
    return 1;                       \
  } while (0)

inline size_t dtype_size(int dt) {
  switch (dt) {
    case B200AG_F64: return 8;
    case B200AG_F32: return 4;
    case B200AG_I64: return 8;
    case B200AG_I32: return 4;
    case B200AG_BOOL: return 1;
  }
  return 0;
}

// Grid sizing for grid-stride kernels: enough CTAs to cover n, capped at a
// multiple of the SM count (148 on B200) so the last wave is full.
inline unsigned grid_for(int64_t work_items, int block, int ctas_per_sm = 8) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

}  // namespace b200ag
