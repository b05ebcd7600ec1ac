// Direct 2-D convolution for autograd.scipy.signal.convolve in its NCHW use
// (examples/convnet.py:83: axes=([2,3],[2,3]), dot_axes=([1],[0])) and the two
// contractions its VJP needs (autograd/scipy/signal.py:153-214).  Channel counts
// are tiny (1, 6, 16) and the arithmetic is FP64, so this is a CUDA-core kernel:
// filters staged once per CTA in shared memory, activations read coalesced.
#include "common.cuh"

namespace b200ag {

// out[n,f,y,x] = sum_{c,i,j} A[n,c,y+i-py,x+j-px] * B[c,f,bi,bj]
//   flip=1: (bi,bj) = (kh-1-i, kw-1-j)  (true convolution, signal.py:49-54)
//   flip=0: (bi,bj) = (i, j)            (correlation)
//   mode valid: py=px=0, Ho=H-kh+1;  mode full: py=kh-1, px=kw-1, Ho=H+kh-1 (zero padded)
template <typename T>
__global__ void __launch_bounds__(256) conv2d_kernel(T* __restrict__ out, const T* __restrict__ A,
                                                      const T* __restrict__ B, int64_t N, int C, int H, int W,
                                                      int F, int kh, int kw, int Ho, int Wo, int py, int px,
                                                      int flip) {
  extern __shared__ unsigned char smem_raw[];
  T* Bs = (T*)smem_raw;  // [C][F][kh][kw], stored pre-flipped so the inner loop indexes (i,j) directly
  int nB = C * F * kh * kw;
  for (int e = threadIdx.x; e < nB; e += blockDim.x) {
    int j = e % kw, i = (e / kw) % kh, cf = e / (kw * kh);
    int bi = flip ? kh - 1 - i : i, bj = flip ? kw - 1 - j : j;
    Bs[e] = B[(cf * kh + bi) * kw + bj];
  }
  __syncthreads();
  int64_t per_img = (int64_t)F * Ho * Wo;
  int64_t total = N * per_img;
  int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += step) {
    int x = (int)(o % Wo);
    int y = (int)((o / Wo) % Ho);
    int f = (int)((o / ((int64_t)Wo * Ho)) % F);
    int64_t n = o / per_img;
    T acc = 0;
    for (int c = 0; c < C; ++c) {
      const T* a = A + (n * C + c) * (int64_t)H * W;
      const T* b = Bs + (c * F + f) * kh * kw;
      for (int i = 0; i < kh; ++i) {
        int yy = y + i - py;
        if (yy < 0 || yy >= H) continue;
        for (int j = 0; j < kw; ++j) {
          int xx = x + j - px;
          if (xx < 0 || xx >= W) continue;
          acc += a[yy * W + xx] * b[i * kw + j];
        }
      }
    }
    out[o] = acc;
  }
}

// dB[c,f,kh-1-i,kw-1-j] = sum_{n,y,x} A[n,c,y+i,x+j] * G[n,f,y,x]   (valid correlation summed over the batch)
// One CTA per (chunk of images); every thread owns a set of (c,f,i,j) outputs and
// accumulates over the chunk from shared-memory copies of A[n] and G[n]; the chunk
// partials are combined with FP64 atomics into a zeroed dB.
template <typename T>
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(T* __restrict__ dB, const T* __restrict__ A,
                                                            const T* __restrict__ G, int64_t N, int C, int H, int W,
                                                            int F, int kh, int kw, int Ho, int Wo,
                                                            int64_t imgs_per_cta) {
  extern __shared__ unsigned char smem_raw[];
  T* As = (T*)smem_raw;               // [C][H][W]
  T* Gs = As + (size_t)C * H * W;     // [F][Ho][Wo]
  const int nout = C * F * kh * kw;
  constexpr int MAX_OWN = 12;         // outputs per thread (nout <= 256*12)
  T acc[MAX_OWN];
#pragma unroll
  for (int k = 0; k < MAX_OWN; ++k) acc[k] = 0;
  int64_t n_lo = (int64_t)blockIdx.x * imgs_per_cta;
  int64_t n_hi = n_lo + imgs_per_cta < N ? n_lo + imgs_per_cta : N;
  for (int64_t n = n_lo; n < n_hi; ++n) {
    const T* a = A + n * (int64_t)C * H * W;
    const T* g = G + n * (int64_t)F * Ho * Wo;
    for (int e = threadIdx.x; e < C * H * W; e += blockDim.x) As[e] = a[e];
    for (int e = threadIdx.x; e < F * Ho * Wo; e += blockDim.x) Gs[e] = g[e];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAX_OWN; ++k) {
      int o = threadIdx.x + k * 256;
      if (o < nout) {
        int j = o % kw, i = (o / kw) % kh, f = (o / (kw * kh)) % F, c = o / (kw * kh * F);
        const T* ap = As + (c * H + i) * W + j;
        const T* gp = Gs + f * Ho * Wo;
        T s = 0;
        for (int y = 0; y < Ho; ++y)
          for (int x = 0; x < Wo; ++x) s += ap[y * W + x] * gp[y * Wo + x];
        acc[k] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < MAX_OWN; ++k) {
    int o = threadIdx.x + k * 256;
    if (o < nout) {
      int j = o % kw, i = (o / kw) % kh, cf = o / (kw * kh);
      atomicAdd(dB + (cf * kh + (kh - 1 - i)) * kw + (kw - 1 - j), acc[k]);
    }
  }
}

}  // namespace b200ag

using namespace b200ag;

extern "C" {

int b200ag_conv2d(void* out, const void* A, const void* B, int dtype, int64_t N, int64_t C, int64_t H, int64_t W,
                  int64_t F, int64_t kh, int64_t kw, int mode, int flip) {
  B200AG_CHECK_INIT();
  int64_t Ho, Wo, py, px;
  if (mode == 0) {
    Ho = H - kh + 1; Wo = W - kw + 1; py = px = 0;
    if (Ho <= 0 || Wo <= 0) B200AG_FAIL("conv2d: filter larger than input in 'valid' mode");
  } else if (mode == 1) {
    Ho = H + kh - 1; Wo = W + kw - 1; py = kh - 1; px = kw - 1;
  } else {
    B200AG_FAIL("conv2d: mode must be 0 (valid) or 1 (full)");
  }
  int64_t total = N * F * Ho * Wo;
  if (total == 0) return 0;
  size_t es = dtype_size(dtype);
  size_t smem = (size_t)(C * F * kh * kw) * es;
  if (smem > 160 * 1024) B200AG_FAIL("conv2d: filter bank does not fit in shared memory");
  unsigned grid = grid_for(total, 256, 8);
  if (dtype == B200AG_F64) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfg = true; }
    conv2d_kernel<double><<<grid, 256, smem, stream()>>>((double*)out, (const double*)A, (const double*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, (int)py, (int)px, flip);
  } else if (dtype == B200AG_F32) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfg = true; }
    conv2d_kernel<float><<<grid, 256, smem, stream()>>>((float*)out, (const float*)A, (const float*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, (int)py, (int)px, flip);
  } else {
    B200AG_FAIL("conv2d: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

int b200ag_conv2d_wgrad(void* dB, const void* A, const void* G, int dtype, int64_t N, int64_t C, int64_t H, int64_t W,
                        int64_t F, int64_t kh, int64_t kw) {
  B200AG_CHECK_INIT();
  int64_t Ho = H - kh + 1, Wo = W - kw + 1;
  if (Ho <= 0 || Wo <= 0) B200AG_FAIL("conv2d_wgrad: filter larger than input");
  int64_t nout = C * F * kh * kw;
  if (nout > 256 * 12) B200AG_FAIL("conv2d_wgrad: more than 3072 filter taps is not supported");
  size_t es = dtype_size(dtype);
  if (b200ag_memset(dB, 0, (size_t)nout * es)) return 1;
  if (N == 0) return 0;
  size_t smem = (size_t)(C * H * W + F * Ho * Wo) * es;
  if (smem > 200 * 1024) B200AG_FAIL("conv2d_wgrad: one image does not fit in shared memory");
  int64_t ctas = (int64_t)sm_count() * 2;
  int64_t per = (N + ctas - 1) / ctas;
  if (per < 1) per = 1;
  unsigned grid = (unsigned)((N + per - 1) / per);
  if (dtype == B200AG_F64) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_wgrad_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
    conv2d_wgrad_kernel<double><<<grid, 256, smem, stream()>>>((double*)dB, (const double*)A, (const double*)G, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, per);
  } else if (dtype == B200AG_F32) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
    conv2d_wgrad_kernel<float><<<grid, 256, smem, stream()>>>((float*)dB, (const float*)A, (const float*)G, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, per);
  } else {
    B200AG_FAIL("conv2d_wgrad: unsupported dtype");
  }
  B200AG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
