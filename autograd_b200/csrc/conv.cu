// 2-D convolution for autograd.scipy.signal.convolve in its NCHW use (examples/convnet.py:83:
// axes=([2,3],[2,3]), dot_axes=([1],[0])) and the two contractions its VJP needs
// (autograd/scipy/signal.py:153-214).
//
// float64 runs as an IMPLICIT GEMM on the DMMA tensor pipe (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind):
//   forward / input-gradient:  rows = output positions (y,x) of one image, cols = output channels f,
//                              k = (c,i,j) filter taps.  Images are staged (zero-padded for 'full') in shared
//                              memory, the A fragment of every MMA is gathered from there through a small
//                              tap-offset table, the filter bank sits in shared memory as the B operand.
//   weight-gradient:           rows = taps (c,i,j), cols = f, k = (n,y,x) positions; every CTA reduces a chunk
//                              of the batch in registers and adds its partial with FP64 atomics.
// float32 keeps the simple CUDA-core kernels (no config uses it).
#include <cuda_pipeline.h>

#include <stdlib.h>

#include "common.cuh"

namespace b200ag {

// conv_direct.cu: register-tiled CUDA-core kernels for layers with few output channels (0 done, 1 error, 2 not covered)
int conv2d_direct_try(double* out, const double* A, const double* B, int64_t N, int C, int H, int W, int F, int kh, int kw,
                      int Ho, int Wo, int py, int px, int flip);
int conv2d_full_lanes_try(double* out, const double* A, const double* B, int64_t N, int C, int H, int W, int F, int kh, int kw,
                          int flip);
int conv2d_wgrad_direct_try(double* dB, const double* A, const double* G, int64_t N, int C, int H, int W, int F, int kh,
                            int kw, int Ho, int Wo);
int g_conv_direct = 1;   // b200ag_conv_mode: 0 = implicit-GEMM DMMA kernels only
static int g_conv_direct_cmax = []() { const char* e = getenv("B200AG_CONV_DIRECT_CMAX"); return e ? atoi(e) : 4; }();

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct ConvArgs {
  double* out; const double* A; const double* B;
  int64_t N;
  int C, H, W, F, kh, kw, Ho, Wo, py, px, flip;
  int Hp, Wp;       // padded image size held in shared memory
  int K, K4;        // taps C*kh*kw and ceil(K/4)
  int P, TPI;       // positions per image Ho*Wo and 8-row tiles per image
  int G;            // images per CTA iteration
  int NBUF;         // 2: double-buffered image staging (cp.async prefetch), 1: single buffer
};

// out[n,f,y,x] = sum_{c,i,j} Apad[n,c,y+i,x+j] * Bk[c,f,i,j], Bk = B flipped when flip=1 (true convolution,
// signal.py:49-54); 'valid': no padding, 'full': Apad = A zero-padded by (kh-1, kw-1).
template <int MT, int NT>
__global__ void __launch_bounds__(256) conv2d_dmma_kernel(ConvArgs g) {
  extern __shared__ double smem[];
  constexpr int PB = NT * 8 + 4;                       // B-operand pitch: 4 mod 16 -> conflict-free fragment reads
  double* Wb = smem;                                   // [K4*4][PB]
  int* koff = (int*)(Wb + g.K4 * 4 * PB);              // [K4*4] tap -> offset inside a padded image
  double* imgs = (double*)(koff + ((g.K4 * 4 + 1) & ~1));   // [G][C*Hp*Wp]
  const int img_elems = g.C * g.Hp * g.Wp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int khw = g.kh * g.kw;

  for (int e = tid; e < g.K4 * 4 * PB; e += 256) {
    int k = e / PB, f = e % PB;
    double v = 0.0;
    if (k < g.K && f < g.F) {
      int c = k / khw, i = (k % khw) / g.kw, j = k % g.kw;
      int bi = g.flip ? g.kh - 1 - i : i, bj = g.flip ? g.kw - 1 - j : j;
      v = g.B[((c * g.F + f) * g.kh + bi) * g.kw + bj];
    }
    Wb[e] = v;
  }
  for (int k = tid; k < g.K4 * 4; k += 256) {
    int kk = k < g.K ? k : 0;
    int c = kk / khw, i = (kk % khw) / g.kw, j = kk % g.kw;
    koff[k] = (c * g.Hp + i) * g.Wp + j;
  }
  for (int e = tid; e < g.NBUF * g.G * img_elems; e += 256) imgs[e] = 0.0;   // borders stay zero for the whole kernel
  __syncthreads();

  const int ipi = (g.TPI + MT - 1) / MT;               // work items per image
  const int64_t groups = (g.N + g.G - 1) / g.G;
  const int hw = g.H * g.W;
  // stage(buf, grp): asynchronous copy (cp.async, 8 B) of a group's images into the interior of buffer `buf`
  auto stage = [&](int buf, int64_t grp) {
    const int64_t n0 = grp * g.G;
    const int ng = (int)((g.N - n0) < g.G ? (g.N - n0) : g.G);
    const double* src = g.A + n0 * g.C * hw;
    double* dst = imgs + (size_t)buf * g.G * img_elems;
    for (int e = tid; e < ng * g.C * hw; e += 256) {
      int x = e % g.W, y = (e / g.W) % g.H, ci = e / hw;     // ci = image*C + channel
      __pipeline_memcpy_async(dst + (ci * g.Hp + y + g.py) * g.Wp + x + g.px, src + e, 8);
    }
    __pipeline_commit();
  };
  int64_t grp = blockIdx.x;
  if (grp < groups) stage(0, grp);
  for (int it = 0; grp < groups; grp += gridDim.x, ++it) {
    const int buf = g.NBUF == 2 ? (it & 1) : 0;
    const int64_t next = grp + gridDim.x;
    if (g.NBUF == 2 && next < groups) {
      stage(buf ^ 1, next);                            // prefetch the next group while this one is multiplied
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int64_t n0 = grp * g.G;
    const int ng = (int)((g.N - n0) < g.G ? (g.N - n0) : g.G);
    const double* gimgs = imgs + (size_t)buf * g.G * img_elems;
    for (int item = warp; item < ng * ipi; item += 8) {
      const int gi = item / ipi, mt0 = (item % ipi) * MT;
      const double* img = gimgs + gi * img_elems;
      int moff[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        int m = (mt0 + t) * 8 + (lane >> 2);
        moff[t] = m < g.P ? (m / g.Wo) * g.Wp + (m % g.Wo) : 0;
      }
      double acc[MT][NT][2];
#pragma unroll
      for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int u = 0; u < NT; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
      for (int k4 = 0; k4 < g.K4; ++k4) {
        const int k = k4 * 4 + (lane & 3);
        const int ko = koff[k];
        double b[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) b[u] = Wb[k * PB + u * 8 + (lane >> 2)];
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const double a = img[moff[t] + ko];
#pragma unroll
          for (int u = 0; u < NT; ++u) dmma(acc[t][u][0], acc[t][u][1], a, b[u]);
        }
      }
      double* o = g.out + (n0 + gi) * (int64_t)g.F * g.P;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int m = (mt0 + t) * 8 + (lane >> 2);
        if (m >= g.P || mt0 + t >= g.TPI) continue;
#pragma unroll
        for (int u = 0; u < NT; ++u) {
          const int f = u * 8 + (lane & 3) * 2;
          if (f < g.F) o[(int64_t)f * g.P + m] = acc[t][u][0];
          if (f + 1 < g.F) o[(int64_t)(f + 1) * g.P + m] = acc[t][u][1];
        }
      }
    }
    __syncthreads();
    if (g.NBUF == 1 && next < groups) stage(0, next);
  }
}

struct WgradArgs {
  double* dB; const double* A; const double* G;
  int64_t N, imgs_per_cta;
  int C, H, W, F, kh, kw, Ho, Wo;
  int K, MTILES;    // taps and ceil(K/8)
  int P, P4;        // positions per image and ceil(P/4)
};

// dB[c,f,kh-1-i,kw-1-j] = sum_{n,y,x} A[n,c,y+i,x+j] * G[n,f,y,x]
template <int NT>
__global__ void __launch_bounds__(256) conv2d_wgrad_dmma_kernel(WgradArgs g) {
  extern __shared__ double smem[];
  const int a_elems = g.C * g.H * g.W + 1;               // +1: a zero cell that padded rows/positions read
  const int a_pitch = (a_elems + 1) & ~1;
  const int PG = g.P4 * 4 + 4;                            // pitch of G rows (positions), 4 mod 16 when P4*4 is 0 mod 16
  const int g_elems = NT * 8 * PG;
  double* As0 = smem;                                     // [2][C*H*W + 1]   (double-buffered)
  double* Gs0 = As0 + 2 * a_pitch;                        // [2][NT*8][PG]
  int* poff = (int*)(Gs0 + 2 * g_elems);                  // [P4*4] position -> offset y*W + x (or -1)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int khw = g.kh * g.kw;
  const int zero_cell = a_elems - 1;

  for (int p = tid; p < g.P4 * 4; p += 256) poff[p] = p < g.P ? (p / g.Wo) * g.W + (p % g.Wo) : -1;
  for (int e = tid; e < 2 * g_elems; e += 256) Gs0[e] = 0.0;   // rows f >= F and positions >= P stay zero
  if (tid < 2) As0[tid * a_pitch + zero_cell] = 0.0;

  // warp -> (row tiles, k-part): enough row tiles: warps stride over them; few row tiles: split the positions
  int ksplit = 1;
  while (g.MTILES * ksplit * 2 <= 8) ksplit *= 2;
  const int mt_first = warp % (8 / ksplit), mt_stride = 8 / ksplit, kpart = warp / (8 / ksplit);
  const int k4_per = (g.P4 + ksplit - 1) / ksplit;
  const int k4_lo = kpart * k4_per, k4_hi = (k4_lo + k4_per < g.P4) ? k4_lo + k4_per : g.P4;
  constexpr int MAXMT = 3;
  int roff[MAXMT];   // statically indexed everywhere so that it and acc[] stay in registers
#pragma unroll
  for (int t = 0; t < MAXMT; ++t) {
    int r = (mt_first + t * mt_stride) * 8 + (lane >> 2);
    roff[t] = -1;
    if (mt_first + t * mt_stride < g.MTILES && r < g.K) {
      int c = r / khw, i = (r % khw) / g.kw, j = r % g.kw;
      roff[t] = (c * g.H + i) * g.W + j;
    }
  }
  int nmt = 0;
  for (int mt = mt_first; mt < g.MTILES && nmt < MAXMT; mt += mt_stride) ++nmt;
  double acc[MAXMT][NT][2];
#pragma unroll
  for (int t = 0; t < MAXMT; ++t)
#pragma unroll
    for (int u = 0; u < NT; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;

  const int64_t n_lo = (int64_t)blockIdx.x * g.imgs_per_cta;
  const int64_t n_hi = (n_lo + g.imgs_per_cta < g.N) ? n_lo + g.imgs_per_cta : g.N;
  const int chw = g.C * g.H * g.W;
  // stage(buf, n): cp.async of image n (activations and output cotangent) into buffer `buf`
  auto stage = [&](int buf, int64_t n) {
    const double* a = g.A + n * chw;
    const double* gg = g.G + n * (int64_t)g.F * g.P;
    double* as = As0 + buf * a_pitch;
    double* gs = Gs0 + buf * g_elems;
    for (int e = tid; e < chw; e += 256) __pipeline_memcpy_async(as + e, a + e, 8);
    for (int e = tid; e < g.F * g.P; e += 256) __pipeline_memcpy_async(gs + (e / g.P) * PG + (e % g.P), gg + e, 8);
    __pipeline_commit();
  };
  __syncthreads();
  if (n_lo < n_hi) stage(0, n_lo);
  for (int64_t n = n_lo; n < n_hi; ++n) {
    const int buf = (int)((n - n_lo) & 1);
    if (n + 1 < n_hi) {
      stage(buf ^ 1, n + 1);                             // prefetch the next image behind this one's MMAs
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const double* As = As0 + buf * a_pitch;
    const double* Gs = Gs0 + buf * g_elems;
    for (int k4 = k4_lo; k4 < k4_hi; ++k4) {
      const int p = k4 * 4 + (lane & 3);
      const int po = poff[p];
      double b[NT];
#pragma unroll
      for (int u = 0; u < NT; ++u) b[u] = Gs[(u * 8 + (lane >> 2)) * PG + p];
#pragma unroll
      for (int t = 0; t < MAXMT; ++t) {
        if (t < nmt) {
          const int ro = roff[t];
          const double av = As[(ro >= 0 && po >= 0) ? ro + po : zero_cell];
#pragma unroll
          for (int u = 0; u < NT; ++u) dmma(acc[t][u][0], acc[t][u][1], av, b[u]);
        }
      }
    }
    __syncthreads();
  }
  // partial result of this CTA's chunk -> global (flipped tap order), FP64 atomics
#pragma unroll
  for (int t = 0; t < MAXMT; ++t) {
    const int mt = mt_first + t * mt_stride;
    const int r = mt * 8 + (lane >> 2);
    if (mt >= g.MTILES || r >= g.K) continue;
    const int c = r / khw, i = (r % khw) / g.kw, j = r % g.kw;
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const int f = u * 8 + (lane & 3) * 2;
      double* base = g.dB + ((int64_t)c * g.F * g.kh + (g.kh - 1 - i)) * g.kw + (g.kw - 1 - j);
      if (f < g.F) atomicAdd(base + (int64_t)f * khw, acc[t][u][0]);
      if (f + 1 < g.F) atomicAdd(base + (int64_t)(f + 1) * khw, acc[t][u][1]);
    }
  }
}

// ---------------------------------------------------------------- float32: plain CUDA-core kernels
template <typename T>
__global__ void __launch_bounds__(256) conv2d_simple_kernel(T* __restrict__ out, const T* __restrict__ A,
                                                             const T* __restrict__ B, int64_t N, int C, int H, int W,
                                                             int F, int kh, int kw, int Ho, int Wo, int py, int px,
                                                             int flip) {
  extern __shared__ unsigned char smem_raw[];
  T* Bs = (T*)smem_raw;
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
    int x = (int)(o % Wo), y = (int)((o / Wo) % Ho), f = (int)((o / ((int64_t)Wo * Ho)) % F);
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

template <typename T>
__global__ void __launch_bounds__(256) conv2d_wgrad_simple_kernel(T* __restrict__ dB, const T* __restrict__ A,
                                                                   const T* __restrict__ G, int64_t N, int C, int H,
                                                                   int W, int F, int kh, int kw, int Ho, int Wo) {
  // one CTA per filter tap (c,f,i,j); threads reduce over (n,y,x)
  __shared__ T part[256];
  int o = blockIdx.x;
  int j = o % kw, i = (o / kw) % kh, f = (o / (kw * kh)) % F, c = o / (kw * kh * F);
  T acc = 0;
  int64_t total = N * Ho * Wo;
  for (int64_t e = threadIdx.x; e < total; e += 256) {
    int x = (int)(e % Wo), y = (int)((e / Wo) % Ho);
    int64_t n = e / ((int64_t)Wo * Ho);
    acc += A[((n * C + c) * H + y + i) * W + x + j] * G[((n * F + f) * Ho + y) * Wo + x];
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dB[((c * F + f) * kh + (kh - 1 - i)) * kw + (kw - 1 - j)] = part[0];
}

template <int NT>
static int launch_conv_dmma(ConvArgs& g, size_t smem) {
  // rows-tiles-per-warp-item: pick the divisor of the per-image tile count that keeps 8 warps evenly busy
  int mt = 8;
  if (g.TPI % 9 == 0) mt = 9;
  else if (g.TPI % 8 == 0) mt = 8;
  else if (g.TPI % 6 == 0) mt = 6;
  else if (g.TPI % 4 == 0) mt = 4;
  int ipi = (g.TPI + mt - 1) / mt;
  // images per iteration: fill the 8 warps, bounded by shared memory
  size_t img_bytes = (size_t)g.C * g.Hp * g.Wp * sizeof(double);
  size_t fixed = smem;
  int G = (8 + ipi - 1) / ipi;
  int nbuf = 2;
  if (fixed + (size_t)2 * G * img_bytes > 200 * 1024) nbuf = 1;   // keep all 8 warps busy before double buffering
  while (G > 1 && fixed + (size_t)nbuf * G * img_bytes > 200 * 1024) --G;
  if (fixed + (size_t)nbuf * G * img_bytes > 200 * 1024) B200AG_FAIL("conv2d: image does not fit in shared memory");
  g.G = G;
  g.NBUF = nbuf;
  size_t total = fixed + (size_t)nbuf * G * img_bytes;
  int64_t groups = (g.N + G - 1) / G;
  int ctas_per_sm = (int)((220 * 1024) / (total + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  unsigned grid = (unsigned)(groups < cap ? groups : cap);
#define B200AG_CONV_CASE(MTV)                                                                                     \
  case MTV: {                                                                                                     \
    static bool cfg = false;                                                                                      \
    if (!cfg) {                                                                                                   \
      B200AG_CUDA(cudaFuncSetAttribute(conv2d_dmma_kernel<MTV, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                       200 * 1024 + 8192));                                                       \
      cfg = true;                                                                                                 \
    }                                                                                                             \
    conv2d_dmma_kernel<MTV, NT><<<grid, 256, total, stream()>>>(g);                                               \
    break;                                                                                                        \
  }
  switch (mt) {
    B200AG_CONV_CASE(9)
    B200AG_CONV_CASE(8)
    B200AG_CONV_CASE(6)
    B200AG_CONV_CASE(4)
  }
#undef B200AG_CONV_CASE
  B200AG_LAUNCH_CHECK();
  return 0;
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
  // few OUTPUT channels: the register-tiled direct kernel.  Forward layers qualify with few input channels (conv1 of
  // LeNet: C = 1); the 'full' correlation that is a layer's INPUT gradient (mode 1: conv2's dX has C = 16 cotangent
  // channels, F = 6) qualifies up to 16 because the direct kernel skips the taps that fall into the zero padding,
  // which the implicit-GEMM formulation multiplies (2.25x the true flops for a 5x5 filter over an 8x8 map).
  // a layer's INPUT gradient over a large batch: lanes = images, the taps that fall into the zero padding are skipped
  // exactly (conv_direct.cu); other shapes / small batches fall through
  if (dtype == B200AG_F64 && g_conv_direct && mode == 1) {
    int rc = conv2d_full_lanes_try((double*)out, (const double*)A, (const double*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh,
                                   (int)kw, flip);
    if (rc != 2) return rc;
  }
  if (dtype == B200AG_F64 && g_conv_direct && F <= 8 && (C <= 4 || (mode == 1 && C <= (g_conv_direct_cmax)))) {
    int rc = conv2d_direct_try((double*)out, (const double*)A, (const double*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh,
                               (int)kw, (int)Ho, (int)Wo, (int)py, (int)px, flip);
    if (rc != 2) return rc;
  }
  if (dtype == B200AG_F64 && F <= 16) {
    ConvArgs g;
    g.out = (double*)out; g.A = (const double*)A; g.B = (const double*)B;
    g.N = N; g.C = (int)C; g.H = (int)H; g.W = (int)W; g.F = (int)F; g.kh = (int)kh; g.kw = (int)kw;
    g.Ho = (int)Ho; g.Wo = (int)Wo; g.py = (int)py; g.px = (int)px; g.flip = flip;
    g.Hp = (int)(H + 2 * py); g.Wp = (int)(W + 2 * px);
    g.K = (int)(C * kh * kw); g.K4 = (g.K + 3) / 4;
    g.P = (int)(Ho * Wo); g.TPI = (g.P + 7) / 8;
    int NT = F <= 8 ? 1 : 2;
    int PB = NT * 8 + 4;
    size_t fixed = (size_t)g.K4 * 4 * PB * sizeof(double) + (size_t)((g.K4 * 4 + 1) & ~1) * sizeof(int);
    return NT == 1 ? launch_conv_dmma<1>(g, fixed) : launch_conv_dmma<2>(g, fixed);
  }
  size_t es = dtype_size(dtype);
  size_t smem = (size_t)(C * F * kh * kw) * es;
  if (smem > 160 * 1024) B200AG_FAIL("conv2d: filter bank does not fit in shared memory");
  unsigned grid = grid_for(total, 256, 8);
  if (dtype == B200AG_F64) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_simple_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfg = true; }
    conv2d_simple_kernel<double><<<grid, 256, smem, stream()>>>((double*)out, (const double*)A, (const double*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, (int)py, (int)px, flip);
  } else if (dtype == B200AG_F32) {
    static bool cfg = false;
    if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_simple_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfg = true; }
    conv2d_simple_kernel<float><<<grid, 256, smem, stream()>>>((float*)out, (const float*)A, (const float*)B, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo, (int)py, (int)px, flip);
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
  size_t es = dtype_size(dtype);
  if (es == 0) B200AG_FAIL("conv2d_wgrad: unsupported dtype");
  if (b200ag_memset(dB, 0, (size_t)nout * es)) return 1;
  if (N == 0 || nout == 0) return 0;
  if (dtype == B200AG_F64 && g_conv_direct && F <= 8 && C <= 4) {
    int rc = conv2d_wgrad_direct_try((double*)dB, (const double*)A, (const double*)G, N, (int)C, (int)H, (int)W, (int)F,
                                     (int)kh, (int)kw, (int)Ho, (int)Wo);
    if (rc != 2) return rc;
  }
  if (dtype == B200AG_F64 && F <= 16 && (C * kh * kw + 7) / 8 <= 24) {
    WgradArgs g;
    g.dB = (double*)dB; g.A = (const double*)A; g.G = (const double*)G;
    g.N = N; g.C = (int)C; g.H = (int)H; g.W = (int)W; g.F = (int)F; g.kh = (int)kh; g.kw = (int)kw;
    g.Ho = (int)Ho; g.Wo = (int)Wo;
    g.K = (int)(C * kh * kw); g.MTILES = (g.K + 7) / 8;
    g.P = (int)(Ho * Wo); g.P4 = (g.P + 3) / 4;
    int NT = F <= 8 ? 1 : 2;
    int PG = g.P4 * 4 + 4;
    size_t a_elems = (size_t)(C * H * W + 1);
    size_t smem = 2 * (((a_elems + 1) & ~size_t(1)) * sizeof(double) + (size_t)NT * 8 * PG * sizeof(double)) +
                  (size_t)g.P4 * 4 * sizeof(int);
    if (smem <= 200 * 1024) {
      int64_t ctas = (int64_t)sm_count() * 2;
      int64_t per = (N + ctas - 1) / ctas;
      if (per < 1) per = 1;
      g.imgs_per_cta = per;
      unsigned grid = (unsigned)((N + per - 1) / per);
      if (NT == 1) {
        static bool cfg = false;
        if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_wgrad_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
        conv2d_wgrad_dmma_kernel<1><<<grid, 256, smem, stream()>>>(g);
      } else {
        static bool cfg = false;
        if (!cfg) { B200AG_CUDA(cudaFuncSetAttribute(conv2d_wgrad_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg = true; }
        conv2d_wgrad_dmma_kernel<2><<<grid, 256, smem, stream()>>>(g);
      }
      B200AG_LAUNCH_CHECK();
      return 0;
    }
  }
  if (dtype == B200AG_F64)
    conv2d_wgrad_simple_kernel<double><<<(unsigned)nout, 256, 0, stream()>>>((double*)dB, (const double*)A, (const double*)G, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo);
  else if (dtype == B200AG_F32)
    conv2d_wgrad_simple_kernel<float><<<(unsigned)nout, 256, 0, stream()>>>((float*)dB, (const float*)A, (const float*)G, N, (int)C, (int)H, (int)W, (int)F, (int)kh, (int)kw, (int)Ho, (int)Wo);
  else
    B200AG_FAIL("conv2d_wgrad: unsupported dtype");
  B200AG_LAUNCH_CHECK();
  return 0;
}

// 1 (default): layers with few output channels use the direct CUDA-core kernels of conv_direct.cu; 0: always the
// implicit-GEMM DMMA kernels.  Returns the previous setting through *previous (may be null).
int b200ag_conv_mode(int direct, int* previous) {
  if (previous) *previous = g_conv_direct;
  if (direct == 0 || direct == 1) g_conv_direct = direct;
  return 0;
}

}  // extern "C"
