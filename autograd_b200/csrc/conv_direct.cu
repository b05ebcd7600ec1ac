// Direct (register-tiled, CUDA-core FP64) convolution kernels for layers with FEW output channels.
//
// The implicit-GEMM DMMA kernels in conv.cu pad the channel dimension to 8 and the tap dimension to a multiple of 4
// and gather their A fragments through an offset table; for a layer like LeNet's conv1 (C=1, F=6, 5x5,
// examples/convnet.py:154-162) that staging, not the math, is the cost (1.39 ms forward / 2.52 ms weight-gradient
// for 65,536 images against an HBM floor of 0.34 ms).  FP64 DMMA and FP64 DFMA have the same peak on B200, so with
// few channels a direct kernel loses nothing and skips the padding:
//   forward / input-gradient: a thread owns TH=4 vertically adjacent output pixels for all F channels
//       (F*TH accumulators); for every (c, j) it reads one column of TH+kh-1 input values (consecutive lanes =
//       consecutive x: conflict-free) and the kh*F weights of that column (broadcast) -> kh*F*TH DFMAs.
//   weight-gradient: a thread owns the F*KW sums of one (c, i) tap row for a slice of output positions and keeps
//       them in registers across ALL images of its CTA; the input window slides along x (one new load per position).
// Same contract as b200ag_conv2d / b200ag_conv2d_wgrad (autograd/scipy/signal.py:10-217); conv.cu calls these first
// and falls back to the DMMA kernels when the shape is outside the instantiated set.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace b200ag {


struct DirectArgs {
  double* out; const double* A; const double* B;
  int64_t N;
  int C, H, W, kh, Ho, Wo, py, px, flip;
  int IMG;          // images per CTA iteration
};

template <int F, int KW, int TH>
__global__ void __launch_bounds__(192) conv2d_direct_kernel(DirectArgs g) {
  extern __shared__ double smem[];
  const int wcount = g.C * KW * g.kh * F;
  const int chw = g.C * g.H * g.W;
  const int buf_elems = (g.IMG * chw + 1) & ~1;
  double* Ws = smem;                                   // [c][j][i][f]
  double* imgs = smem + ((wcount + 1) & ~1);           // [2][IMG][C*H*W]: double-buffered (cp.async prefetch)
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < wcount; e += nthr) {
    int f = e % F, i = (e / F) % g.kh, j = (e / (F * g.kh)) % KW, c = e / (F * g.kh * KW);
    int bi = g.flip ? g.kh - 1 - i : i, bj = g.flip ? KW - 1 - j : j;
    Ws[e] = g.B[((c * F + f) * g.kh + bi) * KW + bj];
  }
  const int strips_y = (g.Ho + TH - 1) / TH;
  const int strips = strips_y * g.Wo;
  const int64_t groups = (g.N + g.IMG - 1) / g.IMG;
  auto stage = [&](int buf, int64_t grp) {
    const int64_t n0 = grp * g.IMG;
    const int ng = (int)((g.N - n0) < g.IMG ? (g.N - n0) : g.IMG);
    const double* src = g.A + n0 * chw;
    double* dst = imgs + buf * buf_elems;
    const int total = ng * chw;
    if ((((n0 * chw) | (int64_t)(buf * buf_elems)) & 1) == 0) {
      for (int e = tid; e < total / 2; e += nthr) __pipeline_memcpy_async(dst + 2 * e, src + 2 * e, 16);
      if ((total & 1) && tid == 0) __pipeline_memcpy_async(dst + total - 1, src + total - 1, 8);
    } else {
      for (int e = tid; e < total; e += nthr) __pipeline_memcpy_async(dst + e, src + e, 8);
    }
    __pipeline_commit();
  };
  int64_t grp = blockIdx.x;
  if (grp < groups) stage(0, grp);
  for (int it = 0; grp < groups; grp += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t next = grp + gridDim.x;
    if (next < groups) {
      stage(buf ^ 1, next);                            // the other buffer was released by the barrier ending the last round
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int64_t n0 = grp * g.IMG;
    const int ng = (int)((g.N - n0) < g.IMG ? (g.N - n0) : g.IMG);
    const double* gimgs = imgs + buf * buf_elems;
    for (int item = tid; item < ng * strips; item += nthr) {
      const int gi = item / strips, s = item - gi * strips;
      const int sy = s / g.Wo, x = s - sy * g.Wo;
      const int y0 = sy * TH;
      const double* img = gimgs + gi * chw;
      double acc[F][TH];
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int t = 0; t < TH; ++t) acc[f][t] = 0.0;
      // filter rows whose whole TH-row window lies in the zero padding contribute nothing ('full' mode: the input
      // gradient of a layer): row i reads input rows y0 + i - py .. y0 + i - py + TH - 1
      int i_lo = g.py - y0 - (TH - 1);
      if (i_lo < 0) i_lo = 0;
      int i_hi = g.H - 1 + g.py - y0;
      if (i_hi > g.kh - 1) i_hi = g.kh - 1;
      for (int c = 0; c < g.C; ++c) {
        const double* plane = img + c * g.H * g.W;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
          const int xx = x + j - g.px;
          if (xx < 0 || xx >= g.W) continue;
          const double* wcol = Ws + (c * KW + j) * g.kh * F;
          double prev[TH];                            // sliding window over the rows: input rows y0+i-py .. +TH-1
#pragma unroll
          for (int t = 0; t < TH; ++t) {
            const int yy = y0 + t + i_lo - g.py;
            prev[t] = (yy >= 0 && yy < g.H) ? plane[yy * g.W + xx] : 0.0;
          }
          for (int i = i_lo; i <= i_hi; ++i) {
            const double* wv = wcol + i * F;
            double w[F];
            if constexpr (F % 2 == 0) {                // [i][f] rows are 16-byte aligned: 128-bit broadcast loads
#pragma unroll
              for (int f = 0; f < F; f += 2) {
                const double2 w2 = *reinterpret_cast<const double2*>(wv + f);
                w[f] = w2.x; w[f + 1] = w2.y;
              }
            } else {
#pragma unroll
              for (int f = 0; f < F; ++f) w[f] = wv[f];
            }
#pragma unroll
            for (int f = 0; f < F; ++f)
#pragma unroll
              for (int t = 0; t < TH; ++t) acc[f][t] = fma(prev[t], w[f], acc[f][t]);
#pragma unroll
            for (int t = 0; t + 1 < TH; ++t) prev[t] = prev[t + 1];
            const int yy = y0 + TH + i - g.py;
            prev[TH - 1] = (yy >= 0 && yy < g.H) ? plane[yy * g.W + xx] : 0.0;
          }
        }
      }
      double* o = g.out + ((n0 + gi) * F) * (int64_t)(g.Ho * g.Wo) + x;
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int t = 0; t < TH; ++t)
          if (y0 + t < g.Ho) o[(int64_t)f * g.Ho * g.Wo + (y0 + t) * g.Wo] = acc[f][t];
    }
    __syncthreads();                                   // this buffer may be refilled by the next round's prefetch
  }
}

template <int F, int KW, int TH>
static int launch_direct(DirectArgs g) {
  const int strips = ((g.Ho + TH - 1) / TH) * g.Wo;
  const size_t wbytes = (size_t)((g.C * KW * g.kh * F + 1) & ~1) * 8;
  const size_t img_bytes = (size_t)g.C * g.H * g.W * 8;
  int img = 1;
  while ((img + 1) * strips <= 192 && wbytes + 2 * (size_t)(img + 1) * img_bytes + 16 <= 64 * 1024) ++img;
  if (wbytes + 2 * (size_t)img * img_bytes + 16 > 200 * 1024) return 2;
  int threads = (img * strips + 31) / 32 * 32;
  if (threads > 192) threads = 192;
  if (threads < 64) threads = 64;
  g.IMG = img;
  const size_t smem = wbytes + 2 * (size_t)((img * (size_t)g.C * g.H * g.W + 1) & ~size_t(1)) * 8;
  static bool cfg = false;
  if (!cfg) {
    B200AG_CUDA(cudaFuncSetAttribute(conv2d_direct_kernel<F, KW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const int64_t groups = (g.N + img - 1) / img;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  conv2d_direct_kernel<F, KW, TH><<<(unsigned)(groups < cap ? groups : cap), threads, smem, stream()>>>(g);
  B200AG_LAUNCH_CHECK();
  return 0;
}

// returns 0 = done, 1 = error, 2 = shape not covered (caller falls back to the implicit-GEMM kernel)
int conv2d_direct_try(double* out, const double* A, const double* B, int64_t N, int C, int H, int W, int F, int kh, int kw,
                      int Ho, int Wo, int py, int px, int flip) {
  DirectArgs g{out, A, B, N, C, H, W, kh, Ho, Wo, py, px, flip, 1};
#define B200AG_DIRECT_CASE(FV, KV) \
  if (F == FV && kw == KV) return (Ho >= 16 && FV <= 6) ? launch_direct<FV, KV, 8>(g) : launch_direct<FV, KV, 4>(g);
  B200AG_DIRECT_CASE(1, 3) B200AG_DIRECT_CASE(2, 3) B200AG_DIRECT_CASE(3, 3) B200AG_DIRECT_CASE(4, 3)
  B200AG_DIRECT_CASE(6, 3) B200AG_DIRECT_CASE(8, 3)
  B200AG_DIRECT_CASE(1, 5) B200AG_DIRECT_CASE(2, 5) B200AG_DIRECT_CASE(3, 5) B200AG_DIRECT_CASE(4, 5)
  B200AG_DIRECT_CASE(6, 5) B200AG_DIRECT_CASE(8, 5)
#undef B200AG_DIRECT_CASE
  return 2;
}

// ---------------------------------------------------------------------------------------------- weight gradient
struct WDirectArgs {
  double* dB; const double* A; const double* G;
  int64_t N;
  int C, H, W, kh, Ho, Wo;
  int SX;            // x-slices per output row
  int64_t imgs_per_cta;
};

// dB[c,f,kh-1-i,kw-1-j] = sum_{n,y,x} A[n,c,y+i,x+j] * G[n,f,y,x]   (gradient of the TRUE convolution: flipped taps)
template <int F, int KW>
__global__ void __launch_bounds__(256) conv2d_wgrad_direct_kernel(WDirectArgs g) {
  extern __shared__ double smem[];
  const int P = g.Ho * g.Wo;
  const int chw = g.C * g.H * g.W;
  const int a_elems = (chw + 1) & ~1;
  const int buf_elems = a_elems + ((F * P + 1) & ~1);  // one staged image: As [C][H][W] then Gs [y][x][F]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int CI = g.C * g.kh;
  const int slices = g.Ho * g.SX;
  const int workers = CI * slices;
  const int ci = tid % CI, sl = tid / CI;               // ci fastest: lanes of a warp share (y, x) -> broadcast G reads
  const int c = ci / g.kh, i = ci - c * g.kh;
  const int y = sl / g.SX, sx = sl - y * g.SX;
  const int xw = (g.Wo + g.SX - 1) / g.SX;
  const int x_lo = sx * xw, x_hi = min(x_lo + xw, g.Wo);
  const bool active = tid < workers;
  double acc[F][KW];
#pragma unroll
  for (int f = 0; f < F; ++f)
#pragma unroll
    for (int j = 0; j < KW; ++j) acc[f][j] = 0.0;
  const int64_t n_begin = (int64_t)blockIdx.x * g.imgs_per_cta;
  const int64_t n_end = min(n_begin + g.imgs_per_cta, g.N);
  auto stage = [&](int buf, int64_t n) {
    double* As = smem + buf * buf_elems;
    double* Gs = As + a_elems;
    const double* a = g.A + n * chw;
    for (int e = tid; e < chw; e += nthr) __pipeline_memcpy_async(As + e, a + e, 8);
    const double* gg = g.G + n * (int64_t)F * P;
    for (int e = tid; e < F * P; e += nthr) {            // coalesced read of [f][p], transposed landing at [p][f]
      const int f = e / P, p = e - f * P;
      __pipeline_memcpy_async(Gs + p * F + f, gg + e, 8);
    }
    __pipeline_commit();
  };
  if (n_begin < n_end) stage(0, n_begin);
  for (int64_t n = n_begin; n < n_end; ++n) {
    const int buf = (int)((n - n_begin) & 1);
    if (n + 1 < n_end) {
      stage(buf ^ 1, n + 1);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    if (active) {
      const double* As = smem + buf * buf_elems;
      const double* Gs = As + a_elems;
      const double* arow = As + (c * g.H + y + i) * g.W;
      double win[KW];
#pragma unroll
      for (int j = 0; j + 1 < KW; ++j) win[j + 1] = arow[x_lo + j];   // win[1..KW-1] = a[x_lo .. x_lo+KW-2]
      for (int x = x_lo; x < x_hi; ++x) {
#pragma unroll
        for (int j = 0; j + 1 < KW; ++j) win[j] = win[j + 1];
        win[KW - 1] = arow[x + KW - 1];
        const double* gv = Gs + (y * g.Wo + x) * F;
        double gval[F];
        if constexpr (F % 2 == 0) {
#pragma unroll
          for (int f = 0; f < F; f += 2) {
            const double2 g2 = *reinterpret_cast<const double2*>(gv + f);
            gval[f] = g2.x; gval[f + 1] = g2.y;
          }
        } else {
#pragma unroll
          for (int f = 0; f < F; ++f) gval[f] = gv[f];
        }
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
          for (int j = 0; j < KW; ++j) acc[f][j] = fma(win[j], gval[f], acc[f][j]);
      }
    }
    __syncthreads();                                   // this buffer may be refilled by the next round's prefetch
  }
  // reduce the slices' partial sums: shared-memory staging, then one FP64 atomic per output and CTA
  double* red = smem;                                   // [workers][F*KW], reuses the staging area
  if (active) {
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int j = 0; j < KW; ++j) red[tid * (F * KW) + f * KW + j] = acc[f][j];
  }
  __syncthreads();
  const int nout = CI * F * KW;
  for (int o = tid; o < nout; o += nthr) {
    const int oci = o / (F * KW), r = o - oci * (F * KW);
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += red[(k * CI + oci) * (F * KW) + r];
    const int oc = oci / g.kh, oi = oci - oc * g.kh, f = r / KW, j = r - f * KW;
    atomicAdd(g.dB + ((oc * F + f) * g.kh + (g.kh - 1 - oi)) * KW + (KW - 1 - j), s);
  }
}

template <int F, int KW>
static int launch_wgrad_direct(WDirectArgs g) {
  const int CI = g.C * g.kh;
  int sx = 1;
  while (CI * g.Ho * (sx * 2) <= 256 && g.Wo / (sx * 2) >= 8) sx *= 2;
  const int workers = CI * g.Ho * sx;
  if (workers > 256 || workers < 32) return 2;
  g.SX = sx;
  const int threads = (workers + 31) / 32 * 32;
  const size_t stage = 2 * (size_t)(((g.C * g.H * g.W + 1) & ~1) + ((F * g.Ho * g.Wo + 1) & ~1)) * 8;
  const size_t red = (size_t)workers * F * KW * 8;
  const size_t smem = stage > red ? stage : red;
  if (smem > 200 * 1024) return 2;
  static bool cfg = false;
  if (!cfg) {
    B200AG_CUDA(cudaFuncSetAttribute(conv2d_wgrad_direct_kernel<F, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int64_t ctas = (int64_t)sm_count() * per_sm;
  int64_t per = (g.N + ctas - 1) / ctas;
  if (per < 1) per = 1;
  g.imgs_per_cta = per;
  conv2d_wgrad_direct_kernel<F, KW><<<(unsigned)((g.N + per - 1) / per), threads, smem, stream()>>>(g);
  B200AG_LAUNCH_CHECK();
  return 0;
}

// dB must be zeroed by the caller.  Returns 0 = done, 1 = error, 2 = shape not covered.
int conv2d_wgrad_direct_try(double* dB, const double* A, const double* G, int64_t N, int C, int H, int W, int F, int kh,
                            int kw, int Ho, int Wo) {
  if (F * kw > 40) return 2;   // accumulators must stay in registers
  WDirectArgs g{dB, A, G, N, C, H, W, kh, Ho, Wo, 1, 1};
#define B200AG_WDIRECT_CASE(FV, KV) \
  if (F == FV && kw == KV) return launch_wgrad_direct<FV, KV>(g);
  B200AG_WDIRECT_CASE(1, 3) B200AG_WDIRECT_CASE(2, 3) B200AG_WDIRECT_CASE(3, 3) B200AG_WDIRECT_CASE(4, 3)
  B200AG_WDIRECT_CASE(6, 3) B200AG_WDIRECT_CASE(8, 3)
  B200AG_WDIRECT_CASE(1, 5) B200AG_WDIRECT_CASE(2, 5) B200AG_WDIRECT_CASE(3, 5) B200AG_WDIRECT_CASE(4, 5)
  B200AG_WDIRECT_CASE(6, 5) B200AG_WDIRECT_CASE(8, 5)
#undef B200AG_WDIRECT_CASE
  return 2;
}

}  // namespace b200ag
