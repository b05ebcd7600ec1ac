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

#include <stdlib.h>

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
    __pipeline_wait_prior(0);                          // this thread's copies of this group (issued a round ago) landed
    __syncthreads();                                   // ... everyone's did, and everyone is done with the previous group,
    if (next < groups) stage(buf ^ 1, next);           // whose buffer the next group now streams into
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
  // several rounds of work items per staged group: the two barriers of a group are paid once per `rounds` items of a
  // thread instead of once per item (register-limited to ~10 warps per SM, the kernel spent as long at barriers as
  // issuing: ncu stall_barrier 1.03 per issue at one round).  C5, scripts/gpu_r2_conv1.sh: 8.68 / 8.53 / 8.60 / 8.59 ms at
  // 1 / 2 / 3 / 4 rounds
  static const int max_rounds = []() { const char* e = getenv("B200AG_CONV_DIRECT_ROUNDS"); return e ? atoi(e) : 2; }();
  int rounds = 1;
  while (rounds < max_rounds && wbytes + 2 * (size_t)(img * (rounds + 1)) * img_bytes + 16 <= 100 * 1024) ++rounds;
  img *= rounds;
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

// ------------------------------------------------------------------------------- 'full' correlation, lane = image
// The INPUT gradient of a convolution layer is the 'full' correlation of the output cotangent with the filters
// (autograd/scipy/signal.py:153-190): out[n,f,y,x] = sum_{c,a,j} A[n,c,y-a,x+j-(KW-1)] * Bk[c,f,KH-1-a,j] with everything
// outside A zero.  For a 5x5 filter over an 8x8 map only 44 % of the (position, tap) pairs touch A at all; a kernel whose
// lanes are neighbouring POSITIONS cannot skip the rest (each lane has its own valid range, the warp runs the union), and
// the implicit-GEMM kernel multiplies the zero padding outright.  Here the 32 lanes of a warp are 32 IMAGES at the same
// output column x: which taps exist is the same for every lane, so the skipping is exact and free of divergence, the
// filter values are warp-wide broadcasts, and the image values are conflict-free (odd image pitch).  A thread owns one
// output column (all F channels) in two passes of HALF rows; the valid (row, tap) pairs of a pass are enumerated at
// compile time.  Channels are staged CC at a time through a two-stage cp.async ring.
struct FullArgs {
  double* out; const double* A; const double* B;
  int64_t N;
  int C, flip;
};

template <int F, int KH, int KW, int H, int W>
struct FullCfg {
  static constexpr int HO = H + KH - 1, WO = W + KW - 1;
  static constexpr int HALF = (HO + 1) / 2;
  static constexpr int CC = 4;                       // channels per stage
  static constexpr int HW = H * W;
  static constexpr int PITCH = CC * HW + 1;          // odd: lane r reads bank (r + const) mod 16
  static constexpr int THREADS = 32 * WO;
};

template <int F, int KH, int KW, int H, int W, int P>
__device__ __forceinline__ void full_lanes_pass(const double* __restrict__ in, const double* __restrict__ Ws, int cc_n,
                                                int c0, int x, double (&acc)[F][FullCfg<F, KH, KW, H, W>::HALF]) {
  using Cfg = FullCfg<F, KH, KW, H, W>;
  constexpr int HALF = Cfg::HALF;
  constexpr int Y_LO = P * HALF, Y_HI = (Y_LO + HALF < Cfg::HO) ? Y_LO + HALF : Cfg::HO;
  constexpr int YY_LO = (Y_LO - (KH - 1) > 0) ? Y_LO - (KH - 1) : 0, YY_HI = (Y_HI < H) ? Y_HI : H;   // input rows used
  for (int cc = 0; cc < cc_n; ++cc) {
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      const int xx = x + j - (KW - 1);
      if (xx < 0 || xx >= W) continue;                // warp-uniform
      const double* col = in + cc * Cfg::HW + xx;
      double v[YY_HI - YY_LO];
#pragma unroll
      for (int r = 0; r < YY_HI - YY_LO; ++r) v[r] = col[(YY_LO + r) * W];
      const double* wj = Ws + (((c0 + cc) * KW + j) * KH) * F;
#pragma unroll
      for (int a = 0; a < KH; ++a) {
        double w[F];
        if constexpr (F % 2 == 0) {
#pragma unroll
          for (int f = 0; f < F; f += 2) {
            const double2 w2 = *reinterpret_cast<const double2*>(wj + a * F + f);
            w[f] = w2.x; w[f + 1] = w2.y;
          }
        } else {
#pragma unroll
          for (int f = 0; f < F; ++f) w[f] = wj[a * F + f];
        }
#pragma unroll
        for (int t = 0; t < HALF; ++t) {
          const int yy = Y_LO + t - a;                // compile-time after unrolling
          if (Y_LO + t < Y_HI && yy >= YY_LO && yy < YY_HI) {
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f][t] = fma(v[yy - YY_LO], w[f], acc[f][t]);
          }
        }
      }
    }
  }
}

template <int F, int KH, int KW, int H, int W>
__global__ void __launch_bounds__(FullCfg<F, KH, KW, H, W>::THREADS, 1) conv2d_full_lanes_kernel(FullArgs g) {
  using Cfg = FullCfg<F, KH, KW, H, W>;
  constexpr int HALF = Cfg::HALF, CC = Cfg::CC, PITCH = Cfg::PITCH, NT = Cfg::THREADS;
  extern __shared__ double smem[];
  const int wcount = g.C * KW * KH * F;
  double* Ws = smem;                                   // [c][j][a][f], a = KH-1-i
  double* bufs = smem + ((wcount + 1) & ~1);           // [2][32][PITCH]
  const int tid = threadIdx.x, lane = tid & 31, x = tid >> 5;
  for (int e = tid; e < wcount; e += NT) {
    const int f = e % F, a = (e / F) % KH, j = (e / (F * KH)) % KW, c = e / (F * KH * KW);
    const int i = KH - 1 - a;
    const int bi = g.flip ? KH - 1 - i : i, bj = g.flip ? KW - 1 - j : j;
    Ws[e] = g.B[((c * F + f) * KH + bi) * KW + bj];
  }
  const int nchunks = (g.C + CC - 1) / CC;
  const int64_t groups = (g.N + 31) / 32;
  const int64_t my_groups = blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t steps = my_groups * 2 * nchunks;       // (group, pass, chunk), chunk fastest
  // stage(step): the CC channels of chunk `step % nchunks` of the step's 32 images -> ring slot step & 1
  auto stage = [&](int64_t step) {
    const int chunk = (int)(step % nchunks);
    const int64_t n0 = (blockIdx.x + (step / (2 * nchunks)) * gridDim.x) * 32;
    const int cc_n = (g.C - chunk * CC) < CC ? (g.C - chunk * CC) : CC;
    double* dst = bufs + (step & 1) * (32 * PITCH);
    const int per_img = cc_n * Cfg::HW;
    for (int e = tid; e < 32 * per_img; e += NT) {
      const int r = e / per_img, o = e - r * per_img;
      if (n0 + r < g.N)
        __pipeline_memcpy_async(dst + r * PITCH + o, g.A + ((n0 + r) * g.C + chunk * CC) * Cfg::HW + o, 8);
    }
    __pipeline_commit();
  };
  if (steps > 0) stage(0);
  double acc[F][HALF];
  for (int64_t step = 0; step < steps; ++step) {
    const int chunk = (int)(step % nchunks), pass = (int)((step / nchunks) & 1);
    const int64_t n0 = (blockIdx.x + (step / (2 * nchunks)) * gridDim.x) * 32;
    __pipeline_wait_prior(0);                          // this thread's copies of step `step` (issued one step ago) landed
    __syncthreads();                                   // ... everyone's did, and everyone is done computing step - 1,
    if (step + 1 < steps) stage(step + 1);             // whose slot the next step's operands now stream into
    if (chunk == 0) {
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int t = 0; t < HALF; ++t) acc[f][t] = 0.0;
    }
    const double* in = bufs + (step & 1) * (32 * PITCH) + lane * PITCH;
    const int cc_n = (g.C - chunk * CC) < CC ? (g.C - chunk * CC) : CC;
    if (pass == 0) full_lanes_pass<F, KH, KW, H, W, 0>(in, Ws, cc_n, chunk * CC, x, acc);
    else full_lanes_pass<F, KH, KW, H, W, 1>(in, Ws, cc_n, chunk * CC, x, acc);
    if (chunk == nchunks - 1 && n0 + lane < g.N) {
      double* o = g.out + (n0 + lane) * (int64_t)(F * Cfg::HO * Cfg::WO) + x;
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int t = 0; t < HALF; ++t) {
          const int y = pass * HALF + t;
          if (y < Cfg::HO) o[(f * Cfg::HO + y) * Cfg::WO] = acc[f][t];
        }
    }
  }
}

template <int F, int KH, int KW, int H, int W>
static int launch_full_lanes(FullArgs g) {
  using Cfg = FullCfg<F, KH, KW, H, W>;
  const size_t wbytes = (size_t)((g.C * KW * KH * F + 1) & ~1) * 8;
  const size_t smem = wbytes + 2 * (size_t)32 * Cfg::PITCH * 8;
  if (smem > 220 * 1024) return 2;
  static bool cfg = false;
  if (!cfg) {
    B200AG_CUDA(cudaFuncSetAttribute(conv2d_full_lanes_kernel<F, KH, KW, H, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     220 * 1024));
    cfg = true;
  }
  const int64_t groups = (g.N + 31) / 32;
  const int64_t cap = sm_count();
  conv2d_full_lanes_kernel<F, KH, KW, H, W><<<(unsigned)(groups < cap ? groups : cap), Cfg::THREADS, smem, stream()>>>(g);
  B200AG_LAUNCH_CHECK();
  return 0;
}

// 'full' correlation with lanes = images; returns 0 = done, 1 = error, 2 = shape not instantiated / batch too small
int conv2d_full_lanes_try(double* out, const double* A, const double* B, int64_t N, int C, int H, int W, int F, int kh, int kw,
                          int flip) {
  static const bool on = []() { const char* e = getenv("B200AG_CONV_FULL_LANES"); return !e || atoi(e) != 0; }();
  if (!on || N < 32 * 8) return 2;                     // small batches: the position-parallel kernels fill the machine better
  FullArgs g{out, A, B, N, C, flip};
#define B200AG_FULL_CASE(FV, KV, HV) \
  if (F == FV && kh == KV && kw == KV && H == HV && W == HV) return launch_full_lanes<FV, KV, KV, HV, HV>(g);
  B200AG_FULL_CASE(6, 5, 8) B200AG_FULL_CASE(4, 5, 8) B200AG_FULL_CASE(8, 5, 8) B200AG_FULL_CASE(3, 5, 8)
  B200AG_FULL_CASE(6, 3, 8) B200AG_FULL_CASE(4, 3, 8) B200AG_FULL_CASE(8, 3, 8)
  B200AG_FULL_CASE(6, 3, 10) B200AG_FULL_CASE(4, 3, 10) B200AG_FULL_CASE(6, 5, 4) B200AG_FULL_CASE(6, 3, 6)
#undef B200AG_FULL_CASE
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
  const int buf_elems = a_elems + ((F * P + 1) & ~1);  // one staged image: As [C][H][W] then Gs [F][y][x]
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
    // [f][p] as it lies in HBM: lanes at neighbouring x then read neighbouring words.  (Landing it transposed, [p][f], for
    // 128-bit reads of the F values was measured slower: the scattered landing and the wide reads kept the shared-memory
    // pipe 72 % busy; C5 8.60 -> 8.46 ms without it.)
    for (int e = tid; e < F * P; e += nthr) __pipeline_memcpy_async(Gs + e, gg + e, 8);
    __pipeline_commit();
  };
  if (n_begin < n_end) stage(0, n_begin);
  for (int64_t n = n_begin; n < n_end; ++n) {
    const int buf = (int)((n - n_begin) & 1);
    __pipeline_wait_prior(0);
    __syncthreads();                                   // image n is complete and everyone is done with image n - 1
    if (n + 1 < n_end) stage(buf ^ 1, n + 1);
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
        double gval[F];
        const double* gv = Gs + y * g.Wo + x;
#pragma unroll
        for (int f = 0; f < F; ++f) gval[f] = gv[f * P];
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
          for (int j = 0; j < KW; ++j) acc[f][j] = fma(win[j], gval[f], acc[f][j]);
      }
    }
  }
  __pipeline_wait_prior(0);
  __syncthreads();                                     // the staging area is reused for the reduction below
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
