// FP32-accurate convolution back end (CUDA-core implicit GEMM) + the small layout kernels.
//
// Role: (1) the fp32 precision mode of PoseHigherResolutionNet.forward -- the reference's
// fp32 tolerance (<=1e-4 of the tensor max) rules out single-pass TF32 (SURVEY.md hard part
// 2), so fp32 activations/weights are multiplied with FFMA and fp32 accumulation; (2) the
// stem conv1 (Cin = 3, bandwidth bound) in both modes; (3) any layer shape the tcgen05
// back end does not take.  Same brtpe_conv_desc contract as conv_umma.cu: NHWC
// activations, a tap table (3x3 / 1x1 / stride 2 / one parity phase of the 4x4 s2
// transposed conv), folded-BN bias, optional residual and ReLU in the epilogue.
#include "conv_common.cuh"
#include <stdlib.h>

#include <algorithm>

namespace brtpe {

constexpr int FBM = 64, FBN = 64, FBK = 16, FTHREADS = 256;

struct FfmaArgs {
  brtpe_conv_desc d;
  const void* in;
  const float* w;      // [ntaps][Cin][Cout]
  const float* bias;
  const void* res;
  void* out;
  int M;               // N*Hm*Wm
};

template <typename TA>
__global__ void __launch_bounds__(FTHREADS) conv_ffma_kernel(FfmaArgs a) {
  __shared__ float As[FBK][FBM + 4];
  __shared__ float Bs[FBK][FBN + 4];
  __shared__ int s_pix_n[FBM], s_pix_y[FBM], s_pix_x[FBM];

  const brtpe_conv_desc& d = a.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FBM;
  const int n0 = blockIdx.y * FBN;
  const TA* __restrict__ in = reinterpret_cast<const TA*>(a.in);

  if (tid < FBM) {
    int m = m0 + tid;
    if (m < a.M) {
      const int hw = d.Hm * d.Wm;
      const int n = m / hw;
      const int r = m - n * hw;
      s_pix_n[tid] = n;
      s_pix_y[tid] = r / d.Wm;
      s_pix_x[tid] = r - (r / d.Wm) * d.Wm;
    } else {
      s_pix_n[tid] = -1;
      s_pix_y[tid] = 0;
      s_pix_x[tid] = 0;
    }
  }
  __syncthreads();

  const int tx = tid & 15;   // 16 column groups of 4 couts
  const int ty = tid >> 4;   // 16 row groups of 4 pixels
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // A loader: thread -> (pixel = tid / 4, k chunk = (tid % 4) * 4)
  const int a_pix = tid >> 2;
  const int a_k = (tid & 3) * 4;
  // B loader: thread -> (k = tid / 16, cout chunk = (tid % 16) * 4)
  const int b_k = tid >> 4;
  const int b_c = (tid & 15) * 4;

  for (int tap = 0; tap < d.ntaps; ++tap) {
    const int pn = s_pix_n[a_pix];
    const int iy = s_pix_y[a_pix] * d.in_stride + d.tap_dy[tap];
    const int ix = s_pix_x[a_pix] * d.in_stride + d.tap_dx[tap];
    const bool pvalid = (pn >= 0) && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win;
    const TA* __restrict__ prow =
        in + (((size_t)(pn < 0 ? 0 : pn) * d.Hin + (pvalid ? iy : 0)) * d.Win + (pvalid ? ix : 0)) *
                 (size_t)d.in_ld + d.in_coff;
    const float* __restrict__ wtap = a.w + (size_t)tap * d.Cin * d.Cout;
    for (int c0 = 0; c0 < d.Cin; c0 += FBK) {
      float av[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + a_k + e;
        av[e] = (pvalid && c < d.Cin) ? to_f32(prow[c]) : 0.0f;
      }
      float bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + b_k;
        const int co = n0 + b_c + e;
        bv[e] = (c < d.Cin && co < d.Cout) ? __ldg(wtap + (size_t)c * d.Cout + co) : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) As[a_k + e][a_pix] = av[e];
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[b_k][b_c + e] = bv[e];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < FBK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
        const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue
  TA* __restrict__ out = reinterpret_cast<TA*>(a.out);
  const TA* __restrict__ res = reinterpret_cast<const TA*>(a.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pl = ty * 4 + i;
    const int pn = s_pix_n[pl];
    if (pn < 0) continue;
    const int oy = s_pix_y[pl] * d.out_scale + d.out_oy;
    const int ox = s_pix_x[pl] * d.out_scale + d.out_ox;
    const size_t opix = ((size_t)pn * d.Hout + oy) * d.Wout + ox;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= d.Cout_store) continue;
      float v = 0.0f;
      if (co < d.Cout) {
        v = acc[i][j];
        if (a.bias) v += __ldg(a.bias + co);
        if (res) v += to_f32(res[opix * d.res_ld + d.res_coff + co]);
        if (d.relu) v = fmaxf(v, 0.0f);
      }
      out[opix * d.out_ld + d.out_coff + co] = from_f32<TA>(v);
    }
  }
}


// ------------------------------------------------------------------------------------------
// fp32 mode, second version: 128 pixels x (64 | 128) output channels per CTA, 8 x (4 | 8) outputs
// per thread, K chunks of 16 double-buffered in shared memory with the next chunk's global loads in
// flight during the FMAs, 16-byte global loads where the layout allows.  Every output accumulates
// its products in the same order as conv_ffma_kernel (taps, then channels ascending): bit-identical
// results.  (The first version -- 64 x 64 tiles, 4 x 4 outputs per thread, two barriers per chunk --
// runs at ~24 TFLOP/s, a third of the FP32 pipe.)
// ------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBK = 16, GTHREADS = 256;

template <int BN>
__global__ void __launch_bounds__(GTHREADS, 2) conv_ffma_big_kernel(FfmaArgs a, int a_vec, int b_vec) {
  constexpr int TN = BN / 16;                 // output channels per thread (4 or 8)
  constexpr int NB4 = TN / 4;                 // float4 groups of output channels per thread
  __shared__ __align__(16) float As[2][GBK][GBM + 4];
  __shared__ __align__(16) float Bs[2][GBK][BN + 4];
  __shared__ int s_pix_n[GBM], s_pix_y[GBM], s_pix_x[GBM];

  const brtpe_conv_desc& d = a.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GBM;
  const int n0 = blockIdx.y * BN;
  const float* __restrict__ in = reinterpret_cast<const float*>(a.in);

  if (tid < GBM) {
    const int m = m0 + tid;
    if (m < a.M) {
      const int hw = d.Hm * d.Wm;
      const int n = m / hw;
      const int r = m - n * hw;
      s_pix_n[tid] = n;
      s_pix_y[tid] = r / d.Wm;
      s_pix_x[tid] = r - (r / d.Wm) * d.Wm;
    } else {
      s_pix_n[tid] = -1;
      s_pix_y[tid] = 0;
      s_pix_x[tid] = 0;
    }
  }
  __syncthreads();

  const int tx = tid & 15;                    // output-channel group: channels tx*4 + g*64 + (0..3)
  const int ty = tid >> 4;                    // pixel group: pixels ty*8 + (0..7)
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  // A loader: thread -> (pixel = tid / 2, 8 channels from (tid % 2) * 8)
  const int a_pix = tid >> 1;
  const int a_k = (tid & 1) * 8;
  const int pn = s_pix_n[a_pix];
  const int py = s_pix_y[a_pix] * d.in_stride, px = s_pix_x[a_pix] * d.in_stride;
  // B loader: thread -> (k = tid / 16, TN output channels: group g at n0 + g*64 + (tid % 16) * 4)
  const int b_k = tid >> 4;
  const int b_c = (tid & 15) * 4;

  const int kchunks = (d.Cin + GBK - 1) / GBK;
  const int total = d.ntaps * kchunks;
  float av[8], bv[TN];

  auto load = [&](int it) {
    const int tap = it / kchunks;
    const int c0 = (it - tap * kchunks) * GBK;
    const int iy = py + d.tap_dy[tap], ix = px + d.tap_dx[tap];
    const bool pvalid = (pn >= 0) && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win;
    const float* __restrict__ prow =
        in + (((size_t)(pn < 0 ? 0 : pn) * d.Hin + (pvalid ? iy : 0)) * d.Win + (pvalid ? ix : 0)) *
                 (size_t)d.in_ld + d.in_coff;
    const int ca = c0 + a_k;
    if (a_vec) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pvalid && ca + 4 * h < d.Cin) v = __ldg(reinterpret_cast<const float4*>(prow + ca + 4 * h));
        av[4 * h] = v.x; av[4 * h + 1] = v.y; av[4 * h + 2] = v.z; av[4 * h + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) av[e] = (pvalid && ca + e < d.Cin) ? __ldg(prow + ca + e) : 0.0f;
    }
    const int cb = c0 + b_k;
    const float* __restrict__ wrow = a.w + ((size_t)tap * d.Cin + cb) * d.Cout;
#pragma unroll
    for (int g = 0; g < NB4; ++g) {
      const int co = n0 + g * 64 + b_c;
      if (b_vec) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb < d.Cin && co < d.Cout) v = __ldg(reinterpret_cast<const float4*>(wrow + co));
        bv[4 * g] = v.x; bv[4 * g + 1] = v.y; bv[4 * g + 2] = v.z; bv[4 * g + 3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          bv[4 * g + e] = (cb < d.Cin && co + e < d.Cout) ? __ldg(wrow + co + e) : 0.0f;
      }
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[buf][a_k + e][a_pix] = av[e];
#pragma unroll
    for (int g = 0; g < NB4; ++g)
      *reinterpret_cast<float4*>(&Bs[buf][b_k][g * 64 + b_c]) =
          make_float4(bv[4 * g], bv[4 * g + 1], bv[4 * g + 2], bv[4 * g + 3]);
  };

  load(0);
  store(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) load(it + 1);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float br[TN];
#pragma unroll
      for (int g = 0; g < NB4; ++g) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        br[4 * g] = b4.x; br[4 * g + 1] = b4.y; br[4 * g + 2] = b4.z; br[4 * g + 3] = b4.w;
      }
      // packed FFMA2 (two IEEE fmas per instruction, sm_100): pairs of neighbouring output channels
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a2 = make_float2(ar[i], ar[i]);
#pragma unroll
        for (int j = 0; j < TN; j += 2) {
          const float2 r = __ffma2_rn(a2, make_float2(br[j], br[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
          acc[i][j] = r.x;
          acc[i][j + 1] = r.y;
        }
      }
    }
    if (it + 1 < total) store(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  float* __restrict__ out = reinterpret_cast<float*>(a.out);
  const float* __restrict__ res = reinterpret_cast<const float*>(a.res);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pl = ty * 8 + i;
    const int qn = s_pix_n[pl];
    if (qn < 0) continue;
    const int oy = s_pix_y[pl] * d.out_scale + d.out_oy;
    const int ox = s_pix_x[pl] * d.out_scale + d.out_ox;
    const size_t opix = ((size_t)qn * d.Hout + oy) * d.Wout + ox;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (co >= d.Cout_store) continue;
      float v = 0.0f;
      if (co < d.Cout) {
        v = acc[i][j];
        if (a.bias) v += __ldg(a.bias + co);
        if (res) v += res[opix * d.res_ld + d.res_coff + co];
        if (d.relu) v = fmaxf(v, 0.0f);
      }
      out[opix * d.out_ld + d.out_coff + co] = v;
    }
  }
}


// Thin layers (Cout_store <= 16: the students' 12-channel dilation branches, 1- and 17/18-channel
// heads take 81-98 % padding in a 64-channel tile): 256 pixels x 16 output channels per CTA, 4 x 4
// outputs per thread, otherwise the scheme of conv_ffma_big_kernel (double-buffered K chunks, same
// accumulation order, bit-identical results).
constexpr int HBM_ = 256;
__global__ void __launch_bounds__(GTHREADS, 2) conv_ffma_thin_kernel(FfmaArgs a, int a_vec) {
  __shared__ __align__(16) float As[2][GBK][HBM_ + 4];
  __shared__ __align__(16) float Bs[2][GBK][16 + 4];
  __shared__ int s_pix_n[HBM_], s_pix_y[HBM_], s_pix_x[HBM_];

  const brtpe_conv_desc& d = a.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * HBM_;
  const float* __restrict__ in = reinterpret_cast<const float*>(a.in);
  {
    const int m = m0 + tid;
    if (m < a.M) {
      const int hw = d.Hm * d.Wm;
      const int n = m / hw;
      const int r = m - n * hw;
      s_pix_n[tid] = n;
      s_pix_y[tid] = r / d.Wm;
      s_pix_x[tid] = r - (r / d.Wm) * d.Wm;
    } else {
      s_pix_n[tid] = -1;
      s_pix_y[tid] = 0;
      s_pix_x[tid] = 0;
    }
  }
  __syncthreads();

  const int tx = tid & 3;                     // output channels tx*4 + (0..3)
  const int ty = tid >> 2;                    // pixels ty*4 + (0..3)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // A loader: thread -> pixel tid, all 16 channels of the chunk; B loader: (k = tid / 16, co = tid % 16)
  const int pn = s_pix_n[tid];
  const int py = s_pix_y[tid] * d.in_stride, px = s_pix_x[tid] * d.in_stride;
  const int b_k = tid >> 4, b_c = tid & 15;
  const int kchunks = (d.Cin + GBK - 1) / GBK;
  const int total = d.ntaps * kchunks;
  float av[16], bv;

  auto load = [&](int it) {
    const int tap = it / kchunks;
    const int c0 = (it - tap * kchunks) * GBK;
    const int iy = py + d.tap_dy[tap], ix = px + d.tap_dx[tap];
    const bool pvalid = (pn >= 0) && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win;
    const float* __restrict__ prow =
        in + (((size_t)(pn < 0 ? 0 : pn) * d.Hin + (pvalid ? iy : 0)) * d.Win + (pvalid ? ix : 0)) *
                 (size_t)d.in_ld + d.in_coff;
    if (a_vec) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pvalid && c0 + 4 * h < d.Cin) v = __ldg(reinterpret_cast<const float4*>(prow + c0 + 4 * h));
        av[4 * h] = v.x; av[4 * h + 1] = v.y; av[4 * h + 2] = v.z; av[4 * h + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) av[e] = (pvalid && c0 + e < d.Cin) ? __ldg(prow + c0 + e) : 0.0f;
    }
    const int cb = c0 + b_k;
    bv = (cb < d.Cin && b_c < d.Cout) ? __ldg(a.w + ((size_t)tap * d.Cin + cb) * d.Cout + b_c) : 0.0f;
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 16; ++e) As[buf][e][tid] = av[e];
    Bs[buf][b_k][b_c] = bv;
  };

  load(0);
  store(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) load(it + 1);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
      const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (it + 1 < total) store(buf ^ 1);
    __syncthreads();
  }

  float* __restrict__ out = reinterpret_cast<float*>(a.out);
  const float* __restrict__ res = reinterpret_cast<const float*>(a.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pl = ty * 4 + i;
    const int qn = s_pix_n[pl];
    if (qn < 0) continue;
    const int oy = s_pix_y[pl] * d.out_scale + d.out_oy;
    const int ox = s_pix_x[pl] * d.out_scale + d.out_ox;
    const size_t opix = ((size_t)qn * d.Hout + oy) * d.Wout + ox;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = tx * 4 + j;
      if (co >= d.Cout_store) continue;
      float v = 0.0f;
      if (co < d.Cout) {
        v = acc[i][j];
        if (a.bias) v += __ldg(a.bias + co);
        if (res) v += res[opix * d.res_ld + d.res_coff + co];
        if (d.relu) v = fmaxf(v, 0.0f);
      }
      out[opix * d.out_ld + d.out_coff + co] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// stem conv1: NCHW float/half image -> 3x3 s2 p1 conv (3 -> Cout) + bias + ReLU -> NHWC
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
stem_conv1_kernel(const TI* __restrict__ img, int N, int H, int W, const float* __restrict__ w,
                  const float* __restrict__ bias, int Cout, TO* __restrict__ out) {
  extern __shared__ float sw[];  // [27][Cout] + [Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = bias ? bias[i] : 0.0f;
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const int groups = Cout / 16;
  const size_t total = (size_t)N * Ho * Wo * groups;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    const size_t pix = t / groups;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((size_t)Wo * Ho));
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = sw[27 * Cout + g * 16 + c];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 + ky - 1;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 + kx - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          float v = 0.0f;
          if (ok) {
            const TI raw = img[(((size_t)n * 3 + ci) * H + iy) * W + ix];
            v = to_f32(raw);
          }
          const float* wr = sw + ((ky * 3 + kx) * 3 + ci) * Cout + g * 16;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
        }
      }
    }
    TO* o = out + pix * Cout + g * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = from_f32<TO>(fmaxf(acc[c], 0.0f));
  }
}

// ------------------------------------------------------------------------------------------
// stem im2col (bf16 / tcgen05 path of conv1): NCHW image -> NHWC bf16 (N, H/2, W/2, 32) where
// channel k = (ky*3 + kx)*3 + ci holds the 3x3 / stride-2 / pad-1 window of output pixel
// (oy, ox) and channels 27..31 are zero; conv1 then is a 1x1 tcgen05 conv with K = 32.
// One thread per output pixel: 27 strided reads (L1 keeps the 2x overlap), one 64-byte write.
// ------------------------------------------------------------------------------------------
// mode bit 1 (flip pair): `img` holds N/2 images; image n >= N/2 of the batch is image n - N/2
// mirrored in x -- the flip-test batch cat(x, flip(x)) without materialising it.  mode bit 2: a
// float32 image is rounded through fp16 first, like the reference's network_to_half wrapper
// (tofp16) does before the network sees it.
template <typename TI> struct Pair2;
template <> struct Pair2<float> { using type = float2; };
template <> struct Pair2<__half> { using type = __half2; };
__device__ __forceinline__ float2 pair_to_f32(float2 v) { return v; }
__device__ __forceinline__ float2 pair_to_f32(__half2 v) { return __half22float2(v); }

// One thread = one output pixel.  The three input columns 2*ox-1, 2*ox, 2*ox+1 of a tap row come
// from ONE aligned 2-pixel load per thread (columns 2*ox, 2*ox+1: a warp reads 64 consecutive
// pixels) plus the neighbouring lane's second pixel through a shuffle (column 2*ox-1); only lane 0
// loads that pixel itself.  (The first version issued 27 scalar loads per thread at a stride of two
// pixels between lanes.)
template <typename TI, bool SPLIT>
__global__ void __launch_bounds__(128)
stem_im2col_kernel(const TI* __restrict__ img, int N, int H, int W, __nv_bfloat16* __restrict__ out,
                   int mode) {
  using P2 = typename Pair2<TI>::type;
  const int Ho = H >> 1, Wo = W >> 1;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = ox < Wo;                  // no early return: the shuffles need whole warps
  const int row = blockIdx.y;                 // n * Ho + oy
  const int n = row / Ho, oy = row - n * Ho;
  const bool mirror = (mode & 2) && n >= (N >> 1);
  const int ns = mirror ? n - (N >> 1) : n;
  const bool via_half = (mode & 4) != 0;
  const int lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int k = 27; k < 32; ++k) v[k] = 0.0f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    const bool yok = iy >= 0 && iy < H;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const TI* __restrict__ r = img + (((size_t)ns * 3 + ci) * H + (yok ? iy : 0)) * W;
      // columns 2*ox (c0) and 2*ox+1 (c1): source pixels (2ox, 2ox+1), mirrored (W-1-2ox, W-2-2ox)
      float2 p = make_float2(0.0f, 0.0f);
      if (live && yok)
        p = pair_to_f32(*reinterpret_cast<const P2*>(r + (mirror ? W - 2 - 2 * ox : 2 * ox)));
      const float c0 = mirror ? p.y : p.x, c1 = mirror ? p.x : p.y;
      // column 2*ox-1 = the previous output pixel's column 2*(ox-1)+1
      float cm = __shfl_up_sync(0xffffffffu, c1, 1);
      if (lane == 0) {
        const int ix = 2 * ox - 1;
        cm = (live && yok && ix >= 0) ? to_f32(r[mirror ? W - 1 - ix : ix]) : 0.0f;
      }
      float t[3] = {cm, c0, c1};
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float tv = t[kx];
        if (via_half) tv = __half2float(__float2half_rn(tv));
        v[(ky * 3 + kx) * 3 + ci] = tv;
      }
    }
  }
  if (!live) return;
  // SPLIT (BRTPE_DT_BF16X2): 64 stored channels per pixel, hi parts in [0, 32), lo parts in [32, 64)
  uint4* o = reinterpret_cast<uint4*>(out + ((size_t)row * Wo + ox) * (SPLIT ? 64 : 32));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4], wl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a0 = v[q * 8 + 2 * e], a1 = v[q * 8 + 2 * e + 1];
      __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);
      w[e] = *reinterpret_cast<uint32_t*>(&pk);
      if (SPLIT) {
        __nv_bfloat162 pl = __floats2bfloat162_rn(a0 - __uint_as_float(w[e] << 16),
                                                  a1 - __uint_as_float(w[e] & 0xffff0000u));
        wl[e] = *reinterpret_cast<uint32_t*>(&pl);
      }
    }
    o[q] = make_uint4(w[0], w[1], w[2], w[3]);
    if (SPLIT) o[4 + q] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  }
}

// ------------------------------------------------------------------------------------------
// cross-resolution fuse: out = relu?( sum_k nearest_up_{2^s_k}(term_k) ), NHWC
// ------------------------------------------------------------------------------------------
struct FuseArgs {
  const void* terms[4];
  int shifts[4];
  int ld[4];
  int nterms, N, H, W, C, out_ld, relu;
  void* out;
};

template <typename T>
__global__ void __launch_bounds__(256) fuse_sum_kernel(FuseArgs a) {
  const int cv = a.C / 4;  // 4 channels per thread
  const size_t total = (size_t)a.N * a.H * a.W * cv;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % cv) * 4;
    const size_t pix = t / cv;
    const int x = (int)(pix % a.W);
    const int y = (int)((pix / a.W) % a.H);
    const int n = (int)(pix / ((size_t)a.W * a.H));
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < a.nterms; ++k) {
      const int sh = a.shifts[k];
      const int hk = a.H >> sh, wk = a.W >> sh;
      const T* p = reinterpret_cast<const T*>(a.terms[k]) +
                   (((size_t)n * hk + (y >> sh)) * wk + (x >> sh)) * a.ld[k] + c;
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = (k == 0) ? to_f32(p[e]) : s[e] + to_f32(p[e]);
    }
    T* o = reinterpret_cast<T*>(a.out) + pix * a.out_ld + c;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = from_f32<T>(a.relu ? fmaxf(s[e], 0.0f) : s[e]);
  }
}

// bf16 fast path: one thread = 8 channels (16 bytes) of one output pixel, one image row per
// blockIdx.y; the same left-to-right fp32 summation order as the generic kernel.
__global__ void __launch_bounds__(256) fuse_sum_bf16x8_kernel(FuseArgs a) {
  const int cv = a.C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.W * cv) return;
  const int x = i / cv, c = (i - x * cv) << 3;
  const int row = blockIdx.y;               // n * H + y
  const int n = row / a.H, y = row - n * a.H;
  float s[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < a.nterms) {
      const int sh = a.shifts[k];
      const int hk = a.H >> sh, wk = a.W >> sh;
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.terms[k]) +
                               (((size_t)n * hk + (y >> sh)) * wk + (x >> sh)) * a.ld[k] + c;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        s[2 * e] = (k == 0) ? lo : s[2 * e] + lo;
        s[2 * e + 1] = (k == 0) ? hi : s[2 * e + 1] + hi;
      }
    }
  }
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float lo = s[2 * e], hi = s[2 * e + 1];
    if (a.relu) { lo = fmaxf(lo, 0.0f); hi = fmaxf(hi, 0.0f); }
    __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
    o[e] = *reinterpret_cast<uint32_t*>(&pk);
  }
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((size_t)row * a.W + x) * a.out_ld + c;
  *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
}


// split (BRTPE_DT_BF16X2) activations: every value is hi + lo (two bf16, the lo half ld / 2 elements
// into the pixel); one thread = 8 channels of one output pixel, rows walked with a grid stride.
// Same left-to-right fp32 summation order as the generic kernel.
__global__ void __launch_bounds__(256) fuse_sum_split_kernel(FuseArgs a) {
  const int cv = a.C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.W * cv) return;
  const int x = i / cv, c = (i - x * cv) << 3;
  for (int row = blockIdx.y; row < a.N * a.H; row += gridDim.y) {
    const int n = row / a.H, y = row - n * a.H;
    float s[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < a.nterms) {
        const int sh = a.shifts[k];
        const int hk = a.H >> sh, wk = a.W >> sh;
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.terms[k]) +
                                 (((size_t)n * hk + (y >> sh)) * wk + (x >> sh)) * a.ld[k] + c;
        const uint4 vh = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 vl = __ldg(reinterpret_cast<const uint4*>(p + (a.ld[k] >> 1)));
        const uint32_t wh[4] = {vh.x, vh.y, vh.z, vh.w}, wl[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = __uint_as_float(wh[e] << 16) + __uint_as_float(wl[e] << 16);
          const float v1 = __uint_as_float(wh[e] & 0xffff0000u) + __uint_as_float(wl[e] & 0xffff0000u);
          s[2 * e] = (k == 0) ? v0 : s[2 * e] + v0;
          s[2 * e + 1] = (k == 0) ? v1 : s[2 * e + 1] + v1;
        }
      }
    }
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = s[2 * e], v1 = s[2 * e + 1];
      if (a.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
      __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
      oh[e] = *reinterpret_cast<uint32_t*>(&pk);
      __nv_bfloat162 pl = __floats2bfloat162_rn(v0 - __uint_as_float(oh[e] << 16),
                                                v1 - __uint_as_float(oh[e] & 0xffff0000u));
      ol[e] = *reinterpret_cast<uint32_t*>(&pl);
    }
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((size_t)row * a.W + x) * a.out_ld + c;
    *reinterpret_cast<uint4*>(op) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(op + (a.out_ld >> 1)) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// bf16 fast path, second version: block = (C/8 channel vectors) x (PX pixels) threads, grid =
// (pixel blocks, rows, images): no integer division, the per-term row pointers are computed once
// per thread, U pixels per thread with all their loads issued before the first add (the first
// version spent ~280 instructions per 16 output bytes, most of them address arithmetic, and ran
// at 41-58 % of the DRAM bandwidth with 62 % of the issue slots busy: gpurun_out/fuse_r01p.ncu-rep).
// Same left-to-right fp32 summation order as the generic kernel.
template <int U, int NT>
__global__ void __launch_bounds__(256) fuse_sum_bf16x8_rows_kernel(FuseArgs a) {
  pdl_wait();                                  // PDL (common.cuh): the terms are the predecessors' outputs
  const int px = blockDim.y;
  const int c = threadIdx.x << 3;
  const int y = blockIdx.y, n = blockIdx.z;
  const __nv_bfloat16* rowp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rowp[k] = nullptr;
    if (k < NT) {
      const int sh = a.shifts[k];
      const int hk = a.H >> sh, wk = a.W >> sh;
      rowp[k] = reinterpret_cast<const __nv_bfloat16*>(a.terms[k]) +
                ((size_t)n * hk + (y >> sh)) * wk * a.ld[k] + c;
    }
  }
  __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(a.out) +
                        ((size_t)n * a.H + y) * a.W * a.out_ld + c;
  const int x0 = blockIdx.x * (U * px) + threadIdx.y;
  uint4 v[4][U];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < NT) {
      const int sh = a.shifts[k], ld = a.ld[k];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int x = x0 + j * px;
        if (x < a.W) v[k][j] = *reinterpret_cast<const uint4*>(rowp[k] + (x >> sh) * ld);
      }
    }
  }
  pdl_trigger();
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const int x = x0 + j * px;
    if (x >= a.W) continue;
    float2 s[4];                              // (even, odd) channel pairs, packed FADD2 adds
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < NT) {
        const uint32_t w[4] = {v[k][j].x, v[k][j].y, v[k][j].z, v[k][j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
          s[e] = (k == 0) ? t : __fadd2_rn(s[e], t);
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo = s[e].x, hi = s[e].y;
      if (a.relu) { lo = fmaxf(lo, 0.0f); hi = fmaxf(hi, 0.0f); }
      __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
      o[e] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(orow + (size_t)x * a.out_ld) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// NHWC (f32 | bf16) -> NCHW (f32 | f16) network outputs
// ------------------------------------------------------------------------------------------
template <typename TS, typename TD, bool SPLIT = false>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const TS* __restrict__ src, int N, int HW, int C, int ld, int coff,
                    TD* __restrict__ dst) {
  __shared__ float tile[64][65];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int pl = i / 64, cl = i % 64;
    const int p = p0 + pl, c = c0 + cl;
    float v = 0.0f;
    if (p < HW && c < C) {
      const TS* sp = src + ((size_t)n * HW + p) * ld + coff + c;
      v = to_f32(sp[0]);
      if (SPLIT) v += to_f32(sp[ld >> 1]);           // hi + lo of a split (bf16x2) value
    }
    tile[pl][cl] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int cl = i / 64, pl = i % 64;
    const int p = p0 + pl, c = c0 + cl;
    if (p < HW && c < C) dst[((size_t)n * C + c) * HW + p] = from_f32<TD>(tile[pl][cl]);
  }
}

// bf16 NHWC -> NCHW (half or float), C <= 64 channels, 16-byte aligned channel slices: the
// network-output case (34 and 17 channels).  128-256 pixels per CTA; the source is read as 16-byte
// vectors (8 channels), transposed through shared memory and written as 16-byte row segments.
template <typename TD>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_bf16_vec_kernel(const __nv_bfloat16* __restrict__ src, int HW, int C, int ld, int coff,
                             TD* __restrict__ dst) {
  constexpr int PX = 512 / (int)sizeof(TD);             // 256 (half) / 128 (float) pixels per CTA
  constexpr int EPV = 16 / (int)sizeof(TD);              // destination elements per 16-byte store
  __shared__ __align__(16) TD tile[64][PX + EPV];        // [channel][pixel], padded rows
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * PX;
  const int cv = (C + 7) >> 3;                           // 16-byte vectors per pixel
  for (int v = threadIdx.x; v < PX * cv; v += 256) {
    const int pl = v / cv, k = v - pl * cv;
    const int p = p0 + pl;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (p < HW)
      q = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)n * HW + p) * ld + coff + k * 8));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = k * 8 + 2 * i;
      if (c < C) tile[c][pl] = from_f32<TD>(__uint_as_float(w[i] << 16));
      if (c + 1 < C) tile[c + 1][pl] = from_f32<TD>(__uint_as_float(w[i] & 0xffff0000u));
    }
  }
  __syncthreads();
  constexpr int SEG = PX / EPV;                          // 16-byte segments per channel row
  for (int v = threadIdx.x; v < C * SEG; v += 256) {
    const int c = v / SEG, sg = v - c * SEG;
    const int p = p0 + sg * EPV;
    TD* d = dst + ((size_t)n * C + c) * HW + p;
    if (p + EPV <= HW && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
      *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(&tile[c][sg * EPV]);
    } else {
      for (int e = 0; e < EPV; ++e)
        if (p + e < HW) d[e] = tile[c][sg * EPV + e];
    }
  }
}

// ------------------------------------------------------------------------------------------
int conv_validate(const brtpe_conv_desc* d) {
  BRTPE_CHECK_ARG(d != nullptr, "conv: null descriptor");
  BRTPE_CHECK_ARG(d->dtype == BRTPE_DT_F32 || d->dtype == BRTPE_DT_BF16 ||
                      d->dtype == BRTPE_DT_BF16X2, "conv: bad dtype %d", d->dtype);
  BRTPE_CHECK_ARG(d->N > 0 && d->Hin > 0 && d->Win > 0 && d->Cin > 0 && d->Cout > 0,
                  "conv: bad tensor sizes");
  BRTPE_CHECK_ARG(d->Hm > 0 && d->Wm > 0 && d->in_stride >= 1 && d->out_scale >= 1,
                  "conv: bad GEMM-M domain");
  BRTPE_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9, "conv: ntaps=%d outside [1,9]", d->ntaps);
  BRTPE_CHECK_ARG(d->in_ld >= d->in_coff + d->Cin, "conv: in_ld %d < in_coff %d + Cin %d", d->in_ld,
                  d->in_coff, d->Cin);
  BRTPE_CHECK_ARG(d->Cout_store >= d->Cout && d->out_ld >= d->out_coff + d->Cout_store,
                  "conv: out_ld %d too small for out_coff %d + Cout_store %d", d->out_ld,
                  d->out_coff, d->Cout_store);
  BRTPE_CHECK_ARG((d->Hm - 1) * d->out_scale + d->out_oy < d->Hout &&
                      (d->Wm - 1) * d->out_scale + d->out_ox < d->Wout && d->out_oy >= 0 &&
                      d->out_ox >= 0,
                  "conv: output mapping leaves the output tensor");
  return BRTPE_OK;
}

double conv_flops(const brtpe_conv_desc* d) {
  return 2.0 * (double)d->N * d->Hm * d->Wm * (double)d->ntaps * d->Cin * d->Cout;
}

int conv_ffma_launch(const brtpe_conv_desc* d, const void* in, const void* weights,
                     const float* bias, const void* residual, void* out, cudaStream_t st) {
  FfmaArgs a;
  a.d = *d;
  a.in = in;
  a.w = reinterpret_cast<const float*>(weights);
  a.bias = bias;
  a.res = residual;
  a.out = out;
  a.M = d->N * d->Hm * d->Wm;
  static int big = -1;
  if (big < 0) {
    const char* e = getenv("BRTPE_FFMA_BIG");
    big = e ? atoi(e) : 1;
  }
  if (big && d->dtype == BRTPE_DT_F32) {
    const int a_vec = (d->Cin % 4 == 0 && d->in_ld % 4 == 0 && d->in_coff % 4 == 0 &&
                       (reinterpret_cast<uintptr_t>(in) & 15) == 0) ? 1 : 0;
    const int b_vec = (d->Cout % 4 == 0 && (reinterpret_cast<uintptr_t>(weights) & 15) == 0) ? 1 : 0;
    if (d->Cout_store <= 16 && big != 2) {
      conv_ffma_thin_kernel<<<ceil_div(a.M, HBM_), GTHREADS, 0, st>>>(a, a_vec);
    } else if (d->Cout_store > 64) {
      dim3 g2(ceil_div(a.M, GBM), ceil_div(d->Cout_store, 128));
      conv_ffma_big_kernel<128><<<g2, GTHREADS, 0, st>>>(a, a_vec, b_vec);
    } else {
      dim3 g2(ceil_div(a.M, GBM), 1);
      conv_ffma_big_kernel<64><<<g2, GTHREADS, 0, st>>>(a, a_vec, b_vec);
    }
    BRTPE_LAUNCH_CHECK();
    return BRTPE_OK;
  }
  dim3 grid(ceil_div(a.M, FBM), ceil_div(d->Cout_store, FBN));
  if (d->dtype == BRTPE_DT_F32)
    conv_ffma_kernel<float><<<grid, FTHREADS, 0, st>>>(a);
  else
    conv_ffma_kernel<__nv_bfloat16><<<grid, FTHREADS, 0, st>>>(a);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

int stem_conv1_launch(const void* img, int img_is_half, int N, int H, int W, const float* w,
                      const float* bias, int Cout, void* out, int out_dtype, cudaStream_t st) {
  BRTPE_CHECK_ARG(img && w && out && N > 0 && H > 0 && W > 0, "stem_conv1: bad arguments");
  BRTPE_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "stem_conv1: H and W must be even");
  BRTPE_CHECK_ARG(Cout % 16 == 0 && Cout <= 256, "stem_conv1: Cout must be a multiple of 16");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (Cout / 16);
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  const size_t smem = (size_t)28 * Cout * sizeof(float);
#define BRTPE_STEM(TI, TO)                                                                        \
  stem_conv1_kernel<TI, TO><<<blocks, 256, smem, st>>>(reinterpret_cast<const TI*>(img), N, H, W, \
                                                      w, bias, Cout, reinterpret_cast<TO*>(out))
  if (img_is_half) {
    if (out_dtype == BRTPE_DT_F32) BRTPE_STEM(__half, float);
    else BRTPE_STEM(__half, __nv_bfloat16);
  } else {
    if (out_dtype == BRTPE_DT_F32) BRTPE_STEM(float, float);
    else BRTPE_STEM(float, __nv_bfloat16);
  }
#undef BRTPE_STEM
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

int stem_im2col_launch(const void* img, int img_mode, int N, int H, int W, void* out,
                       cudaStream_t st) {
  // img_mode: bit 0 = half input, bit 1 = flip pair, bit 2 = round float32 through fp16,
  // bit 3 = split (bf16x2) output: 64 stored channels per pixel
  const int img_is_half = img_mode & 1;
  const int kmode = img_mode & 6;
  const bool split = (img_mode & 8) != 0;
  BRTPE_CHECK_ARG(img && out && N > 0 && H > 0 && W > 0, "stem_im2col: bad arguments");
  BRTPE_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "stem_im2col: H and W must be even");
  BRTPE_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "stem_im2col: out must be 16-byte aligned");
  BRTPE_CHECK_ARG(!(kmode & 2) || (N % 2) == 0, "stem_im2col: flip-pair mode needs an even N");
  BRTPE_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & (img_is_half ? 3 : 7)) == 0,
                  "stem_im2col: the image must be aligned to two pixels (8 bytes float32, 4 bytes half)");
  const int rows = N * (H / 2);
  dim3 grid(ceil_div(W / 2, 128), rows);
  if (rows > 65535) {
    BRTPE_CHECK_ARG(!(kmode & 2), "stem_im2col: flip-pair mode needs N * H / 2 <= 65535");
    // blockIdx.y is limited to 65535: split the batch
    const int per = 65535 / (H / 2);
    BRTPE_CHECK_ARG(per >= 1, "stem_im2col: image too tall");
    const size_t isz = img_is_half ? 2 : 4;
    for (int n0 = 0; n0 < N; n0 += per) {
      const int nn = std::min(per, N - n0);
      int rc = stem_im2col_launch(reinterpret_cast<const char*>(img) + (size_t)n0 * 3 * H * W * isz,
                                  img_mode, nn, H, W,
                                  reinterpret_cast<__nv_bfloat16*>(out) +
                                      (size_t)n0 * (H / 2) * (W / 2) * (split ? 64 : 32),
                                  st);
      if (rc) return rc;
    }
    return BRTPE_OK;
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (img_is_half) {
    const __half* im = reinterpret_cast<const __half*>(img);
    if (split) stem_im2col_kernel<__half, true><<<grid, 128, 0, st>>>(im, N, H, W, o, kmode & 2);
    else stem_im2col_kernel<__half, false><<<grid, 128, 0, st>>>(im, N, H, W, o, kmode & 2);
  } else {
    const float* im = reinterpret_cast<const float*>(img);
    if (split) stem_im2col_kernel<float, true><<<grid, 128, 0, st>>>(im, N, H, W, o, kmode);
    else stem_im2col_kernel<float, false><<<grid, 128, 0, st>>>(im, N, H, W, o, kmode);
  }
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

int fuse_sum_launch(int dtype, int nterms, const void* const* terms, const int32_t* shifts,
                    const int32_t* term_ld, int N, int H, int W, int C, void* out, int out_ld,
                    int relu, cudaStream_t st) {
  BRTPE_CHECK_ARG(nterms >= 1 && nterms <= 4 && terms && shifts && term_ld && out,
                  "fuse_sum: bad arguments");
  BRTPE_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0, "fuse_sum: C must be a multiple of 4");
  FuseArgs a;
  for (int k = 0; k < 4; ++k) {
    a.terms[k] = k < nterms ? terms[k] : nullptr;
    a.shifts[k] = k < nterms ? shifts[k] : 0;
    a.ld[k] = k < nterms ? term_ld[k] : 0;
    if (k < nterms) {
      BRTPE_CHECK_ARG(terms[k] != nullptr, "fuse_sum: null term %d", k);
      BRTPE_CHECK_ARG(shifts[k] >= 0 && (H % (1 << shifts[k])) == 0 && (W % (1 << shifts[k])) == 0,
                      "fuse_sum: term %d shift %d does not divide %dx%d", k, shifts[k], H, W);
    }
  }
  a.nterms = nterms; a.N = N; a.H = H; a.W = W; a.C = C; a.out_ld = out_ld; a.relu = relu;
  a.out = out;
  if (dtype == BRTPE_DT_BF16X2) {
    bool ok = (C % 8) == 0 && (out_ld % 16) == 0 && out_ld / 2 >= C &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int k = 0; k < nterms; ++k)
      ok = ok && (term_ld[k] % 16) == 0 && term_ld[k] / 2 >= C &&
           (reinterpret_cast<uintptr_t>(terms[k]) & 15) == 0;
    BRTPE_CHECK_ARG(ok, "fuse_sum: split (bf16x2) tensors need C %% 8 == 0, strides %% 16 == 0 and "
                        "16-byte aligned pointers");
    dim3 grid(ceil_div(W * (C / 8), 256), std::min(N * H, 65535));
    fuse_sum_split_kernel<<<grid, 256, 0, st>>>(a);
    BRTPE_LAUNCH_CHECK();
    return BRTPE_OK;
  }
  const size_t total = (size_t)N * H * W * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  bool fast = dtype == BRTPE_DT_BF16 && (C % 8) == 0 && (out_ld % 8) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int k = 0; k < nterms; ++k)
    fast = fast && (term_ld[k] % 8) == 0 && (reinterpret_cast<uintptr_t>(terms[k]) & 15) == 0;
  static int fuse_v2 = -1;
  if (fuse_v2 < 0) {
    const char* e = getenv("BRTPE_FUSE_V2");
    fuse_v2 = e ? atoi(e) : 2;                 // 0: first version; 1 / 2 / 4: pixels per thread
  }
  const int cv = C / 8;
  if (fast && fuse_v2 && cv <= 256 && H <= 65535 && N <= 65535) {
    const int px = 256 / cv;
    dim3 block(cv, px);
    // pixels per thread, measured on the 23 fuse launches of a 64-forward plan: 1.66 ms (first
    // version) -> 1.33 (U = 4) -> 1.24 (U = 2); U = 8: 2.6 ms (166 registers)
    const int U = fuse_v2 == 1 ? 1 : fuse_v2 == 4 ? 4 : 2;
    dim3 grid(ceil_div(W, U * px), H, N);
#define BRTPE_FUSE(UU)                                                                         \
  switch (nterms) {                                                                            \
    case 1: launch_ex(fuse_sum_bf16x8_rows_kernel<UU, 1>, grid, block, 0, st, 0, true, a); break;              \
    case 2: launch_ex(fuse_sum_bf16x8_rows_kernel<UU, 2>, grid, block, 0, st, 0, true, a); break;              \
    case 3: launch_ex(fuse_sum_bf16x8_rows_kernel<UU, 3>, grid, block, 0, st, 0, true, a); break;              \
    default: launch_ex(fuse_sum_bf16x8_rows_kernel<UU, 4>, grid, block, 0, st, 0, true, a); break;             \
  }
    if (U == 1) { BRTPE_FUSE(1) }
    else if (U == 4) { BRTPE_FUSE(4) }
    else { BRTPE_FUSE(2) }
#undef BRTPE_FUSE
  } else if (fast && (long long)N * H <= 65535) {
    dim3 grid(ceil_div(W * (C / 8), 256), N * H);
    fuse_sum_bf16x8_kernel<<<grid, 256, 0, st>>>(a);
  } else if (dtype == BRTPE_DT_F32) {
    fuse_sum_kernel<float><<<blocks, 256, 0, st>>>(a);
  } else {
    fuse_sum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(a);
  }
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

int nhwc_to_nchw_launch(int dtype, const void* src, int N, int H, int W, int C, int ld, int coff,
                        void* dst, int dst_is_half, cudaStream_t st) {
  BRTPE_CHECK_ARG(src && dst && N > 0 && H > 0 && W > 0 && C > 0 && ld >= coff + C,
                  "nhwc_to_nchw: bad arguments");
  if (dtype == BRTPE_DT_BF16X2) {
    BRTPE_CHECK_ARG((ld % 2) == 0 && ld / 2 >= coff + C, "nhwc_to_nchw: split tensor needs ld / 2 >= coff + C");
    dim3 sgrid(ceil_div(H * W, 64), ceil_div(C, 64), N);
    if (dst_is_half)
      nhwc_to_nchw_kernel<__nv_bfloat16, __half, true><<<sgrid, 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(src), N, H * W, C, ld, coff, reinterpret_cast<__half*>(dst));
    else
      nhwc_to_nchw_kernel<__nv_bfloat16, float, true><<<sgrid, 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(src), N, H * W, C, ld, coff, reinterpret_cast<float*>(dst));
    BRTPE_LAUNCH_CHECK();
    return BRTPE_OK;
  }
  if (dtype == BRTPE_DT_BF16 && C <= 64 && (ld % 8) == 0 && (coff % 8) == 0 && N <= 65535 &&
      (reinterpret_cast<uintptr_t>(src) & 15) == 0 && coff + ((C + 7) / 8) * 8 <= ld) {
    dim3 vgrid(ceil_div(H * W, dst_is_half ? 256 : 128), N);
    if (dst_is_half)
      nhwc_to_nchw_bf16_vec_kernel<__half><<<vgrid, 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(src), H * W, C, ld, coff, reinterpret_cast<__half*>(dst));
    else
      nhwc_to_nchw_bf16_vec_kernel<float><<<vgrid, 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(src), H * W, C, ld, coff, reinterpret_cast<float*>(dst));
    BRTPE_LAUNCH_CHECK();
    return BRTPE_OK;
  }
  dim3 grid(ceil_div(H * W, 64), ceil_div(C, 64), N);
#define BRTPE_T(TS, TD)                                                                    \
  nhwc_to_nchw_kernel<TS, TD><<<grid, 256, 0, st>>>(reinterpret_cast<const TS*>(src), N, H * W, C, \
                                                   ld, coff, reinterpret_cast<TD*>(dst))
  if (dtype == BRTPE_DT_F32) {
    if (dst_is_half) BRTPE_T(float, __half);
    else BRTPE_T(float, float);
  } else {
    if (dst_is_half) BRTPE_T(__nv_bfloat16, __half);
    else BRTPE_T(__nv_bfloat16, float);
  }
#undef BRTPE_T
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

}  // namespace brtpe
