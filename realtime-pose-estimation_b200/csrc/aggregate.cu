// Heat-map / tag aggregation between the network and the parser.
//
//  * brtpe_bilinear_resize : F.interpolate(..., mode="bilinear") of the in-tree path
//    (validate_hhrnet.py:94-98, align_corners=True to the original image size); the
//    destination addressing lets the tag result land directly in the (N,A,h,w,T) tensor
//    the parser takes (the reference's aes.unsqueeze(-1)).
//  * brtpe_aggregate_scale : flip-test + multi-scale mode (upstream HigherHRNet
//    get_multi_stage_outputs / aggregate_results, used by legacy/valid_ae_avg.py:176-185):
//    stage average, H/4 -> H/2 -> base-size bilinear cascade (align_corners=False),
//    flip-back with left/right channel permutation, flip average, scale accumulation and
//    the final division, fused into one pass that writes det and tag exactly once.
//
// Both are write-bound streaming kernels: one thread per output pixel, consecutive
// threads on consecutive x, sources are 4-16x smaller than the output and stay in L1/L2.
#include "common.cuh"
#include "umma_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace brtpe {

// PyTorch's area_pixel_compute_source_index (float32 arithmetic).
__device__ __forceinline__ float src_index(float scale, int dst, bool align_corners) {
  if (align_corners) return scale * (float)dst;
  const float s = scale * ((float)dst + 0.5f) - 0.5f;
  return s < 0.0f ? 0.0f : s;
}
static inline float resize_scale(int in, int out, bool align_corners) {
  if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  return (float)in / (float)out;
}

struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp make_lerp(float scale, int dst, int in, bool ac) {
  Lerp l;
  const float s = src_index(scale, dst, ac);
  l.i0 = (int)s;
  if (l.i0 > in - 1) l.i0 = in - 1;
  l.i1 = l.i0 + ((l.i0 < in - 1) ? 1 : 0);
  l.w1 = s - (float)l.i0;
  l.w0 = 1.0f - l.w1;
  return l;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ p, int W, const Lerp& ly,
                                        const Lerp& lx) {
  const float v00 = __ldg(p + (size_t)ly.i0 * W + lx.i0), v01 = __ldg(p + (size_t)ly.i0 * W + lx.i1);
  const float v10 = __ldg(p + (size_t)ly.i1 * W + lx.i0), v11 = __ldg(p + (size_t)ly.i1 * W + lx.i1);
  return ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
}

__global__ void bilinear_resize_kernel(const float* __restrict__ src, long long src_plane_stride,
                                       int planes, int Hi, int Wi, float* __restrict__ dst, int Ho,
                                       int Wo, int dst_inner, int dst_off, float sy, float sx,
                                       bool ac) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= Wo) return;
  const Lerp ly = make_lerp(sy, y, Hi, ac);
  const Lerp lx = make_lerp(sx, x, Wi, ac);
  for (int p = blockIdx.z; p < planes; p += gridDim.z) {
    const float v = bilerp(src + (size_t)p * src_plane_stride, Wi, ly, lx);
    dst[((size_t)p * Ho * Wo + (size_t)y * Wo + x) * dst_inner + dst_off] = v;
  }
}

struct AggArgs {
  const float *y0, *y1, *y0f, *y1f;
  float *det, *tag;
  int N, J, A, H4, W4, H2, W2, Hb, Wb;
  int accumulate;
  float final_div;
  float s42y, s42x, s2by, s2bx;
  int flip_index[BRTPE_MAX_JOINTS];
};

// value of bilinear_{H2->Hb}(G) at (yb, xb) where G[y2][x2] is produced on the fly:
//   G = (bilinear_{H4->H2}(q0)[y2][x2'] + q1[y2][x2']) * 0.5   (q1 != null: heat-maps)
//   G =  bilinear_{H4->H2}(q0)[y2][x2']                        (q1 == null: tags)
// with x2' = mirror ? W2-1-x2 : x2.
__device__ __forceinline__ float cascade_sample(const AggArgs& a, const float* __restrict__ q0,
                                                const float* __restrict__ q1, const Lerp& by,
                                                const Lerp& bx, bool mirror) {
  float g[2][2];
#pragma unroll
  for (int iy = 0; iy < 2; ++iy) {
    const int y2 = iy ? by.i1 : by.i0;
    const Lerp ly = make_lerp(a.s42y, y2, a.H4, false);
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      int x2 = ix ? bx.i1 : bx.i0;
      if (mirror) x2 = a.W2 - 1 - x2;
      const Lerp lx = make_lerp(a.s42x, x2, a.W4, false);
      float v = bilerp(q0, a.W4, ly, lx);
      if (q1) v = (v + __ldg(q1 + (size_t)y2 * a.W2 + x2)) * 0.5f;
      g[iy][ix] = v;
    }
  }
  return by.w0 * (bx.w0 * g[0][0] + bx.w1 * g[0][1]) + by.w1 * (bx.w0 * g[1][0] + bx.w1 * g[1][1]);
}

// ---- tiled version: the stage-average map G (at 1/2 resolution) of the tile's footprint is
// built once in shared memory (for the image and for the mirrored image), then every output
// pixel is one 4-tap blend of G.  Same per-element arithmetic as the direct kernel below;
// ~10x fewer instructions per output, so the kernel is bound by its det/tag writes.
constexpr int AG_TH = 32, AG_TW = 128, AG_THREADS = 256;
constexpr int AG_MAXP = 2560;             // floats per G patch (two patches in smem)

__device__ __forceinline__ float g_value(const AggArgs& a, const float* __restrict__ q0,
                                         const float* __restrict__ q1, int y2, int x2, bool mirror) {
  if (mirror) x2 = a.W2 - 1 - x2;
  const Lerp ly = make_lerp(a.s42y, y2, a.H4, false);
  const Lerp lx = make_lerp(a.s42x, x2, a.W4, false);
  float v = bilerp(q0, a.W4, ly, lx);
  if (q1) v = (v + __ldg(q1 + (size_t)y2 * a.W2 + x2)) * 0.5f;
  return v;
}

__global__ void __launch_bounds__(AG_THREADS) aggregate_scale_tiled_kernel(AggArgs a) {
  __shared__ float G[2][AG_MAXP];
  const int xb0 = blockIdx.x * AG_TW, yb0 = blockIdx.y * AG_TH;
  const int yb1 = min(yb0 + AG_TH, a.Hb) - 1, xb1 = min(xb0 + AG_TW, a.Wb) - 1;
  const int y2lo = make_lerp(a.s2by, yb0, a.H2, false).i0, y2hi = make_lerp(a.s2by, yb1, a.H2, false).i1;
  const int x2lo = make_lerp(a.s2bx, xb0, a.W2, false).i0, x2hi = make_lerp(a.s2bx, xb1, a.W2, false).i1;
  const int PH = y2hi - y2lo + 1, PW = x2hi - x2lo + 1;
  const int C0 = a.J + a.A;
  const size_t p4 = (size_t)a.H4 * a.W4, p2 = (size_t)a.H2 * a.W2, pb = (size_t)a.Hb * a.Wb;
  const bool flip = a.y0f != nullptr;
  const int T = flip ? 2 : 1;
  const int nch = a.J + (a.tag ? a.A : 0);
  const int tid = threadIdx.x;
  const int cx = (tid & 31) * 4, ry = tid >> 5;             // 4 columns x rows ry, ry+8, ...
  // per-thread column interpolation (fixed for the whole kernel)
  int ci0[4], ci1[4];
  float cw0[4], cw1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int xb = min(xb0 + cx + q, a.Wb - 1);
    const Lerp bx = make_lerp(a.s2bx, xb, a.W2, false);
    ci0[q] = bx.i0 - x2lo; ci1[q] = bx.i1 - x2lo; cw0[q] = bx.w0; cw1[q] = bx.w1;
  }
  const bool col_ok = xb0 + cx < a.Wb;
  const bool vec4 = ((a.Wb & 3) == 0);

  for (int nc = blockIdx.z; nc < a.N * nch; nc += gridDim.z) {
    const int n = nc / nch, c = nc - n * nch;
    const bool is_det = c < a.J;
    const float *q0, *q1 = nullptr, *q0f = nullptr, *q1f = nullptr;
    if (is_det) {
      q0 = a.y0 + ((size_t)n * C0 + c) * p4;
      q1 = a.y1 + ((size_t)n * a.J + c) * p2;
      if (flip) {
        const int cf = a.flip_index[c];
        q0f = a.y0f + ((size_t)n * C0 + cf) * p4;
        q1f = a.y1f + ((size_t)n * a.J + cf) * p2;
      }
    } else {
      const int t = c - a.J;
      q0 = a.y0 + ((size_t)n * C0 + a.J + t) * p4;
      if (flip) {
        const int tf = (a.A == a.J) ? a.flip_index[t] : t;
        q0f = a.y0f + ((size_t)n * C0 + a.J + tf) * p4;
      }
    }
    __syncthreads();                                        // previous channel's readers are done
    for (int i = tid; i < PH * PW; i += AG_THREADS) {
      const int r = i / PW, s_ = i - r * PW;
      G[0][i] = g_value(a, q0, q1, y2lo + r, x2lo + s_, false);
      if (flip) G[1][i] = g_value(a, q0f, q1f, y2lo + r, x2lo + s_, true);
    }
    __syncthreads();
    if (!col_ok) continue;
    for (int yb = yb0 + ry; yb <= yb1; yb += AG_THREADS / 32) {
      const Lerp by = make_lerp(a.s2by, yb, a.H2, false);
      const float* g0 = &G[0][(by.i0 - y2lo) * PW];
      const float* g1 = &G[0][(by.i1 - y2lo) * PW];
      float h[4], hf[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = by.w0 * (cw0[q] * g0[ci0[q]] + cw1[q] * g0[ci1[q]]) +
               by.w1 * (cw0[q] * g1[ci0[q]] + cw1[q] * g1[ci1[q]]);
        hf[q] = 0.0f;
      }
      if (flip) {
        const float* f0 = g0 + AG_MAXP;
        const float* f1 = g1 + AG_MAXP;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          hf[q] = by.w0 * (cw0[q] * f0[ci0[q]] + cw1[q] * f0[ci1[q]]) +
                  by.w1 * (cw0[q] * f1[ci0[q]] + cw1[q] * f1[ci1[q]]);
      }
      const int xb = xb0 + cx;
      if (is_det) {
        float* d = a.det + ((size_t)n * a.J + c) * pb + (size_t)yb * a.Wb + xb;
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = flip ? (h[q] + hf[q]) * 0.5f : h[q];
          if (a.accumulate && xb + q < a.Wb) v = d[q] + v;
          if (a.final_div != 0.0f) v = __fdiv_rn(v, a.final_div);
          o[q] = v;
        }
        if (vec4) {
          *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (xb + q < a.Wb) d[q] = o[q];
        }
      } else {
        const int t = c - a.J;
        float* o = a.tag + (((size_t)n * a.A + t) * pb + (size_t)yb * a.Wb + xb) * T;
        if (flip) {
          if (vec4) {
            reinterpret_cast<float4*>(o)[0] = make_float4(h[0], hf[0], h[1], hf[1]);
            reinterpret_cast<float4*>(o)[1] = make_float4(h[2], hf[2], h[3], hf[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (xb + q < a.Wb) *reinterpret_cast<float2*>(o + 2 * q) = make_float2(h[q], hf[q]);
          }
        } else {
          if (vec4) {
            *reinterpret_cast<float4*>(o) = make_float4(h[0], h[1], h[2], h[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (xb + q < a.Wb) o[q] = h[q];
          }
        }
      }
    }
  }
}

// worst-case G patch of a tile, computed with the kernel's own float formulas
static bool agg_tiled_fits(int H2, int W2, int Hb, int Wb) {
  auto span = [](int in, int out, int tile) {
    const float sc = resize_scale(in, out, false);
    int worst = 0;
    for (int t0 = 0; t0 < out; t0 += tile) {
      const int t1 = (t0 + tile < out ? t0 + tile : out) - 1;
      auto idx0 = [&](int dst) {
        float s = sc * ((float)dst + 0.5f) - 0.5f;
        if (s < 0.0f) s = 0.0f;
        int i0 = (int)s;
        if (i0 > in - 1) i0 = in - 1;
        return i0;
      };
      const int lo = idx0(t0);
      int hi = idx0(t1);
      hi = hi + ((hi < in - 1) ? 1 : 0);
      if (hi - lo + 1 > worst) worst = hi - lo + 1;
    }
    return worst;
  };
  return (long)span(H2, Hb, AG_TH) * span(W2, Wb, AG_TW) <= AG_MAXP;
}

__global__ void __launch_bounds__(256) aggregate_scale_kernel(AggArgs a) {
  const int xb = blockIdx.x * blockDim.x + threadIdx.x;
  const int yb = blockIdx.y;
  if (xb >= a.Wb) return;
  const Lerp by = make_lerp(a.s2by, yb, a.H2, false);
  const Lerp bx = make_lerp(a.s2bx, xb, a.W2, false);
  const int C0 = a.J + a.A;
  const size_t p4 = (size_t)a.H4 * a.W4, p2 = (size_t)a.H2 * a.W2, pb = (size_t)a.Hb * a.Wb;
  const bool flip = a.y0f != nullptr;
  const int T = flip ? 2 : 1;
  const int nch = a.J + (a.tag ? a.A : 0);
  for (int nc = blockIdx.z; nc < a.N * nch; nc += gridDim.z) {
    const int n = nc / nch, c = nc - n * nch;
    if (c < a.J) {
      float h = cascade_sample(a, a.y0 + ((size_t)n * C0 + c) * p4, a.y1 + ((size_t)n * a.J + c) * p2,
                               by, bx, false);
      if (flip) {
        const int cf = a.flip_index[c];
        const float hf = cascade_sample(a, a.y0f + ((size_t)n * C0 + cf) * p4,
                                        a.y1f + ((size_t)n * a.J + cf) * p2, by, bx, true);
        h = (h + hf) * 0.5f;
      }
      float* d = a.det + ((size_t)n * a.J + c) * pb + (size_t)yb * a.Wb + xb;
      if (a.accumulate) h = *d + h;
      if (a.final_div != 0.0f) h = __fdiv_rn(h, a.final_div);
      *d = h;
    } else {
      const int t = c - a.J;
      float* o = a.tag + (((size_t)n * a.A + t) * pb + (size_t)yb * a.Wb + xb) * T;
      const float v = cascade_sample(a, a.y0 + ((size_t)n * C0 + a.J + t) * p4, nullptr, by, bx, false);
      if (flip) {
        const int tf = (a.A == a.J) ? a.flip_index[t] : t;
        const float vf = cascade_sample(a, a.y0f + ((size_t)n * C0 + a.J + tf) * p4, nullptr, by, bx, true);
        *reinterpret_cast<float2*>(o) = make_float2(v, vf);
      } else {
        *o = v;
      }
    }
  }
}

// ---- exact 2x -> 2x cascade (H2 = 2*H4, Hb = 2*H2: heat-maps projected to the network input
// size, the configuration of the flip-test benchmark).  With align_corners=False a 2x bilinear
// upsample has the fixed weights (0.25, 0.75) / (0.75, 0.25) and clamped neighbours, so the whole
// cascade of one 4x4 output block is a separable stencil on a 3x3 patch of the 1/4-resolution map
// and a 4x4 patch of the 1/2-resolution map.  One thread = one 1/4-resolution pixel = one 4x4
// output block per channel: ~20 instructions per output value instead of ~75 in the generic
// tiled kernel, float4 stores, every source value loaded once per thread.
// (Border values use 0.25*v + 0.75*v where PyTorch uses 1*v + 0*w: equal up to one rounding.)
__device__ __forceinline__ void up2_4(const float a, const float b, const float c, float (&o)[4]) {
  // neighbours (k-1, k, k+1) of a 1-D signal -> samples 2k-1, 2k, 2k+1, 2k+2 of its 2x upsample
  o[0] = 0.75f * a + 0.25f * b;
  o[1] = 0.25f * a + 0.75f * b;
  o[2] = 0.75f * b + 0.25f * c;
  o[3] = 0.25f * b + 0.75f * c;
}
__device__ __forceinline__ void up2_mid(const float g0, const float g1, const float g2, const float g3,
                                        float (&o)[4]) {
  // samples 2k-1 .. 2k+2 of a signal -> samples 4k .. 4k+3 of its 2x upsample
  o[0] = 0.25f * g0 + 0.75f * g1;
  o[1] = 0.75f * g1 + 0.25f * g2;
  o[2] = 0.25f * g1 + 0.75f * g2;
  o[3] = 0.75f * g2 + 0.25f * g3;
}

struct X4Idx {
  int r4[3], c4[3];     // clamped rows / columns of the 1/4-resolution map
  int r2[4], c2[4];     // clamped rows / columns of the 1/2-resolution map
  int c4f[3], c2f[4];   // the same columns in the mirrored (flipped-image) maps
};

// G (4x4 patch of the stage-average map at 1/2 resolution, rows 2by-1..2by+2, cols 2bx-1..2bx+2)
__device__ __forceinline__ void x4_patch(const float* __restrict__ q0, const float* __restrict__ q1,
                                         int W4, int W2, const int (&r4)[3], const int (&c4)[3],
                                         const int (&r2)[4], const int (&c2)[4], float (&G)[4][4]) {
  float h[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* row = q0 + (size_t)r4[r] * W4;
    up2_4(__ldg(row + c4[0]), __ldg(row + c4[1]), __ldg(row + c4[2]), h[r]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float col[4];
    up2_4(h[0][i], h[1][i], h[2][i], col);
#pragma unroll
    for (int j = 0; j < 4; ++j) G[j][i] = col[j];
  }
  if (q1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* row = q1 + (size_t)r2[j] * W2;
#pragma unroll
      for (int i = 0; i < 4; ++i) G[j][i] = (G[j][i] + __ldg(row + c2[i])) * 0.5f;
    }
  }
}

// 4x4 patch of G -> the 4x4 output block
__device__ __forceinline__ void x4_out(const float (&G)[4][4], float (&O)[4][4]) {
  float t[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) up2_mid(G[j][0], G[j][1], G[j][2], G[j][3], t[j]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float col[4];
    up2_mid(t[0][i], t[1][i], t[2][i], t[3][i], col);
#pragma unroll
    for (int j = 0; j < 4; ++j) O[j][i] = col[j];
  }
}

template <bool FLIP>
__global__ void __launch_bounds__(256, 2) aggregate_x4_kernel(AggArgs a) {
  const int bx = blockIdx.x * 64 + (threadIdx.x & 63);
  const int by = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int n = blockIdx.z;
  if (bx >= a.W4 || by >= a.H4) return;
  const int W4 = a.W4, H4 = a.H4, W2 = a.W2, H2 = a.H2, Wb = a.Wb, Hb = a.Hb;
  X4Idx ix;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ix.r4[k] = min(max(by - 1 + k, 0), H4 - 1);
    ix.c4[k] = min(max(bx - 1 + k, 0), W4 - 1);
    ix.c4f[k] = W4 - 1 - ix.c4[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ix.r2[k] = min(max(2 * by - 1 + k, 0), H2 - 1);
    ix.c2[k] = min(max(2 * bx - 1 + k, 0), W2 - 1);
    ix.c2f[k] = W2 - 1 - ix.c2[k];
  }
  const int C0 = a.J + a.A;
  const size_t p4 = (size_t)H4 * W4, p2 = (size_t)H2 * W2, pb = (size_t)Hb * Wb;
  const size_t obase = (size_t)(4 * by) * Wb + 4 * bx;
  constexpr int T = FLIP ? 2 : 1;

  // ---- heat-maps: det = up2( G ) or up2( (G + mirror(G_flip)) / 2 )
  for (int c = 0; c < a.J; ++c) {
    float G[4][4], O[4][4];
    x4_patch(a.y0 + ((size_t)n * C0 + c) * p4, a.y1 + ((size_t)n * a.J + c) * p2, W4, W2, ix.r4,
             ix.c4, ix.r2, ix.c2, G);
    if (FLIP) {
      const int cf = a.flip_index[c];
      float Gf[4][4];
      x4_patch(a.y0f + ((size_t)n * C0 + cf) * p4, a.y1f + ((size_t)n * a.J + cf) * p2, W4, W2,
               ix.r4, ix.c4f, ix.r2, ix.c2f, Gf);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) G[j][i] = (G[j][i] + Gf[j][i]) * 0.5f;
    }
    x4_out(G, O);
    float* d = a.det + ((size_t)n * a.J + c) * pb + obase;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 v = make_float4(O[j][0], O[j][1], O[j][2], O[j][3]);
      float4* dp = reinterpret_cast<float4*>(d + (size_t)j * Wb);
      if (a.accumulate) {
        const float4 o = *dp;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      if (a.final_div != 0.0f) {
        v.x = __fdiv_rn(v.x, a.final_div); v.y = __fdiv_rn(v.y, a.final_div);
        v.z = __fdiv_rn(v.z, a.final_div); v.w = __fdiv_rn(v.w, a.final_div);
      }
      *dp = v;
    }
  }
  // ---- tags: slot 0 = up2(up2(y0[J+t])), slot 1 = the mirrored flipped-image tags
  if (a.tag == nullptr) return;
  for (int t = 0; t < a.A; ++t) {
    float G[4][4], O[4][4];
    x4_patch(a.y0 + ((size_t)n * C0 + a.J + t) * p4, nullptr, W4, W2, ix.r4, ix.c4, ix.r2, ix.c2, G);
    x4_out(G, O);
    float* o = a.tag + (((size_t)n * a.A + t) * pb + obase) * T;
    if (FLIP) {
      const int tf = (a.A == a.J) ? a.flip_index[t] : t;
      float Gf[4][4], Of[4][4];
      x4_patch(a.y0f + ((size_t)n * C0 + a.J + tf) * p4, nullptr, W4, W2, ix.r4, ix.c4f, ix.r2,
               ix.c2f, Gf);
      x4_out(Gf, Of);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4* op = reinterpret_cast<float4*>(o + (size_t)j * Wb * 2);
        op[0] = make_float4(O[j][0], Of[j][0], O[j][1], Of[j][1]);
        op[1] = make_float4(O[j][2], Of[j][2], O[j][3], Of[j][3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(o + (size_t)j * Wb) = make_float4(O[j][0], O[j][1], O[j][2], O[j][3]);
    }
  }
}

// ---- the same exact 2x -> 2x cascade with TMA-staged sources --------------------------------
// aggregate_x4_kernel above is bound by the latency of its ~50 scattered source loads per channel
// (ncu r01k: 71 % of the stall samples are long-scoreboard, 21 % of the warp slots active).  Here
// a CTA owns a 32 x 8 block of 1/4-resolution pixels of one image and walks the channels; for
// every channel a producer lane fetches the 1/4- and 1/2-resolution source tiles (with their
// one-pixel aprons, zero-filled outside the map and never read there) of the image and of the
// mirrored flipped image as four cp.async.bulk.tensor boxes into a ring of shared-memory stages
// (full/empty mbarriers, 4 channels in flight), and the 8 consumer warps compute the 4x4 output
// blocks from shared memory with the arithmetic of x4_patch / x4_out.
constexpr int XT_W = 32, XT_H = 8;
// boxes start 4 columns left of the tile: the innermost TMA coordinate must stay 16-byte aligned
// (a box starting at column -1 faults), so the one-column apron costs a 4-column margin per side
constexpr int XQ_X0 = 4;
constexpr int XQ0_W = XT_W + 2 * XQ_X0, XQ0_H = XT_H + 2;
constexpr int XQ1_W = 2 * XT_W + 2 * XQ_X0, XQ1_H = 2 * XT_H + 2;
constexpr int XQ0_BYTES = XQ0_W * XQ0_H * 4, XQ1_BYTES = XQ1_W * XQ1_H * 4;
constexpr int XQ0_SLOT = (XQ0_BYTES + 127) / 128 * 128, XQ1_SLOT = (XQ1_BYTES + 127) / 128 * 128;
constexpr int X4_STAGE_BYTES = 2 * (XQ0_SLOT + XQ1_SLOT);
constexpr int X4_NSTG = 4;
constexpr int X4_CONSUMERS = XT_W * XT_H / 32;
constexpr int X4T_THREADS = (X4_CONSUMERS + 1) * 32;

struct X4Maps {
  CUtensorMap y0, y1, y0f, y1f;
};

__device__ __forceinline__ void x4_patch_s(const float* __restrict__ q0, const float* __restrict__ q1,
                                           const int (&r0)[3], const int (&c0)[3],
                                           const int (&r1)[4], const int (&c1)[4],
                                           float (&G)[4][4]) {
  float h[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    up2_4(q0[r0[r] + c0[0]], q0[r0[r] + c0[1]], q0[r0[r] + c0[2]], h[r]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float col[4];
    up2_4(h[0][i], h[1][i], h[2][i], col);
#pragma unroll
    for (int j = 0; j < 4; ++j) G[j][i] = col[j];
  }
  if (q1) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) G[j][i] = (G[j][i] + q1[r1[j] + c1[i]]) * 0.5f;
  }
}

template <bool FLIP>
__global__ void __launch_bounds__(X4T_THREADS, 2)
aggregate_x4_tma_kernel(AggArgs a, const __grid_constant__ X4Maps maps) {
  extern __shared__ unsigned char x4_dyn[];
  // TMA destinations must be 128-byte aligned: align by hand (the launch adds 128 bytes of slack)
  // (offset arithmetic, not a pointer round trip, so the loads below stay LDS and not generic LD)
  unsigned char* x4_raw = x4_dyn + ((128u - (smem_u32(x4_dyn) & 127u)) & 127u);
  __shared__ __align__(8) unsigned long long full_bar[X4_NSTG];
  __shared__ __align__(8) unsigned long long empty_bar[X4_NSTG];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bx0 = blockIdx.x * XT_W, by0 = blockIdx.y * XT_H;
  const int n = blockIdx.z;
  const int W4 = a.W4, H4 = a.H4, W2 = a.W2, H2 = a.H2, Wb = a.Wb;
  const int C0 = a.J + a.A;
  const int nitems = a.J + (a.tag ? a.A : 0);

  if (tid == 0) {
    for (int s = 0; s < X4_NSTG; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), X4_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == X4_CONSUMERS) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.y0);
      tma_prefetch_desc(&maps.y1);
      if (FLIP) {
        tma_prefetch_desc(&maps.y0f);
        tma_prefetch_desc(&maps.y1f);
      }
      int slot = 0;
      uint32_t ephase = 1;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t ring0 = smem_u32(x4_raw);
      for (int it = 0; it < nitems; ++it) {
        mbar_wait(empty0 + 8u * slot, ephase);
        const uint32_t bar = full0 + 8u * slot;
        const uint32_t base = ring0 + (uint32_t)slot * X4_STAGE_BYTES;
        const bool is_det = it < a.J;
        const int c = is_det ? it : it - a.J;
        const int cf = (is_det || a.A == a.J) ? a.flip_index[c] : c;
        const int ch0 = n * C0 + (is_det ? c : a.J + c);
        const int ch0f = n * C0 + (is_det ? cf : a.J + cf);
        const uint32_t bytes = (uint32_t)((FLIP ? 2 : 1) * (XQ0_BYTES + (is_det ? XQ1_BYTES : 0)));
        mbar_expect_tx(bar, bytes);
        tma_load_3d(base, &maps.y0, bar, bx0 - XQ_X0, by0 - 1, ch0);
        if (is_det)
          tma_load_3d(base + XQ0_SLOT, &maps.y1, bar, 2 * bx0 - XQ_X0, 2 * by0 - 1, n * a.J + c);
        if (FLIP) {
          tma_load_3d(base + XQ0_SLOT + XQ1_SLOT, &maps.y0f, bar, W4 - bx0 - (XT_W + XQ_X0), by0 - 1,
                      ch0f);
          if (is_det)
            tma_load_3d(base + 2 * XQ0_SLOT + XQ1_SLOT, &maps.y1f, bar,
                        W2 - 2 * bx0 - (2 * XT_W + XQ_X0), 2 * by0 - 1, n * a.J + cf);
        }
        if (++slot == X4_NSTG) { slot = 0; ephase ^= 1u; }
      }
    }
    return;
  }

  // ---- consumers
  const int bx = bx0 + (tid & 31), by = by0 + (tid >> 5);
  const bool active = bx < W4 && by < H4;
  int r0[3], c0[3], c0f[3], r1[4], c1[4], c1f[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r0[k] = (min(max(by - 1 + k, 0), H4 - 1) - (by0 - 1)) * XQ0_W;
    c0[k] = min(max(bx - 1 + k, 0), W4 - 1) - (bx0 - XQ_X0);
    c0f[k] = (XQ0_W - 1) - c0[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r1[k] = (min(max(2 * by - 1 + k, 0), H2 - 1) - (2 * by0 - 1)) * XQ1_W;
    c1[k] = min(max(2 * bx - 1 + k, 0), W2 - 1) - (2 * bx0 - XQ_X0);
    c1f[k] = (XQ1_W - 1) - c1[k];
  }
  if (!active) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { r0[k] = 0; c0[k] = 0; c0f[k] = 0; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { r1[k] = 0; c1[k] = 0; c1f[k] = 0; }
  }
  const size_t pb = (size_t)a.Hb * Wb;
  const size_t obase = (size_t)(4 * by) * Wb + 4 * bx;
  constexpr int T = FLIP ? 2 : 1;
  int slot = 0;
  uint32_t phase = 0;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  for (int it = 0; it < nitems; ++it) {
    mbar_wait(full0 + 8u * slot, phase);
    const float* q0 = reinterpret_cast<const float*>(x4_raw + (size_t)slot * X4_STAGE_BYTES);
    const float* q1 = q0 + XQ0_SLOT / 4;
    const float* q0f = q1 + XQ1_SLOT / 4;
    const float* q1f = q0f + XQ0_SLOT / 4;
    if (active) {
      if (it < a.J) {
        const int c = it;
        float G[4][4], O[4][4];
        x4_patch_s(q0, q1, r0, c0, r1, c1, G);
        if (FLIP) {
          float Gf[4][4];
          x4_patch_s(q0f, q1f, r0, c0f, r1, c1f, Gf);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) G[j][i] = (G[j][i] + Gf[j][i]) * 0.5f;
        }
        x4_out(G, O);
        float* d = a.det + ((size_t)n * a.J + c) * pb + obase;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 v = make_float4(O[j][0], O[j][1], O[j][2], O[j][3]);
          float4* dp = reinterpret_cast<float4*>(d + (size_t)j * Wb);
          if (a.accumulate) {
            const float4 o = *dp;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          if (a.final_div != 0.0f) {
            v.x = __fdiv_rn(v.x, a.final_div); v.y = __fdiv_rn(v.y, a.final_div);
            v.z = __fdiv_rn(v.z, a.final_div); v.w = __fdiv_rn(v.w, a.final_div);
          }
          *dp = v;
        }
      } else {
        const int t = it - a.J;
        float G[4][4], O[4][4];
        x4_patch_s(q0, nullptr, r0, c0, r1, c1, G);
        x4_out(G, O);
        float* o = a.tag + (((size_t)n * a.A + t) * pb + obase) * T;
        if (FLIP) {
          float Gf[4][4], Of[4][4];
          x4_patch_s(q0f, nullptr, r0, c0f, r1, c1f, Gf);
          x4_out(Gf, Of);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4* op = reinterpret_cast<float4*>(o + (size_t)j * Wb * 2);
            op[0] = make_float4(O[j][0], Of[j][0], O[j][1], Of[j][1]);
            op[1] = make_float4(O[j][2], Of[j][2], O[j][3], Of[j][3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(o + (size_t)j * Wb) = make_float4(O[j][0], O[j][1], O[j][2], O[j][3]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8u * slot);
    if (++slot == X4_NSTG) { slot = 0; phase ^= 1u; }
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 agg_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// (W, H, planes) fp32 tensor map with a (bw, bh, 1) box
static bool agg_make_map(CUtensorMap* m, const float* ptr, int W, int H, long long planes, int bw,
                         int bh) {
  auto encode = agg_encode_fn();
  if (!encode) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns 1 when launched, 0 when the shape / alignment does not qualify, < 0 on error
static int launch_x4_tma(const AggArgs& a, cudaStream_t st) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.W4 % 4 != 0 || a.W4 < XQ0_W || a.H4 < XQ0_H) return 0;
  if (!al16(a.y0) || !al16(a.y1) || (a.y0f && (!al16(a.y0f) || !al16(a.y1f)))) return 0;
  X4Maps maps;
  const long long p0 = (long long)a.N * (a.J + a.A), p1 = (long long)a.N * a.J;
  if (!agg_make_map(&maps.y0, a.y0, a.W4, a.H4, p0, XQ0_W, XQ0_H)) return 0;
  if (!agg_make_map(&maps.y1, a.y1, a.W2, a.H2, p1, XQ1_W, XQ1_H)) return 0;
  if (a.y0f) {
    if (!agg_make_map(&maps.y0f, a.y0f, a.W4, a.H4, p0, XQ0_W, XQ0_H)) return 0;
    if (!agg_make_map(&maps.y1f, a.y1f, a.W2, a.H2, p1, XQ1_W, XQ1_H)) return 0;
  } else {
    maps.y0f = maps.y0;
    maps.y1f = maps.y1;
  }
  const size_t smem = (size_t)X4_NSTG * X4_STAGE_BYTES + 128;
  dim3 grid(ceil_div(a.W4, XT_W), ceil_div(a.H4, XT_H), a.N);
  if (a.y0f) {
    BRTPE_CUDA(cudaFuncSetAttribute(aggregate_x4_tma_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aggregate_x4_tma_kernel<true><<<grid, X4T_THREADS, smem, st>>>(a, maps);
  } else {
    BRTPE_CUDA(cudaFuncSetAttribute(aggregate_x4_tma_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aggregate_x4_tma_kernel<false><<<grid, X4T_THREADS, smem, st>>>(a, maps);
  }
  return 1;
}

}  // namespace brtpe

using namespace brtpe;

extern "C" int brtpe_bilinear_resize(const float* src, long long src_plane_stride, int planes,
                                     int Hi, int Wi, float* dst, int Ho, int Wo, int dst_inner,
                                     int dst_off, int align_corners, void* stream) {
  BRTPE_CHECK_ARG(src && dst && planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0,
                  "brtpe_bilinear_resize: bad arguments");
  BRTPE_CHECK_ARG(dst_inner >= 1 && dst_off >= 0 && dst_off < dst_inner,
                  "brtpe_bilinear_resize: bad destination addressing");
  BRTPE_CHECK_ARG(Ho <= 65535, "brtpe_bilinear_resize: Ho too large");
  const bool ac = align_corners != 0;
  dim3 block(128);
  int gz = planes < 64 ? planes : 64;
  dim3 grid(ceil_div(Wo, 128), Ho, gz);
  bilinear_resize_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      src, src_plane_stride, planes, Hi, Wi, dst, Ho, Wo, dst_inner, dst_off,
      resize_scale(Hi, Ho, ac), resize_scale(Wi, Wo, ac), ac);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

extern "C" int brtpe_aggregate_scale(const float* y0, const float* y1, const float* y0f,
                                     const float* y1f, int N, int J, int A, int H4, int W4, int H2,
                                     int W2, int Hb, int Wb, const int32_t* flip_index_host,
                                     int accumulate, float final_div, float* det, float* tag_out,
                                     void* stream) {
  BRTPE_CHECK_ARG(y0 && y1 && det, "brtpe_aggregate_scale: null y0/y1/det");
  BRTPE_CHECK_ARG((y0f == nullptr) == (y1f == nullptr),
                  "brtpe_aggregate_scale: y0f and y1f must both be given or both be NULL");
  BRTPE_CHECK_ARG(N > 0 && J > 0 && J <= BRTPE_MAX_JOINTS && A >= 0 && A <= BRTPE_MAX_JOINTS,
                  "brtpe_aggregate_scale: bad channel counts");
  BRTPE_CHECK_ARG(H4 > 0 && W4 > 0 && H2 > 0 && W2 > 0 && Hb > 0 && Wb > 0 && Hb <= 65535,
                  "brtpe_aggregate_scale: bad sizes");
  BRTPE_CHECK_ARG(!y0f || flip_index_host, "brtpe_aggregate_scale: flip test needs flip_index");
  BRTPE_CHECK_ARG(!tag_out || A > 0, "brtpe_aggregate_scale: tag_out given but A == 0");
  if (tag_out && y0f)
    BRTPE_CHECK_ARG((reinterpret_cast<uintptr_t>(tag_out) & 7) == 0,
                    "brtpe_aggregate_scale: tag_out must be 8-byte aligned");
  AggArgs a;
  a.y0 = y0; a.y1 = y1; a.y0f = y0f; a.y1f = y1f; a.det = det; a.tag = tag_out;
  a.N = N; a.J = J; a.A = A; a.H4 = H4; a.W4 = W4; a.H2 = H2; a.W2 = W2; a.Hb = Hb; a.Wb = Wb;
  a.accumulate = accumulate;
  a.final_div = final_div;
  a.s42y = resize_scale(H4, H2, false);
  a.s42x = resize_scale(W4, W2, false);
  a.s2by = resize_scale(H2, Hb, false);
  a.s2bx = resize_scale(W2, Wb, false);
  for (int i = 0; i < BRTPE_MAX_JOINTS; ++i)
    a.flip_index[i] = (flip_index_host && i < J) ? flip_index_host[i] : i;
  for (int i = 0; i < J; ++i)
    BRTPE_CHECK_ARG(a.flip_index[i] >= 0 && a.flip_index[i] < J,
                    "brtpe_aggregate_scale: flip_index[%d] out of range", i);
  const int nch = N * (J + (tag_out ? A : 0));
  const bool aligned = ((reinterpret_cast<uintptr_t>(det) & 15) == 0) &&
                       (!tag_out || (reinterpret_cast<uintptr_t>(tag_out) & 15) == 0);
  const bool x4 = aligned && H2 == 2 * H4 && W2 == 2 * W4 && Hb == 2 * H2 && Wb == 2 * W2 &&
                  N <= 65535 && ceil_div(H4, 4) <= 65535 && !getenv("BRTPE_AGG_GENERIC");
  int tma = 0;
  if (x4 && !getenv("BRTPE_AGG_X4_V1")) {
    tma = launch_x4_tma(a, (cudaStream_t)stream);
    if (tma < 0) return tma;
  }
  if (tma == 1) {
    // launched above
  } else if (x4) {
    dim3 grid(ceil_div(W4, 64), ceil_div(H4, 4), N);
    if (y0f) aggregate_x4_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else aggregate_x4_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  } else if (aligned && agg_tiled_fits(H2, W2, Hb, Wb) && ceil_div(Hb, AG_TH) <= 65535) {
    dim3 grid(ceil_div(Wb, AG_TW), ceil_div(Hb, AG_TH), nch < 65535 ? nch : 65535);
    aggregate_scale_tiled_kernel<<<grid, AG_THREADS, 0, (cudaStream_t)stream>>>(a);
  } else {
    dim3 grid(ceil_div(Wb, 256), Hb, nch < 32 ? nch : 32);
    aggregate_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  }
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}
