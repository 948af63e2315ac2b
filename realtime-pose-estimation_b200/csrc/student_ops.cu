// Small NHWC kernels of the context-aware-module student (rtpe/students.py, config 4):
//   * AvgPool2d(3, stride 2, padding 1, count_include_pad=False)   (students.py:657-658)
//   * SELayer: global average pool -> Linear -> ReLU -> Linear -> Sigmoid, returned as a gate
//     per (image, channel)                                          (students.py:118-142)
//   * ContextAwareModule tail: relu(residual + hdc * gate)          (students.py:197-201)
//   * attention injection: stem_out + sigmoid(att / 20)             (students.py:752-753)
//   * NHWC bilinear resize (MultistageStudent's out_hw path)        (students.py:481-498)
//   * AttentionStudentSteps: image -> NHWC (bilinear resize / space-to-depth), NHWC
//     space-to-depth (5x5 stride-2 convs run as 3x3 stride-1 convs on the 2x2 pixel-parity
//     planes), stem_out * sigmoid(att / divisor)                    (students.py:966-1018)
// All are bandwidth-bound streaming kernels over NHWC activations (fp32 or bf16, the precision
// mode of the plan); reductions are two-stage and deterministic (no atomics).
#include "conv_common.cuh"
#include <string.h>

namespace brtpe {

enum AuxKind { AUX_AVGPOOL = 1, AUX_SE_PARTIAL = 2, AUX_SE_GATE = 3, AUX_CAM_MIX = 4, AUX_ATT_ADD = 5,
               AUX_RESIZE_NHWC = 6, AUX_IMAGE_NHWC = 7, AUX_S2D = 8, AUX_ATT_MUL = 9 };

// PyTorch's area_pixel_compute_source_index / compute_scale in float32 (as aggregate.cu).
struct Lerp1 {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp1 lerp1(float scale, int dst, int in, bool ac) {
  float s = ac ? scale * (float)dst : scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.0f) s = 0.0f;
  Lerp1 l;
  l.i0 = (int)s;
  if (l.i0 > in - 1) l.i0 = in - 1;
  l.i1 = l.i0 + ((l.i0 < in - 1) ? 1 : 0);
  l.w1 = s - (float)l.i0;
  l.w0 = 1.0f - l.w1;
  return l;
}
static inline float lerp_scale(int in, int out, bool ac) {
  if (ac) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  return (float)in / (float)out;
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C,
                  int in_ld, int out_ld) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t p = i / C;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float s = 0.0f;
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * yo + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = 2 * xo + dx;
        if (x < 0 || x >= W) continue;
        s += to_f32(in[(((size_t)n * H + y) * W + x) * in_ld + c]);
        ++cnt;
      }
    }
    out[(((size_t)n * Ho + yo) * Wo + xo) * out_ld + c] = from_f32<T>(s / (float)cnt);
  }
}

// partial[n][chunk][c] = sum over the chunk's pixels of in[n][p][c]
template <typename T>
__global__ void __launch_bounds__(256)
se_partial_kernel(const T* __restrict__ in, float* __restrict__ partial, int HW, int C, int ld,
                  int chunks) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  __shared__ float red[256];
  // thread t owns channel t % C of pixel rows t / C, t / C + rows, ...
  const int rows = max(1, 256 / C);
  const int c = threadIdx.x % C, r = threadIdx.x / C;
  float s = 0.0f;
  if (r < rows)
    for (int p = p0 + r; p < p1; p += rows) s += to_f32(in[((size_t)n * HW + p) * ld + c]);
  red[threadIdx.x] = (r < rows) ? s : 0.0f;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.0f;
    for (int k = 0; k < rows; ++k) t += red[k * C + threadIdx.x];   // fixed order
    partial[((size_t)n * chunks + chunk) * C + threadIdx.x] = t;
  }
}

// gate[n][c] = sigmoid(W2 relu(W1 mean + b1) + b2); fc = [W1 (hid x Cin), b1 (hid), W2 (C x hid), b2 (C)]
// (Cin = pooled channels as they are stored, C = gate channels in module order)
__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ partial, const float* __restrict__ fc,
               float* __restrict__ gate, int Cin, int C, int hid, int chunks, int HW) {
  const int n = blockIdx.x;
  __shared__ float mean[256], h[256];
  const int t = threadIdx.x;
  if (t < Cin) {
    float s = 0.0f;
    for (int k = 0; k < chunks; ++k) s += partial[((size_t)n * chunks + k) * Cin + t];
    mean[t] = s / (float)HW;
  }
  __syncthreads();
  const float* w1 = fc;
  const float* b1 = w1 + (size_t)hid * Cin;
  const float* w2 = b1 + hid;
  const float* b2 = w2 + (size_t)C * hid;
  if (t < hid) {
    float s = b1[t];
    for (int k = 0; k < Cin; ++k) s += w1[t * Cin + k] * mean[k];
    h[t] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  if (t < C) {
    float s = b2[t];
    for (int k = 0; k < hid; ++k) s += w2[t * hid + k] * h[k];
    gate[(size_t)n * C + t] = 1.0f / (1.0f + expf(-s));
  }
}

// out = relu(res + hdc * gate) for c < C; channels [C, Cz) of out are written as zeros (Cz = C: none)
template <typename T>
__global__ void __launch_bounds__(256)
cam_mix_kernel(const T* __restrict__ res, const T* __restrict__ hdc, const float* __restrict__ gate,
               T* __restrict__ out, int N, int HW, int C, int ld_res, int ld_hdc, int ld_out, int Cz) {
  const size_t total = (size_t)N * HW * Cz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cz);
    const size_t p = i / Cz;
    const int n = (int)(p / HW);
    float v = 0.0f;
    if (c < C)
      v = fmaxf(to_f32(res[p * ld_res + c]) + to_f32(hdc[p * ld_hdc + c]) * gate[(size_t)n * C + c], 0.0f);
    out[p * ld_out + c] = from_f32<T>(v);
  }
}

// att_out[p] = sigmoid(att[p][0] / 20)  (float, = NCHW with one channel);
// out[p][c] = stem[p][c] + att_out[p]
template <typename T>
__global__ void __launch_bounds__(256)
att_add_kernel(const T* __restrict__ att, const T* __restrict__ stem, T* __restrict__ out,
               float* __restrict__ att_out, size_t P, int C, int ld_att, int ld_stem, int ld_out) {
  const size_t total = P * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const float a = 1.0f / (1.0f + expf(-to_f32(att[p * ld_att]) / 20.0f));
    if (c == 0) att_out[p] = a;
    out[p * ld_out + c] = from_f32<T>(to_f32(stem[p * ld_stem + c]) + a);
  }
}

// NHWC bilinear resize of C channels: in (N,H,W,in_ld) -> out (N,Ho,Wo,out_ld) + out_coff
template <typename T>
__global__ void __launch_bounds__(256)
resize_nhwc_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C,
                   int in_ld, int out_ld, int Ho, int Wo, float sy, float sx, bool ac, int out_coff) {
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t p = i / C;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    const Lerp1 ly = lerp1(sy, yo, H, ac), lx = lerp1(sx, xo, W, ac);
    const T* b = in + (size_t)n * H * W * in_ld + c;
    const float v00 = to_f32(b[((size_t)ly.i0 * W + lx.i0) * in_ld]);
    const float v01 = to_f32(b[((size_t)ly.i0 * W + lx.i1) * in_ld]);
    const float v10 = to_f32(b[((size_t)ly.i1 * W + lx.i0) * in_ld]);
    const float v11 = to_f32(b[((size_t)ly.i1 * W + lx.i1) * in_ld]);
    const float v = ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
    out[(((size_t)n * Ho + yo) * Wo + xo) * out_ld + out_coff + c] = from_f32<T>(v);
  }
}

// NCHW image (float / half) -> NHWC slice.  s2d == 0: bilinear resize (align_corners=False) to
// (Ho, Wo), channels [out_coff, out_coff + C) get the image, [out_coff + C, out_coff + Cz) zeros.
// s2d == 1: space-to-depth by 2 (Ho = H/2, Wo = W/2), channel (ry*2+rx)*C + c of output pixel
// (y, x) = img[c][2y+ry][2x+rx]; channels [4C, Cz) zeros.
template <typename T, typename I>
__global__ void __launch_bounds__(256)
image_nhwc_kernel(const I* __restrict__ img, T* __restrict__ out, int N, int H, int W, int C,
                  int out_ld, int Ho, int Wo, float sy, float sx, int out_coff, int Cz, int s2d) {
  const size_t total = (size_t)N * Ho * Wo * Cz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cz);
    size_t p = i / Cz;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float v = 0.0f;
    if (s2d) {
      if (c < 4 * C) {
        const int blk = c / C, ci = c % C;
        const I* b = img + ((size_t)n * C + ci) * H * W;
        v = to_f32(b[(size_t)(2 * yo + (blk >> 1)) * W + 2 * xo + (blk & 1)]);
      }
    } else if (c < C) {
      const Lerp1 ly = lerp1(sy, yo, H, false), lx = lerp1(sx, xo, W, false);
      const I* b = img + ((size_t)n * C + c) * H * W;
      const float v00 = to_f32(b[(size_t)ly.i0 * W + lx.i0]), v01 = to_f32(b[(size_t)ly.i0 * W + lx.i1]);
      const float v10 = to_f32(b[(size_t)ly.i1 * W + lx.i0]), v11 = to_f32(b[(size_t)ly.i1 * W + lx.i1]);
      v = ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
    }
    out[(((size_t)n * Ho + yo) * Wo + xo) * out_ld + out_coff + c] = from_f32<T>(v);
  }
}

// NHWC space-to-depth by 2: out (N,H/2,W/2,out_ld), channel (ry*2+rx)*C + c = in[2y+ry][2x+rx][c]
template <typename T>
__global__ void __launch_bounds__(256)
s2d_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int in_ld,
           int out_ld) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * 4 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t p = i / C;
    const int blk = (int)(p % 4);
    p /= 4;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    out[(((size_t)n * Ho + yo) * Wo + xo) * out_ld + blk * C + c] =
        in[(((size_t)n * H + 2 * yo + (blk >> 1)) * W + 2 * xo + (blk & 1)) * in_ld + c];
  }
}

// att_out[p] = sigmoid(att[p][0] / div); out[p][c] = x[p][c] * att_out[p]   (students.py:1001-1018)
template <typename T>
__global__ void __launch_bounds__(256)
att_mul_kernel(const T* __restrict__ att, const T* __restrict__ x, T* __restrict__ out,
               float* __restrict__ att_out, size_t P, int C, int ld_att, int ld_x, int ld_out,
               float div) {
  const size_t total = P * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const float a = 1.0f / (1.0f + expf(-(to_f32(att[p * ld_att]) / div)));
    if (c == 0) att_out[p] = a;
    out[p * ld_out + c] = from_f32<T>(to_f32(x[p * ld_x + c]) * a);
  }
}

// ---- bf16 fast paths: one thread = 8 consecutive channels (one 16-byte load / store) ---------
struct BF8 {
  uint4 u;
};
__device__ __forceinline__ void bf8_to_f32(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 f32_to_bf8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ld8(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

__global__ void __launch_bounds__(256)
avgpool3s2_bf8_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H,
                      int W, int C8, int in_ld, int out_ld) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    size_t p = i / C8;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * yo + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = 2 * xo + dx;
        if (x < 0 || x >= W) continue;
        float f[8];
        bf8_to_f32(ld8(in + (((size_t)n * H + y) * W + x) * in_ld + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
        ++cnt;
      }
    }
    const float inv = (float)cnt;
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = s[j] / inv;
    st8(out + (((size_t)n * Ho + yo) * Wo + xo) * out_ld + c, f32_to_bf8(s));
  }
}

// partial[n][chunk][c]: thread = 8 channels of one pixel row slot; rows = 256 / C8 pixels in flight;
// the per-row partial sums are reduced in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
se_partial_bf8_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ partial, int HW, int C8,
                      int ld, int chunks) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  __shared__ float red[256 * 8];
  const int rows = 256 / C8;
  const int c8 = threadIdx.x % C8, r = threadIdx.x / C8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows)
    for (int p = p0 + r; p < p1; p += rows) {
      float f[8];
      bf8_to_f32(ld8(in + ((size_t)n * HW + p) * ld + c8 * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = (r < rows) ? s[j] : 0.0f;
  __syncthreads();
  const int C = C8 * 8;
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.0f;
    for (int k = 0; k < rows; ++k) t += red[(k * C8 + c / 8) * 8 + (c & 7)];
    partial[((size_t)n * chunks + chunk) * C + c] = t;
  }
}

__global__ void __launch_bounds__(256)
cam_mix_bf8_kernel(const __nv_bfloat16* __restrict__ res, const __nv_bfloat16* __restrict__ hdc,
                   const float* __restrict__ gate, __nv_bfloat16* __restrict__ out, int N, int HW, int C,
                   int ld_res, int ld_hdc, int ld_out, int Cz8) {
  const size_t total = (size_t)N * HW * Cz8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cz8) * 8;
    const size_t p = i / Cz8;
    const int n = (int)(p / HW);
    float a[8], b[8], v[8];
    bf8_to_f32(ld8(res + p * ld_res + c), a);
    bf8_to_f32(ld8(hdc + p * ld_hdc + c), b);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (c + j < C) ? fmaxf(a[j] + b[j] * __ldg(gate + (size_t)n * C + c + j), 0.0f) : 0.0f;
    st8(out + p * ld_out + c, f32_to_bf8(v));
  }
}

// MUL = false: out = x + sigmoid(att / div)  (AttentionStudent); MUL = true: out = x * sigmoid(att / div)
template <bool MUL>
__global__ void __launch_bounds__(256)
att_apply_bf8_kernel(const __nv_bfloat16* __restrict__ att, const __nv_bfloat16* __restrict__ x,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ att_out, size_t P, int C8,
                     int ld_att, int ld_x, int ld_out, float div) {
  const size_t total = P * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const size_t p = i / C8;
    const float a = 1.0f / (1.0f + expf(-(__bfloat162float(att[p * ld_att]) / div)));
    if (c == 0) att_out[p] = a;
    float f[8];
    bf8_to_f32(ld8(x + p * ld_x + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = MUL ? f[j] * a : f[j] + a;
    st8(out + p * ld_out + c, f32_to_bf8(f));
  }
}

__global__ void __launch_bounds__(256)
s2d_bf8_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H, int W,
               int C8, int in_ld, int out_ld) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Ho * Wo * 4 * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    size_t p = i / C8;
    const int blk = (int)(p % 4);
    p /= 4;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const int n = (int)(p / Ho);
    st8(out + (((size_t)n * Ho + yo) * Wo + xo) * out_ld + blk * C8 * 8 + c,
        ld8(in + (((size_t)n * H + 2 * yo + (blk >> 1)) * W + 2 * xo + (blk & 1)) * in_ld + c));
  }
}

static int grid_for(size_t total) {
  size_t b = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// ip: kind-specific integer parameters (see the brtpe_aux_run documentation in brtpe.h)
int aux_launch(int kind, const void* in0, const void* in1, const void* in2, void* out,
               const int32_t* ip, cudaStream_t st) {
  const bool bf = ip[0] == BRTPE_DT_BF16;
  switch (kind) {
    case AUX_AVGPOOL: {
      const int N = ip[1], H = ip[2], W = ip[3], C = ip[4], ild = ip[5], old = ip[6];
      BRTPE_CHECK_ARG(N > 0 && H >= 2 && W >= 2 && !(H & 1) && !(W & 1) && C > 0, "avgpool: bad shape");
      const int g = grid_for((size_t)N * (H / 2) * (W / 2) * C);
      if (bf && C % 8 == 0 && ild % 8 == 0 && old % 8 == 0)
        avgpool3s2_bf8_kernel<<<grid_for((size_t)N * (H / 2) * (W / 2) * (C / 8)), 256, 0, st>>>((const __nv_bfloat16*)in0, (__nv_bfloat16*)out, N, H, W, C / 8, ild, old);
      else if (bf) avgpool3s2_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (__nv_bfloat16*)out, N, H, W, C, ild, old);
      else avgpool3s2_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (float*)out, N, H, W, C, ild, old);
      break;
    }
    case AUX_SE_PARTIAL: {
      const int N = ip[1], HW = ip[2], C = ip[3], ld = ip[4], chunks = ip[5];
      BRTPE_CHECK_ARG(N > 0 && HW > 0 && C > 0 && C <= 256 && chunks > 0, "se_partial: bad shape");
      dim3 grid(chunks, N);
      if (bf && C % 8 == 0 && ld % 8 == 0)
        se_partial_bf8_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in0, (float*)out, HW, C / 8, ld, chunks);
      else if (bf) se_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in0, (float*)out, HW, C, ld, chunks);
      else se_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)in0, (float*)out, HW, C, ld, chunks);
      break;
    }
    case AUX_SE_GATE: {
      const int N = ip[1], C = ip[2], hid = ip[3], chunks = ip[4], HW = ip[5];
      const int Cin = ip[6] > 0 ? ip[6] : C;             // pooled (stored) channels
      BRTPE_CHECK_ARG(N > 0 && C > 0 && C <= 256 && Cin <= 256 && hid > 0 && hid <= 256, "se_gate: bad shape");
      se_gate_kernel<<<N, 256, 0, st>>>((const float*)in0, (const float*)in1, (float*)out, Cin, C, hid, chunks, HW);
      break;
    }
    case AUX_CAM_MIX: {
      const int N = ip[1], HW = ip[2], C = ip[3];
      const int Cz = ip[7] > C ? ip[7] : C;              // zero-fill the pad channels [C, Cz)
      BRTPE_CHECK_ARG(Cz <= ip[6], "cam_mix: zero fill beyond ld_out");
      const int g = grid_for((size_t)N * HW * Cz);
      if (bf && Cz % 8 == 0 && ip[4] % 8 == 0 && ip[5] % 8 == 0 && ip[6] % 8 == 0 && ip[4] >= Cz && ip[5] >= Cz)
        cam_mix_bf8_kernel<<<grid_for((size_t)N * HW * (Cz / 8)), 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (const float*)in2, (__nv_bfloat16*)out, N, HW, C, ip[4], ip[5], ip[6], Cz / 8);
      else if (bf) cam_mix_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (const float*)in2, (__nv_bfloat16*)out, N, HW, C, ip[4], ip[5], ip[6], Cz);
      else cam_mix_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (const float*)in1, (const float*)in2, (float*)out, N, HW, C, ip[4], ip[5], ip[6], Cz);
      break;
    }
    case AUX_ATT_ADD: {
      const size_t P = (size_t)ip[1] * ip[2];
      const int C = ip[3];
      const int g = grid_for(P * C);
      float* att_out = (float*)const_cast<void*>(in2);
      if (bf && C % 8 == 0 && ip[5] % 8 == 0 && ip[6] % 8 == 0)
        att_apply_bf8_kernel<false><<<grid_for(P * (C / 8)), 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (__nv_bfloat16*)out, att_out, P, C / 8, ip[4], ip[5], ip[6], 20.0f);
      else if (bf) att_add_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (__nv_bfloat16*)out, att_out, P, C, ip[4], ip[5], ip[6]);
      else att_add_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (const float*)in1, (float*)out, att_out, P, C, ip[4], ip[5], ip[6]);
      break;
    }
    case AUX_RESIZE_NHWC: {
      const int N = ip[1], H = ip[2], W = ip[3], C = ip[4], ild = ip[5], old = ip[6], Ho = ip[7],
                Wo = ip[8], coff = ip[10];
      const bool ac = ip[9] != 0;
      BRTPE_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && C <= ild && coff >= 0 &&
                      coff + C <= old, "resize_nhwc: bad shape");
      const int g = grid_for((size_t)N * Ho * Wo * C);
      const float sy = lerp_scale(H, Ho, ac), sx = lerp_scale(W, Wo, ac);
      if (bf) resize_nhwc_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (__nv_bfloat16*)out, N, H, W, C, ild, old, Ho, Wo, sy, sx, ac, coff);
      else resize_nhwc_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (float*)out, N, H, W, C, ild, old, Ho, Wo, sy, sx, ac, coff);
      break;
    }
    case AUX_IMAGE_NHWC: {
      const int N = ip[1], H = ip[2], W = ip[3], C = ip[4], half = ip[5], old = ip[6], Ho = ip[7],
                Wo = ip[8], coff = ip[9], Cz = ip[10], s2d = ip[11];
      BRTPE_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && coff >= 0 &&
                      Cz >= (s2d ? 4 * C : C) && coff + Cz <= old, "image_nhwc: bad shape");
      BRTPE_CHECK_ARG(!s2d || (Ho * 2 == H && Wo * 2 == W), "image_nhwc: space-to-depth needs Ho = H/2, Wo = W/2");
      const int g = grid_for((size_t)N * Ho * Wo * Cz);
      const float sy = lerp_scale(H, Ho, false), sx = lerp_scale(W, Wo, false);
      if (bf) {
        if (half) image_nhwc_kernel<__nv_bfloat16, __half><<<g, 256, 0, st>>>((const __half*)in0, (__nv_bfloat16*)out, N, H, W, C, old, Ho, Wo, sy, sx, coff, Cz, s2d);
        else image_nhwc_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const float*)in0, (__nv_bfloat16*)out, N, H, W, C, old, Ho, Wo, sy, sx, coff, Cz, s2d);
      } else {
        if (half) image_nhwc_kernel<float, __half><<<g, 256, 0, st>>>((const __half*)in0, (float*)out, N, H, W, C, old, Ho, Wo, sy, sx, coff, Cz, s2d);
        else image_nhwc_kernel<float, float><<<g, 256, 0, st>>>((const float*)in0, (float*)out, N, H, W, C, old, Ho, Wo, sy, sx, coff, Cz, s2d);
      }
      break;
    }
    case AUX_S2D: {
      const int N = ip[1], H = ip[2], W = ip[3], C = ip[4], ild = ip[5], old = ip[6];
      BRTPE_CHECK_ARG(N > 0 && H >= 2 && W >= 2 && !(H & 1) && !(W & 1) && C > 0 && C <= ild &&
                      4 * C <= old, "space_to_depth: bad shape");
      const int g = grid_for((size_t)N * (H / 2) * (W / 2) * 4 * C);
      if (bf && C % 8 == 0 && ild % 8 == 0 && old % 8 == 0)
        s2d_bf8_kernel<<<grid_for((size_t)N * (H / 2) * (W / 2) * 4 * (C / 8)), 256, 0, st>>>((const __nv_bfloat16*)in0, (__nv_bfloat16*)out, N, H, W, C / 8, ild, old);
      else if (bf) s2d_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (__nv_bfloat16*)out, N, H, W, C, ild, old);
      else s2d_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (float*)out, N, H, W, C, ild, old);
      break;
    }
    case AUX_ATT_MUL: {
      const size_t P = (size_t)ip[1] * ip[2];
      const int C = ip[3];
      float div;
      memcpy(&div, &ip[7], sizeof(float));
      BRTPE_CHECK_ARG(P > 0 && C > 0 && in1 && in2 && div != 0.0f, "att_mul: bad arguments");
      const int g = grid_for(P * C);
      float* att_out = (float*)const_cast<void*>(in2);
      if (bf && C % 8 == 0 && ip[5] % 8 == 0 && ip[6] % 8 == 0)
        att_apply_bf8_kernel<true><<<grid_for(P * (C / 8)), 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (__nv_bfloat16*)out, att_out, P, C / 8, ip[4], ip[5], ip[6], div);
      else if (bf) att_mul_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)in0, (const __nv_bfloat16*)in1, (__nv_bfloat16*)out, att_out, P, C, ip[4], ip[5], ip[6], div);
      else att_mul_kernel<float><<<g, 256, 0, st>>>((const float*)in0, (const float*)in1, (float*)out, att_out, P, C, ip[4], ip[5], ip[6], div);
      break;
    }
    default:
      set_error("brtpe_aux_run: unknown kind %d", kind);
      return BRTPE_EINVAL;
  }
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

}  // namespace brtpe

extern "C" int brtpe_aux_run(int kind, const void* in0, const void* in1, const void* in2, void* out,
                             const int32_t* iparams, int nparams, void* stream) {
  BRTPE_CHECK_ARG(in0 && out && iparams && nparams >= 4 && nparams <= 12, "brtpe_aux_run: bad arguments");
  int32_t ip[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nparams; ++i) ip[i] = iparams[i];
  return brtpe::aux_launch(kind, in0, in1, in2, out, ip, (cudaStream_t)stream);
}
