// Image pre-processing of the teacher-inference driver on the GPU:
//   cv2.warpAffine(image, trans, size)             (rtpe/third_party/transforms.py:183-190)
//   torchvision ToTensor() + Normalize(mean, std)  (teacher_inference.py:70-73, :79)
// fused into one pass that reads the uint8 HxWx3 image and writes the normalised float32 CHW
// tensor the network takes (and/or the warped uint8 image).
//
// OpenCV's 8-bit INTER_LINEAR / BORDER_CONSTANT warp is fixed point (imgwarp.cpp,
// WarpAffineInvoker + remapBilinear): the inverse matrix is applied in double with 10 fractional
// bits (adelta/bdelta per column, X0/Y0 per row, round_delta = 16), coordinates are cut to 1/32
// pixel, the four weights are (32-fx)(32-fy)*32 ... (sum 2^15) and the result is
// (sum + 2^14) >> 15.  Reproduced bit for bit, so the kernel equals the reference's CPU path on
// the byte level (oracle/preprocess_ref.py, pinned against cv2 4.13 in this image).
// HBM-bound: reads Hs*Ws*3 bytes (taps hit L1/L2), writes 12 bytes per output pixel.
#include "common.cuh"

namespace brtpe {

struct WarpArgs {
  const uint8_t* img;
  int Hs, Ws, pitch;
  double m[6];            // INVERTED matrix (dst -> src), row major 2x3
  int Ho, Wo;
  float mean[3], std[3];
  float* out_f32;         // (3, Ho, Wo) or null
  uint8_t* out_u8;        // (Ho, Wo, 3) or null
};

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

__global__ void __launch_bounds__(256) warp_normalize_kernel(WarpArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= a.Wo) return;
  // explicit _rn operations: no FMA contraction, like the scalar C++ of the reference build
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(a.m[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(a.m[3], (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(a.m[1], (double)y), a.m[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(a.m[4], (double)y), a.m[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
  const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0ok = sx >= 0 && sx < a.Ws, x1ok = sx + 1 >= 0 && sx + 1 < a.Ws;
  const bool y0ok = sy >= 0 && sy < a.Hs, y1ok = sy + 1 >= 0 && sy + 1 < a.Hs;
  const uint8_t* r0 = a.img + (size_t)(y0ok ? sy : 0) * a.pitch;
  const uint8_t* r1 = a.img + (size_t)(y1ok ? sy + 1 : 0) * a.pitch;
  const int c0 = (x0ok ? sx : 0) * 3, c1 = (x1ok ? sx + 1 : 0) * 3;
  const size_t plane = (size_t)a.Ho * a.Wo;
  const size_t o = (size_t)y * a.Wo + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v00 = (y0ok && x0ok) ? __ldg(r0 + c0 + c) : 0;
    const int v01 = (y0ok && x1ok) ? __ldg(r0 + c1 + c) : 0;
    const int v10 = (y1ok && x0ok) ? __ldg(r1 + c0 + c) : 0;
    const int v11 = (y1ok && x1ok) ? __ldg(r1 + c1 + c) : 0;
    int v = (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
    v = max(0, min(255, v));
    if (a.out_u8) a.out_u8[o * 3 + c] = (uint8_t)v;
    if (a.out_f32)
      a.out_f32[c * plane + o] =
          __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), a.mean[c]), a.std[c]);
  }
}

}  // namespace brtpe

using namespace brtpe;

extern "C" int brtpe_preprocess_warp_normalize(const uint8_t* img, int Hs, int Ws, int src_pitch,
                                               const double* trans_host, int Ho, int Wo,
                                               const float* mean_host, const float* std_host,
                                               float* out_f32, uint8_t* out_u8, void* stream) {
  BRTPE_CHECK_ARG(img && trans_host, "brtpe_preprocess_warp_normalize: null image / matrix");
  BRTPE_CHECK_ARG(Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0 && Ho <= 65535 && src_pitch >= 3 * Ws,
                  "brtpe_preprocess_warp_normalize: bad sizes");
  BRTPE_CHECK_ARG(out_f32 || out_u8, "brtpe_preprocess_warp_normalize: no output requested");
  BRTPE_CHECK_ARG(!out_f32 || (mean_host && std_host),
                  "brtpe_preprocess_warp_normalize: float output needs mean and std");
  WarpArgs a;
  a.img = img; a.Hs = Hs; a.Ws = Ws; a.pitch = src_pitch; a.Ho = Ho; a.Wo = Wo;
  a.out_f32 = out_f32; a.out_u8 = out_u8;
  // cv::warpAffine inverts the 2x3 matrix in double (no WARP_INVERSE_MAP in the reference's call)
  double m[6];
  for (int i = 0; i < 6; ++i) m[i] = trans_host[i];
  double d = m[0] * m[4] - m[1] * m[3];
  d = d != 0.0 ? 1.0 / d : 0.0;
  const double a11 = m[4] * d, a22 = m[0] * d;
  m[0] = a11; m[1] *= -d; m[3] *= -d; m[4] = a22;
  const double b1 = -m[0] * m[2] - m[1] * m[5];
  const double b2 = -m[3] * m[2] - m[4] * m[5];
  m[2] = b1; m[5] = b2;
  for (int i = 0; i < 6; ++i) a.m[i] = m[i];
  for (int c = 0; c < 3; ++c) {
    a.mean[c] = mean_host ? mean_host[c] : 0.0f;
    a.std[c] = std_host ? std_host[c] : 1.0f;
  }
  dim3 grid(ceil_div(Wo, 256), Ho);
  warp_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}
