// brtpe_prepack_weights: fold an eval-mode BatchNorm into its convolution and lay the weights out
// for the convolution engine, ONE launch per layer (SURVEY.md 8b minimum export set).
//
// Replaces, per conv+BN pair of the network (rtpe/third_party/pose_higher_hrnet.py conv+bn pairs,
// e.g. :57-64, :96-106, :202-229, :514-521), the chain of ATen element-wise launches
//   scale = gamma / sqrt(var + eps); w' = w * scale; b' = beta - mean * scale (+ bias * scale);
//   gather taps; pad channels; cast to bf16; transpose
// with the same float32 arithmetic (IEEE sqrt / div / mul, no FMA contraction), so the packed
// weights are bit-identical to that chain.
#include "common.cuh"

namespace brtpe {

struct PrepackParams {
  brtpe_prepack_desc d;
  const void* w;
  const void* conv_bias;
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  const int32_t* cin_index;
  void* packed;
  float* bias_out;
  int bias_len;
  long long total;
};

__device__ __forceinline__ float load_w(const void* p, int dtype, long long i) {
  if (dtype == BRTPE_WT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == BRTPE_WT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

__device__ __forceinline__ float bn_scale(const PrepackParams& p, int co) {
  if (p.gamma == nullptr) return 1.0f;
  return __fdiv_rn(p.gamma[co], __fsqrt_rn(__fadd_rn(p.var[co], p.d.bn_eps)));
}

__global__ void __launch_bounds__(256) prepack_weights_kernel(const PrepackParams p) {
  const brtpe_prepack_desc& d = p.d;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // ---- bias (the first bias_len threads of the grid)
  if (idx < p.bias_len && p.bias_out != nullptr) {
    const int co = (int)idx;
    float b = 0.0f;
    if (co < d.Cout) {
      const float cb = p.conv_bias ? load_w(p.conv_bias, d.w_dtype, co) : 0.0f;
      if (p.gamma != nullptr) {
        const float s = bn_scale(p, co);
        b = __fsub_rn(p.beta[co], __fmul_rn(p.mean[co], s));
        if (p.conv_bias) b = __fadd_rn(b, __fmul_rn(cb, s));
      } else {
        b = cb;
      }
    }
    p.bias_out[co] = b;
  }
  if (idx >= p.total) return;
  // ---- one packed weight element
  int t, co, ci;
  if (d.layout == BRTPE_PACK_KMAJOR_BF16) {          // [ntaps][cout_pad][cin_pad]
    ci = (int)(idx % d.cin_pad);
    const long long r = idx / d.cin_pad;
    co = (int)(r % d.cout_pad);
    t = (int)(r / d.cout_pad);
  } else {                                            // [ntaps][Cin_store][Cout_pack]
    co = (int)(idx % d.Cout_pack);
    const long long r = idx / d.Cout_pack;
    ci = (int)(r % d.Cin_store);
    t = (int)(r / d.Cin_store);
  }
  // split (BRTPE_DT_BF16X2) layers: K = [w_hi | w_hi | w_lo], each segment padded to 64 channels
  int seg = 0;
  if (d.split) {
    const int seg_pad = d.cin_pad / 3;
    seg = ci / seg_pad;
    ci -= seg * seg_pad;
  }
  float v = 0.0f;
  if (co < d.Cout && ci < d.Cin_store) {
    int src = p.cin_index ? p.cin_index[ci] : ci;
    int kh = d.tap_kh[t], kw = d.tap_kw[t];
    bool ok = src >= 0;
    if (d.im2col && ok) {                             // stored channel = (kh*KW + kw)*Cin + c
      const int kpos = src / d.Cin;
      src -= kpos * d.Cin;
      kh = kpos / d.KW;
      kw = kpos - kh * d.KW;
      ok = kpos < d.KH * d.KW;
    }
    if (ok && src < d.Cin) {
      const long long wi = d.transposed
                               ? (((long long)src * d.Cout + co) * d.KH + kh) * d.KW + kw
                               : (((long long)co * d.Cin + src) * d.KH + kh) * d.KW + kw;
      v = __fmul_rn(load_w(p.w, d.w_dtype, wi), bn_scale(p, co));
      if (seg == 2) v = __fsub_rn(v, __bfloat162float(__float2bfloat16_rn(v)));
    }
  }
  if (d.layout == BRTPE_PACK_KMAJOR_BF16) {
    reinterpret_cast<__nv_bfloat16*>(p.packed)[idx] = __float2bfloat16_rn(v);
  } else {
    if (d.round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
    reinterpret_cast<float*>(p.packed)[idx] = v;
  }
}

}  // namespace brtpe

using namespace brtpe;

extern "C" int brtpe_prepack_weights(const brtpe_prepack_desc* d, const void* w,
                                     const void* conv_bias, const float* bn_gamma,
                                     const float* bn_beta, const float* bn_mean,
                                     const float* bn_var, const int32_t* cin_index, void* packed,
                                     float* bias_out, int bias_len, void* stream) {
  BRTPE_CHECK_ARG(d && w && packed, "brtpe_prepack_weights: null argument");
  BRTPE_CHECK_ARG(d->w_dtype >= BRTPE_WT_F32 && d->w_dtype <= BRTPE_WT_F16,
                  "brtpe_prepack_weights: w_dtype %d", d->w_dtype);
  BRTPE_CHECK_ARG(d->Cout > 0 && d->Cin > 0 && d->KH > 0 && d->KW > 0 && d->Cin_store > 0,
                  "brtpe_prepack_weights: bad module dimensions");
  BRTPE_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9, "brtpe_prepack_weights: ntaps %d", d->ntaps);
  for (int t = 0; t < d->ntaps; ++t)
    BRTPE_CHECK_ARG(d->tap_kh[t] >= 0 && d->tap_kh[t] < d->KH && d->tap_kw[t] >= 0 &&
                        d->tap_kw[t] < d->KW,
                    "brtpe_prepack_weights: tap %d = (%d, %d) outside the %dx%d kernel", t,
                    d->tap_kh[t], d->tap_kw[t], d->KH, d->KW);
  BRTPE_CHECK_ARG(!d->im2col || d->ntaps == 1, "brtpe_prepack_weights: im2col needs ntaps == 1");
  const bool bn = bn_gamma != nullptr;
  BRTPE_CHECK_ARG(!bn || (bn_beta && bn_mean && bn_var),
                  "brtpe_prepack_weights: BatchNorm needs gamma, beta, mean and var");
  BRTPE_CHECK_ARG(bias_out == nullptr || bias_len >= d->Cout,
                  "brtpe_prepack_weights: bias_len %d < Cout %d", bias_len, d->Cout);
  long long total;
  if (d->layout == BRTPE_PACK_KMAJOR_BF16) {
    BRTPE_CHECK_ARG(!d->split || (d->cin_pad % 3 == 0 && d->cin_pad / 3 >= d->Cin_store),
                    "brtpe_prepack_weights: split layers need cin_pad = 3 x padded Cin_store");
    BRTPE_CHECK_ARG(d->cin_pad >= d->Cin_store && d->cout_pad >= d->Cout,
                    "brtpe_prepack_weights: padded dims (%d, %d) smaller than (%d, %d)", d->cin_pad,
                    d->cout_pad, d->Cin_store, d->Cout);
    total = (long long)d->ntaps * d->cout_pad * d->cin_pad;
  } else {
    BRTPE_CHECK_ARG(!d->split, "brtpe_prepack_weights: split needs the K-major bf16 layout");
    BRTPE_CHECK_ARG(d->layout == BRTPE_PACK_CIN_COUT_F32 && d->Cout_pack >= d->Cout,
                    "brtpe_prepack_weights: bad layout %d / Cout_pack %d", d->layout, d->Cout_pack);
    total = (long long)d->ntaps * d->Cin_store * d->Cout_pack;
  }
  PrepackParams p;
  p.d = *d;
  p.w = w; p.conv_bias = conv_bias;
  p.gamma = bn_gamma; p.beta = bn_beta; p.mean = bn_mean; p.var = bn_var;
  p.cin_index = cin_index;
  p.packed = packed; p.bias_out = bias_out; p.bias_len = bias_out ? bias_len : 0;
  p.total = total;
  const long long work = total > p.bias_len ? total : p.bias_len;
  const int blocks = (int)((work + 255) / 256);
  prepack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}
