// Shared declarations of the convolution engine (FFMA + tcgen05 back ends, plan executor).
#pragma once
#include "common.cuh"

namespace brtpe {

int conv_validate(const brtpe_conv_desc* d);
int conv_ffma_launch(const brtpe_conv_desc* d, const void* in, const void* weights,
                     const float* bias, const void* residual, void* out, cudaStream_t st);
double conv_flops(const brtpe_conv_desc* d);

// ---- tcgen05 back end (conv_umma.cu)
struct UmmaConvPrepared;  // tensor maps + launch geometry, built once per plan op
bool umma_conv_supported(const brtpe_conv_desc* d, const char** why);
UmmaConvPrepared* umma_conv_prepare(const brtpe_conv_desc* d, const void* in, const void* weights);
void umma_conv_release(UmmaConvPrepared*);
int umma_conv_launch(const UmmaConvPrepared* p, const float* bias, const void* residual, void* out,
                     cudaStream_t st, const void* const* add_ptrs = nullptr, void* out2 = nullptr);

// ---- tcgen05 halo-tile back end for 3x3 stride-1 convs (conv_halo.cu)
struct HaloConvPrepared;
bool halo_conv_supported(const brtpe_conv_desc* d);
HaloConvPrepared* halo_conv_prepare(const brtpe_conv_desc* d, const void* in, const void* weights,
                                    void* out);
void halo_conv_release(HaloConvPrepared*);
int halo_conv_launch(const HaloConvPrepared* p, const float* bias, const void* residual, void* out,
                     cudaStream_t st, const void* const* add_ptrs = nullptr, void* out2 = nullptr);

// chain mode of the halo engine: conv1 -> conv2 (+ residual) of a BasicBlock in one launch whose work
// list keeps a sub-batch's tensors in L2 (conv_halo_impl.cuh); nullptr when the pair is not chainable
HaloConvPrepared* halo_chain_prepare(const brtpe_conv_desc* d0, const void* in, const void* w0, void* mid,
                                     const brtpe_conv_desc* d1, const void* w1, void* out);
int halo_chain_launch(const HaloConvPrepared* p, const float* bias0, const float* bias1,
                      const void* residual, void* out, cudaStream_t st);

int stem_conv1_launch(const void* img, int img_is_half, int N, int H, int W, const float* w,
                      const float* bias, int Cout, void* out, int out_dtype, cudaStream_t st);
int stem_im2col_launch(const void* img, int img_is_half, int N, int H, int W, void* out,
                       cudaStream_t st);
int fuse_sum_launch(int dtype, int nterms, const void* const* terms, const int32_t* shifts,
                    const int32_t* term_ld, int N, int H, int W, int C, void* out, int out_ld,
                    int relu, cudaStream_t st);
int nhwc_to_nchw_launch(int dtype, const void* src, int N, int H, int W, int C, int ld, int coff,
                        void* dst, int dst_is_half, cudaStream_t st);

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

}  // namespace brtpe
