// Shared epilogue of the tcgen05 convolution kernels: TMEM accumulator -> + folded-BN bias ->
// + residual -> ReLU -> bf16 -> global (NHWC).  One thread owns one output pixel (one TMEM
// lane) and walks the Cout tile in chunks of 16 columns = 32 bytes = one DRAM sector, moved
// with single 256-bit LDG/STG when the tensors allow it.  Everything is statically indexed
// (no local-memory arrays); the residual of the first EPI_PRE chunks is fetched BEFORE the
// thread waits for the accumulator so that its latency hides behind the MMAs.
#pragma once
#include "conv_common.cuh"
#include "umma_ptx.cuh"

namespace brtpe {

constexpr int EPI_PRE = 6;        // chunks (x16 channels) of residual prefetched per tile
constexpr int EPI_MAX_CHUNKS = 16;

struct EpiParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  int out_ld, out_coff, res_ld, res_coff, Cout, Cout_store, relu;
  int vec32;                      // 1: every 16-channel chunk is 32-byte aligned in out and res
};

struct Chunk32 {
  uint32_t w[8];
};

__device__ __forceinline__ Chunk32 ld_chunk32(const __nv_bfloat16* p, bool vec32) {
  Chunk32 c;
  if (vec32) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]),
                   "=r"(c.w[5]), "=r"(c.w[6]), "=r"(c.w[7])
                 : "l"(p));
  } else {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    c.w[0] = a.x; c.w[1] = a.y; c.w[2] = a.z; c.w[3] = a.w;
    c.w[4] = b.x; c.w[5] = b.y; c.w[6] = b.z; c.w[7] = b.w;
  }
  return c;
}
__device__ __forceinline__ void st_chunk32(__nv_bfloat16* p, const Chunk32& c, bool vec32) {
  if (vec32) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(c.w[0]),
                 "r"(c.w[1]), "r"(c.w[2]), "r"(c.w[3]), "r"(c.w[4]), "r"(c.w[5]), "r"(c.w[6]),
                 "r"(c.w[7])
                 : "memory");
  } else {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(c.w[4], c.w[5], c.w[6], c.w[7]);
  }
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct ResPrefetch {
  Chunk32 c[EPI_PRE];
};

__device__ __forceinline__ void epi_prefetch(ResPrefetch& rp, const EpiParams& e, bool valid,
                                             size_t opix, int co0, int nchunks) {
#pragma unroll
  for (int c = 0; c < EPI_PRE; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) rp.c[c].w[i] = 0u;
    if (e.res != nullptr && valid && c < nchunks && co0 + c * 16 + 16 <= e.Cout)
      rp.c[c] = ld_chunk32(e.res + opix * e.res_ld + e.res_coff + co0 + c * 16, e.vec32 != 0);
  }
}

// Drains one accumulator tile.  `arrive_bar`: tmem_empty barrier, arrived on (one lane per
// warp) as soon as the last TMEM read of this tile has completed.
__device__ __forceinline__ void epi_drain(const EpiParams& e, const float* __restrict__ bias_s,
                                          const ResPrefetch& rp, uint32_t t_addr, int nchunks,
                                          int co0, bool valid, size_t opix, uint32_t arrive_bar,
                                          int lane) {
  const bool vec32 = e.vec32 != 0;
#pragma unroll
  for (int c = 0; c < EPI_MAX_CHUNKS; ++c) {
    if (c >= nchunks) break;
    const int co = co0 + c * 16;
    const bool live = valid && co < e.Cout_store;
    const bool whole = co + 16 <= e.Cout;            // chunk entirely made of real channels
    Chunk32 rc;
#pragma unroll
    for (int i = 0; i < 8; ++i) rc.w[i] = 0u;
    if (c < EPI_PRE) {
      rc = rp.c[c < EPI_PRE ? c : 0];
    } else if (e.res != nullptr && live && whole) {
      rc = ld_chunk32(e.res + opix * e.res_ld + e.res_coff + co, vec32);
    }
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15;
    tmem_ld16s(t_addr + (uint32_t)(c * 16), r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12,
               r13, r14, r15);
    tmem_ld_wait();
    if (c == nchunks - 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(arrive_bar);
    }
    if (!live) continue;
    float v[16] = {__uint_as_float(r0),  __uint_as_float(r1),  __uint_as_float(r2),
                   __uint_as_float(r3),  __uint_as_float(r4),  __uint_as_float(r5),
                   __uint_as_float(r6),  __uint_as_float(r7),  __uint_as_float(r8),
                   __uint_as_float(r9),  __uint_as_float(r10), __uint_as_float(r11),
                   __uint_as_float(r12), __uint_as_float(r13), __uint_as_float(r14),
                   __uint_as_float(r15)};
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += bias_s[co + i];
    if (e.res != nullptr) {
      if (whole) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[2 * i] += bf16lo(rc.w[i]);
          v[2 * i + 1] += bf16hi(rc.w[i]);
        }
      } else {
        const __nv_bfloat16* rptr = e.res + opix * e.res_ld + e.res_coff + co;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (co + i < e.Cout) v[i] += __bfloat162float(rptr[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (e.relu) v[i] = fmaxf(v[i], 0.0f);
      if (co + i >= e.Cout) v[i] = 0.0f;
    }
    __nv_bfloat16* op = e.out + opix * e.out_ld + e.out_coff + co;
    if (co + 16 <= e.Cout_store) {
      Chunk32 oc;
#pragma unroll
      for (int i = 0; i < 8; ++i) oc.w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      st_chunk32(op, oc, vec32);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (co + i < e.Cout_store) op[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

static inline int epi_vec32_ok(const brtpe_conv_desc* d) {
  const bool out_ok = (d->out_ld % 16 == 0) && (d->out_coff % 16 == 0);
  const bool res_ok = (d->res_ld % 16 == 0) && (d->res_coff % 16 == 0);
  return (out_ok && res_ok) ? 1 : 0;
}

}  // namespace brtpe
