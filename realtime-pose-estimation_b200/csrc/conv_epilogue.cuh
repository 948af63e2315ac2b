// Shared epilogue of the tcgen05 convolution kernels: TMEM accumulator -> + folded-BN bias ->
// + residual -> ReLU -> bf16 -> global (NHWC).  One thread owns one output pixel (one TMEM
// lane) and walks the Cout tile in chunks of 16 columns = 32 bytes = one DRAM sector.
//
// Two paths:
//  * epi_fast<RES, RELU>: every chunk is 16 real channels, 32-byte aligned in out and res
//    (all BasicBlock / Bottleneck / transition / fuse convs of the W48 network).  Two chunks
//    per tcgen05.ld (x32), 256-bit LDG/STG, packed f32x2 adds, ReLU folded into the bf16x2
//    conversion; the residual of the NEXT chunk pair is in flight while the current one is
//    processed, and the first pair is fetched before the thread waits for the accumulator.
//    ~50 instructions per 16 channels.
//  * epi_drain: general path (ragged channel counts, unaligned tensors: the 34/17-channel
//    heads and the concat writer).
#pragma once
#include "conv_common.cuh"
#include "umma_ptx.cuh"

namespace brtpe {

constexpr int EPI_PRE = 6;        // chunks (x16 channels) of residual prefetched per tile (general path)
constexpr int EPI_MAX_CHUNKS = 16;

struct EpiParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  int out_ld, out_coff, res_ld, res_coff, Cout, Cout_store, relu;
  int vec32;                      // 1: every 16-channel chunk is 32-byte aligned in out and res
  int fast;                       // 1: vec32 and Cout == Cout_store is a multiple of 16
  int split;                      // 1: BRTPE_DT_BF16X2 activations (hi / lo bf16 pairs)
  int out_lo, res_lo;             // split: element offset of the lo half inside a pixel (ld / 2)
  // HRNet fuse addends (brtpe_conv_desc.n_add): term k read at (y >> shift, x >> shift)
  const __nv_bfloat16* add[3];
  int add_ld[3], add_shift[3], n_add;
  int Hout, Wout;                 // output tensor geometry (addend indexing)
  __nv_bfloat16* out2;            // second output (fused y_i) or nullptr
  int out2_ld;
};

// q = x / d for 0 <= x < 2^32 / d  (mul = floor(2^32 / d) + 1; mul == 0 selects plain division)
struct FastDiv {
  uint32_t mul, d;
};
static inline FastDiv make_fastdiv(uint32_t d, uint64_t max_x) {
  FastDiv f;
  f.d = d;
  f.mul = (d > 1 && max_x * (uint64_t)d < (1ull << 32)) ? (uint32_t)((1ull << 32) / d) + 1u : 0u;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) {
  return f.mul ? __umulhi(x, f.mul) : (f.d == 1 ? x : x / f.d);
}

struct Chunk32 {
  uint32_t w[8];
};

__device__ __forceinline__ Chunk32 ld_chunk32(const __nv_bfloat16* p, bool vec32) {
  Chunk32 c;
  if (vec32) {
    // (not .nc: under programmatic dependent launch the predecessor may still be writing other
    // buffers while this kernel is resident -- .nc promises read-only data for the kernel's lifetime)
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]), "=r"(c.w[4]),
                   "=r"(c.w[5]), "=r"(c.w[6]), "=r"(c.w[7])
                 : "l"(p));
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint4 b = *(reinterpret_cast<const uint4*>(p) + 1);
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
// (lo, hi) -> bf16x2 with round-to-nearest-even, optionally max(.,0) first (one F2FP)
template <bool RELU>
__device__ __forceinline__ uint32_t pack_bf16x2_act(float lo, float hi) {
  uint32_t r;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// (x0, x1) += (y0, y1) as one packed fp32 add (sm_100 FADD2): same rounding as two FADDs
__device__ __forceinline__ void add2(float& x0, float& x1, float y0, float y1) {
  asm("{\n\t.reg .b64 a, b;\n\t"
      "mov.b64 a, {%0, %1};\n\t"
      "mov.b64 b, {%2, %3};\n\t"
      "add.rn.f32x2 a, a, b;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "+f"(x0), "+f"(x1)
      : "f"(y0), "f"(y1));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// sum of the fuse addends of one 16-channel chunk at output pixel (n, y, x), channel co
template <int K0 = 0>
__device__ __forceinline__ void epi_terms16(const EpiParams& e, int n, int y, int x, int co,
                                            float (&s)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = 0.0f;
#pragma unroll
  for (int k = K0; k < 3; ++k) {
    if (k < e.n_add) {
      const int sh = e.add_shift[k];
      const size_t pix = ((size_t)n * (e.Hout >> sh) + (size_t)(y >> sh)) * (size_t)(e.Wout >> sh) +
                         (size_t)(x >> sh);
      const Chunk32 c = ld_chunk32(e.add[k] + pix * e.add_ld[k] + co, true);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[2 * i] += bf16lo(c.w[i]);
        s[2 * i + 1] += bf16hi(c.w[i]);
      }
    }
  }
}
// address of addend k's chunk for output pixel (n, y, x)
__device__ __forceinline__ const __nv_bfloat16* epi_term_ptr(const EpiParams& e, int k, int n, int y,
                                                             int x, int co) {
  const int sh = e.add_shift[k];
  const size_t pix = ((size_t)n * (e.Hout >> sh) + (size_t)(y >> sh)) * (size_t)(e.Wout >> sh) +
                     (size_t)(x >> sh);
  return e.add[k] + pix * e.add_ld[k] + co;
}
// s += the 16 bf16 values of a prefetched addend chunk
__device__ __forceinline__ void epi_acc_chunk(float (&s)[16], const Chunk32& c) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[2 * i] += bf16lo(c.w[i]);
    s[2 * i + 1] += bf16hi(c.w[i]);
  }
}
// single-output form: the addends join the accumulator (out = act(acc + sum + bias + res))
template <int OFF>
__device__ __forceinline__ void epi_add_terms(uint32_t (&a)[32], const float (&s)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) a[OFF + i] = __float_as_uint(__uint_as_float(a[OFF + i]) + s[i]);
}
// two-output form: op gets x = act(acc + bias + res) (the branch output), op2 gets relu(x + sum)
template <bool RES, bool RELU, int OFF>
__device__ __forceinline__ void epi_dual_chunk(const uint32_t (&a)[32], uint32_t bias16_smem,
                                               const Chunk32& rc, const float (&s)[16],
                                               __nv_bfloat16* op, __nv_bfloat16* op2) {
  Chunk32 oc, oc2;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 b;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "r"(bias16_smem + 16u * q));
    float v[4] = {__uint_as_float(a[OFF + 4 * q]) + b.x, __uint_as_float(a[OFF + 4 * q + 1]) + b.y,
                  __uint_as_float(a[OFF + 4 * q + 2]) + b.z, __uint_as_float(a[OFF + 4 * q + 3]) + b.w};
    if (RES) {
      v[0] += bf16lo(rc.w[2 * q]); v[1] += bf16hi(rc.w[2 * q]);
      v[2] += bf16lo(rc.w[2 * q + 1]); v[3] += bf16hi(rc.w[2 * q + 1]);
    }
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    oc.w[2 * q] = pack_bf16x2(v[0], v[1]);
    oc.w[2 * q + 1] = pack_bf16x2(v[2], v[3]);
    oc2.w[2 * q] = pack_bf16x2_act<true>(v[0] + s[4 * q], v[1] + s[4 * q + 1]);
    oc2.w[2 * q + 1] = pack_bf16x2_act<true>(v[2] + s[4 * q + 2], v[3] + s[4 * q + 3]);
  }
  st_chunk32(op, oc, true);
  st_chunk32(op2, oc2, true);
}

// one 16-channel chunk: acc (16 raw fp32 words at a[OFF..OFF+15]) + bias + residual -> out
template <bool RES, bool RELU, int OFF>
__device__ __forceinline__ void epi_fast_chunk(const uint32_t (&a)[32], uint32_t bias16_smem,
                                               const Chunk32& rc, __nv_bfloat16* op) {
  Chunk32 oc;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 b;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"      // bias_s is written once, before the
        : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)   // first CTA-wide sync: safe to hoist
        : "r"(bias16_smem + 16u * q));
    float v0 = __uint_as_float(a[OFF + 4 * q]), v1 = __uint_as_float(a[OFF + 4 * q + 1]);
    float v2 = __uint_as_float(a[OFF + 4 * q + 2]), v3 = __uint_as_float(a[OFF + 4 * q + 3]);
    add2(v0, v1, b.x, b.y);
    add2(v2, v3, b.z, b.w);
    if (RES) {
      add2(v0, v1, bf16lo(rc.w[2 * q]), bf16hi(rc.w[2 * q]));
      add2(v2, v3, bf16lo(rc.w[2 * q + 1]), bf16hi(rc.w[2 * q + 1]));
    }
    oc.w[2 * q] = pack_bf16x2_act<RELU>(v0, v1);
    oc.w[2 * q + 1] = pack_bf16x2_act<RELU>(v2, v3);
  }
  st_chunk32(op, oc, true);
}

// same arithmetic, result kept in registers (staged through shared memory + TMA store by the caller)
template <bool RES, bool RELU, int OFF>
__device__ __forceinline__ void epi_pack_chunk(const uint32_t (&a)[32], uint32_t bias16_smem,
                                               const Chunk32& rc, Chunk32& oc) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 b;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"      // bias_s is written once, before the
        : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)   // first CTA-wide sync: safe to hoist
        : "r"(bias16_smem + 16u * q));
    float v0 = __uint_as_float(a[OFF + 4 * q]), v1 = __uint_as_float(a[OFF + 4 * q + 1]);
    float v2 = __uint_as_float(a[OFF + 4 * q + 2]), v3 = __uint_as_float(a[OFF + 4 * q + 3]);
    add2(v0, v1, b.x, b.y);
    add2(v2, v3, b.z, b.w);
    if (RES) {
      add2(v0, v1, bf16lo(rc.w[2 * q]), bf16hi(rc.w[2 * q]));
      add2(v2, v3, bf16lo(rc.w[2 * q + 1]), bf16hi(rc.w[2 * q + 1]));
    }
    oc.w[2 * q] = pack_bf16x2_act<RELU>(v0, v1);
    oc.w[2 * q + 1] = pack_bf16x2_act<RELU>(v2, v3);
  }
}

// Waits for the accumulator (tfull_bar / parity), drains it and arrives on `arrive_bar`
// (tmem_empty, one lane per warp) as soon as the last TMEM read of the tile has completed.
template <bool RES, bool RELU, bool ADD = false>
__device__ __forceinline__ void epi_fast(const EpiParams& e, const float* __restrict__ bias_s,
                                         uint32_t t_addr, int nchunks, int co0, bool valid,
                                         size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                         uint32_t arrive_bar, int lane, int pn = 0, int py = 0,
                                         int px = 0) {
  const __nv_bfloat16* rp = RES ? e.res + opix * e.res_ld + e.res_coff + co0 : nullptr;
  __nv_bfloat16* op = e.out + opix * e.out_ld + e.out_coff + co0;
  const uint32_t bias_u32 = smem_u32(bias_s);
  Chunk32 r0, r1;
#pragma unroll
  for (int i = 0; i < 8; ++i) r0.w[i] = r1.w[i] = 0u;
  if (RES && valid) {
    r0 = ld_chunk32(rp, true);
    if (nchunks > 1) r1 = ld_chunk32(rp + 16, true);
  }
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2) {
    const bool two = c + 1 < nchunks;
    uint32_t a[32];
    if (two) tmem_ld32(t_addr + (uint32_t)(c * 16), a);
    else tmem_ld16_lo(t_addr + (uint32_t)(c * 16), a);
    Chunk32 n0 = r0, n1 = r1;
    if (RES && valid) {
      if (c + 2 < nchunks) n0 = ld_chunk32(rp + (c + 2) * 16, true);
      if (c + 3 < nchunks) n1 = ld_chunk32(rp + (c + 3) * 16, true);
    }
    tmem_ld_wait();
    if (c + 2 >= nchunks) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(arrive_bar);
    }
    if (valid) {
      const uint32_t bs = bias_u32 + (uint32_t)((co0 + c * 16) * 4);
      if (ADD) {                                   // HRNet fuse addends (nearest-upsampled terms)
        float s[16];
        epi_terms16(e, pn, py, px, co0 + c * 16, s);
        epi_add_terms<0>(a, s);
        if (two) {
          epi_terms16(e, pn, py, px, co0 + c * 16 + 16, s);
          epi_add_terms<16>(a, s);
        }
      }
      epi_fast_chunk<RES, RELU, 0>(a, bs, r0, op + c * 16);
      if (two) epi_fast_chunk<RES, RELU, 16>(a, bs + 64u, r1, op + c * 16 + 16);
    }
    r0 = n0;
    r1 = n1;
  }
}

// ---- split (BRTPE_DT_BF16X2) epilogue: the float32 result v is stored as hi = bf16(v) and
// lo = bf16(v - hi); the residual is read back as hi + lo (exact in float32).
template <bool RES, bool RELU, int OFF>
__device__ __forceinline__ void epi_split_chunk(const uint32_t (&a)[32], uint32_t bias16_smem,
                                                const Chunk32& rh, const Chunk32& rl,
                                                __nv_bfloat16* op_hi, __nv_bfloat16* op_lo) {
  Chunk32 oh, ol;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 b;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "r"(bias16_smem + 16u * q));
    float v[4] = {__uint_as_float(a[OFF + 4 * q]) + b.x, __uint_as_float(a[OFF + 4 * q + 1]) + b.y,
                  __uint_as_float(a[OFF + 4 * q + 2]) + b.z, __uint_as_float(a[OFF + 4 * q + 3]) + b.w};
    if (RES) {
      v[0] += bf16lo(rh.w[2 * q]) + bf16lo(rl.w[2 * q]);
      v[1] += bf16hi(rh.w[2 * q]) + bf16hi(rl.w[2 * q]);
      v[2] += bf16lo(rh.w[2 * q + 1]) + bf16lo(rl.w[2 * q + 1]);
      v[3] += bf16hi(rh.w[2 * q + 1]) + bf16hi(rl.w[2 * q + 1]);
    }
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]);
    oh.w[2 * q] = h0;
    oh.w[2 * q + 1] = h1;
    ol.w[2 * q] = pack_bf16x2(v[0] - bf16lo(h0), v[1] - bf16hi(h0));
    ol.w[2 * q + 1] = pack_bf16x2(v[2] - bf16lo(h1), v[3] - bf16hi(h1));
  }
  st_chunk32(op_hi, oh, true);
  st_chunk32(op_lo, ol, true);
}

template <bool RES, bool RELU>
__device__ __forceinline__ void epi_split(const EpiParams& e, const float* __restrict__ bias_s,
                                          uint32_t t_addr, int nchunks, int co0, bool valid,
                                          size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                          uint32_t arrive_bar, int lane) {
  const __nv_bfloat16* rp = RES ? e.res + opix * e.res_ld + e.res_coff + co0 : nullptr;
  __nv_bfloat16* op = e.out + opix * e.out_ld + e.out_coff + co0;
  const uint32_t bias_u32 = smem_u32(bias_s);
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2) {
    const bool two = c + 1 < nchunks;
    Chunk32 rh0, rl0, rh1, rl1;
#pragma unroll
    for (int i = 0; i < 8; ++i) rh0.w[i] = rl0.w[i] = rh1.w[i] = rl1.w[i] = 0u;
    if (RES && valid) {
      rh0 = ld_chunk32(rp + c * 16, true);
      rl0 = ld_chunk32(rp + e.res_lo + c * 16, true);
      if (two) {
        rh1 = ld_chunk32(rp + c * 16 + 16, true);
        rl1 = ld_chunk32(rp + e.res_lo + c * 16 + 16, true);
      }
    }
    uint32_t a[32];
    if (two) tmem_ld32(t_addr + (uint32_t)(c * 16), a);
    else tmem_ld16_lo(t_addr + (uint32_t)(c * 16), a);
    tmem_ld_wait();
    if (c + 2 >= nchunks) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(arrive_bar);
    }
    if (valid) {
      const uint32_t bs = bias_u32 + (uint32_t)((co0 + c * 16) * 4);
      epi_split_chunk<RES, RELU, 0>(a, bs, rh0, rl0, op + c * 16, op + e.out_lo + c * 16);
      if (two)
        epi_split_chunk<RES, RELU, 16>(a, bs + 64u, rh1, rl1, op + c * 16 + 16,
                                       op + e.out_lo + c * 16 + 16);
    }
  }
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

// General path.  Drains one accumulator tile.  `arrive_bar`: tmem_empty barrier, arrived on
// (one lane per warp) as soon as the last TMEM read of this tile has completed.
static __device__ __noinline__ void epi_drain(const EpiParams& e, const float* __restrict__ bias_s,
                                       uint32_t t_addr, int nchunks, int co0, bool valid,
                                       size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                       uint32_t arrive_bar, int lane) {
  const bool vec32 = e.vec32 != 0;
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const int co = co0 + c * 16;
    const bool live = valid && co < e.Cout_store;
    const bool whole = co + 16 <= e.Cout;            // chunk entirely made of real channels
    Chunk32 rc;
#pragma unroll
    for (int i = 0; i < 8; ++i) rc.w[i] = 0u;
    if (e.res != nullptr && live && whole)
      rc = ld_chunk32(e.res + opix * e.res_ld + e.res_coff + co, vec32);
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

// Dispatch (warp-uniform): waits for the accumulator and drains it.
__device__ __forceinline__ void epi_tile(const EpiParams& e, const float* __restrict__ bias_s,
                                         uint32_t t_addr, int nchunks, int co0, bool valid,
                                         size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                         uint32_t arrive_bar, int lane) {
  if (e.fast) {
    if (e.res != nullptr) {
      if (e.relu) epi_fast<true, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
      else epi_fast<true, false>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
    } else {
      if (e.relu) epi_fast<false, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
      else epi_fast<false, false>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
    }
  } else {
    epi_drain(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
  }
}

// Dispatch of the kernels with HRNet fuse addends (fast-path layers only, checked at prepare)
__device__ __forceinline__ void epi_tile_add(const EpiParams& e, const float* __restrict__ bias_s,
                                             uint32_t t_addr, int nchunks, int co0, bool valid,
                                             size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                             uint32_t arrive_bar, int lane, int pn, int py, int px) {
  if (e.res != nullptr) {
    if (e.relu) epi_fast<true, true, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane, pn, py, px);
    else epi_fast<true, false, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane, pn, py, px);
  } else {
    if (e.relu) epi_fast<false, true, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane, pn, py, px);
    else epi_fast<false, false, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane, pn, py, px);
  }
}

// Dispatch of the split (BRTPE_DT_BF16X2) kernels: split layers always satisfy the fast-path
// conditions (checked at prepare).
__device__ __forceinline__ void epi_tile_split(const EpiParams& e, const float* __restrict__ bias_s,
                                               uint32_t t_addr, int nchunks, int co0, bool valid,
                                               size_t opix, uint32_t tfull_bar, uint32_t tfull_parity,
                                               uint32_t arrive_bar, int lane) {
  if (e.res != nullptr) {
    if (e.relu) epi_split<true, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
    else epi_split<true, false>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
  } else {
    if (e.relu) epi_split<false, true>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
    else epi_split<false, false>(e, bias_s, t_addr, nchunks, co0, valid, opix, tfull_bar, tfull_parity, arrive_bar, lane);
  }
}

static inline int epi_vec32_ok(const brtpe_conv_desc* d) {
  const bool out_ok = (d->out_ld % 16 == 0) && (d->out_coff % 16 == 0);
  const bool res_ok = (d->res_ld % 16 == 0) && (d->res_coff % 16 == 0);
  return (out_ok && res_ok) ? 1 : 0;
}
static inline int epi_fast_ok(const brtpe_conv_desc* d) {
  return (epi_vec32_ok(d) && d->Cout % 16 == 0 && d->Cout_store == d->Cout) ? 1 : 0;
}
// split (BRTPE_DT_BF16X2) layers: the lo halves (offset ld / 2) must be 32-byte aligned as well
static inline bool conv_is_split(const brtpe_conv_desc* d) { return d->dtype == BRTPE_DT_BF16X2; }
static inline int epi_split_ok(const brtpe_conv_desc* d) {
  return (epi_fast_ok(d) && d->out_ld % 32 == 0 && d->res_ld % 32 == 0 && d->in_ld % 32 == 0 &&
          d->in_coff % 8 == 0) ? 1 : 0;
}
// fuse addends need the fast epilogue, 32-byte aligned chunks in every addend and shifts that
// divide the output size
static inline bool epi_add_ok(const brtpe_conv_desc* d) {
  if (d->n_add == 0 && d->out2_ld == 0) return true;
  if (d->n_add < 0 || d->n_add > 3 || !epi_fast_ok(d) || d->dtype != BRTPE_DT_BF16) return false;
  if (d->out2_ld && (d->out2_ld % 16 || d->out2_ld < d->Cout)) return false;
  for (int k = 0; k < d->n_add; ++k) {
    if (d->add_ld[k] % 16 || d->add_ld[k] < d->Cout || d->add_shift[k] < 0 || d->add_shift[k] > 4) return false;
    if ((d->Hout % (1 << d->add_shift[k])) || (d->Wout % (1 << d->add_shift[k]))) return false;
  }
  return true;
}
static inline void epi_set_add(EpiParams* e, const brtpe_conv_desc* d) {
  e->n_add = d->n_add;
  for (int k = 0; k < 3; ++k) {
    e->add[k] = nullptr;
    e->add_ld[k] = k < d->n_add ? d->add_ld[k] : 0;
    e->add_shift[k] = k < d->n_add ? d->add_shift[k] : 0;
  }
  e->Hout = d->Hout; e->Wout = d->Wout;
  e->out2 = nullptr; e->out2_ld = d->out2_ld;
}
static inline void epi_set_split(EpiParams* e, const brtpe_conv_desc* d) {
  e->split = conv_is_split(d) ? 1 : 0;
  e->out_lo = d->out_ld / 2;
  e->res_lo = d->res_ld / 2;
}

}  // namespace brtpe
