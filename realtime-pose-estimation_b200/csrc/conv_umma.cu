// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate).
//
// Replaces the cuDNN conv + BatchNorm + ReLU + residual-add launches of
// PoseHigherResolutionNet.forward (rtpe/third_party/pose_higher_hrnet.py:40-43 conv3x3,
// :46-75 BasicBlock, :78-116 Bottleneck, :202-229 fuse convs, :460-482 final 1x1 convs,
// :514-521 the 4x4/s2 transposed conv as four parity phases, :558-580 transitions).
//
// GEMM view per launch: D[M = output pixels, N = Cout] = sum over taps of
// A_tap[M, Cin] * W_tap[Cin, Cout]; activations are NHWC bf16, so for a fixed tap the A
// operand of a (TN x TH x TW)-pixel tile is a dense 5-D TMA box of the input tensor shifted
// by the tap offset -- out-of-image pixels are zero-filled by TMA, which IS the conv
// padding.  Stride-2 convs address the input through a (channel-pair, W/2, row parity, H/2,
// N) view so that each tap is again a dense box.  Every K step (one tap, one 64-channel
// block) is one TMA load of A (128 rows x 128 B, SWIZZLE_128B) + one of W and up to four
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Cout tile, K=16) issued by a single thread,
// accumulating in TMEM.
//
// Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
// issuer + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld of their TMEM lane quarter ->
// + folded-BN bias -> + residual -> ReLU -> bf16 -> 16-byte global stores).  smem ring of
// `stages` A/B slots (full/empty mbarriers) and two TMEM accumulator stages (tmem_full /
// tmem_empty), so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>
#include <algorithm>

namespace brtpe {

constexpr int UM_THREADS = 192;
constexpr int UM_MAX_STAGES = 8;
constexpr int UM_A_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int UM_MAX_COUT_PAD = 512;

struct alignas(64) UmmaParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  int N, Hm, Wm;
  int TW, TH, TN, tiles_x, tiles_y, tiles_n;
  int n_tiles, BN, total_tiles;
  int ntaps, num_kb, last_k16;
  int tap_c[9], tap_x[9], tap_p[9], tap_y[9];
  int stages, b_stage_bytes, tmem_cols, acc_cols;
  uint32_t idesc;
  void* out;
  const void* res;
  const float* bias;
  int Hout, Wout, out_scale, out_oy, out_ox, out_ld, out_coff, res_ld, res_coff;
  int Cout, Cout_store, relu;
};

struct UmmaConvPrepared {
  UmmaParams p;
  int grid;
  size_t smem;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 22)) __trap();  // a protocol bug becomes an error, not a hung GPU
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) |
//   [32,46) SBO >> 4 (8 rows x 128 B = 1024) | [46,48) version = 1 | [61,64) layout = 2.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UM_THREADS)
conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stage_bytes = UM_A_BYTES + p.b_stage_bytes;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + UM_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + UM_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // UM_MAX_COUT_PAD floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.BN; i += UM_THREADS)
    bias_s[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.0f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_xy = p.tiles_x * p.tiles_y;
  const uint32_t tx_bytes = (uint32_t)(p.TW * p.TH * p.TN * 128 + p.BN * 128);

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int x0 = (mt % p.tiles_x) * p.TW;
        const int y0 = ((mt % tiles_xy) / p.tiles_x) * p.TH;
        const int n0 = (mt / tiles_xy) * p.TN;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, tx_bytes);
            const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
            tma_load_5d(a_dst, &p.tmap_a, fb, p.tap_c[tap] + kb * 64, x0 + p.tap_x[tap],
                        p.tap_p[tap], y0 + p.tap_y[tap], n0);
            tma_load_3d(a_dst + UM_A_BYTES, &p.tmap_b, fb, kb * 64, nt * p.BN, tap);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_cols);
        uint32_t accumulate = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
            const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + UM_A_BYTES);
            const int k16 = (kb == p.num_kb - 1) ? p.last_k16 : 4;
            for (int k = 0; k < k16; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in >>4 units
              umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                       accumulate);
              accumulate = 1;
            }
            umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot when MMAs retire
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit(smem_u32(&tfull_bar[as]));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const int tw = m % p.TW;
    const int th = (m / p.TW) % p.TH;
    const int tn = m / (p.TW * p.TH);
    __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(p.out);
    const __nv_bfloat16* __restrict__ res = reinterpret_cast<const __nv_bfloat16*>(p.res);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int xm = (mt % p.tiles_x) * p.TW + tw;
      const int ym = ((mt % tiles_xy) / p.tiles_x) * p.TH + th;
      const int n = (mt / tiles_xy) * p.TN + tn;
      const bool valid = (tn < p.TN) && xm < p.Wm && ym < p.Hm && n < p.N;
      const size_t opix =
          valid ? ((size_t)n * p.Hout + (size_t)(ym * p.out_scale + p.out_oy)) * p.Wout +
                      (size_t)(xm * p.out_scale + p.out_ox)
                : 0;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&tfull_bar[as]), aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * p.acc_cols);
      const int co0 = nt * p.BN;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_addr + (uint32_t)c, r);
        tmem_ld_wait();
        if (c + 16 >= p.BN) {
          // all TMEM reads of this tile are done: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
        }
        const int co = co0 + c;
        if (!valid || co >= p.Cout_store) continue;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bias_s[co + i];
        const bool full_chunk = (co + 16 <= p.Cout_store);
        if (res) {
          const __nv_bfloat16* rp = res + opix * p.res_ld + p.res_coff + co;
          if (full_chunk) {
            uint4 q[2];
            q[0] = __ldg(reinterpret_cast<const uint4*>(rp));
            q[1] = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
            const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(q);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __bfloat162float(rb[i]);
          } else {
            for (int i = 0; i < 16 && co + i < p.Cout; ++i) v[i] += __bfloat162float(rp[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (p.relu) v[i] = fmaxf(v[i], 0.0f);
          if (co + i >= p.Cout) v[i] = 0.0f;
        }
        __nv_bfloat16* op = out + opix * p.out_ld + p.out_coff + co;
        if (full_chunk) {
          uint4 q[2];
          __nv_bfloat162* qb = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
          for (int i = 0; i < 8; ++i) qb[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          reinterpret_cast<uint4*>(op)[0] = q[0];
          reinterpret_cast<uint4*>(op)[1] = q[1];
        } else {
          for (int i = 0; i < 16 && co + i < p.Cout_store; ++i) op[i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

static void umma_n_tiling(int cout_store, int* n_tiles, int* bn) {
  const int cp = (cout_store + 15) / 16 * 16;
  *n_tiles = (cp + 255) / 256;
  *bn = ((cp + *n_tiles - 1) / *n_tiles + 15) / 16 * 16;
}

bool umma_conv_supported(const brtpe_conv_desc* d, const char** why) {
  static const char* msg = "";
  auto fail = [&](const char* m) {
    msg = m;
    if (why) *why = msg;
    return false;
  };
  if (d->dtype != BRTPE_DT_BF16) return fail("tcgen05 path needs bf16 activations");
  if (d->Cin % 16) return fail("Cin must be a multiple of 16");
  if (d->in_ld % 8 || d->in_coff % 8) return fail("in_ld / in_coff must be multiples of 8");
  if (d->out_ld % 8 || d->out_coff % 8) return fail("out_ld / out_coff must be multiples of 8");
  if (d->res_ld % 8 || d->res_coff % 8) return fail("res_ld / res_coff must be multiples of 8");
  if (d->in_stride != 1 && d->in_stride != 2) return fail("in_stride must be 1 or 2");
  if (d->in_stride == 2 && ((d->Hin | d->Win) & 1)) return fail("stride 2 needs even Hin/Win");
  for (int t = 0; t < d->ntaps; ++t)
    if (d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1)
      return fail("tap offsets must be in [-1, 1]");
  int nt, bn;
  umma_n_tiling(d->Cout_store, &nt, &bn);
  if (nt * bn > UM_MAX_COUT_PAD) return fail("Cout too large");
  if (!get_encode_fn()) return fail("cuTensorMapEncodeTiled not available");
  return true;
}

static bool g_attr_set = false;

UmmaConvPrepared* umma_conv_prepare(const brtpe_conv_desc* d, const void* in, const void* weights) {
  const char* why = nullptr;
  if (!umma_conv_supported(d, &why)) {
    set_error("tcgen05 conv: unsupported layer: %s", why);
    return nullptr;
  }
  UmmaConvPrepared* P = new UmmaConvPrepared();
  UmmaParams& p = P->p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.Hm = d->Hm; p.Wm = d->Wm;

  // ---- M tile geometry: fewest tiles of <= 128 pixels, then widest rows, then TN = 1
  long best_tiles = -1;
  int bw = 1, bh = 1, bnn = 1;
  for (int tw = 1; tw <= d->Wm && tw <= 128; ++tw)
    for (int th = 1; th <= d->Hm && tw * th <= 128; ++th)
      for (int tn = 1; tn <= d->N && tw * th * tn <= 128; ++tn) {
        const long tiles = (long)ceil_div(d->Wm, tw) * ceil_div(d->Hm, th) * ceil_div(d->N, tn);
        bool better = best_tiles < 0 || tiles < best_tiles;
        if (!better && tiles == best_tiles) {
          if (tn < bnn) better = true;
          else if (tn == bnn && tw > bw) better = true;
        }
        if (better) { best_tiles = tiles; bw = tw; bh = th; bnn = tn; }
      }
  p.TW = bw; p.TH = bh; p.TN = bnn;
  p.tiles_x = ceil_div(d->Wm, bw); p.tiles_y = ceil_div(d->Hm, bh); p.tiles_n = ceil_div(d->N, bnn);
  umma_n_tiling(d->Cout_store, &p.n_tiles, &p.BN);
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;

  p.ntaps = d->ntaps;
  p.num_kb = ceil_div(d->Cin, 64);
  p.last_k16 = (d->Cin - (p.num_kb - 1) * 64) / 16;
  const int s = d->in_stride;
  for (int t = 0; t < d->ntaps; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (s == 1) {
      p.tap_c[t] = d->in_coff; p.tap_x[t] = dx; p.tap_p[t] = 0; p.tap_y[t] = dy;
    } else {
      const int px = (dx != 0), py = (dy != 0);
      p.tap_c[t] = px * d->in_ld + d->in_coff;
      p.tap_x[t] = (dx < 0) ? -1 : 0;
      p.tap_p[t] = py;
      p.tap_y[t] = (dy < 0) ? -1 : 0;
    }
  }
  p.b_stage_bytes = (int)align_up((size_t)p.BN * 128, 1024);
  const int stage_bytes = UM_A_BYTES + p.b_stage_bytes;
  p.acc_cols = p.BN;
  int cols = 32;
  while (cols < 2 * p.acc_cols) cols *= 2;
  p.tmem_cols = cols;
  const int tail = 4096;  // barriers + bias + alignment slack
  const bool two_per_sm = (cols <= 256);
  const int budget = two_per_sm ? 110 * 1024 : 220 * 1024;
  int stages = (budget - tail - 1024) / stage_bytes;
  if (stages > UM_MAX_STAGES) stages = UM_MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  P->smem = (size_t)stages * stage_bytes + tail + 1024;
  P->grid = (int)std::min<long>(p.total_tiles, (long)num_sms() * (two_per_sm ? 2 : 1));

  // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);

  p.out = nullptr; p.res = nullptr; p.bias = nullptr;
  p.Hout = d->Hout; p.Wout = d->Wout; p.out_scale = d->out_scale; p.out_oy = d->out_oy;
  p.out_ox = d->out_ox; p.out_ld = d->out_ld; p.out_coff = d->out_coff; p.res_ld = d->res_ld;
  p.res_coff = d->res_coff; p.Cout = d->Cout; p.Cout_store = d->Cout_store; p.relu = d->relu;

  // ---- tensor maps
  auto encode = get_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->in_ld * 2;
  cuuint64_t gdim[5], gstr[4];
  if (s == 1) {
    gdim[0] = (cuuint64_t)(d->in_coff + d->Cin); gdim[1] = d->Win; gdim[2] = 1; gdim[3] = d->Hin;
    gdim[4] = d->N;
    gstr[0] = ld_b; gstr[1] = ld_b * d->Win; gstr[2] = ld_b * d->Win;
    gstr[3] = ld_b * d->Win * d->Hin;
  } else {
    gdim[0] = (cuuint64_t)(d->in_ld + d->in_coff + d->Cin); gdim[1] = d->Win / 2; gdim[2] = 2;
    gdim[3] = d->Hin / 2; gdim[4] = d->N;
    gstr[0] = 2 * ld_b; gstr[1] = ld_b * d->Win; gstr[2] = 2 * ld_b * d->Win;
    gstr[3] = ld_b * d->Win * d->Hin;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&p.tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), gdim,
                      gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(A) failed with %d (dims %llu %llu %llu %llu %llu)", (int)r,
              (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
              (unsigned long long)gdim[3], (unsigned long long)gdim[4]);
    delete P;
    return nullptr;
  }
  const int cin_pad = p.num_kb * 64;
  const int cout_pad = p.n_tiles * p.BN;
  cuuint64_t wdim[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)d->ntaps};
  cuuint64_t wstr[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * 2 * cout_pad};
  cuuint32_t wbox[3] = {64, (cuuint32_t)p.BN, 1};
  cuuint32_t westr[3] = {1, 1, 1};
  r = encode(&p.tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(weights), wdim, wstr,
             wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    delete P;
    return nullptr;
  }
  if (!g_attr_set) {
    if (cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_umma_kernel) failed: %s",
                cudaGetErrorString(cudaGetLastError()));
      delete P;
      return nullptr;
    }
    g_attr_set = true;
  }
  return P;
}

void umma_conv_release(UmmaConvPrepared* p) { delete p; }

int umma_conv_launch(const UmmaConvPrepared* P, const float* bias, const void* residual, void* out,
                     cudaStream_t st) {
  UmmaParams p = P->p;
  p.bias = bias;
  p.res = residual;
  p.out = out;
  conv_umma_kernel<<<P->grid, UM_THREADS, P->smem, st>>>(p);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

}  // namespace brtpe

extern "C" int brtpe_has_umma(void) { return 1; }

// Packed-weight geometry the tcgen05 path expects: [ntaps][cout_pad][cin_pad] bf16.
extern "C" int brtpe_umma_weight_dims(int Cin, int Cout_store, int* cin_pad, int* cout_pad) {
  int nt, bn;
  brtpe::umma_n_tiling(Cout_store, &nt, &bn);
  if (cin_pad) *cin_pad = (Cin + 63) / 64 * 64;
  if (cout_pad) *cout_pad = nt * bn;
  return BRTPE_OK;
}
