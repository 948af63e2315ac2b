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
#include "umma_ptx.cuh"
#include "conv_epilogue.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace brtpe {

constexpr int UM_THREADS = 320;           // TMA warp, MMA warp, 2 x 4 epilogue warps (UmmaParams::epi_groups == 2)
// Four epilogue warps (192 threads, two CTAs per SM by registers) unless the layer is a 1x1 conv with a Cout
// tile of >= 64 channels on a large map: those are pure streaming layers (K <= 256: a handful of MMAs per
// tile, then 128 pixels x BN channels of residual + output), bound by the bytes their epilogue warps keep in
// flight; one CTA per SM with eight warps, each group draining half of the channel chunks, measured (64
// images): 64 -> 256 + residual @160^2 0.427 -> 0.404 ms, 256 -> 64 0.201 -> 0.193.  Everywhere else the
// second CTA per SM is worth more (48 -> 32 @320^2 0.245 -> 0.419 ms with 320 threads).
constexpr int UM_THREADS_NARROW = 192;
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
  int nkb_seg, lo_off;    // split (BRTPE_DT_BF16X2): num_kb = 3 segments [hi | lo | hi] of nkb_seg
                          // 64-channel blocks, the lo half starts lo_off channels into the pixel
  int tap_c[9], tap_x[9], tap_p[9], tap_y[9];
  int stages, b_stage_bytes, tmem_cols, acc_cols;
  uint32_t idesc;
  FastDiv fd_nt, fd_xy, fd_x, fd_tw, fd_twh;
  const float* bias;
  int Hout, Wout, out_scale, out_oy, out_ox;
  int res_l2_prefetch;    // epilogue prefetches the next tile's residual rows into L2
  int epi_groups;         // 1: four epilogue warps (192 threads), 2: eight (320 threads)
  EpiParams epi;
};

struct UmmaConvPrepared {
  UmmaParams p;
  int grid;
  size_t smem;
};

__device__ __forceinline__ uint32_t um_desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
}

template <int K16>
__device__ __forceinline__ void um_issue_k(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                           uint32_t idesc, uint32_t first) {
#pragma unroll
  for (int k = 0; k < K16; ++k)
    umma_f16_lohi(d_tmem, a_lo + 2u * k, hi, b_lo + 2u * k, hi, idesc, (first | (uint32_t)k) ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <bool SPLIT, bool ADD>
__global__ void __launch_bounds__(UM_THREADS)
conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stages = p.stages, ntaps = p.ntaps, num_kb = p.num_kb, last_k16 = p.last_k16;
  const int nkb_seg = p.nkb_seg, lo_off = p.lo_off;
  const int total_tiles = p.total_tiles, n_tiles = p.n_tiles, BN = p.BN;
  const int tiles_x = p.tiles_x, tiles_xy = p.tiles_x * p.tiles_y;
  const int TW = p.TW, TH = p.TH, TN = p.TN;
  const uint32_t idesc = p.idesc;
  const int stage_bytes = UM_A_BYTES + p.b_stage_bytes;
  uint8_t* tail = smem + (size_t)stages * stage_bytes;
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
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4u * (uint32_t)p.epi_groups);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_tiles * BN; i += blockDim.x)
    bias_s[i] = (p.bias && i < p.epi.Cout) ? p.bias[i] : 0.0f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tx_bytes = (uint32_t)(TW * TH * TN * 128 + BN * 128);
  if (threadIdx.x == 0) pdl_trigger();              // PDL (common.cuh): successor may start its prologue
  if (warp != 1) pdl_wait();                        // producer + epilogue touch the predecessor's output

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform, elected lane issues) ================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = (int)fdiv((uint32_t)tile, p.fd_nt);
      const int nt = tile - mt * n_tiles;
      const int tn_ = (int)fdiv((uint32_t)mt, p.fd_xy);
      const int rem = mt - tn_ * tiles_xy;
      const int ty_ = (int)fdiv((uint32_t)rem, p.fd_x);
      const int x0 = (rem - ty_ * tiles_x) * TW;
      const int y0 = ty_ * TH;
      const int n0 = tn_ * TN;
#pragma unroll 1
      for (int tap = 0; tap < ntaps; ++tap) {
        const int tc = p.tap_c[tap], tx = x0 + p.tap_x[tap], tp = p.tap_p[tap], ty = y0 + p.tap_y[tap];
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          // channel block kb of the reduction: segment (hi, lo, hi again) and block inside it
          const int seg = (kb >= nkb_seg) + (kb >= 2 * nkb_seg);
          const int coff = (kb - seg * nkb_seg) * 64 + (seg == 1 ? lo_off : 0);
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, tx_bytes);
            const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
            tma_load_5d(a_dst, &p.tmap_a, fb, tc + coff, tx, tp, ty, n0);
            tma_load_3d(a_dst + UM_A_BYTES, &p.tmap_b, fb, kb * 64, nt * BN, tap);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform, elected lane issues) ==================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t ring_lo = um_desc_lo(smem_u32(smem));
    const uint32_t stage_lo = (uint32_t)(stage_bytes >> 4);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_cols);
      uint32_t first = 0;
#pragma unroll 1
      for (int tap = 0; tap < ntaps; ++tap) {
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + (uint32_t)stage * stage_lo;
          const uint32_t b_lo = a_lo + (uint32_t)(UM_A_BYTES >> 4);
          const int seg = (kb >= nkb_seg) + (kb >= 2 * nkb_seg);
          const int k16 = (kb - seg * nkb_seg == nkb_seg - 1) ? last_k16 : 4;
          if (elect_one()) {
            // no run-time branch between the MMAs of a K block (tools/exp_umma_seq.cu: a partial
            // block issued through `if (k < k16)` costs ~10 cycles per MMA)
            switch (k16) {
              case 4: um_issue_k<4>(d_tmem, a_lo, b_lo, HI, idesc, first); break;
              case 3: um_issue_k<3>(d_tmem, a_lo, b_lo, HI, idesc, first); break;
              case 2: um_issue_k<2>(d_tmem, a_lo, b_lo, HI, idesc, first); break;
              default: um_issue_k<1>(d_tmem, a_lo, b_lo, HI, idesc, first); break;
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          first = 1;
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(smem_u32(&tfull_bar[as]));
      __syncwarp();
    }
  } else {
    // ===================== epilogue (1 or 2 groups of 4 warps, one TMEM lane quarter each) ===
    // two groups: group g drains the channel chunks [cbeg, cend) of every tile
    const int lg = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int m = lg * 32 + lane;
    const int tn = (int)fdiv((uint32_t)m, p.fd_twh);
    const int th = (int)fdiv((uint32_t)(m - tn * TW * TH), p.fd_tw);
    const int tw = m - tn * TW * TH - th * TW;
    const EpiParams e = p.epi;
    const int allchunks = BN >> 4;
    const bool two_groups = p.epi_groups == 2;
    const int csplit = two_groups ? (((allchunks + 3) >> 2) << 1) : allchunks;   // even: chunks go in pairs
    const int cbeg = grp == 0 ? 0 : (csplit < allchunks ? csplit : allchunks);
    const int cend = grp == 0 ? (csplit < allchunks ? csplit : allchunks) : allchunks;
    const int nchunks = cend - cbeg;
    const int Wm = p.Wm, Hm = p.Hm, N = p.N, Hout = p.Hout, Wout = p.Wout;
    const int out_scale = p.out_scale, out_oy = p.out_oy, out_ox = p.out_ox, acc_cols = p.acc_cols;
    int it = 0;
    // output pixel / channel offset of this thread in a tile (valid = inside the tensor)
    int pn = 0, py = 0, px = 0;                    // output pixel of this thread (fuse addends)
    auto locate = [&](int tile, size_t& opix, int& co0) -> bool {
      const int mt = (int)fdiv((uint32_t)tile, p.fd_nt);
      const int nt = tile - mt * n_tiles;
      const int tn_ = (int)fdiv((uint32_t)mt, p.fd_xy);
      const int rem = mt - tn_ * tiles_xy;
      const int ty_ = (int)fdiv((uint32_t)rem, p.fd_x);
      const int xm = (rem - ty_ * tiles_x) * TW + tw;
      const int ym = ty_ * TH + th;
      const int n = tn_ * TN + tn;
      const bool valid = (tn < TN) && xm < Wm && ym < Hm && n < N;
      opix = valid ? ((size_t)n * Hout + (size_t)(ym * out_scale + out_oy)) * Wout +
                         (size_t)(xm * out_scale + out_ox)
                   : 0;
      co0 = nt * BN;
      pn = n; py = ym * out_scale + out_oy; px = xm * out_scale + out_ox;
      return valid;
    };
    const bool res_pf = p.res_l2_prefetch != 0 && e.res != nullptr;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      size_t opix;
      int co0;
      if (res_pf && tile + (int)gridDim.x < total_tiles) {
        // the residual row of this thread's pixel in the NEXT tile -> L2 (no registers held): the
        // HBM-bound 1x1 layers wait on exactly these loads (ncu: 70 % long-scoreboard)
        size_t opix_n;
        int co0_n;
        if (locate(tile + (int)gridDim.x, opix_n, co0_n)) {
          const char* rp = reinterpret_cast<const char*>(e.res + opix_n * e.res_ld + e.res_coff + co0_n);
          for (int b = cbeg * 32; b < cend * 32; b += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + b));
        }
      }
      const bool valid = locate(tile, opix, co0);    // (after the prefetch: it leaves pn / py / px)
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * acc_cols) +
                              (uint32_t)(cbeg * 16);
      co0 += cbeg * 16;
      if (nchunks <= 0) {                            // (host never selects two groups for such tiles)
        mbar_wait(smem_u32(&tfull_bar[as]), aphase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
        continue;
      }
      if (SPLIT)
        epi_tile_split(e, bias_s, t_addr, nchunks, co0, valid, opix, smem_u32(&tfull_bar[as]), aphase,
                       smem_u32(&tempty_bar[as]), lane);
      else if (ADD)
        epi_tile_add(e, bias_s, t_addr, nchunks, co0, valid, opix, smem_u32(&tfull_bar[as]), aphase,
                     smem_u32(&tempty_bar[as]), lane, pn, py, px);
      else
        epi_tile(e, bias_s, t_addr, nchunks, co0, valid, opix, smem_u32(&tfull_bar[as]), aphase,
                 smem_u32(&tempty_bar[as]), lane);
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
  // Cout tiles of at most 128 channels: two accumulator stages then fit in 256 TMEM columns, so
  // two CTAs (8 epilogue warps) share an SM.  The wide 1x1 layers (64 -> 256 + residual) are
  // epilogue / HBM bound, not MMA bound: BRTPE_UMMA_BN_MAX=256 restores one 256-wide tile.
  static int bn_max = 0;
  if (!bn_max) {
    const char* e = getenv("BRTPE_UMMA_BN_MAX");
    bn_max = e ? atoi(e) : 128;
    if (bn_max < 16 || bn_max > 256) bn_max = 128;
  }
  *n_tiles = (cp + bn_max - 1) / bn_max;
  *bn = ((cp + *n_tiles - 1) / *n_tiles + 15) / 16 * 16;
}

// Per-layer Cout tiling.  Layers with a long reduction (taps x Cin >= 768: the 3x3 layers) are MMA
// bound, not epilogue bound: tiles of up to 192 channels (one CTA per SM, 384 TMEM columns) re-load
// the pixel operand fewer times -- 3x3 384 -> 384 on 20x20 maps: 2 x 192 instead of 3 x 128,
// 0.0645 -> 0.0536 ms per 64 images (profiles/r01q_conv_384.md).  Only tilings that keep the packed
// weight geometry of umma_n_tiling (brtpe_umma_weight_dims) are taken.  BRTPE_UMMA_WIDE=0 disables.
static void umma_n_tiling_layer(const brtpe_conv_desc* d, int* n_tiles, int* bn) {
  umma_n_tiling(d->Cout_store, n_tiles, bn);
  static int wide = -1;
  if (wide < 0) {
    const char* e = getenv("BRTPE_UMMA_WIDE");
    wide = e ? atoi(e) : 192;
    if (wide == 1) wide = 192;
    if (wide != 0 && (wide < 128 || wide > 256)) wide = 192;
  }
  if (!wide || d->ntaps * d->Cin < 768) return;
  const int cp = (d->Cout_store + 15) / 16 * 16;
  const int nt2 = (cp + wide - 1) / wide;
  const int bn2 = ((cp + nt2 - 1) / nt2 + 15) / 16 * 16;
  if (nt2 < *n_tiles && nt2 * bn2 == *n_tiles * *bn) {
    *n_tiles = nt2;
    *bn = bn2;
  }
}

bool umma_conv_supported(const brtpe_conv_desc* d, const char** why) {
  static const char* msg = "";
  auto fail = [&](const char* m) {
    msg = m;
    if (why) *why = msg;
    return false;
  };
  if (d->dtype != BRTPE_DT_BF16 && d->dtype != BRTPE_DT_BF16X2)
    return fail("tcgen05 path needs bf16 (or split bf16x2) activations");
  if (conv_is_split(d) && !epi_split_ok(d))
    return fail("split (bf16x2) layers need Cout == Cout_store, a multiple of 16, and pixel strides "
                "that are multiples of 32");
  if (conv_is_split(d) && d->in_ld / 2 < d->in_coff + d->Cin) return fail("split: in_ld / 2 < in_coff + Cin");
  if (d->Cin % 16) return fail("Cin must be a multiple of 16");
  if (d->in_ld % 8 || d->in_coff % 8) return fail("in_ld / in_coff must be multiples of 8");
  if (d->out_ld % 8 || d->out_coff % 8) return fail("out_ld / out_coff must be multiples of 8");
  if (d->res_ld % 8 || d->res_coff % 8) return fail("res_ld / res_coff must be multiples of 8");
  if (d->in_stride != 1 && d->in_stride != 2) return fail("in_stride must be 1 or 2");
  if (d->in_stride == 2 && ((d->Hin | d->Win) & 1)) return fail("stride 2 needs even Hin/Win");
  // stride 1: a tap is just a shifted TMA box (zero fill outside the map), so dilated 3x3
  // convolutions run here too; stride 2 maps taps onto pixel parities and needs [-1, 1]
  const int tmax = d->in_stride == 1 ? 64 : 1;
  for (int t = 0; t < d->ntaps; ++t)
    if (d->tap_dy[t] < -tmax || d->tap_dy[t] > tmax || d->tap_dx[t] < -tmax || d->tap_dx[t] > tmax)
      return fail("tap offsets must be in [-1, 1] (stride 2) / [-64, 64] (stride 1)");
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
  umma_n_tiling_layer(d, &p.n_tiles, &p.BN);
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;

  p.ntaps = d->ntaps;
  const bool split = conv_is_split(d);
  p.nkb_seg = ceil_div(d->Cin, 64);
  p.num_kb = split ? 3 * p.nkb_seg : p.nkb_seg;
  p.last_k16 = (d->Cin - (p.nkb_seg - 1) * 64) / 16;
  p.lo_off = d->in_ld / 2;
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
  {
    static int eg = -1;                            // BRTPE_UMMA_EPI8=0: four epilogue warps everywhere
    if (eg < 0) eg = getenv("BRTPE_UMMA_EPI8") ? atoi(getenv("BRTPE_UMMA_EPI8")) : 1;
    const bool stream_1x1 = d->ntaps == 1 && d->in_stride == 1 && p.BN >= 64 && (p.BN >> 4) % 2 == 0 &&
                            d->Cin >= 64 && d->Cin <= 256 && (long)p.total_tiles >= 8L * num_sms() &&
                            d->n_add == 0;     // (the stem's K = 32 GEMM @320^2 loses: 0.431 -> 0.488 ms)
    p.epi_groups = (eg && stream_1x1) ? 2 : 1;
  }

  // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);

  p.bias = nullptr;
  p.Hout = d->Hout; p.Wout = d->Wout; p.out_scale = d->out_scale; p.out_oy = d->out_oy;
  p.out_ox = d->out_ox;
  {
    // on for residual layers unless BRTPE_UMMA_RES_PREFETCH=0 (A/B switch)
    const char* e = getenv("BRTPE_UMMA_RES_PREFETCH");
    p.res_l2_prefetch = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  p.epi.out = nullptr; p.epi.res = nullptr;
  p.epi.out_ld = d->out_ld; p.epi.out_coff = d->out_coff; p.epi.res_ld = d->res_ld;
  p.epi.res_coff = d->res_coff; p.epi.Cout = d->Cout; p.epi.Cout_store = d->Cout_store;
  p.epi.relu = d->relu; p.epi.vec32 = epi_vec32_ok(d);
  // the fast epilogue walks whole 16-channel chunks of every Cout tile without a bound check: only
  // for tilings that cover the channels exactly (176 = 2 x 96 would write 16 channels too many)
  p.epi.fast = (epi_fast_ok(d) && p.n_tiles * p.BN == d->Cout) ? 1 : 0;
  epi_set_split(&p.epi, d);
  epi_set_add(&p.epi, d);
  if ((d->n_add > 0 || d->out2_ld > 0) && (!epi_add_ok(d) || !p.epi.fast || d->out2_ld > 0)) {
    set_error("tcgen05 conv (per-tap engine): fuse addends need the fast epilogue, aligned addends and "
              "no second output");
    delete P;
    return nullptr;
  }
  if (split && !p.epi.fast) {
    set_error("tcgen05 conv: split layers need a Cout tiling that covers the channels exactly "
              "(Cout %d, %d x %d)", d->Cout, p.n_tiles, p.BN);
    delete P;
    return nullptr;
  }
  p.fd_nt = make_fastdiv((uint32_t)p.n_tiles, (uint64_t)p.total_tiles + 1);
  p.fd_xy = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y), (uint64_t)p.total_tiles + 1);
  p.fd_x = make_fastdiv((uint32_t)p.tiles_x, (uint64_t)p.tiles_x * p.tiles_y);
  p.fd_tw = make_fastdiv((uint32_t)p.TW, 128);
  p.fd_twh = make_fastdiv((uint32_t)(p.TW * p.TH), 128);

  // ---- tensor maps
  auto encode = get_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->in_ld * 2;
  cuuint64_t gdim[5], gstr[4];
  if (s == 1) {
    gdim[0] = (cuuint64_t)(split ? d->in_ld : d->in_coff + d->Cin); gdim[1] = d->Win; gdim[2] = 1;
    gdim[3] = d->Hin;
    gdim[4] = d->N;
    gstr[0] = ld_b; gstr[1] = ld_b * d->Win; gstr[2] = ld_b * d->Win;
    gstr[3] = ld_b * d->Win * d->Hin;
  } else {
    gdim[0] = (cuuint64_t)(split ? 2 * d->in_ld : d->in_ld + d->in_coff + d->Cin);
    gdim[1] = d->Win / 2; gdim[2] = 2;
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
    if (cudaFuncSetAttribute(conv_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
                     cudaStream_t st, const void* const* add_ptrs, void* out2) {
  UmmaParams p = P->p;
  p.bias = bias;
  p.epi.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.epi.out = reinterpret_cast<__nv_bfloat16*>(out);
  for (int k = 0; k < p.epi.n_add; ++k) {
    if (!add_ptrs || !add_ptrs[k]) {
      set_error("tcgen05 conv: fuse addend %d is null", k);
      return BRTPE_EINVAL;
    }
    p.epi.add[k] = reinterpret_cast<const __nv_bfloat16*>(add_ptrs[k]);
  }
  (void)out2;
  cudaError_t e;
  if (p.epi.split) e = launch_ex(conv_umma_kernel<true, false>, dim3(P->grid), dim3(p.epi_groups == 2 ? UM_THREADS : UM_THREADS_NARROW), P->smem, st, 0, true, p);
  else if (p.epi.n_add > 0) e = launch_ex(conv_umma_kernel<false, true>, dim3(P->grid), dim3(p.epi_groups == 2 ? UM_THREADS : UM_THREADS_NARROW), P->smem, st, 0, true, p);
  else e = launch_ex(conv_umma_kernel<false, false>, dim3(P->grid), dim3(p.epi_groups == 2 ? UM_THREADS : UM_THREADS_NARROW), P->smem, st, 0, true, p);
  if (e != cudaSuccess) {
    set_error("conv_umma_kernel launch failed: %s", cudaGetErrorString(e));
    return BRTPE_ECUDA;
  }
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
BRTPE_MBAR_DEBUG_EXPORT(brtpe_debug_mbar_umma)
