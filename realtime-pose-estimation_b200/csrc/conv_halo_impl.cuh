// tcgen05 implicit-GEMM kernel for the 3x3 / stride-1 / pad-1 convolutions of HigherHRNet
// (BasicBlock conv1/conv2, Bottleneck conv2, transition1[0]: 87 % of the network's FLOPs,
// rtpe/third_party/pose_higher_hrnet.py:40-43, :46-75, :85-87, :558-563).
//
// A operand: the 8x16-pixel output tile reads ONE (16+2)x(8+2)-pixel halo tile per 64-channel
// block, loaded once by TMA (zero fill = padding) as 180 rows of 128 B with SWIZZLE_128B.
// tcgen05 swizzles on absolute shared-memory address bits (profiles/r01_exp_umma_shift.md),
// so tap (kh, kw) is the same tile read through a descriptor whose start address is advanced
// by (kh*10 + kw)*128 B and whose 8-row-group pitch (SBO) is 10*128 B: the nine taps cost no
// extra global/L2 traffic.
//
// Work decomposition (version 2; measurements in profiles/r01d_halo_anatomy.md):
//  * persistent, ONE CTA per SM, 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9
//    two epilogue groups;
//  * one work item = a PAIR of pixel tiles (2 x 128 output pixels) x one Cout tile.  Every
//    weight stage that reaches shared memory feeds both tiles (two TMEM accumulators), which
//    halves the L2->SMEM weight traffic and the number of barrier hand-offs per MMA -- the two
//    things that bounded version 1 (one hand-off costs ~115 cycles, a 12 KB TMA stage ~228);
//  * weights stay resident for the whole kernel when they fit (Cin <= 64: 48/64-channel
//    layers), otherwise they stream in stages of 9, 3 or 1 taps;
//  * accumulators are double-buffered in TMEM when 4*BN <= 512 columns, so the epilogue of
//    item i overlaps the MMAs of item i+1; epilogue group g drains tile g of the pair;
//  * the epilogue runs one 64-channel round ahead with its residual loads, across tile
//    boundaries, so their L2/HBM latency is off the critical path.
#include "conv_common.cuh"
#include "umma_ptx.cuh"
#include "conv_epilogue.cuh"

#include <cudaTypedefs.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace brtpe {

constexpr int HL_THREADS = 352;                          // 11 warps: TMA, MMA, 8 epilogue, second MMA
constexpr int HL_TW = 8, HL_TH = 16;
constexpr int HL_PITCH = HL_TW + 2;                       // halo row pitch in pixels
constexpr int HL_HROWS = (HL_TH + 2) * HL_PITCH;          // 180 pixel rows of 128 B
constexpr int HL_A_BYTES = HL_HROWS * 128;                // 23040
constexpr int HL_A_TILE = 23552;                          // rounded to 1024
constexpr int HL_MAX_A = 8, HL_MAX_B = 8;                 // an A stage holds the tpc (1 or 2) tiles of one item
constexpr int HL_MAX_BN = 256;
constexpr int HL_SMEM_MAX = 227 * 1024;
constexpr int HL_TAIL = 4608;                             // barriers + TMEM slot + bias
constexpr int HL_STAGE_OUT = 2048;                        // per epilogue warp: 32 px x 32 ch bf16
constexpr int HL_STAGE_BYTES = 8 * HL_STAGE_OUT;          // x out_slabs (1 or 2 slabs per warp)

struct alignas(64) HaloParams {
  CUtensorMap tmap_a;           // stride 1: the halo tile; stride 2: pixel-parity plane (even row, even col)
  CUtensorMap tmap_a2[3];       // stride 2: planes (even, odd col), (odd row, even), (odd, odd)
  CUtensorMap tmap_b;
  CUtensorMap tmap_o;           // output (C, W, H, N), box 32 ch x 8 px x 4 rows, SWIZZLE_64B
  CUtensorMap tmap_r;           // residual (C, W, H, N), box BN ch x 8 px x 16 rows: L2 prefetch only
  int res_prefetch;             // 1: the producer prefetches every item's residual tiles into L2
  CUtensorMap tmap_r4;          // residual (C, W, H, N), box BN ch x 8 px x 4 rows = one epilogue warp's pixels
  int res_staged;               // 1: every epilogue warp pulls the residual of its 32 pixels of the NEXT item
                                // into its own shared-memory slab by TMA (narrow resident-weight layers: the
                                // register-prefetched global loads were latency bound, profiles/r02_mma_issue.md)
  int res_slab_bytes;
  int N, H, W;
  int tiles_x, tiles_y, m_tiles, n_tiles, BN, num_items;
  int num_kb, last_k16, in_coff;
  int nkb_seg, lo_off;          // split (BRTPE_DT_BF16X2): num_kb = 3 segments [hi | lo | hi] of
                                // nkb_seg 64-channel blocks; the lo half starts lo_off channels in
  int a_stages, b_stages, b_stage_bytes, a_stage_bytes;
  int reverse;                  // 1: tiles are walked from the last image to the first (L2 reuse between
                                // consecutive layers, brtpe_conv_desc.reverse_order)
  int s2_ld;                    // stride 2: input pixel stride (the second pixel of a column pair)
  int s2;                       // 1: 3x3 / stride 2 (four pixel-parity planes per tile instead of one halo tile)
  int a_tile_bytes;             // shared-memory bytes of one pixel tile's activations (1024-aligned)
  int a_tx_bytes;               // bytes TMA delivers per pixel tile
  int sub_off[4];               // stride 2: byte offset of the parity planes inside a tile
  uint32_t tap_aoff[9];         // per tap: descriptor start offset inside the tile (bytes >> 4) ...
  uint32_t tap_ahi[9];          // ... and descriptor high word (SBO = rows-of-8 pitch of the plane the tap reads)
  int tps, b_groups;            // taps per weight stage, stages per channel block (tps*b_groups = 9)
  int resident;                 // 1: all weights stay in smem for the whole kernel
  int cg;                       // 1: one CTA per MMA; 2: CTA pair (tcgen05 cta_group::2): M = 256 over
                                // two SMs, each CTA holds its 2 pixel tiles and HALF of the weight rows
  int tpc;                      // pixel tiles per CTA and work item: 2, or 1 for wide Cout tiles (two
                                // accumulators of BN columns would leave no room for double buffering)
  int num_units;                // groups of tpc*cg pixel tiles
  int acc_stages, tmem_cols;
  int dual;                     // 1: two MMA-issuing warps take alternate work items (needs acc_stages == 2)
  int out_slabs;                // staging slabs per epilogue warp (2 when shared memory allows)
  int tma_out;                  // 1: output through staging slabs + TMA stores, 0: direct stores
  uint32_t idesc;
  FastDiv fd_xy, fd_x, fd_nt;
  const float* bias;
  EpiParams epi;
  // Chain mode (conv1 -> conv2 of a BasicBlock in ONE launch, halo_chain_prepare): the work list
  // interleaves sub-batches of the two layers -- [L0 s0][L0 s1][L1 s0][L0 s2][L1 s1] ... -- so that a
  // sub-batch's block input, intermediate and output (3 x ch_ip*2 tiles) are still in L2 when
  // layer 1 reads them; a layer-1 tile waits for the layer-0 tiles of ITS image (ch_count).
  int chain;                    // 1: chain mode
  int ch_ip, ch_S;              // work items per phase (= one sub-batch of one layer), sub-batches
  FastDiv fd_ip;
  CUtensorMap tmap_a1;          // layer 1 reads the intermediate tensor
  CUtensorMap tmap_b1;          // layer-1 weights (resident next to layer 0's)
  int in_coff1;
  const float* bias1;
  __nv_bfloat16* out1;          // layer-1 output / residual (p.epi describes layer 0)
  const __nv_bfloat16* res1;
  int out1_ld, out1_coff, res1_ld, res1_coff;
  int* ch_count;                // [N] epilogue-warp completions of layer 0 per image, [N] = CTAs done
  int ch_need;                  // completions that make an image ready: tiles per image x 4 warps
  int ch_defer;                 // 1: an epilogue warp counts item i when it starts item i+1 (its stores have
                                // landed by then, the fence costs nothing); needs phases of >= 2 grid rounds
  int ch_pf;                    // layer-0 activation tiles are prefetched into L2 this many items ahead
  int ch_dbg;                   // debug (BRTPE_CHAIN_DBG, timing only, results may be wrong): 1 no
                                // completion signal / fence, 2 no dependency wait
  long long* prof;              // debug: per-CTA role/wait cycle counters
  int dbg;                      // debug (BRTPE_HALO_DBG): 1 no MMA issue, 2 no epilogue work,
                                // 4 no weight loads, 8 no activation loads (results are garbage)
};

// Debug instrumentation (brtpe_debug_halo_prof): 16 counters per CTA, in SM clock cycles.
//  0 kernel total (thread 0)   1 prologue (until the first CTA-wide sync)
//  2 producer loop total       3 producer waiting for a free A stage   4 ... free B stage
//  5 MMA loop total            6 MMA waiting for A data   7 ... B data   8 ... a free accumulator
//  9 epilogue warp 2 loop total  10 epilogue waiting for the accumulator  11 items of this CTA
constexpr int HL_PROF_SLOTS = 16;
extern long long* g_halo_prof;       // defined in conv_halo.cu
extern int g_halo_prof_ctas;

#define HL_TIMED(slot, stmt)                   \
  do {                                         \
    if (PROF) {                                \
      const long long t0__ = clock64();        \
      stmt;                                    \
      pc[slot] += clock64() - t0__;            \
    } else {                                   \
      stmt;                                    \
    }                                          \
  } while (0)

struct HaloConvPrepared {
  HaloParams p;
  int grid;
  size_t smem;
  brtpe_conv_desc d;            // to (re-)encode the output / residual tensor maps
  const void* out_encoded;      // output pointer tmap_o was encoded for
  const void* res_encoded;      // residual pointer tmap_r was encoded for (mutable cache)
  CUtensorMap tmap_r_cache;
  const void* res4_encoded = nullptr;   // same for tmap_r4 (staged residual)
  CUtensorMap tmap_r4_cache;
  // chain mode (halo_chain_prepare): second layer + the intermediate tensor + the image counters
  brtpe_conv_desc d1;
  void* mid = nullptr;
  int* counters = nullptr;
  ~HaloConvPrepared() {
    if (counters) cudaFree(counters);
  }
};

// descriptor halves (see make_kmajor_sw128_desc): lo = start>>4 | LBO(1)<<16, hi = SBO>>4 |
// version(1)<<14 | SWIZZLE_128B(2)<<29
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
}
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
}

// tile index -> image / origin of the 8x16 output tile (n == p.N for tiles past the end:
// their TMA box is fully out of bounds and reads zeros, their outputs are never stored)
struct TileOrg {
  int n, y0, x0;
};
__device__ __forceinline__ TileOrg tile_origin(const HaloParams& p, int mt) {
  TileOrg o;
  const int tiles_xy = p.tiles_x * p.tiles_y;
  const bool past_end = mt >= p.m_tiles;
  if (p.reverse && !past_end) mt = p.m_tiles - 1 - mt;
  o.n = (int)fdiv((uint32_t)mt, p.fd_xy);
  const int rem = mt - o.n * tiles_xy;
  const int ty = (int)fdiv((uint32_t)rem, p.fd_x);
  o.y0 = ty * HL_TH;
  o.x0 = (rem - ty * p.tiles_x) * HL_TW;
  if (past_end) o.n = p.N;
  return o;
}

// work item -> (layer, unit of tpc*cg pixel tiles).  Plain kernels: unit = item / n_tiles.  Chain
// mode (n_tiles == 1): phase ph = item / ch_ip of the interleaved list (see HaloParams).
template <bool CHAIN>
__device__ __forceinline__ int item_unit(const HaloParams& p, int item, int& layer) {
  if constexpr (CHAIN) {
    const int ph = (int)fdiv((uint32_t)item, p.fd_ip);
    const int r = item - ph * p.ch_ip;
    int s;
    if (ph == 0) { layer = 0; s = 0; }
    else if (ph == 2 * p.ch_S - 1) { layer = 1; s = p.ch_S - 1; }
    else if (ph & 1) { layer = 0; s = (ph + 1) >> 1; }
    else { layer = 1; s = (ph >> 1) - 1; }
    return s * p.ch_ip + r;
  } else {
    layer = 0;
    return (int)fdiv((uint32_t)item, p.fd_nt);
  }
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ---- epilogue, fast path: one 64-channel round of residual in flight ahead of the math
struct EpiPix {
  bool valid;
  size_t opix;
  int n, y, x;                   // output pixel (fuse addends are indexed at (y >> s, x >> s))
};
template <bool CHAIN = false>
__device__ __forceinline__ EpiPix epi_pixel(const HaloParams& p, int item, int tile, int m,
                                            int rank, int* layer_out = nullptr) {
  EpiPix e;
  e.valid = false;
  e.opix = 0;
  e.n = e.y = e.x = 0;
  if (layer_out) *layer_out = 0;
  if (item >= p.num_items) return e;
  int layer = 0;
  const int unit = item_unit<CHAIN>(p, item, layer);
  if (layer_out) *layer_out = layer;
  const int mt = unit * (p.tpc * p.cg) + rank * p.tpc + tile;
  const TileOrg o = tile_origin(p, mt);
  const int y = o.y0 + (m >> 3), x = o.x0 + (m & 7);
  e.valid = mt < p.m_tiles && y < p.H && x < p.W;
  e.opix = e.valid ? ((size_t)o.n * p.H + y) * p.W + x : 0;
  e.n = o.n; e.y = y; e.x = x;
  return e;
}

template <bool RES, bool CHAIN = false>
__device__ __forceinline__ void epi_fetch_round(Chunk32 (&r)[4], const __nv_bfloat16* rp,
                                                bool valid, int c0, int nchunks) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (RES && valid && c0 + j < nchunks) {
      r[j] = ld_chunk32(rp + (c0 + j) * 16, true);
    } else if (CHAIN) {                       // a layer without residual follows one with: zeros
#pragma unroll
      for (int i = 0; i < 8; ++i) r[j].w[i] = 0u;
    }
  }
}

template <bool RES, bool RELU, bool ADD, bool CHAIN = false, bool STAGED = false>
__device__ __forceinline__ void halo_epilogue_fast(const HaloParams& p, const float* bias_s,
                                                   uint32_t tmem_base, uint64_t* tfull,
                                                   uint32_t tempty_addr, int group, int lg,
                                                   int lane, int rank, int item0, int istep,
                                                   uint32_t stage_buf, bool PROF, long long* pc,
                                                   uint32_t res_full = 0, uint32_t res_empty = 0) {
  const EpiParams& e = p.epi;
  const int m = lg * 32 + lane;
  const int BN = p.BN, n_tiles = p.n_tiles, num_items = p.num_items;
  // two tiles per CTA: group g drains tile g; one tile per CTA: both groups drain it, group g the
  // channel chunks [cbeg, cend)
  const int tile_sel = (p.tpc == 2) ? group : 0;
  const int allchunks = BN >> 4;
  // one-tile mode: group 0 takes the chunks [0, csplit), group 1 [csplit, allchunks); csplit is even
  // (chunks are drained in pairs, and a 32-channel TMA store must not straddle the two groups)
  const int csplit = ((allchunks + 3) >> 2) << 1;
  const int cbeg = (p.tpc == 2) ? 0 : (group == 0 ? 0 : (csplit < allchunks ? csplit : allchunks));
  const int nchunks = (p.tpc == 2) ? allchunks : (group == 0 ? (csplit < allchunks ? csplit : allchunks) : allchunks);
  const int acc_stages = p.acc_stages;
  const int acc_cols = p.tpc * BN;
  const uint32_t bias_u32 = smem_u32(bias_s);
  Chunk32 nxt[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) nxt[j].w[i] = 0u;

  constexpr bool remote = (HL_CG == 2);
  const bool two_slabs = p.out_slabs == 2;
  const bool tma_out = p.tma_out != 0;
  int slab_sel = 0;   // the accumulator-free barrier lives in the leader CTA
  int item = item0;
  int pend_n = -1;              // chain mode: image whose completion count this warp still owes
  int lay = 0;                  // chain mode: layer of the current item (0: no residual, out = the
                                // intermediate tensor; 1: residual + the block's output)
  EpiPix px = epi_pixel<CHAIN>(p, item, tile_sel, m, rank, &lay);
  int nt = CHAIN ? 0 : item - (int)fdiv((uint32_t)item, p.fd_nt) * n_tiles;
  // STAGED: stage_buf is this warp's residual slab (32 pixels x BN channels, pixel-major); the warp
  // itself requests the slab of its next item as soon as it has copied the current one to registers
  auto res_request = [&](int it_) {
    int lu = 0;
    const TileOrg o = tile_origin(p, item_unit<CHAIN>(p, it_, lu) * (p.tpc * p.cg) + rank * p.tpc + tile_sel);
    mbar_expect_tx(res_full, (uint32_t)(32 * BN * 2));
    tma_load_4d(stage_buf, &p.tmap_r4, res_full, e.res_coff, o.x0, o.y0 + 4 * lg, o.n);
  };
  if constexpr (STAGED) {
    if (lane == 0 && item < num_items) res_request(item);
  } else if constexpr (CHAIN)
    epi_fetch_round<RES, true>(nxt, p.res1 + px.opix * p.res1_ld + p.res1_coff, px.valid && lay == 1,
                               cbeg, nchunks);
  else
    epi_fetch_round<RES>(nxt, RES ? e.res + px.opix * e.res_ld + e.res_coff + nt * BN : nullptr,
                         px.valid, cbeg, nchunks);
  int it = 0;
  for (; item < num_items; item += istep, ++it) {
    const int next_item = item + istep;
    int nlay = 0;
    const EpiPix npx = epi_pixel<CHAIN>(p, next_item, tile_sel, m, rank, &nlay);
    const int nnt = CHAIN ? 0 : next_item - (int)fdiv((uint32_t)next_item, p.fd_nt) * n_tiles;
    const int co0 = nt * BN;
    const __nv_bfloat16* rp;
    const __nv_bfloat16* nrp;
    bool rvalid = px.valid, nrvalid = npx.valid;
    if constexpr (CHAIN) {
      rp = p.res1 + px.opix * p.res1_ld + p.res1_coff;
      nrp = p.res1 + npx.opix * p.res1_ld + p.res1_coff;
      rvalid = px.valid && lay == 1;
      nrvalid = npx.valid && nlay == 1;
    } else {
      rp = RES ? e.res + px.opix * e.res_ld + e.res_coff + co0 : nullptr;
      nrp = RES ? e.res + npx.opix * e.res_ld + e.res_coff + nnt * BN : nullptr;
    }
    // origin of this warp's 8 x 4 pixel slab (TMA clips what lies outside the image / batch)
    int lay_unused = 0;
    const TileOrg org = tile_origin(
        p, item_unit<CHAIN>(p, item, lay_unused) * (p.tpc * p.cg) + rank * p.tpc + tile_sel);
    const int oy = org.y0 + 4 * lg;
    const int acc = (acc_stages == 2) ? (it & 1) : 0;
    const uint32_t accph = (acc_stages == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
    const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) +
                            (uint32_t)(acc * acc_cols + tile_sel * BN);
    if constexpr (STAGED) {
      // residual of this item: slab -> registers, then the slab is free for the next item's TMA (while
      // the MMAs of this item are still running)
      mbar_wait(res_full, (uint32_t)it & 1u);
      const uint32_t row = stage_buf + (uint32_t)(lane * BN * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nchunks) {
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(nxt[j].w[0]), "=r"(nxt[j].w[1]), "=r"(nxt[j].w[2]), "=r"(nxt[j].w[3])
                       : "r"(row + (uint32_t)(j * 32)));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(nxt[j].w[4]), "=r"(nxt[j].w[5]), "=r"(nxt[j].w[6]), "=r"(nxt[j].w[7])
                       : "r"(row + (uint32_t)(j * 32 + 16)));
        }
      }
      mbar_arrive(res_empty);                           // all 32 lanes: their reads are done
      if (lane == 0 && next_item < num_items) {
        mbar_wait(res_empty, (uint32_t)it & 1u);
        res_request(next_item);
      }
      __syncwarp();
    }
    HL_TIMED(10, mbar_wait(smem_u32(&tfull[acc]), accph));
    if (PROF) pc[11] += 1;
    tc_fence_after();
    if constexpr (CHAIN) {
      // deferred completion count of the previous (layer-0) item: its stores were issued one work
      // item ago, so the fence does not wait, and none of this item's stores is in flight yet
      if (pend_n >= 0) {
        __threadfence();
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(p.ch_count + pend_n, 1);
        pend_n = -1;
      }
    }
    if (cbeg >= nchunks) {                             // this group has no chunk of the tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (remote) mbar_arrive_cluster(tempty_addr + 8u * acc);
        else mbar_arrive(tempty_addr + 8u * acc);
      }
    }
#pragma unroll 1
    for (int c0 = cbeg; c0 < nchunks; c0 += 4) {
      Chunk32 cur[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
      if constexpr (!STAGED) {
        if (c0 + 4 < nchunks) epi_fetch_round<RES, CHAIN>(nxt, rp, rvalid, c0 + 4, nchunks);
        else epi_fetch_round<RES, CHAIN>(nxt, nrp, nrvalid, cbeg, nchunks);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 2 * h;
        if (c < nchunks) {
          const bool two = c + 1 < nchunks;
          uint32_t a[32];
          if (two) tmem_ld32(t_addr + (uint32_t)(c * 16), a);
          else tmem_ld16_lo(t_addr + (uint32_t)(c * 16), a);
          tmem_ld_wait();
          if (c + 2 >= nchunks) {                  // last TMEM read of this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (remote) mbar_arrive_cluster(tempty_addr + 8u * acc);
              else mbar_arrive(tempty_addr + 8u * acc);
            }
          }
          const Chunk32& r0 = cur[2 * h];
          const Chunk32& r1 = cur[2 * h + 1];
          // HRNet fuse addends (loaded here, not prefetched: a register-prefetched variant with in-place
          // refill spilled and was 4 % slower on the whole forward, profiles/r02_fuse_epilogue.md):
          // single output -> they join the accumulator; two outputs (direct stores only) ->
          // x = act(acc + bias + res) goes to out, relu(x + sum) to out2
          const bool dual = ADD && e.out2 != nullptr;
          // sum of the addends of chunk c + k (k = 0, 1), loaded at their point of use (register
          // prefetch variants -- in-place refill a round ahead, early loads before the accumulator wait
          // -- spilled / indexed local memory and were 3-4 % slower: profiles/r02_fuse_epilogue.md)
          auto terms = [&](int k, float (&s)[16]) {
            epi_terms16(e, px.n, px.y, px.x, co0 + (c + k) * 16, s);
          };
          if (ADD && !dual && px.valid) {
            float s[16];
            terms(0, s);
            epi_add_terms<0>(a, s);
            if (two) {
              terms(1, s);
              epi_add_terms<16>(a, s);
            }
          }
          if (!tma_out || (!two && n_tiles > 1)) {
            // 16-channel tail of a Cout tile that has a neighbour: the 32-channel store box
            // would spill into the neighbour's channels, so this chunk is stored directly
            if (px.valid) {
              uint32_t bs = bias_u32 + (uint32_t)((co0 + c * 16) * 4);
              __nv_bfloat16* op = e.out + px.opix * e.out_ld + e.out_coff + co0 + c * 16;
              if constexpr (CHAIN) {
                if (lay == 1) {
                  bs += (uint32_t)(BN * 4);
                  op = p.out1 + px.opix * p.out1_ld + p.out1_coff + c * 16;
                }
              }
              if (dual) {
                __nv_bfloat16* op2 = e.out2 + px.opix * e.out2_ld + co0 + c * 16;
                float s[16];
                terms(0, s);
                epi_dual_chunk<RES, RELU, 0>(a, bs, r0, s, op, op2);
                if (two) {
                  terms(1, s);
                  epi_dual_chunk<RES, RELU, 16>(a, bs + 64u, r1, s, op + 16, op2 + 16);
                }
              } else {
                epi_fast_chunk<RES, RELU, 0>(a, bs, r0, op);
                if (two) epi_fast_chunk<RES, RELU, 16>(a, bs + 64u, r1, op + 16);
              }
            }
          } else {
            // 32 (or 16) channels of 32 pixels -> bf16 in registers -> this warp's staging
            // slab (64-byte rows, SWIZZLE_64B so the 16-byte stores are conflict free) -> one
            // TMA store.  One line request per lane and access was the limit of the direct
            // 32-byte global stores (profiles/r01e_halo_pair.md).
            const uint32_t bs = bias_u32 + (uint32_t)((co0 + c * 16) * 4);
            Chunk32 o0, o1;
            epi_pack_chunk<RES, RELU, 0>(a, bs, r0, o0);
            if (two) epi_pack_chunk<RES, RELU, 16>(a, bs + 64u, r1, o1);
            // slab (re-)use: with two slabs per warp only the store before the previous one
            // must have finished reading
            if (lane == 0) {
              if (two_slabs) bulk_wait_read1();
              else bulk_wait_read0();
            }
            __syncwarp();
            const uint32_t slab = stage_buf + (two_slabs ? (uint32_t)(slab_sel * HL_STAGE_OUT) : 0u);
            slab_sel ^= 1;
            const uint32_t row = slab + (uint32_t)lane * 64u;
            const uint32_t sw = (uint32_t)((lane >> 1) & 3);
            st_shared_v4(row + ((0u ^ sw) << 4), o0.w[0], o0.w[1], o0.w[2], o0.w[3]);
            st_shared_v4(row + ((1u ^ sw) << 4), o0.w[4], o0.w[5], o0.w[6], o0.w[7]);
            if (two) {
              st_shared_v4(row + ((2u ^ sw) << 4), o1.w[0], o1.w[1], o1.w[2], o1.w[3]);
              st_shared_v4(row + ((3u ^ sw) << 4), o1.w[4], o1.w[5], o1.w[6], o1.w[7]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&p.tmap_o, slab, e.out_coff + co0 + c * 16, org.x0, oy, org.n);
              bulk_commit();
            }
          }
        }
      }
    }
    if constexpr (CHAIN) {
      // layer 0: this warp's 32 pixels of the intermediate tensor are stored -> count them on the
      // tile's image (release: the stores of all 32 lanes; the layer-1 producer that waits for
      // the image reads them through TMA = the async proxy)
      if (lay == 0 && !(p.ch_dbg & 1) && org.n < p.N) {
        if (p.ch_defer) {
          pend_n = org.n;
        } else {
          __threadfence();
          fence_proxy_async_all();
          __syncwarp();
          if (lane == 0) red_release_gpu_add(p.ch_count + org.n, 1);
        }
      }
    }
    px = npx;
    nt = nnt;
    lay = nlay;
  }
  if constexpr (CHAIN) {
    if (pend_n >= 0) {
      __threadfence();
      fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) red_release_gpu_add(p.ch_count + pend_n, 1);
    }
  }
  if (lane == 0) bulk_wait0();                      // every TMA store of this warp has landed
  __syncwarp();
}

// MMA issue of `tps` taps (starting at tap `tap0`) of one 64-channel block for the work item's one or
// two pixel tiles: K16 MMAs per tap and tile, no branch between them.  Tap (kh, kw) reads the halo
// tile through a descriptor advanced by (kh*PITCH + kw) rows of 128 B.
template <int K16, bool CG2>
__device__ __forceinline__ void halo_issue_taps(const HaloParams& p, uint32_t d0, uint32_t d1,
                                                uint32_t a0, uint32_t a1, uint32_t b_lo,
                                                uint32_t b_tap_lo, uint32_t b_hi, uint32_t idesc,
                                                int tap0, int tps, int tpc, uint32_t first_base) {
#pragma unroll 1
  for (int t = 0; t < tps; ++t) {
    // stride 1: the tile read at (kh*PITCH + kw) rows of 128 B; stride 2: the tap's parity plane
    const uint32_t a_off = p.tap_aoff[tap0 + t];
    const uint32_t a_hi = p.tap_ahi[tap0 + t];
    const uint32_t bt = b_lo + (uint32_t)t * b_tap_lo;
    const uint32_t first = first_base | (uint32_t)t;
#pragma unroll
    for (int k = 0; k < K16; ++k) {
      if (CG2) umma_f16_lohi_cg2(d0, a0 + a_off + 2u * k, a_hi, bt + 2u * k, b_hi, idesc, (first | (uint32_t)k) ? 1u : 0u);
      else umma_f16_lohi(d0, a0 + a_off + 2u * k, a_hi, bt + 2u * k, b_hi, idesc, (first | (uint32_t)k) ? 1u : 0u);
    }
    if (tpc == 2) {
#pragma unroll
      for (int k = 0; k < K16; ++k) {
        if (CG2) umma_f16_lohi_cg2(d1, a1 + a_off + 2u * k, a_hi, bt + 2u * k, b_hi, idesc, (first | (uint32_t)k) ? 1u : 0u);
        else umma_f16_lohi(d1, a1 + a_off + 2u * k, a_hi, bt + 2u * k, b_hi, idesc, (first | (uint32_t)k) ? 1u : 0u);
      }
    }
  }
}

// ---- split (BRTPE_DT_BF16X2) epilogue: float32 result -> (hi, lo) bf16 pair, residual = hi + lo.
// The MMAs of a split layer take three times as long per tile, so the plain form (residual loads
// issued right before the TMEM read, direct 32-byte stores) stays off the critical path.
template <bool RES, bool RELU>
__device__ __forceinline__ void halo_epilogue_split(const HaloParams& p, const float* bias_s,
                                                    uint32_t tmem_base, uint64_t* tfull,
                                                    uint32_t tempty_addr, int group, int lg,
                                                    int lane, int rank, int item0, int istep) {
  const EpiParams& e = p.epi;
  const int m = lg * 32 + lane;
  const int BN = p.BN, n_tiles = p.n_tiles, num_items = p.num_items;
  const int tile_sel = (p.tpc == 2) ? group : 0;
  const int allchunks = BN >> 4;
  // one-tile mode: group 0 takes the chunks [0, csplit), group 1 [csplit, allchunks); csplit is even
  // (chunks are drained in pairs, and a 32-channel TMA store must not straddle the two groups)
  const int csplit = ((allchunks + 3) >> 2) << 1;
  const int cbeg = (p.tpc == 2) ? 0 : (group == 0 ? 0 : (csplit < allchunks ? csplit : allchunks));
  const int nchunks = (p.tpc == 2) ? allchunks : (group == 0 ? (csplit < allchunks ? csplit : allchunks) : allchunks);
  const int acc_stages = p.acc_stages;
  const int acc_cols = p.tpc * BN;
  const uint32_t bias_u32 = smem_u32(bias_s);
  constexpr bool remote = (HL_CG == 2);
  int it = 0;
  for (int item = item0; item < num_items; item += istep, ++it) {
    const EpiPix px = epi_pixel(p, item, tile_sel, m, rank);
    const int nt = item - (int)fdiv((uint32_t)item, p.fd_nt) * n_tiles;
    const int co0 = nt * BN;
    const __nv_bfloat16* rp = RES ? e.res + px.opix * e.res_ld + e.res_coff + co0 : nullptr;
    __nv_bfloat16* op = e.out + px.opix * e.out_ld + e.out_coff + co0;
    const int acc = (acc_stages == 2) ? (it & 1) : 0;
    const uint32_t accph = (acc_stages == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
    const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) +
                            (uint32_t)(acc * acc_cols + tile_sel * BN);
    mbar_wait(smem_u32(&tfull[acc]), accph);
    tc_fence_after();
    if (cbeg >= nchunks) {                             // this group has no chunk of the tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (remote) mbar_arrive_cluster(tempty_addr + 8u * acc);
        else mbar_arrive(tempty_addr + 8u * acc);
      }
    }
#pragma unroll 1
    for (int c = cbeg; c < nchunks; c += 2) {
      const bool two = c + 1 < nchunks;
      Chunk32 rh0, rl0, rh1, rl1;
#pragma unroll
      for (int i = 0; i < 8; ++i) rh0.w[i] = rl0.w[i] = rh1.w[i] = rl1.w[i] = 0u;
      if (RES && px.valid) {
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
      if (c + 2 >= nchunks) {                    // last TMEM read of this tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (remote) mbar_arrive_cluster(tempty_addr + 8u * acc);
          else mbar_arrive(tempty_addr + 8u * acc);
        }
      }
      if (px.valid) {
        const uint32_t bs = bias_u32 + (uint32_t)((co0 + c * 16) * 4);
        epi_split_chunk<RES, RELU, 0>(a, bs, rh0, rl0, op + c * 16, op + e.out_lo + c * 16);
        if (two)
          epi_split_chunk<RES, RELU, 16>(a, bs + 64u, rh1, rl1, op + c * 16 + 16,
                                         op + e.out_lo + c * 16 + 16);
      }
    }
  }
}

// Instantiations (the kernel is ~8000 instructions with every path in it, and the hot loops are
// sensitive to code size: an extra, unused epilogue variant cost 5 % on the 48-channel layers,
// profiles/r02_mma_issue.md): SPLIT = BRTPE_DT_BF16X2 epilogue; PROFT = cycle counters compiled in;
// EPI = -1: epilogue chosen at run time (general path included; only ONE instantiation may call the
// non-inlined epi_drain: ptxas 12.9 crashes on two), -2: run-time choice among the fast epilogues,
// 0..3: fast epilogue with RES = EPI >> 1, RELU = EPI & 1 only.
template <bool SPLIT, bool PROFT, int EPI>
__global__ void __launch_bounds__(HL_THREADS, 1)
HL_NAME(conv_halo_kernel)(const __grid_constant__ HaloParams p) {
  const bool PROF = PROFT && p.prof != nullptr;   // debug counters (brtpe_debug_halo_prof), warp-uniform
  constexpr bool CHAIN = (EPI == 8);              // two layers per launch (HaloParams::chain)
  constexpr bool STAGED = (EPI == 9);             // residual through per-warp TMA slabs (HaloParams::res_staged)
  extern __shared__ uint8_t smem_raw[];
  long long pc[HL_PROF_SLOTS];
#pragma unroll
  for (int i = 0; i < HL_PROF_SLOTS; ++i) pc[i] = 0;
  const long long t_start = PROF ? clock64() : 0;
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // hoist everything the hot loops need out of the constant bank
  const int a_stages = p.a_stages, b_stages = p.b_stages, b_stage_bytes = p.b_stage_bytes;
  const int num_kb = p.num_kb, last_k16 = p.last_k16, num_items = p.num_items;
  const int nkb_seg = p.nkb_seg, lo_off = p.lo_off;
  const int n_tiles = p.n_tiles, BN = p.BN;
  const int tps = p.tps, b_groups = p.b_groups, acc_stages = p.acc_stages;
  const bool resident = p.resident != 0;
  const uint32_t idesc = p.idesc;
  const int dbg = p.dbg;
  // CTA pair (cta_group::2): rank 0 is the leader (issues the MMAs, owns the pair's barriers)
  constexpr bool cg2 = (HL_CG == 2);
  const int rank = cg2 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int item0 = cg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int istep = cg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tpc = p.tpc;                            // pixel tiles per CTA and work item (1 or 2)
  const int tpi = tpc * p.cg;                       // pixel tiles per work item
  const int BNh = cg2 ? (BN >> 1) : BN;             // weight rows this CTA holds

  uint8_t* a_ring = smem;
  const int a_stage_bytes = p.a_stage_bytes;
  uint8_t* b_ring = smem + (size_t)a_stages * a_stage_bytes;
  uint8_t* stage_out = b_ring + (size_t)b_stages * b_stage_bytes;     // 8 epilogue-warp slabs
  // (the residual slabs of the staged epilogue take the place of the output slabs: never both)
  uint8_t* tail = stage_out + (size_t)p.out_slabs * HL_STAGE_BYTES + (STAGED ? 8 * (size_t)p.res_slab_bytes : 0);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_a = full_a + HL_MAX_A;
  uint64_t* full_b = empty_a + HL_MAX_A;
  uint64_t* empty_b = full_b + HL_MAX_B;
  uint64_t* tfull = empty_b + HL_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint64_t* full_r = tempty + 2;                   // staged residual: one pair per epilogue warp
  uint64_t* empty_r = full_r + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_r + 8);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    tma_prefetch_desc(&p.tmap_o);
    if constexpr (CHAIN) {
      tma_prefetch_desc(&p.tmap_a1);
      tma_prefetch_desc(&p.tmap_b1);
    }
    for (int s = 0; s < HL_MAX_A; ++s) {
      mbar_init(smem_u32(&full_a[s]), 1);
      mbar_init(smem_u32(&empty_a[s]), 1);
    }
    for (int s = 0; s < HL_MAX_B; ++s) {
      mbar_init(smem_u32(&full_b[s]), 1);
      mbar_init(smem_u32(&empty_b[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), (uint32_t)(8 * p.cg));   // 8 epilogue warps per CTA
    }
    if constexpr (STAGED) {
      tma_prefetch_desc(&p.tmap_r4);
      for (int s = 0; s < 8; ++s) {
        mbar_init(smem_u32(&full_r[s]), 1);
        mbar_init(smem_u32(&empty_r[s]), 32);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_tiles * BN; i += HL_THREADS)
    bias_s[i] = (p.bias && i < p.epi.Cout) ? p.bias[i] : 0.0f;
  if constexpr (CHAIN) {
    for (int i = threadIdx.x; i < BN; i += HL_THREADS)
      bias_s[BN + i] = (p.bias1 && i < p.epi.Cout) ? p.bias1[i] : 0.0f;
  }
  if (warp == 1) {
    if (cg2) tmem_alloc_cg2(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    else tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (cg2) cluster_sync_all();      // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_loop = PROF ? clock64() : 0;
  // PDL: the next kernel of the lane may be scheduled as soon as every CTA of this one is past its
  // prologue (its CTAs take the SMs this kernel's CTAs leave and run THEIR prologue under our tail)
  if (threadIdx.x == 0) pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) ========
    // Pair mode: both CTAs load their own activation tiles and their half of the weight rows;
    // completion is counted on the LEADER's full barrier, which alone expects the bytes of both.
    int as_ = 0, bs_ = 0;
    uint32_t aph = 0, bph = 0;
    const uint32_t stage_b_bytes = (uint32_t)(tps * BN * 128);          // both CTAs together
    const int in_coff = p.in_coff;
    const int brow0 = rank * BNh;
    const bool res_prefetch = p.res_prefetch != 0;
    auto sig = [&](const uint64_t* bar) -> uint32_t {                   // barrier the TMA signals
      const uint32_t a = smem_u32(bar);
      return cg2 ? mapa_u32(a, 0u) : a;
    };
    if (resident) {
      // every tap of every channel block, once per kernel
      if (elect_one()) {
        const uint32_t fb = smem_u32(&full_b[0]);
        if (leader) mbar_expect_tx(fb, (uint32_t)((CHAIN ? 2 : 1) * num_kb * 9 * BN * 128));
        const uint32_t fbs = sig(&full_b[0]);
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t dst = smem_u32(b_ring) + (uint32_t)(kb * 9 * BNh * 128);
          if (cg2) tma_load_3d_cg2(dst, &p.tmap_b, fbs, kb * 64, brow0, 0);
          else tma_load_3d(dst, &p.tmap_b, fb, kb * 64, 0, 0);
        }
        if constexpr (CHAIN)                        // (num_kb == 1) layer 1 behind layer 0
          tma_load_3d(smem_u32(b_ring) + (uint32_t)(9 * BNh * 128), &p.tmap_b1, fb, 0, 0, 0);
      }
      __syncwarp();
    }
    pdl_wait();                                     // activations / residual: the predecessor's output
    // The activation tiles of step s+1 (a step = one channel block of one item) are requested as
    // soon as their stage is free, in between the weight stages of step s: the weight loads
    // block on the MMA's progress, and the activation load must neither queue behind all of
    // them nor hold them up.
    int nx_item = item0, nx_kb = 0;                 // next activation load to issue
    auto load_a = [&](bool block) -> bool {
      if (nx_item >= num_items) return true;
      const uint32_t eb = smem_u32(&empty_a[as_]);
      if (block) {
        HL_TIMED(3, mbar_wait(eb, aph ^ 1u));
      } else if (!mbar_test(eb, aph ^ 1u)) {
        return false;
      }
      int layer = 0;
      const int unit = item_unit<CHAIN>(p, nx_item, layer);
      const TileOrg o0 = tile_origin(p, unit * tpi + rank * tpc);
      const TileOrg o1 = tile_origin(p, unit * tpi + rank * tpc + 1);      // unused when tpc == 1
      // channel block nx_kb of the reduction: segment (hi, lo, hi again) and block inside it
      const int seg = (nx_kb >= nkb_seg) + (nx_kb >= 2 * nkb_seg);
      int ch0 = in_coff + (nx_kb - seg * nkb_seg) * 64 + (seg == 1 ? lo_off : 0);
      const CUtensorMap* amap = &p.tmap_a;
      if constexpr (CHAIN) {
        if (layer == 1) {
          amap = &p.tmap_a1;
          ch0 = p.in_coff1;
        }
      }
      if (elect_one()) {
        if constexpr (CHAIN) {
          if (layer == 1 && !(p.ch_dbg & 2)) {
            // the intermediate tiles of these images (halo included) must be complete: every
            // epilogue warp that stored a layer-0 tile of image n has counted itself on ch_count[n]
            if (o0.n < p.N)
              while (ld_acquire_gpu(p.ch_count + o0.n) < p.ch_need) __nanosleep(64);
            if (o1.n < p.N && o1.n != o0.n)
              while (ld_acquire_gpu(p.ch_count + o1.n) < p.ch_need) __nanosleep(64);
            fence_proxy_async_all();
          }
        }
        if (res_prefetch && nx_kb == 0 && (!CHAIN || layer == 1)) {
          // the residual tiles of this item will be read by the epilogue one to two items from
          // now: pull them into L2 so that those loads do not pay the HBM latency
          const int rc0 = CHAIN ? p.res1_coff : p.epi.res_coff + (nx_item - unit * n_tiles) * BN;
          if (o0.n < p.N) tma_prefetch_l2_4d(&p.tmap_r, rc0, o0.x0, o0.y0, o0.n);
          if (o1.n < p.N) tma_prefetch_l2_4d(&p.tmap_r, rc0, o1.x0, o1.y0, o1.n);
        }
        if constexpr (CHAIN) {
          // two activation stages cannot cover the HBM latency of the block input: pull the
          // layer-0 tiles of a later item into L2 now (layer 1 reads what was just written)
          const int pf_item = nx_item + p.ch_pf * istep;
          if (p.ch_pf > 0 && pf_item < num_items) {
            int pl = 0;
            const int pu = item_unit<true>(p, pf_item, pl);
            if (pl == 0) {
              const TileOrg q0 = tile_origin(p, pu * tpi);
              const TileOrg q1 = tile_origin(p, pu * tpi + 1);
              tma_prefetch_l2_5d(&p.tmap_a, in_coff, q0.x0 - 1, 0, q0.y0 - 1, q0.n);
              tma_prefetch_l2_5d(&p.tmap_a, in_coff, q1.x0 - 1, 0, q1.y0 - 1, q1.n);
            }
          }
        }
        const uint32_t fa = smem_u32(&full_a[as_]);
        if (dbg & 8) {
          if (leader) mbar_arrive(fa);
        } else {
          const uint32_t dst = smem_u32(a_ring + (size_t)as_ * a_stage_bytes);
          if (leader) mbar_expect_tx(fa, (uint32_t)(tpi * p.a_tx_bytes));
          const uint32_t fas = cg2 ? sig(&full_a[as_]) : fa;
          if (p.s2) {
            // the four pixel-parity planes of the (2*16+1) x (2*8+1) input window (tpc == 1): plane
            // (rp, cp) holds input rows 2*j + rp, columns 2*i + cp; the odd planes start one
            // element earlier (input row / column -1 = padding, zero-filled by TMA)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int rp = q >> 1, cp = q & 1;
              const CUtensorMap* m = q == 0 ? &p.tmap_a : &p.tmap_a2[q - 1];
              const int c0 = cp * p.s2_ld + ch0;
              if (cg2) tma_load_5d_cg2(dst + (uint32_t)p.sub_off[q], m, fas, c0, o0.x0 - cp, rp, o0.y0 - rp, o0.n);
              else tma_load_5d(dst + (uint32_t)p.sub_off[q], m, fas, c0, o0.x0 - cp, rp, o0.y0 - rp, o0.n);
            }
          } else if (cg2) {
            tma_load_5d_cg2(dst, &p.tmap_a, fas, ch0, o0.x0 - 1, 0, o0.y0 - 1, o0.n);
            if (tpc == 2)
              tma_load_5d_cg2(dst + (uint32_t)p.a_tile_bytes, &p.tmap_a, fas, ch0, o1.x0 - 1, 0, o1.y0 - 1, o1.n);
          } else {
            tma_load_5d(dst, amap, fa, ch0, o0.x0 - 1, 0, o0.y0 - 1, o0.n);
            if (tpc == 2)
              tma_load_5d(dst + (uint32_t)p.a_tile_bytes, amap, fa, ch0, o1.x0 - 1, 0, o1.y0 - 1, o1.n);
          }
        }
      }
      __syncwarp();
      if (++as_ == a_stages) { as_ = 0; aph ^= 1u; }
      if (++nx_kb == num_kb) { nx_kb = 0; nx_item += istep; }
      return true;
    };
    load_a(true);                                   // step 0
    for (int item = item0; item < num_items; item += istep) {
      const int nt = item - (int)fdiv((uint32_t)item, p.fd_nt) * n_tiles;
      for (int kb = 0; kb < num_kb; ++kb) {
        // invariant: the activation load of this step has been issued
        bool next_done = load_a(false);
        if (!resident) {
#pragma unroll 1
          for (int g = 0; g < b_groups; ++g) {
            HL_TIMED(4, mbar_wait(smem_u32(&empty_b[bs_]), bph ^ 1u));
            if (elect_one()) {
              const uint32_t fb = smem_u32(&full_b[bs_]);
              if (dbg & 4) {
                if (leader) mbar_arrive(fb);
              } else {
                const uint32_t dst = smem_u32(b_ring + (size_t)bs_ * b_stage_bytes);
                if (leader) mbar_expect_tx(fb, stage_b_bytes);
                if (cg2) tma_load_3d_cg2(dst, &p.tmap_b, sig(&full_b[bs_]), kb * 64,
                                         nt * BN + brow0, g * tps);
                else tma_load_3d(dst, &p.tmap_b, fb, kb * 64, nt * BN, g * tps);
              }
            }
            __syncwarp();
            if (++bs_ == b_stages) { bs_ = 0; bph ^= 1u; }
            if (!next_done) next_done = load_a(false);
          }
        }
        if (!next_done) load_a(true);
      }
    }
    if (PROF && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * HL_PROF_SLOTS;
      o[2] = clock64() - t_loop; o[3] = pc[3]; o[4] = pc[4];
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers (leader CTA only; one elected lane issues) ===========
    // tcgen05.mma issue is nearly synchronous (the queue behind the issuing thread holds one or two
    // instructions), so every barrier wait / commit between two work items is lost tensor-pipe
    // time: ~220 cycles per item for two satisfied mbarrier waits alone (tools/exp_umma_seq.cu,
    // profiles/r02_mma_issue.md).  In dual mode warp 1 and warp 10 therefore take ALTERNATE work
    // items -- item it uses accumulator it & 1, so each warp owns one accumulator -- and one
    // warp's hand-offs overlap the other's MMAs.  Ring positions are functions of the item
    // ordinal alone, so each warp steps its counters over the other warp's item.
    const bool dual = p.dual != 0;
    const int mw = (warp == 10) ? 1 : 0;               // MMA warp index
    if (leader && (mw == 0 || dual)) {
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      const int it_step = dual ? 2 : 1;
      // Step the ring positions over one (other warp's) item.  A parity wait can only tell the
      // NEXT phase of a barrier from the current one, so this warp must observe every fill of every
      // stage in order -- also the fills the other warp consumes: it waits for them here, off the
      // critical path (the other warp is issuing MMAs meanwhile).
      auto skip_item = [&]() {
        for (int k = 0; k < num_kb; ++k) {
          mbar_wait(smem_u32(&full_a[as_]), aph);
          if (++as_ == a_stages) { as_ = 0; aph ^= 1u; }
          if (!resident)
            for (int g = 0; g < b_groups; ++g) {
              mbar_wait(smem_u32(&full_b[bs_]), bph);
              if (++bs_ == b_stages) { bs_ = 0; bph ^= 1u; }
            }
        }
      };
      if (mw == 1) {
        it = 1;
        if (item0 + istep < num_items) skip_item();    // item 0 is the other warp's
      }
      constexpr uint32_t B_HI = desc_hi(1024);
      const uint32_t a_ring_lo = desc_lo(smem_u32(a_ring));
      const uint32_t b_ring_lo = desc_lo(smem_u32(b_ring));
      const uint32_t a_stage_lo = (uint32_t)(a_stage_bytes >> 4);
      const uint32_t a_tile_lo = (uint32_t)(p.a_tile_bytes >> 4);
      const uint32_t b_stage_lo = (uint32_t)(b_stage_bytes >> 4);
      const uint32_t b_tap_lo = (uint32_t)((BNh * 128) >> 4);
      const bool no_mma = (dbg & 1) != 0;
      constexpr uint16_t PAIR = 0x3;
      if (resident) {
        HL_TIMED(7, mbar_wait(smem_u32(&full_b[0]), 0));
      }
      for (int item = item0 + it * istep; item < num_items; item += it_step * istep, it += it_step) {
        const int acc = (acc_stages == 2) ? (it & 1) : 0;
        const uint32_t accph = (acc_stages == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
        if (cg2) {
          HL_TIMED(8, mbar_wait_cluster(smem_u32(&tempty[acc]), accph ^ 1u));
        } else {
          HL_TIMED(8, mbar_wait(smem_u32(&tempty[acc]), accph ^ 1u));
        }
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * tpc * BN);
        const uint32_t d1 = d0 + (uint32_t)BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          HL_TIMED(6, mbar_wait(smem_u32(&full_a[as_]), aph));
          const uint32_t a0 = a_ring_lo + (uint32_t)as_ * a_stage_lo;
          const uint32_t a1 = a0 + a_tile_lo;
          const int seg = (kb >= nkb_seg) + (kb >= 2 * nkb_seg);
          const int k16 = (kb - seg * nkb_seg == nkb_seg - 1) ? last_k16 : 4;
#pragma unroll 1
          for (int g = 0; g < b_groups; ++g) {
            uint32_t b_lo;
            if (resident) {
              b_lo = b_ring_lo + (uint32_t)(kb * 9) * b_tap_lo;
              if constexpr (CHAIN) {
                int layer = 0;
                item_unit<true>(p, item, layer);
                b_lo += (uint32_t)(layer * 9) * b_tap_lo;
              }
            } else {
              HL_TIMED(7, mbar_wait(smem_u32(&full_b[bs_]), bph));
              b_lo = b_ring_lo + (uint32_t)bs_ * b_stage_lo;
            }
            if (elect_one()) {
              if (!no_mma) {
                // K16 is a template parameter: a run-time `if (k < k16)` between the MMAs of a tap
                // costs ~10 cycles per MMA for K = 48 (3 x K16: 57.7 instead of 48.5 cycles per MMA
                // at N = 48, 73.1 instead of 59.8 at N = 96; tools/exp_umma_seq.cu, profiles/r02_mma_issue.md)
                const uint32_t first = (uint32_t)(kb | g);
                switch (k16) {
                  case 4: halo_issue_taps<4, cg2>(p, d0, d1, a0, a1, b_lo, b_tap_lo, B_HI, idesc, g * tps, tps, tpc, first); break;
                  case 3: halo_issue_taps<3, cg2>(p, d0, d1, a0, a1, b_lo, b_tap_lo, B_HI, idesc, g * tps, tps, tpc, first); break;
                  case 2: halo_issue_taps<2, cg2>(p, d0, d1, a0, a1, b_lo, b_tap_lo, B_HI, idesc, g * tps, tps, tpc, first); break;
                  default: halo_issue_taps<1, cg2>(p, d0, d1, a0, a1, b_lo, b_tap_lo, B_HI, idesc, g * tps, tps, tpc, first); break;
                }
              }
              if (!resident) {
                if (cg2) umma_commit_cg2(smem_u32(&empty_b[bs_]), PAIR);
                else umma_commit(smem_u32(&empty_b[bs_]));
              }
            }
            __syncwarp();
            if (!resident) {
              if (++bs_ == b_stages) { bs_ = 0; bph ^= 1u; }
            }
          }
          if (elect_one()) {
            if (cg2) umma_commit_cg2(smem_u32(&empty_a[as_]), PAIR);
            else umma_commit(smem_u32(&empty_a[as_]));
          }
          __syncwarp();
          if (++as_ == a_stages) { as_ = 0; aph ^= 1u; }
        }
        if (elect_one()) {
          if (cg2) umma_commit_cg2(smem_u32(&tfull[acc]), PAIR);
          else umma_commit(smem_u32(&tfull[acc]));
        }
        __syncwarp();
        if (dual && item + it_step * istep < num_items) skip_item();   // over the other warp's item
      }
      if (PROF && lane == 0 && mw == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * HL_PROF_SLOTS;
        o[5] = clock64() - t_loop; o[6] = pc[6]; o[7] = pc[7]; o[8] = pc[8];
      }
    }
  } else {
    // ===================== epilogue: group g drains tile g of this CTA =======================
    const int group = (warp - 2) >> 2;
    const int lg = warp & 3;                       // TMEM lane quarter this warp may read
    const EpiParams& e = p.epi;
    // the accumulator-free barrier of the pair lives in the leader CTA
    const uint32_t tempty_addr = cg2 ? mapa_u32(smem_u32(&tempty[0]), 0u) : smem_u32(&tempty[0]);
    const uint32_t stage_slab = smem_u32(stage_out) + (uint32_t)((warp - 2) * p.out_slabs * HL_STAGE_OUT);
    pdl_wait();                                     // residual loads and output stores
    if (dbg & 2) {
      int it = 0;
      for (int item = item0; item < num_items; item += istep, ++it) {
        const int acc = (acc_stages == 2) ? (it & 1) : 0;
        const uint32_t accph = (acc_stages == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
        mbar_wait(smem_u32(&tfull[acc]), accph);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cg2) mbar_arrive_cluster(tempty_addr + 8u * acc);
          else mbar_arrive(tempty_addr + 8u * acc);
        }
      }
    } else if (SPLIT) {
      if (e.res != nullptr) {
        if (e.relu) halo_epilogue_split<true, true>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep);
        else halo_epilogue_split<true, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep);
      } else {
        if (e.relu) halo_epilogue_split<false, true>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep);
        else halo_epilogue_split<false, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep);
      }
    } else if (STAGED) {
      halo_epilogue_fast<true, true, false, false, true>(
          p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep,
          smem_u32(stage_out) + (uint32_t)((warp - 2) * p.res_slab_bytes), PROF, pc,
          smem_u32(&full_r[warp - 2]), smem_u32(&empty_r[warp - 2]));
    } else if (CHAIN) {
      halo_epilogue_fast<true, true, false, true>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
    } else if (EPI >= 0) {
      halo_epilogue_fast<((EPI >> 1) & 1) != 0, (EPI & 1) != 0, (EPI >= 4)>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
    } else if (e.fast) {
      if (e.res != nullptr) {
        if (e.relu) halo_epilogue_fast<true, true, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
        else halo_epilogue_fast<true, false, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
      } else {
        if (e.relu) halo_epilogue_fast<false, true, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
        else halo_epilogue_fast<false, false, false>(p, bias_s, tmem_base, tfull, tempty_addr, group, lg, lane, rank, item0, istep, stage_slab, PROF, pc);
      }
    } else if (EPI == -1) {
      // general epilogue (ragged channel counts): single-CTA mode only (host guarantees cg == 1)
      const int m = lg * 32 + lane;
      const int nchunks = BN >> 4;
      int it = 0;
      for (int item = item0; item < num_items; item += istep, ++it) {
        const int unit = (int)fdiv((uint32_t)item, p.fd_nt);
        const int nt = item - unit * n_tiles;
        const EpiPix px = epi_pixel(p, item, group, m, rank);
        const int acc = (acc_stages == 2) ? (it & 1) : 0;
        const uint32_t accph = (acc_stages == 2) ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) +
                                (uint32_t)(acc * tpc * BN + group * BN);
        epi_drain(e, bias_s, t_addr, nchunks, nt * BN, px.valid, px.opix, smem_u32(&tfull[acc]),
                  accph, smem_u32(&tempty[acc]), lane);
      }
    }
    if (PROF && warp == 2 && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * HL_PROF_SLOTS;
      o[9] = clock64() - t_loop; o[10] = pc[10]; o[11] = pc[11];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cg2) cluster_sync_all();      // nobody leaves while the pair's MMAs / commits may still touch it
  if (warp == 1) {
    tc_fence_after();
    if (cg2) tmem_dealloc_cg2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if constexpr (CHAIN) {
    // the last CTA to finish zeroes the image counters for the next launch (every other CTA is past
    // its last counter read: it has counted itself as done after its final CTA-wide sync)
    if (threadIdx.x == 0) {
      __threadfence();
      const int done = atomicAdd(p.ch_count + p.N, 1);
      if (done == (int)gridDim.x - 1) {
        for (int n = 0; n <= p.N; ++n) p.ch_count[n] = 0;
        __threadfence();
      }
    }
  }
  if (PROF && threadIdx.x == 0) {
    long long* o = p.prof + (size_t)blockIdx.x * HL_PROF_SLOTS;
    o[0] = clock64() - t_start;
    o[1] = t_loop - t_start;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 halo_encode_fn() {
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

// cycles of one M=128, K=16 MMA with both operands in shared memory (measured:
// profiles/r01_exp_umma_shift.md): max(N/2, 32 + N/4)
static inline double mma_cycles(int bn) { return std::max(bn / 2.0, 32.0 + bn / 4.0); }

// Cout tiling: the fewest tiles that fit (BN <= 256), unless a finer split fills the SMs
// better (small maps: few pixel-tile pairs) -- estimated as rounds * per-item MMA time.
static void halo_n_tiling(int cout_store, int pairs, int workers, int* n_tiles, int* bn) {
  const int cp = (cout_store + 15) / 16 * 16;
  const int sms = workers;
  int best_nt = 0, best_bn = 0;
  double best = 0;
  const int nt0 = (cp + HL_MAX_BN - 1) / HL_MAX_BN;
  // tilings that cover the channels exactly (n_tiles * BN == Cout) keep the fast epilogue, which
  // the one-tile and the CTA-pair modes require: when one exists, only those compete
  bool have_exact = false;
  for (int nt = nt0; nt <= nt0 + 3; ++nt) {
    const int b = ((cp + nt - 1) / nt + 15) / 16 * 16;
    if (nt > nt0 && b < 96) break;
    if (nt * b == cp) have_exact = true;
  }
  for (int nt = nt0; nt <= nt0 + 3; ++nt) {
    const int b = ((cp + nt - 1) / nt + 15) / 16 * 16;
    if (nt > nt0 && b < 96) break;
    if (have_exact && nt * b != cp) continue;
    const long long items = (long long)pairs * nt;
    const double rounds = (double)((items + sms - 1) / sms);
    const double cost = rounds * mma_cycles(b) + 0.05 * nt;   // tie -> fewer tiles
    if (best_nt == 0 || cost < best) {
      best = cost; best_nt = nt; best_bn = b;
    }
  }
  *n_tiles = best_nt;
  *bn = best_bn;
}

static bool HL_NAME(g_halo_attr_set) = false;

// kernel instantiations: 0..3 fast epilogue with (RES, RELU) = (v >> 1, v & 1); 4 run-time epilogue
// choice (general path); 5 split (BRTPE_DT_BF16X2); 6 with the debug cycle counters; 7..10 = 0..3 with
// the HRNet fuse addends
constexpr int HL_NUM_VARIANTS = (HL_CG == 1) ? 13 : 11;     // 11: chain mode, 12: staged residual (single-CTA kernel only)
#define HL_FOR_VARIANT(v, CALL)                                                   \
  switch (v) {                                                                    \
    case 0: CALL((HL_NAME(conv_halo_kernel)<false, false, 0>)); break;            \
    case 1: CALL((HL_NAME(conv_halo_kernel)<false, false, 1>)); break;            \
    case 2: CALL((HL_NAME(conv_halo_kernel)<false, false, 2>)); break;            \
    case 3: CALL((HL_NAME(conv_halo_kernel)<false, false, 3>)); break;            \
    case 4: CALL((HL_NAME(conv_halo_kernel)<false, false, -1>)); break;           \
    case 5: CALL((HL_NAME(conv_halo_kernel)<true, false, -2>)); break;            \
    case 7: CALL((HL_NAME(conv_halo_kernel)<false, false, 4>)); break;            \
    case 8: CALL((HL_NAME(conv_halo_kernel)<false, false, 5>)); break;            \
    case 9: CALL((HL_NAME(conv_halo_kernel)<false, false, 6>)); break;            \
    case 10: CALL((HL_NAME(conv_halo_kernel)<false, false, 7>)); break;           \
    case 11: CALL((HL_NAME(conv_halo_kernel)<false, false, (HL_CG == 1 ? 8 : 7)>)); break; \
    case 12: CALL((HL_NAME(conv_halo_kernel)<false, false, (HL_CG == 1 ? 9 : 7)>)); break; \
    default: CALL((HL_NAME(conv_halo_kernel)<false, true, -2>)); break;           \
  }

// Output tensor map of the epilogue's TMA stores: (C = out_coff + Cout, W, H, N) with the pixel
// stride of the output buffer; box = 32 channels x 8 px x 4 rows (one epilogue warp's slab),
// SWIZZLE_64B.  Stores beyond C / W / H / N are clipped by the hardware: that is how ragged
// tiles, the 16-channel tail of odd channel counts and tiles past the end of the batch are
// handled.
// Residual tensor map (L2 prefetch only): (C = res_coff + Cout, W, H, N), box BN ch x 8 x 16.
static bool halo_encode_res(const brtpe_conv_desc* d, const void* res, int bn, CUtensorMap* map) {
  auto encode = halo_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->res_ld * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)(d->res_coff + d->Cout), (cuuint64_t)d->Wout,
                        (cuuint64_t)d->Hout, (cuuint64_t)d->N};
  cuuint64_t gstr[3] = {ld_b, ld_b * d->Wout, ld_b * d->Wout * d->Hout};
  cuuint32_t box[4] = {(cuuint32_t)bn, (cuuint32_t)HL_TW, (cuuint32_t)HL_TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(res), gdim, gstr,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(residual) failed with %d", (int)r);
    return false;
  }
  return true;
}

// Residual tensor map of the staged epilogue: box BN ch x 8 px x 4 rows = the 32 pixels of one epilogue warp
static bool halo_encode_res4(const brtpe_conv_desc* d, const void* res, int bn, CUtensorMap* map) {
  auto encode = halo_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->res_ld * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)(d->res_coff + d->Cout), (cuuint64_t)d->Wout,
                        (cuuint64_t)d->Hout, (cuuint64_t)d->N};
  cuuint64_t gstr[3] = {ld_b, ld_b * d->Wout, ld_b * d->Wout * d->Hout};
  cuuint32_t box[4] = {(cuuint32_t)bn, (cuuint32_t)HL_TW, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(res), gdim, gstr,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(residual slab) failed with %d", (int)r);
    return false;
  }
  return true;
}

static bool halo_encode_out(const brtpe_conv_desc* d, void* out, CUtensorMap* map) {
  auto encode = halo_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->out_ld * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)(d->out_coff + d->Cout), (cuuint64_t)d->Wout,
                        (cuuint64_t)d->Hout, (cuuint64_t)d->N};
  cuuint64_t gstr[3] = {ld_b, ld_b * d->Wout, ld_b * d->Wout * d->Hout};
  cuuint32_t box[4] = {32, (cuuint32_t)HL_TW, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    return false;
  }
  return true;
}

HaloConvPrepared* HL_NAME(halo_conv_prepare)(const brtpe_conv_desc* d, const void* in,
                                             const void* weights, void* out) {
  HaloConvPrepared* P = new HaloConvPrepared();
  P->d = *d;
  P->out_encoded = nullptr;
  P->res_encoded = nullptr;
  HaloParams& p = P->p;
  memset(&p, 0, sizeof(p));
  {
    static int rev = -1;                     // BRTPE_HALO_REVERSE=0 ignores reverse_order (A/B switch)
    if (rev < 0) rev = getenv("BRTPE_HALO_REVERSE") ? atoi(getenv("BRTPE_HALO_REVERSE")) : 1;
    p.reverse = (rev && d->reverse_order) ? 1 : 0;
  }
  p.s2 = d->in_stride == 2 ? 1 : 0;
  p.s2_ld = d->in_ld;
  p.N = d->N; p.H = d->Hm; p.W = d->Wm;             // tiles walk the OUTPUT pixels (== input for stride 1)
  p.tiles_x = ceil_div(p.W, HL_TW);
  p.tiles_y = ceil_div(p.H, HL_TH);
  p.m_tiles = p.tiles_x * p.tiles_y * p.N;
  // CTA-pair mode (tcgen05 cta_group::2): needs the chunk-aligned fast epilogue and an even
  // number of SMs; BRTPE_HALO_CG=1 forces the single-CTA kernel
  p.cg = HL_CG;
  if (p.cg == 2 && (!epi_fast_ok(d) || (num_sms() & 1) || p.m_tiles < 4)) {
    delete P;
    return nullptr;
  }
  // one tile per CTA when two accumulators of BN columns could not be double-buffered in the 512
  // TMEM columns (BN > 128): the epilogue of item i then overlaps the MMAs of item i+1 again and
  // the work items are finer (better balance over the SMs); needs the fast epilogue
  const int cp0 = (d->Cout_store + 15) / 16 * 16;
  // (whole 64-channel rounds per epilogue group only: 192 and 256 channels, the widths this mode
  // was built and verified for; 160 / 224 channels gave wrong results / no valid tiling)
  p.tpc = (cp0 > 128 && cp0 <= HL_MAX_BN && epi_fast_ok(d) && cp0 % 64 == 0) ? 1 : 2;
  if (getenv("BRTPE_HALO_TPC")) {
    const int v = atoi(getenv("BRTPE_HALO_TPC"));
    if (v == 2 || (v == 1 && epi_fast_ok(d) && (cp0 / 16) % 2 == 0)) p.tpc = v;
  }
  if (p.s2) {
    // the input window of a stride-2 tile is 33 x 17 pixels (72 KB per 64-channel block): one tile
    // per CTA and item; needs the fast epilogue (the two epilogue groups share the one tile)
    if (!epi_fast_ok(d) || cp0 > HL_MAX_BN) {
      delete P;
      set_error("halo conv (stride 2): needs Cout == Cout_store <= %d, a multiple of 16", HL_MAX_BN);
      return nullptr;
    }
    p.tpc = 1;
  }
  const int tpi = p.tpc * p.cg;
  const int workers = num_sms() / p.cg;               // CTAs (cg 1) or CTA pairs (cg 2)
  const int pairs = ceil_div(p.m_tiles, tpi);         // work units of 2*cg pixel tiles
  p.num_units = pairs;
  halo_n_tiling(d->Cout_store, pairs, workers, &p.n_tiles, &p.BN);
  if (getenv("BRTPE_HALO_NT")) {
    const int nt = atoi(getenv("BRTPE_HALO_NT"));
    const int cp = (d->Cout_store + 15) / 16 * 16;
    const int b = ((cp + nt - 1) / std::max(nt, 1) + 15) / 16 * 16;
    if (nt >= 1 && b <= HL_MAX_BN && b >= 16) { p.n_tiles = nt; p.BN = b; }
  }
  p.num_items = pairs * p.n_tiles;

  const bool split = conv_is_split(d);
  p.nkb_seg = ceil_div(d->Cin, 64);
  p.num_kb = split ? 3 * p.nkb_seg : p.nkb_seg;
  p.last_k16 = (d->Cin - (p.nkb_seg - 1) * 64) / 16;
  p.lo_off = d->in_ld / 2;
  p.in_coff = d->in_coff;
  p.dbg = getenv("BRTPE_HALO_DBG") ? atoi(getenv("BRTPE_HALO_DBG")) : 0;
  p.acc_stages = (2 * p.tpc * p.BN <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * p.tpc * p.BN) cols *= 2;
  p.tmem_cols = cols;
  // two MMA-issuing warps need one accumulator each; BRTPE_HALO_DUAL=0 keeps a single issuer
  p.dual = (p.acc_stages == 2) ? 1 : 0;
  if (getenv("BRTPE_HALO_DUAL")) p.dual = (atoi(getenv("BRTPE_HALO_DUAL")) != 0 && p.acc_stages == 2) ? 1 : 0;

  // shared-memory plan: [A ring][B ring or resident weights][tail]; in pair mode every CTA
  // holds half of the weight rows
  const int tap_bytes = (p.BN / p.cg) * 128;
  // The 48/64-channel layers are bound by shared-memory bandwidth (MMA operand reads); staging
  // their output through shared memory costs more than it saves there: direct 32-byte stores.
  p.tma_out = (p.BN >= 96) ? 1 : 0;
  if (getenv("BRTPE_HALO_TMA_OUT")) p.tma_out = atoi(getenv("BRTPE_HALO_TMA_OUT")) ? 1 : 0;
  if (split) p.tma_out = 0;                         // the split epilogue stores directly
  // two output slabs per epilogue warp unless that would push resident weights out / leave the
  // weight ring with fewer than two stages
  // activation tile in shared memory: stride 1 = the 18 x 10 halo tile; stride 2 = four parity planes
  // (even/odd rows x even/odd columns) of 16|17 rows x 8|9 pixels, each 1024-aligned
  if (p.s2) {
    const int rows[2] = {HL_TH, HL_TH + 1}, cols[2] = {HL_TW, HL_TW + 1};
    int off = 0;
    p.a_tx_bytes = 0;
    for (int q = 0; q < 4; ++q) {
      p.sub_off[q] = off;
      const int bytes = rows[q >> 1] * cols[q & 1] * 128;
      p.a_tx_bytes += bytes;
      off += (int)align_up((size_t)bytes, 1024);
    }
    p.a_tile_bytes = off;
    for (int t = 0; t < 9; ++t) {
      const int kh = t / 3, kw = t % 3;
      const int q = ((kh != 1) ? 2 : 0) + ((kw != 1) ? 1 : 0);
      const int pitch = cols[q & 1];
      p.tap_aoff[t] = (uint32_t)((p.sub_off[q] + (((kh == 2) ? pitch : 0) + ((kw == 2) ? 1 : 0)) * 128) >> 4);
      p.tap_ahi[t] = desc_hi((uint32_t)(pitch * 128));
    }
  } else {
    p.a_tile_bytes = HL_A_TILE;
    p.a_tx_bytes = HL_A_BYTES;
    for (int t = 0; t < 9; ++t) {
      p.tap_aoff[t] = (uint32_t)((((t / 3) * HL_PITCH + t % 3) * 128) >> 4);
      p.tap_ahi[t] = desc_hi(HL_PITCH * 128);
    }
  }
  p.a_stage_bytes = p.tpc * p.a_tile_bytes;
  const int A2 = 2 * p.a_stage_bytes;                 // the minimum: two activation stages
  p.out_slabs = 2;
  {
    const int b2 = HL_SMEM_MAX - HL_TAIL - 1024 - 2 * HL_STAGE_BYTES;
    const int tb = (p.BN / p.cg) * 128;
    const bool res1 = p.n_tiles == 1 && p.num_kb * 9 * tb + A2 <= b2 + HL_STAGE_BYTES;
    const bool res2 = p.n_tiles == 1 && p.num_kb * 9 * tb + A2 <= b2;
    if ((res1 && !res2) || (!res2 && 2 * tb > b2 - A2)) p.out_slabs = 1;
  }
  if (!p.tma_out) p.out_slabs = 0;
  const int budget = HL_SMEM_MAX - HL_TAIL - 1024 - p.out_slabs * HL_STAGE_BYTES;
  const int resident_bytes = p.num_kb * 9 * tap_bytes;
  p.resident = (p.n_tiles == 1 && resident_bytes + A2 <= budget) ? 1 : 0;
  if (getenv("BRTPE_HALO_NO_RESIDENT")) p.resident = 0;
  // activation stages wanted: with two MMA-issuing warps two items (num_kb stages each) are in flight
  // and the producer should be one item ahead; two stages is the minimum
  const int a_want = std::min(HL_MAX_A, std::max(3, 2 * p.num_kb + p.num_kb));
  const int a_env = getenv("BRTPE_HALO_A_STAGES") ? atoi(getenv("BRTPE_HALO_A_STAGES")) : 0;
  if (p.resident) {
    p.tps = 9; p.b_groups = 1; p.b_stages = 1;
    p.b_stage_bytes = (int)align_up((size_t)resident_bytes, 1024);
    p.a_stages = std::min(a_want, (budget - p.b_stage_bytes) / p.a_stage_bytes);
    if (a_env) p.a_stages = std::max(2, std::min((budget - p.b_stage_bytes) / p.a_stage_bytes, std::min(HL_MAX_A, a_env)));
  } else {
    // the largest tap group whose stage fits twice next to the minimum of activation stages
    int tps = 9;
    if (2 * 9 * tap_bytes > budget - A2) tps = 3;
    if (tps == 3 && 2 * 3 * tap_bytes > budget - A2) tps = 1;
    if (getenv("BRTPE_HALO_TPS")) {
      const int t = atoi(getenv("BRTPE_HALO_TPS"));
      if ((t == 1 || t == 3 || t == 9) && 2 * t * tap_bytes <= budget - A2) tps = t;
    }
    p.tps = tps;
    p.b_groups = 9 / tps;
    p.b_stage_bytes = (int)align_up((size_t)tps * tap_bytes, 1024);
    if (2 * p.b_stage_bytes > budget - A2) {
      set_error("halo conv: weight stage of %d bytes does not fit twice", p.b_stage_bytes);
      delete P;
      return nullptr;
    }
    // Two stages of each ring are the minimum; what is left is dealt out in the order that was
    // measured best for the weight-streaming layers (192 -> 192 on 40 x 40, 64 images: 2 + 2 stages
    // 0.0643 ms, 2 A + 3 B 0.0612, 3 + 3 0.0606, 5 A + 2 B 0.0673): a third weight stage, a third
    // activation stage, a fourth weight stage, then the rest of the activation stages wanted.
    p.a_stages = 2;
    p.b_stages = 2;
    int rem = budget - 2 * p.a_stage_bytes - 2 * p.b_stage_bytes;
    auto more_b = [&](int upto) { while (p.b_stages < upto && rem >= p.b_stage_bytes) { ++p.b_stages; rem -= p.b_stage_bytes; } };
    auto more_a = [&](int upto) { while (p.a_stages < upto && rem >= p.a_stage_bytes) { ++p.a_stages; rem -= p.a_stage_bytes; } };
    if (a_env) {
      more_a(std::min(HL_MAX_A, a_env));
      more_b(HL_MAX_B);
    } else {
      more_b(3);
      more_a(3);
      more_b(4);
      more_a(a_want);
      more_b(HL_MAX_B);
    }
  }
  // Two issuing warps observe each other's activation fills through parity waits, which can only tell a
  // barrier's next phase from the current one.  When one work item takes more fills than the ring has
  // stages (split 3x3 / stride-2 layers: 3 channel blocks, two 72 KB stages), two fills of the OTHER warp's
  // item land in the same stage and a warp that is still issuing its own item can miss one: it then waits for
  // a fill that needs its own progress.  Seen as an intermittent launch failure (spin-limit trap) of the
  // fp32 mode, 48 -> 48 stride 2, on some boxes only, ~1 in 300 launches inside the network and never
  // stand-alone; gone with one issuer (tools/stress_fp32.py, profiles/r02_fp32_hang.md).
  if (p.dual && p.a_stages < p.num_kb) p.dual = 0;
  P->smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes +
            (size_t)p.out_slabs * HL_STAGE_BYTES + HL_TAIL + 1024;
  // Staged residual (opt-in, BRTPE_HALO_RES_STAGED=1): narrow single-CTA layers with direct stores whose
  // shared memory still has room for eight slabs of 32 pixels x BN channels (48 channels: 3 activation
  // stages + resident weights + 24 KB).  Measured neutral (48 -> 48 + residual, 64 images: 0.0964 vs
  // 0.0963-0.0977 ms): conv2 of a 48-channel block is bound by its 472 MB of HBM traffic (76 % of the copy
  // bandwidth), not by the latency of its residual loads (profiles/r02_chain.md).
  p.res_staged = 0;
  p.res_slab_bytes = (int)align_up((size_t)32 * p.BN * 2, 128);
  {
    static int rs = -1;
    if (rs < 0) rs = getenv("BRTPE_HALO_RES_STAGED") ? atoi(getenv("BRTPE_HALO_RES_STAGED")) : 0;
    const bool fast_exact = epi_fast_ok(d) && p.n_tiles * p.BN == d->Cout;
    if (rs && HL_CG == 1 && p.tpc == 2 && p.n_tiles == 1 && p.BN <= 64 && !p.tma_out && !split && !p.s2 &&
        fast_exact && d->n_add == 0 && d->out2_ld == 0 && d->relu && d->res_ld > 0 && p.dbg == 0 &&
        (d->res_ld * 2) % 16 == 0 && (d->res_coff * 2) % 16 == 0 && (p.BN * 2) % 16 == 0 &&
        P->smem + 8 * (size_t)p.res_slab_bytes <= (size_t)HL_SMEM_MAX) {
      p.res_staged = 1;
      P->smem += 8 * (size_t)p.res_slab_bytes;
    }
  }
  P->grid = p.cg * std::max(1, std::min(p.num_items, workers));

  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)((128 * p.cg) >> 4) << 24);
  p.epi.out = nullptr; p.epi.res = nullptr;
  p.epi.out_ld = d->out_ld; p.epi.out_coff = d->out_coff; p.epi.res_ld = d->res_ld;
  p.epi.res_coff = d->res_coff; p.epi.Cout = d->Cout; p.epi.Cout_store = d->Cout_store;
  p.epi.relu = d->relu; p.epi.vec32 = epi_vec32_ok(d);
  // the fast epilogue walks whole 16-channel chunks of real channels
  p.epi.fast = (epi_fast_ok(d) && p.n_tiles * p.BN == d->Cout) ? 1 : 0;
  epi_set_split(&p.epi, d);
  epi_set_add(&p.epi, d);
  if (d->n_add > 0 || d->out2_ld > 0) {
    // two outputs go through the direct-store epilogue (Cout tiles <= 64 channels)
    if (!epi_add_ok(d) || !p.epi.fast || split || d->n_add == 0 || (d->out2_ld > 0 && p.tma_out)) {
      set_error("halo conv: fuse addends need the fast bf16 epilogue (two outputs: Cout tile <= 64)");
      delete P;
      return nullptr;
    }
  }
  if (split && !p.epi.fast) {
    set_error("halo conv: split layers need Cout = n_tiles * BN (Cout %d, %d x %d)", d->Cout, p.n_tiles, p.BN);
    delete P;
    return nullptr;
  }
  if (p.tpc == 1 && !p.epi.fast) {
    set_error("halo conv: one-tile mode needs Cout = n_tiles * BN (Cout %d, %d x %d)", d->Cout, p.n_tiles, p.BN);
    delete P;
    return nullptr;
  }
  if (p.cg == 2 && !p.epi.fast) {
    set_error("halo conv: pair mode needs Cout = n_tiles * BN (Cout %d, %d x %d)", d->Cout, p.n_tiles, p.BN);
    delete P;
    return nullptr;
  }
  p.fd_nt = make_fastdiv((uint32_t)p.n_tiles, (uint64_t)p.num_items + 2ull * num_sms() + 2);
  p.fd_xy = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y), (uint64_t)p.m_tiles + 8ull * num_sms() + 8);
  p.fd_x = make_fastdiv((uint32_t)p.tiles_x, (uint64_t)p.tiles_x * p.tiles_y);

  auto encode = halo_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->in_ld * 2;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  if (!p.s2) {
    cuuint64_t gdim[5] = {(cuuint64_t)(split ? d->in_ld : d->in_coff + d->Cin), (cuuint64_t)d->Win, 1,
                          (cuuint64_t)d->Hin, (cuuint64_t)d->N};
    cuuint64_t gstr[4] = {ld_b, ld_b * d->Win, ld_b * d->Win, ld_b * d->Win * d->Hin};
    cuuint32_t box[5] = {64, (cuuint32_t)HL_PITCH, 1, (cuuint32_t)(HL_TH + 2), 1};
    r = encode(&p.tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), gdim, gstr, box,
               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // (channel of a column PAIR, W/2, row parity, H/2, N) view: coordinate (c + cp*ld, j, rp, i, n) is
    // input pixel (row 2i + rp, column 2j + cp); one map per parity plane (the box sizes differ)
    cuuint64_t gdim[5] = {(cuuint64_t)(split ? 2 * d->in_ld : d->in_ld + d->in_coff + d->Cin),
                          (cuuint64_t)(d->Win / 2), 2, (cuuint64_t)(d->Hin / 2), (cuuint64_t)d->N};
    cuuint64_t gstr[4] = {2 * ld_b, ld_b * d->Win, 2 * ld_b * d->Win, ld_b * d->Win * d->Hin};
    for (int q = 0; q < 4 && r == CUDA_SUCCESS; ++q) {
      cuuint32_t box[5] = {64, (cuuint32_t)(HL_TW + (q & 1)), 1, (cuuint32_t)(HL_TH + (q >> 1)), 1};
      r = encode(q == 0 ? &p.tmap_a : &p.tmap_a2[q - 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                 const_cast<void*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
  }
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    delete P;
    return nullptr;
  }
  // packed weights: bf16 [9][cout_pad][cin_pad] (brtpe_umma_weight_dims); rows beyond cout_pad
  // (when n_tiles*BN rounds up past it) are out of bounds and read as zeros
  int cin_pad = 0, cout_pad = 0;
  brtpe_umma_weight_dims(d->Cin, d->Cout_store, &cin_pad, &cout_pad);
  if (split) cin_pad *= 3;                          // [w_hi | w_hi | w_lo] segments
  cuuint64_t wdim[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, 9};
  cuuint64_t wstr[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * 2 * cout_pad};
  cuuint32_t wbox[3] = {64, (cuuint32_t)(p.BN / p.cg), (cuuint32_t)p.tps};
  cuuint32_t westr[3] = {1, 1, 1};
  r = encode(&p.tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(weights), wdim, wstr,
             wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    delete P;
    return nullptr;
  }
  if (out != nullptr && p.epi.fast) {
    if (!halo_encode_out(d, out, &p.tmap_o)) {
      delete P;
      return nullptr;
    }
    P->out_encoded = out;
  }
  if (!HL_NAME(g_halo_attr_set)) {
    for (int v = 0; v < HL_NUM_VARIANTS; ++v) {
      cudaError_t ae = cudaSuccess;
#define HL_SET_ATTR(K) ae = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_MAX)
      HL_FOR_VARIANT(v, HL_SET_ATTR)
#undef HL_SET_ATTR
      if (ae != cudaSuccess) {
        set_error("cudaFuncSetAttribute(HL_NAME(conv_halo_kernel)) failed");
        delete P;
        return nullptr;
      }
    }
    HL_NAME(g_halo_attr_set) = true;
  }
  return P;
}

int HL_NAME(halo_conv_launch)(const HaloConvPrepared* P, const float* bias, const void* residual, void* out,
                     cudaStream_t st, const void* const* add_ptrs, void* out2) {
  HaloParams p = P->p;
  p.bias = bias;
  p.epi.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.epi.out = reinterpret_cast<__nv_bfloat16*>(out);
  for (int k = 0; k < p.epi.n_add; ++k) {
    if (!add_ptrs || !add_ptrs[k]) {
      set_error("halo conv: fuse addend %d is null", k);
      return BRTPE_EINVAL;
    }
    p.epi.add[k] = reinterpret_cast<const __nv_bfloat16*>(add_ptrs[k]);
  }
  if (p.epi.out2_ld > 0) {
    if (!out2) {
      set_error("halo conv: second output is null");
      return BRTPE_EINVAL;
    }
    p.epi.out2 = reinterpret_cast<__nv_bfloat16*>(out2);
  }
  if (p.epi.fast && out != P->out_encoded) {       // output buffer changed since prepare
    if (!halo_encode_out(&P->d, out, &p.tmap_o)) return BRTPE_ECUDA;
  }
  p.res_prefetch = 0;
  // L2 prefetch of the residual tiles by the TMA producer.  With two MMA-issuing warps the narrow
  // layers wait for their epilogue, which waits for the residual (HBM latency, one tile of loads in
  // flight per warp): 48 channels 0.0970 -> 0.0933 ms with the prefetch, but 96 channels 0.0629 ->
  // 0.0675 (the extra TMA traffic delays the operand loads): on for Cout tiles <= 64 channels.
  // BRTPE_HALO_RES_PREFETCH=0 / 1 forces it off / on.
  static int res_pf = -1;
  if (res_pf < 0) {
    const char* e = getenv("BRTPE_HALO_RES_PREFETCH");
    res_pf = e ? (atoi(e) ? 1 : 0) : 2;
  }
  if (residual != nullptr && p.epi.fast && !p.epi.split && !p.s2 &&
      (res_pf == 1 || (res_pf == 2 && p.BN <= 64))) {
    HaloConvPrepared* PM = const_cast<HaloConvPrepared*>(P);     // cache of the encoded map
    if (residual != PM->res_encoded) {
      if (!halo_encode_res(&P->d, residual, p.BN, &PM->tmap_r_cache)) return BRTPE_ECUDA;
      PM->res_encoded = residual;
    }
    p.tmap_r = PM->tmap_r_cache;
    p.res_prefetch = 1;
  }
  const bool prof = g_halo_prof != nullptr && P->grid <= g_halo_prof_ctas;
  p.prof = prof ? g_halo_prof : nullptr;
  // staged residual: only the residual + ReLU instantiation exists, and not with the debug counters
  const bool staged = p.res_staged && residual != nullptr && p.epi.relu && p.epi.fast && p.prof == nullptr &&
                      p.epi.n_add == 0 && p.dbg == 0;
  p.res_staged = staged ? 1 : 0;
  if (staged) {
    HaloConvPrepared* PM = const_cast<HaloConvPrepared*>(P);
    if (residual != PM->res4_encoded) {
      if (!halo_encode_res4(&P->d, residual, p.BN, &PM->tmap_r4_cache)) return BRTPE_ECUDA;
      PM->res4_encoded = residual;
    }
    p.tmap_r4 = PM->tmap_r4_cache;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(P->grid);
  cfg.blockDim = dim3(HL_THREADS);
  cfg.dynamicSmemBytes = P->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  int variant;
  if (p.epi.split) variant = 5;
  else if (staged) variant = 12;
  else if (p.prof != nullptr && p.epi.fast) variant = 6;
  else if (p.epi.n_add > 0) variant = 7 + (p.epi.res != nullptr ? 2 : 0) + (p.epi.relu ? 1 : 0);
  else if (p.epi.fast && p.dbg == 0) variant = (p.epi.res != nullptr ? 2 : 0) + (p.epi.relu ? 1 : 0);
  else variant = 4;
  cudaError_t e = cudaSuccess;
#define HL_LAUNCH(K) e = cudaLaunchKernelEx(&cfg, K, p)
  HL_FOR_VARIANT(variant, HL_LAUNCH)
#undef HL_LAUNCH
  if (e != cudaSuccess) {
    set_error("HL_NAME(conv_halo_kernel) launch failed: %s", cudaGetErrorString(e));
    return BRTPE_ECUDA;
  }
  return BRTPE_OK;
}

#if HL_CG == 1
// ------------------------------------------------------------------------------------------
// chain mode: conv1 -> conv2 (+ residual) of a BasicBlock (pose_higher_hrnet.py:46-75) in one launch
// ------------------------------------------------------------------------------------------
// At the bench's 64 forwards per plan the 48-channel tensors (157 MB at 160^2, 629 MB at 320^2) do
// not fit the 126 MB L2, so the two launches of a block move five tensor passes through HBM (conv1:
// x in, t out; conv2: t + x in, y out) and conv2 is HBM-bound.  One launch that walks the batch in
// sub-batches, layer 1 one phase behind layer 0, finds t and x of a sub-batch still in L2: three
// passes (x in, t and y written back).  Both weight sets stay resident (2 x 54 KB at 48 channels),
// which leaves two activation stages.
bool halo_chain_eligible(const HaloConvPrepared* P0, const brtpe_conv_desc* d0, const brtpe_conv_desc* d1) {
  const HaloParams& p = P0->p;
  if (p.cg != 1 || p.tpc != 2 || p.n_tiles != 1 || !p.resident || p.num_kb != 1 || p.s2 || p.tma_out ||
      !p.epi.fast || p.epi.split || p.epi.n_add || d0->out2_ld > 0 || p.dbg)
    return false;
  if (!d0->relu || !d1->relu) return false;
  if (!halo_conv_supported(d1) || conv_is_split(d1) || d1->in_stride != 1 || d1->n_add || d1->out2_ld > 0 ||
      !epi_fast_ok(d1))
    return false;
  if (d1->N != d0->N || d1->Hm != d0->Hm || d1->Wm != d0->Wm || d1->Cin != d0->Cout ||
      d1->Cout != d0->Cout || d1->Cout_store != d0->Cout_store || d1->in_ld != d0->out_ld ||
      d1->in_coff != d0->out_coff || d1->Cin > 64)
    return false;
  return (p.m_tiles & 1) == 0;
}

HaloConvPrepared* halo_chain_prepare(const brtpe_conv_desc* d0, const void* in, const void* w0, void* mid,
                                     const brtpe_conv_desc* d1, const void* w1, void* out) {
  HaloConvPrepared* P = halo_conv_prepare_cg1(d0, in, w0, mid);
  if (!P) return nullptr;
  if (!halo_chain_eligible(P, d0, d1)) {
    delete P;
    set_error("halo chain: the two layers are not a chainable pair");
    return nullptr;
  }
  HaloParams& p = P->p;
  P->d1 = *d1;
  P->mid = mid;
  // shared memory: two resident weight sets + at least two activation stages
  const int tap_bytes = p.BN * 128;
  const int resident2 = 2 * 9 * tap_bytes;
  const int budget = HL_SMEM_MAX - HL_TAIL - 1024;
  p.b_stage_bytes = (int)align_up((size_t)resident2, 1024);
  p.a_stages = std::min(HL_MAX_A, (budget - p.b_stage_bytes) / p.a_stage_bytes);
  if (p.a_stages < 2 || 2 * p.BN * 4 + 64 > HL_TAIL - 512) {
    delete P;
    set_error("halo chain: two weight sets of %d bytes leave fewer than two activation stages", resident2 / 2);
    return nullptr;
  }
  P->smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stage_bytes + HL_TAIL + 1024;
  // images per sub-batch: the smallest divisor G of N whose phase (G images of one layer) is a whole
  // number of tile pairs and at least two rounds of the grid (a layer-1 item then depends on items
  // that were handed out >= two rounds earlier), within ~16 MB per tensor and sub-batch
  const int tiles_xy = p.tiles_x * p.tiles_y;
  const int workers = num_sms();
  const long long img_bytes = (long long)p.H * p.W * d0->out_ld * 2;
  int G = 0;
  for (int g = 1; g <= p.N; ++g) {
    if (p.N % g || ((g * tiles_xy) & 1)) continue;
    if (G && g * img_bytes > (16ll << 20)) break;
    G = g;
    if (g * tiles_xy / 2 >= 2 * workers) break;
  }
  if (getenv("BRTPE_CHAIN_G")) {
    const int g = atoi(getenv("BRTPE_CHAIN_G"));
    if (g >= 1 && p.N % g == 0 && ((g * tiles_xy) & 1) == 0) G = g;
  }
  if (!G) {
    delete P;
    set_error("halo chain: no sub-batch size divides the batch into whole tile pairs");
    return nullptr;
  }
  p.chain = 1;
  p.ch_dbg = getenv("BRTPE_CHAIN_DBG") ? atoi(getenv("BRTPE_CHAIN_DBG")) : 0;
  p.ch_ip = G * tiles_xy / 2;
  p.ch_defer = (p.ch_ip >= 2 * workers) ? 1 : 0;
  if (getenv("BRTPE_CHAIN_DEFER") && atoi(getenv("BRTPE_CHAIN_DEFER")) == 0) p.ch_defer = 0;
  p.ch_pf = getenv("BRTPE_CHAIN_PF") ? atoi(getenv("BRTPE_CHAIN_PF")) : 2;
  p.ch_S = p.N / G;
  p.fd_ip = make_fastdiv((uint32_t)p.ch_ip, 2ull * p.num_units + 4ull * workers + 4);
  p.num_items = 2 * p.ch_S * p.ch_ip;
  p.ch_need = tiles_xy * 4;
  P->grid = std::max(1, std::min(p.num_items, workers));
  p.in_coff1 = d1->in_coff;
  p.out1_ld = d1->out_ld; p.out1_coff = d1->out_coff;
  p.res1_ld = d1->res_ld; p.res1_coff = d1->res_coff;
  // tensor maps of layer 1: the intermediate tensor as activation operand, its weights
  auto encode = halo_encode_fn();
  {
    const cuuint64_t ld_b = (cuuint64_t)d1->in_ld * 2;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gdim[5] = {(cuuint64_t)(d1->in_coff + d1->Cin), (cuuint64_t)d1->Win, 1,
                          (cuuint64_t)d1->Hin, (cuuint64_t)d1->N};
    cuuint64_t gstr[4] = {ld_b, ld_b * d1->Win, ld_b * d1->Win, ld_b * d1->Win * d1->Hin};
    cuuint32_t box[5] = {64, (cuuint32_t)HL_PITCH, 1, (cuuint32_t)(HL_TH + 2), 1};
    CUresult r = encode(&p.tmap_a1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, mid, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int cin_pad = 0, cout_pad = 0;
    brtpe_umma_weight_dims(d1->Cin, d1->Cout_store, &cin_pad, &cout_pad);
    cuuint64_t wdim[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, 9};
    cuuint64_t wstr[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * 2 * cout_pad};
    cuuint32_t wbox[3] = {64, (cuuint32_t)p.BN, 9};
    cuuint32_t westr[3] = {1, 1, 1};
    if (r == CUDA_SUCCESS)
      r = encode(&p.tmap_b1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w1), wdim, wstr, wbox,
                 westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("halo chain: cuTensorMapEncodeTiled failed with %d", (int)r);
      delete P;
      return nullptr;
    }
  }
  (void)out;
  if (cudaMalloc(&P->counters, (size_t)(p.N + 1) * sizeof(int)) != cudaSuccess ||
      cudaMemset(P->counters, 0, (size_t)(p.N + 1) * sizeof(int)) != cudaSuccess) {
    set_error("halo chain: counter allocation failed");
    delete P;
    return nullptr;
  }
  p.ch_count = P->counters;
  return P;
}

int halo_chain_launch(const HaloConvPrepared* P, const float* bias0, const float* bias1,
                      const void* residual, void* out, cudaStream_t st) {
  HaloParams p = P->p;
  if (!p.chain || !residual || !out) {
    set_error("halo chain: not a prepared chain / null tensor");
    return BRTPE_EINVAL;
  }
  p.bias = bias0;
  p.bias1 = bias1;
  p.epi.res = nullptr;
  p.epi.out = reinterpret_cast<__nv_bfloat16*>(P->mid);
  p.res1 = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out1 = reinterpret_cast<__nv_bfloat16*>(out);
  p.res_prefetch = 0;
  static int res_pf = -1;
  if (res_pf < 0) {
    const char* e = getenv("BRTPE_HALO_RES_PREFETCH");
    res_pf = e ? (atoi(e) ? 1 : 0) : 0;       // the residual was read as layer 0's input one phase ago
  }
  if (res_pf) {
    HaloConvPrepared* PM = const_cast<HaloConvPrepared*>(P);
    if (residual != PM->res_encoded) {
      if (!halo_encode_res(&P->d1, residual, p.BN, &PM->tmap_r_cache)) return BRTPE_ECUDA;
      PM->res_encoded = residual;
    }
    p.tmap_r = PM->tmap_r_cache;
    p.res_prefetch = 1;
  }
  p.prof = nullptr;
  cudaError_t e = launch_ex(conv_halo_kernel_cg1<false, false, 8>, dim3(P->grid), dim3(HL_THREADS), P->smem,
                            st, 1, true, p);
  if (e != cudaSuccess) {
    set_error("conv_halo_kernel (chain) launch failed: %s", cudaGetErrorString(e));
    return BRTPE_ECUDA;
  }
  return BRTPE_OK;
}
#endif  // HL_CG == 1

}  // namespace brtpe
