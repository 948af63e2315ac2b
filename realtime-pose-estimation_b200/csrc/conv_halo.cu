// tcgen05 implicit-GEMM kernel for the 3x3 / stride-1 / pad-1 convolutions of HigherHRNet
// (BasicBlock conv1/conv2, Bottleneck conv2, transition1[0]: 87 % of the network's FLOPs,
// rtpe/third_party/pose_higher_hrnet.py:40-43, :46-75, :85-87, :558-563).
//
// Difference to conv_umma.cu (one TMA box per tap): the A operand of an 8x16-pixel output tile
// is ONE (16+2)x(8+2)-pixel halo tile per 64-channel block, loaded once by TMA (zero fill =
// padding) as 180 rows of 128 B with SWIZZLE_128B.  Because tcgen05 swizzles on absolute
// shared-memory address bits (profiles/r01_exp_umma_shift.md), tap (kh, kw) is the same tile
// read through a descriptor whose start address is advanced by (kh*10 + kw)*128 B and whose
// 8-row-group pitch (SBO) is 10*128 B: the nine taps cost no extra global/L2 traffic.
// The weight tile of a (tap, channel-block) is fetched once per CLUSTER: each CTA of the
// cluster loads 1/CS of the rows and multicasts them to its peers, the MMA warps release the
// slot with a multicast tcgen05.commit.  A ring (private) and B ring (cluster shared) are
// separate mbarrier pipelines; two TMEM accumulator stages overlap the epilogue with the next
// tile; the epilogue prefetches the residual before it waits for the accumulator.
#include "conv_common.cuh"
#include "umma_ptx.cuh"
#include "conv_epilogue.cuh"

#include <cudaTypedefs.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace brtpe {

constexpr int HL_THREADS = 192;
constexpr int HL_TW = 8, HL_TH = 16;
constexpr int HL_PITCH = HL_TW + 2;                       // halo row pitch in pixels
constexpr int HL_HROWS = (HL_TH + 2) * HL_PITCH;          // 180 pixel rows of 128 B
constexpr int HL_A_BYTES = HL_HROWS * 128;                // 23040
constexpr int HL_A_STAGE = 23552;                         // rounded to 1024
constexpr int HL_MAX_A = 4, HL_MAX_B = 8;
constexpr int HL_MAX_COUT_PAD = 512;
constexpr int HL_PREFETCH_CHUNKS = 6;                     // residual columns prefetched: 96

struct alignas(64) HaloParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  int N, H, W;
  int tiles_x, tiles_y, m_tiles, m_groups, n_tiles, BN, num_items;
  int num_kb, last_k16, in_coff;
  int a_stages, b_stages, b_stage_bytes, slice_rows, cs;
  int resident;                 // 1: all nine weight taps stay in smem for the whole kernel
  int tmem_cols, acc_cols;
  uint32_t idesc;
  FastDiv fd_xy, fd_x, fd_nt;
  const float* bias;
  EpiParams epi;
};

struct HaloConvPrepared {
  HaloParams p;
  int grid;
  size_t smem;
};

// descriptor halves (see make_kmajor_sw128_desc): lo = start>>4 | LBO(1)<<16, hi = SBO>>4 |
// version(1)<<14 | SWIZZLE_128B(2)<<29
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
  return ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
}
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
}

__global__ void __launch_bounds__(HL_THREADS)
conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // hoist everything the hot loops need out of the constant bank
  const int a_stages = p.a_stages, b_stages = p.b_stages, b_stage_bytes = p.b_stage_bytes;
  const int num_kb = p.num_kb, last_k16 = p.last_k16, num_items = p.num_items;
  const int n_tiles = p.n_tiles, BN = p.BN, cs = p.cs, tiles_x = p.tiles_x;
  const int tiles_xy = p.tiles_x * p.tiles_y, m_tiles = p.m_tiles;
  const bool resident = p.resident != 0;
  const uint32_t idesc = p.idesc;

  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)a_stages * HL_A_STAGE;
  uint8_t* tail = b_ring + (size_t)b_stages * b_stage_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_a = full_a + HL_MAX_A;
  uint64_t* full_b = empty_a + HL_MAX_A;
  uint64_t* empty_b = full_b + HL_MAX_B;
  uint64_t* tfull = empty_b + HL_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (cs > 1) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / cs;
  const int num_clusters = gridDim.x / cs;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < a_stages; ++s) {
      mbar_init(smem_u32(&full_a[s]), 1);
      mbar_init(smem_u32(&empty_a[s]), 1);
    }
    for (int s = 0; s < b_stages; ++s) {
      mbar_init(smem_u32(&full_b[s]), 1);
      mbar_init(smem_u32(&empty_b[s]), (uint32_t)cs);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_tiles * BN; i += HL_THREADS)
    bias_s[i] = (p.bias && i < p.epi.Cout) ? p.bias[i] : 0.0f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // peers' barriers exist before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) ========
    int as_ = 0, bs_ = 0;
    uint32_t aph = 0, bph = 0;
    const uint32_t b_bytes = (uint32_t)(BN * 128);
    const uint32_t slice_off = (uint32_t)(rank * p.slice_rows * 128);
    const int slice_row0 = rank * p.slice_rows;
    const int in_coff = p.in_coff;
    bool first = true;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int mg = (int)fdiv((uint32_t)item, p.fd_nt);
      const int nt = item - mg * n_tiles;
      const int mt = mg * cs + rank;
      int n = (int)fdiv((uint32_t)mt, p.fd_xy);
      const int rem = mt - n * tiles_xy;
      const int ty = (int)fdiv((uint32_t)rem, p.fd_x);
      const int y0 = ty * HL_TH;
      const int x0 = (rem - ty * tiles_x) * HL_TW;
      if (mt >= m_tiles) n = p.N;                       // fully out of bounds -> zeros
      if (resident && first) {
        // all nine taps of the (single) channel block, once per kernel
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_b[0]);
          mbar_expect_tx(fb, 9u * b_bytes);
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t dst = smem_u32(b_ring) + (uint32_t)(tap * BN * 128) + slice_off;
            if (cs > 1) tma_load_3d_mcast(dst, &p.tmap_b, fb, 0, slice_row0, tap, cmask);
            else tma_load_3d(dst, &p.tmap_b, fb, 0, 0, tap);
          }
        }
        __syncwarp();
        first = false;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_a[as_]), aph ^ 1u);
        if (elect_one()) {
          const uint32_t fa = smem_u32(&full_a[as_]);
          mbar_expect_tx(fa, (uint32_t)HL_A_BYTES);
          tma_load_5d(smem_u32(a_ring + (size_t)as_ * HL_A_STAGE), &p.tmap_a, fa,
                      in_coff + kb * 64, x0 - 1, 0, y0 - 1, n);
        }
        __syncwarp();
        if (++as_ == a_stages) { as_ = 0; aph ^= 1u; }
        if (!resident) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(smem_u32(&empty_b[bs_]), bph ^ 1u);
            if (elect_one()) {
              const uint32_t fb = smem_u32(&full_b[bs_]);
              mbar_expect_tx(fb, b_bytes);
              const uint32_t dst = smem_u32(b_ring + (size_t)bs_ * b_stage_bytes) + slice_off;
              if (cs > 1)
                tma_load_3d_mcast(dst, &p.tmap_b, fb, kb * 64, nt * BN + slice_row0, tap, cmask);
              else
                tma_load_3d(dst, &p.tmap_b, fb, kb * 64, nt * BN, tap);
            }
            __syncwarp();
            if (++bs_ == b_stages) { bs_ = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) ==========
    int as_ = 0, bs_ = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    constexpr uint32_t A_HI = desc_hi(HL_PITCH * 128);
    constexpr uint32_t B_HI = desc_hi(1024);
    const uint32_t a_ring_lo = desc_lo(smem_u32(a_ring));
    const uint32_t b_ring_lo = desc_lo(smem_u32(b_ring));
    const uint32_t a_stage_lo = (uint32_t)(HL_A_STAGE >> 4);
    const uint32_t b_stage_lo = (uint32_t)(b_stage_bytes >> 4);
    const uint32_t b_tap_lo = (uint32_t)((BN * 128) >> 4);
    if (resident) {
      mbar_wait(smem_u32(&full_b[0]), 0);
      tc_fence_after();
    }
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t accph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&tempty[acc]), accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_a[as_]), aph);
        tc_fence_after();
        const uint32_t a_lo = a_ring_lo + (uint32_t)as_ * a_stage_lo;
        const int k16 = (kb == num_kb - 1) ? last_k16 : 4;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          // tap (kh, kw): same halo tile, start advanced by (kh*PITCH + kw) rows of 128 B
          const uint32_t a_tap = a_lo + (uint32_t)((((tap / 3) * HL_PITCH + (tap % 3)) * 128) >> 4);
          uint32_t b_lo;
          if (resident) {
            b_lo = b_ring_lo + (uint32_t)tap * b_tap_lo;
          } else {
            mbar_wait(smem_u32(&full_b[bs_]), bph);
            tc_fence_after();
            b_lo = b_ring_lo + (uint32_t)bs_ * b_stage_lo;
          }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < k16)
                umma_f16_lohi(d_tmem, a_tap + 2u * k, A_HI, b_lo + 2u * k, B_HI, idesc,
                              (kb | tap | k) ? 1u : 0u);
            if (!resident) {
              if (cs > 1) umma_commit_mcast(smem_u32(&empty_b[bs_]), cmask);
              else umma_commit(smem_u32(&empty_b[bs_]));
            }
          }
          __syncwarp();
          if (!resident) {
            if (++bs_ == b_stages) { bs_ = 0; bph ^= 1u; }
          }
        }
        if (elect_one()) umma_commit(smem_u32(&empty_a[as_]));
        __syncwarp();
        if (++as_ == a_stages) { as_ = 0; aph ^= 1u; }
      }
      if (elect_one()) umma_commit(smem_u32(&tfull[acc]));
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const int th = m >> 3, tw = m & 7;
    const EpiParams e = p.epi;
    const int nchunks = BN >> 4;
    const int H = p.H, W = p.W, acc_cols = p.acc_cols;
    int it = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const int mg = (int)fdiv((uint32_t)item, p.fd_nt);
      const int nt = item - mg * n_tiles;
      const int mt = mg * cs + rank;
      const int n = (int)fdiv((uint32_t)mt, p.fd_xy);
      const int rem = mt - n * tiles_xy;
      const int ty = (int)fdiv((uint32_t)rem, p.fd_x);
      const int y = ty * HL_TH + th;
      const int x = (rem - ty * tiles_x) * HL_TW + tw;
      const bool valid = mt < m_tiles && y < H && x < W;
      const size_t opix = valid ? ((size_t)n * H + y) * W + x : 0;
      const int co0 = nt * BN;
      const int acc = it & 1;
      const uint32_t accph = (uint32_t)(it >> 1) & 1u;
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * acc_cols);
      epi_tile(e, bias_s, t_addr, nchunks, co0, valid, opix, smem_u32(&tfull[acc]), accph,
               smem_u32(&tempty[acc]), lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // nobody exits while a peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

static void halo_n_tiling(int cout_store, int* n_tiles, int* bn) {
  const int cp = (cout_store + 15) / 16 * 16;
  *n_tiles = (cp + 255) / 256;
  *bn = ((cp + *n_tiles - 1) / *n_tiles + 15) / 16 * 16;
}

bool halo_conv_supported(const brtpe_conv_desc* d) {
  if (d->dtype != BRTPE_DT_BF16 || d->ntaps != 9 || d->in_stride != 1 || d->out_scale != 1) return false;
  if (d->out_oy || d->out_ox || d->Hm != d->Hin || d->Wm != d->Win || d->Hout != d->Hin ||
      d->Wout != d->Win)
    return false;
  for (int t = 0; t < 9; ++t)
    if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1) return false;
  if (d->Cin % 16 || d->in_ld % 8 || d->in_coff % 8 || d->out_ld % 8 || d->out_coff % 8 ||
      d->res_ld % 8 || d->res_coff % 8)
    return false;
  int nt, bn;
  halo_n_tiling(d->Cout_store, &nt, &bn);
  if (nt * bn > HL_MAX_COUT_PAD) return false;
  return halo_encode_fn() != nullptr;
}

static bool g_halo_attr_set = false;

HaloConvPrepared* halo_conv_prepare(const brtpe_conv_desc* d, const void* in, const void* weights) {
  if (!halo_conv_supported(d)) {
    set_error("halo conv: unsupported layer");
    return nullptr;
  }
  HaloConvPrepared* P = new HaloConvPrepared();
  HaloParams& p = P->p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->Hin; p.W = d->Win;
  p.tiles_x = ceil_div(p.W, HL_TW);
  p.tiles_y = ceil_div(p.H, HL_TH);
  p.m_tiles = p.tiles_x * p.tiles_y * p.N;
  halo_n_tiling(d->Cout_store, &p.n_tiles, &p.BN);

  // cluster size: weight tile rows are split over the CTAs of a cluster and multicast
  int cs = 1;
  const char* env = getenv("BRTPE_HALO_CS");
  if (env) cs = atoi(env);
  if (cs != 1 && cs != 2 && cs != 4) cs = 1;
  while (cs > 1 && ((p.BN / cs) % 8 != 0 || p.BN % cs != 0 || p.m_tiles < cs)) cs >>= 1;
  p.cs = cs;
  p.slice_rows = p.BN / cs;
  p.m_groups = ceil_div(p.m_tiles, cs);
  p.num_items = p.m_groups * p.n_tiles;

  p.num_kb = ceil_div(d->Cin, 64);
  p.last_k16 = (d->Cin - (p.num_kb - 1) * 64) / 16;
  p.in_coff = d->in_coff;
  p.b_stage_bytes = (int)align_up((size_t)p.BN * 128, 1024);
  p.acc_cols = p.BN;
  int cols = 32;
  while (cols < 2 * p.acc_cols) cols *= 2;
  p.tmem_cols = cols;
  const int tail = 4096;
  bool two_per_sm = cols <= 256;
  // weights resident for the whole kernel when the layer has one channel block and one N tile
  p.resident = (p.num_kb == 1 && p.n_tiles == 1 && 9 * p.BN * 128 <= 80 * 1024) ? 1 : 0;
  if (getenv("BRTPE_HALO_NO_RESIDENT")) p.resident = 0;
  if (p.resident) {
    p.b_stage_bytes = (int)align_up((size_t)9 * p.BN * 128, 1024);
    p.b_stages = 1;
    if (p.b_stage_bytes + 2 * HL_A_STAGE + tail + 1024 > 113 * 1024) two_per_sm = false;
    const int budget = (two_per_sm ? 113 : 222) * 1024;
    int ast = (budget - tail - 1024 - p.b_stage_bytes) / HL_A_STAGE;
    if (ast > HL_MAX_A) ast = HL_MAX_A;
    p.a_stages = ast;
  } else {
    const int budget = (two_per_sm ? 113 : 222) * 1024;
    p.a_stages = two_per_sm ? 2 : 3;
    int bst = (budget - tail - 1024 - p.a_stages * HL_A_STAGE) / p.b_stage_bytes;
    if (bst > HL_MAX_B) bst = HL_MAX_B;
    if (bst < 2) bst = 2;
    p.b_stages = bst;
  }
  P->smem = (size_t)p.a_stages * HL_A_STAGE + (size_t)p.b_stages * p.b_stage_bytes + tail + 1024;
  const int max_ctas = num_sms() * (two_per_sm ? 2 : 1);
  int clusters = std::min(p.num_items, max_ctas / cs);
  if (clusters < 1) clusters = 1;
  P->grid = clusters * cs;

  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  p.epi.out = nullptr; p.epi.res = nullptr;
  p.epi.out_ld = d->out_ld; p.epi.out_coff = d->out_coff; p.epi.res_ld = d->res_ld;
  p.epi.res_coff = d->res_coff; p.epi.Cout = d->Cout; p.epi.Cout_store = d->Cout_store;
  p.epi.relu = d->relu; p.epi.vec32 = epi_vec32_ok(d); p.epi.fast = epi_fast_ok(d);
  p.fd_nt = make_fastdiv((uint32_t)p.n_tiles, (uint64_t)p.num_items + 1);
  p.fd_xy = make_fastdiv((uint32_t)(p.tiles_x * p.tiles_y), (uint64_t)p.m_groups * cs + cs);
  p.fd_x = make_fastdiv((uint32_t)p.tiles_x, (uint64_t)p.tiles_x * p.tiles_y);

  auto encode = halo_encode_fn();
  const cuuint64_t ld_b = (cuuint64_t)d->in_ld * 2;
  cuuint64_t gdim[5] = {(cuuint64_t)(d->in_coff + d->Cin), (cuuint64_t)d->Win, 1,
                        (cuuint64_t)d->Hin, (cuuint64_t)d->N};
  cuuint64_t gstr[4] = {ld_b, ld_b * d->Win, ld_b * d->Win, ld_b * d->Win * d->Hin};
  cuuint32_t box[5] = {64, (cuuint32_t)HL_PITCH, 1, (cuuint32_t)(HL_TH + 2), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&p.tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), gdim,
                      gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    delete P;
    return nullptr;
  }
  const int cin_pad = p.num_kb * 64;
  const int cout_pad = p.n_tiles * p.BN;
  cuuint64_t wdim[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, 9};
  cuuint64_t wstr[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * 2 * cout_pad};
  cuuint32_t wbox[3] = {64, (cuuint32_t)p.slice_rows, 1};
  cuuint32_t westr[3] = {1, 1, 1};
  r = encode(&p.tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(weights), wdim, wstr,
             wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("halo conv: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    delete P;
    return nullptr;
  }
  if (!g_halo_attr_set) {
    if (cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_halo_kernel) failed");
      delete P;
      return nullptr;
    }
    g_halo_attr_set = true;
  }
  return P;
}

void halo_conv_release(HaloConvPrepared* p) { delete p; }

int halo_conv_launch(const HaloConvPrepared* P, const float* bias, const void* residual, void* out,
                     cudaStream_t st) {
  HaloParams p = P->p;
  p.bias = bias;
  p.epi.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.epi.out = reinterpret_cast<__nv_bfloat16*>(out);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(P->grid);
  cfg.blockDim = dim3(HL_THREADS);
  cfg.dynamicSmemBytes = P->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_halo_kernel, p);
  if (e != cudaSuccess) {
    set_error("conv_halo_kernel launch failed: %s", cudaGetErrorString(e));
    return BRTPE_ECUDA;
  }
  return BRTPE_OK;
}

}  // namespace brtpe
