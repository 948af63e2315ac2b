// Fused NMS + per-(image, joint) top-K + tag gather.
//
// Replaces HeatmapParser.nms / HeatmapParser.top_k of the reference
// (rtpe/third_party/group.py:134-138, :144-179): MaxPool2d(k,1,p) -> eq -> mul ->
// view(N,J,HW).topk(K) -> gather(tag) -> ind % W, ind / W, which the reference runs as
// ~8 library kernels with 5 full-map passes plus a D2H sync.
//
// Design (HBM-bound: the heat map is read from DRAM exactly once, nothing map-sized is
// written):
//   * a warp owns a (row band) x (120 output columns) window of one (image, joint)
//     plane.  Each lane streams one float4 column strip top-to-bottom keeping the last
//     2R+1 rows in registers; the vertical max is taken in registers, the horizontal
//     max through warp shuffles of the neighbouring lanes' vertical maxima (lanes 0 and
//     31 only carry halo columns).  No shared memory in the streaming loop.
//   * positive peaks are pushed into a warp-resident sorted list (one 64-bit key per
//     lane and list; key = order(value) << 32 | ~index, so "value descending, index
//     ascending" is a plain integer maximum).  A running threshold (the list minimum)
//     rejects almost every later peak with one float compare.
//   * warps of a CTA merge through shared memory; planes split over several CTAs merge
//     in a tiny second kernel, which also zero-fills (first zero-valued positions in
//     index order, like topk over the NMS'd map), gathers the tags and writes (x, y).
#include "common.cuh"
#include "umma_ptx.cuh"

#include <stdlib.h>

namespace brtpe {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_WARPS = TOPK_THREADS / 32;
constexpr int TOPK_OUT_COLS = 120;  // output columns per warp window (lanes 1..30, 4 each)

template <int S>
struct TopList {
  unsigned long long k[S];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int s = 0; s < S; ++s) k[s] = 0ull;
  }
  __device__ __forceinline__ unsigned long long min_key() const {
    return __shfl_sync(FULL_MASK, k[S - 1], 31);
  }
  // warp-uniform insert of key c (all lanes pass the same c)
  __device__ __forceinline__ void insert(unsigned long long c, int lane) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      unsigned gt = __ballot_sync(FULL_MASK, k[s] > c);
      int pos = __popc(gt);
      if (pos < 32) {
        unsigned long long ev = __shfl_sync(FULL_MASK, k[s], 31);
        unsigned long long up = __shfl_up_sync(FULL_MASK, k[s], 1);
        if (lane > pos) k[s] = up;
        else if (lane == pos) k[s] = c;
        c = ev;
      }
    }
  }
};

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// NMS'd value of one pixel, straight from global memory (slow path helper).
__device__ float nms_value_direct(const float* __restrict__ plane, int H, int W, int R, int idx) {
  int y = idx / W, x = idx - y * W;
  float c = __ldg(plane + idx);
  float m = c;
  for (int dy = -R; dy <= R; ++dy) {
    int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -R; dx <= R; ++dx) {
      int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      m = fmaxf(m, __ldg(plane + yy * W + xx));
    }
  }
  return (m == c) ? c : c * 0.0f;
}

struct TopkOut {
  const float* tag;   // (N, Jt, HW, T)
  float* val_k;       // (planes, K)
  int32_t* ind_k;     // (planes, K)
  int64_t* loc_k;     // (planes, K, 2) or null
  float* tag_k;       // (planes, K, T)
  int J, Jt, T, K;
};

// One warp: write the final top-K of `plane` from a sorted list of positive peaks.
template <int S>
__device__ void topk_finalize(TopList<S>& L, const float* __restrict__ plane_ptr, int plane,
                              int H, int W, int R, const TopkOut& o, int lane) {
  const int HW = H * W;
  const int K = o.K;
  // positives already sorted: slot index of lane within list s is lane + 32*s
  int npos = 0;
#pragma unroll
  for (int s = 0; s < S; ++s) npos += __popc(__ballot_sync(FULL_MASK, L.k[s] != 0ull));
  if (npos > K) npos = K;

  const int n = plane / o.J, j = plane - n * o.J;
  const int jt = (o.Jt == o.J) ? j : 0;
  const float* tag_plane = o.tag ? o.tag + ((size_t)n * o.Jt + jt) * (size_t)HW * o.T : nullptr;

  auto emit = [&](int slot, float v, int idx) {
    size_t oidx = (size_t)plane * K + slot;
    o.val_k[oidx] = v;
    o.ind_k[oidx] = idx;
    if (o.loc_k) {
      o.loc_k[oidx * 2 + 0] = (int64_t)(idx % W);
      o.loc_k[oidx * 2 + 1] = (int64_t)(idx / W);
    }
    if (tag_plane) {
      for (int t = 0; t < o.T; ++t)
        o.tag_k[oidx * o.T + t] = __ldg(tag_plane + (size_t)idx * o.T + t);
    }
  };

#pragma unroll
  for (int s = 0; s < S; ++s) {
    int slot = lane + 32 * s;
    if (slot < npos) emit(slot, sel_key_value(L.k[s]), (int)sel_key_index(L.k[s]));
  }
  if (npos >= K) return;

  // ---- slow path: fewer than K positive peaks.  Next come zero-valued positions of
  // the NMS'd map in ascending index order, then negative peaks (value desc, idx asc).
  int need = K - npos;
  int found = 0;
  for (int base = 0; base < HW && found < need; base += 32) {
    int idx = base + lane;
    bool z = false;
    float nv = 0.0f;
    if (idx < HW) {
      nv = nms_value_direct(plane_ptr, H, W, R, idx);
      z = (nv == 0.0f);
    }
    unsigned m = __ballot_sync(FULL_MASK, z);
    int rank = __popc(m & ((1u << lane) - 1u));
    if (z && found + rank < need) emit(npos + found + rank, nv, idx);
    found += __popc(m);
  }
  if (found >= need) return;
  int base_slot = npos + found;
  need -= found;
  TopList<S> NL;
  NL.clear();
  unsigned long long mink = 0ull;
  for (int base = 0; base < HW; base += 32) {
    int idx = base + lane;
    bool c = false;
    unsigned long long key = 0ull;
    if (idx < HW) {
      float nv = nms_value_direct(plane_ptr, H, W, R, idx);
      if (nv < 0.0f) {
        key = make_sel_key(nv, (uint32_t)idx);
        c = key > mink;
      }
    }
    unsigned m = __ballot_sync(FULL_MASK, c);
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      unsigned long long kk = __shfl_sync(FULL_MASK, key, src);
      if (kk > mink) {
        NL.insert(kk, lane);
        mink = NL.min_key();
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s) {
    int slot = lane + 32 * s;
    if (slot < need && NL.k[s] != 0ull)
      emit(base_slot + slot, sel_key_value(NL.k[s]), (int)sel_key_index(NL.k[s]));
  }
}

template <int R, bool VEC>
__device__ __forceinline__ float4 load_row4(const float* __restrict__ plane, int H, int W, int y,
                                            int xb) {
  const float ninf = neg_inf();
  float4 v = make_float4(ninf, ninf, ninf, ninf);
  if (y < 0 || y >= H) return v;
  const float* row = plane + (size_t)y * W;
  if (VEC) {
    if (xb >= 0 && xb + 3 < W) v = __ldg(reinterpret_cast<const float4*>(row + xb));
  } else {
    if (xb >= 0 && xb < W) v.x = __ldg(row + xb);
    if (xb + 1 >= 0 && xb + 1 < W) v.y = __ldg(row + xb + 1);
    if (xb + 2 >= 0 && xb + 2 < W) v.z = __ldg(row + xb + 2);
    if (xb + 3 >= 0 && xb + 3 < W) v.w = __ldg(row + xb + 3);
  }
  return v;
}

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

template <int R, int S, bool VEC>
__global__ void __launch_bounds__(TOPK_THREADS)
nms_topk_kernel(const float* __restrict__ det, int H, int W, int band_h, int nbands,
                int ncolgroups, int splits, unsigned long long* __restrict__ ws_keys,
                TopkOut out) {
  constexpr int WIN = 2 * R + 1;
  const int plane = blockIdx.x;
  const int split = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const float* __restrict__ plane_ptr = det + (size_t)plane * H * W;

  const int items = nbands * ncolgroups;
  const int i0 = (int)(((long long)items * split) / splits);
  const int i1 = (int)(((long long)items * (split + 1)) / splits);

  TopList<S> L;
  L.clear();
  unsigned long long mink = 0ull;
  float thr = 0.0f;

  for (int item = i0 + warp; item < i1; item += TOPK_WARPS) {
    const int band = item / ncolgroups;
    const int cg = item - band * ncolgroups;
    const int y0 = band * band_h;
    const int y1 = min(H, y0 + band_h);
    const int xb = cg * TOPK_OUT_COLS - 4 + 4 * lane;
    const bool out_lane = (lane >= 1) && (lane <= 30);

    float4 win[WIN];
#pragma unroll
    for (int i = 0; i < WIN - 1; ++i) win[i] = load_row4<R, VEC>(plane_ptr, H, W, y0 - R + i, xb);

    for (int ybase = y0; ybase < y1; ybase += WIN) {
#pragma unroll
      for (int u = 0; u < WIN; ++u) {
        const int yc = ybase + u;
        if (yc < y1) {
          win[(WIN - 1 + u) % WIN] = load_row4<R, VEC>(plane_ptr, H, W, yc + R, xb);
          float4 vm = win[0];
#pragma unroll
          for (int i = 1; i < WIN; ++i) vm = max4(vm, win[i]);
          const float4 c = win[(R + u) % WIN];
          // a[0..11]: left neighbour strip, own strip, right neighbour strip
          float a[12];
          a[4] = vm.x; a[5] = vm.y; a[6] = vm.z; a[7] = vm.w;
          a[0] = a[1] = a[2] = a[3] = neg_inf();
          a[8] = a[9] = a[10] = a[11] = neg_inf();
          if (R >= 1) {
            a[3] = __shfl_up_sync(FULL_MASK, vm.w, 1);
            a[8] = __shfl_down_sync(FULL_MASK, vm.x, 1);
          }
          if (R >= 2) {
            a[2] = __shfl_up_sync(FULL_MASK, vm.z, 1);
            a[9] = __shfl_down_sync(FULL_MASK, vm.y, 1);
          }
          if (R >= 3) {
            a[1] = __shfl_up_sync(FULL_MASK, vm.y, 1);
            a[10] = __shfl_down_sync(FULL_MASK, vm.z, 1);
          }
          const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float hm = a[4 + q];
#pragma unroll
            for (int d = 1; d <= R; ++d) hm = fmaxf(hm, fmaxf(a[4 + q - d], a[4 + q + d]));
            const float v = cc[q];
            const bool cand = out_lane && (xb + q < W) && (hm == v) && (v > 0.0f) && (v >= thr);
            unsigned m = __ballot_sync(FULL_MASK, cand);
            if (m) {
              const unsigned long long key =
                  cand ? make_sel_key(v, (uint32_t)(yc * W + xb + q)) : 0ull;
              while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long kk = __shfl_sync(FULL_MASK, key, src);
                if (kk > mink) {
                  L.insert(kk, lane);
                  mink = L.min_key();
                  thr = (mink == 0ull) ? 0.0f : sel_key_value(mink);
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- merge the warps of this CTA
  __shared__ unsigned long long sh[TOPK_WARPS][32 * S];
#pragma unroll
  for (int s = 0; s < S; ++s) sh[warp][lane + 32 * s] = L.k[s];
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < TOPK_WARPS; ++w) {
    for (int e = 0; e < 32 * S; ++e) {
      const unsigned long long kk = sh[w][e];
      if (kk == 0ull || kk <= mink) break;  // lists are sorted descending
      L.insert(kk, lane);
      mink = L.min_key();
    }
  }
  if (splits == 1) {
    topk_finalize<S>(L, plane_ptr, plane, H, W, R, out, lane);
  } else {
    unsigned long long* dst = ws_keys + ((size_t)plane * splits + split) * (32 * S);
#pragma unroll
    for (int s = 0; s < S; ++s) dst[lane + 32 * s] = L.k[s];
  }
}

template <int S>
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ det, int planes, int H, int W, int R, int splits,
                  const unsigned long long* __restrict__ ws_keys, TopkOut out) {
  const int lane = threadIdx.x & 31;
  const int plane = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (plane >= planes) return;
  TopList<S> L;
  L.clear();
  unsigned long long mink = 0ull;
  for (int sp = 0; sp < splits; ++sp) {
    const unsigned long long* src = ws_keys + ((size_t)plane * splits + sp) * (32 * S);
    for (int e = 0; e < 32 * S; ++e) {
      const unsigned long long kk = src[e];
      if (kk == 0ull || kk <= mink) break;
      L.insert(kk, lane);
      mink = L.min_key();
    }
  }
  topk_finalize<S>(L, det + (size_t)plane * H * W, plane, H, W, R, out, lane);
}

// ---- streaming version (the default whenever rows are 16-byte aligned) ----------------
// The row band of a CTA is pulled through a ring of shared-memory stages by 1-D bulk async
// copies (cp.async.bulk + mbarrier complete_tx) issued by a producer warp: stage = SR whole
// rows = one contiguous range of the plane, NS stages, full/empty mbarriers, no CTA-wide
// barrier in the loop.  NS-3 stages (tens of KB per CTA) are in flight whatever the register
// budget is -- the v1 kernel above had one 16-byte load per thread in flight and ran at 22 %
// of the HBM peak.  A consumer warp takes (row, 128-column window) items of the current stage.
// Fast path per item: one LDS.128, max of the four values, one vote against the running
// threshold (the largest K-th best any warp of the CTA -- or any CTA of the plane, through a
// global word -- has proven so far).  Only row segments that still hold a candidate compute
// the (2R+1)^2 maximum, from shared memory, with no halo lanes; survivors go into the warp's
// register-resident sorted list (a shared list under a lock was 3x slower: r01k notes).
// NC consumer warps + producer warp + selector warp
template <int R, int S, int NC>
__global__ void __launch_bounds__((NC + 2) * 32)
nms_topk_stream_kernel(const float* __restrict__ det, int H, int W, int splits, int sr_shift,
                       int ns_shift, unsigned long long* __restrict__ ws_keys,
                       unsigned int* __restrict__ gthr, TopkOut out) {
  extern __shared__ __align__(128) unsigned char tk_ring_raw[];
  __shared__ unsigned long long sh[NC][32 * S];
  __shared__ __align__(8) unsigned long long full_bar[8];
  __shared__ __align__(8) unsigned long long empty_bar[8];
  __shared__ unsigned int s_thr;
  __shared__ unsigned int s_ver[NC];      // seqlock per published list
  __shared__ int s_done;

  float* ring = reinterpret_cast<float*>(tk_ring_raw);
  const int plane = blockIdx.x;
  const int split = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ plane_ptr = det + (size_t)plane * H * W;
  const int SR = 1 << sr_shift, NS = 1 << ns_shift, ns_mask = NS - 1;
  const int Y0 = (int)(((long long)H * split) / splits);
  const int Y1 = (int)(((long long)H * (split + 1)) / splits);
  const int ybase = max(0, Y0 - R), yend = min(H, Y1 + R);
  const int nst = (yend - ybase + SR - 1) >> sr_shift;
  const int stage_floats = SR * W;
  const int ncw = (W + 127) >> 7;
  const int inv_ncw = (65536 + ncw - 1) / ncw;     // it / ncw == (it * inv_ncw) >> 16 for it < 2^11
  const float ninf = neg_inf();

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NC);
    }
    s_thr = gthr ? *(volatile unsigned int*)(gthr + plane) : 0u;
    s_done = 0;
    fence_mbar_init();
  }
  if (tid < NC) s_ver[tid] = 0u;
  for (int e = tid; e < NC * 32 * S; e += ((NC + 2) * 32)) (&sh[0][0])[e] = 0ull;
  __syncthreads();

  TopList<S> L;
  L.clear();
  unsigned long long mink = 0ull;
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  if (warp == NC) {
    // ---- producer: one lane keeps the ring full
    if (lane == 0) {
      for (int k = 0; k < nst; ++k) {
        const int slot = k & ns_mask;
        if (k >= NS) mbar_wait(empty0 + 8u * slot, ((k >> ns_shift) - 1) & 1);
        const int r0 = ybase + (k << sr_shift);
        const uint32_t bytes = (uint32_t)(min(SR, yend - r0) * W) * 4u;
        const uint32_t bar = full0 + 8u * slot;
        mbar_expect_tx(bar, bytes);
        bulk_load_1d(smem_u32(ring) + (uint32_t)slot * (uint32_t)stage_floats * 4u,
                     plane_ptr + (size_t)r0 * W, bytes, bar);
      }
      *(volatile int*)&s_done = 1;
    }
  } else if (warp == NC + 1) {
    // ---- selector: K-th best key over the lists the consumer warps have published = a CTA-wide
    // threshold, much sharper than any single warp's own K-th best; exchanged with the other
    // CTAs of the plane through gthr
    while (*(volatile int*)&s_done == 0) {
      unsigned hi[NC * S];
#pragma unroll
      for (int w = 0; w < NC; ++w) {
        const unsigned v1 = *(volatile unsigned int*)&s_ver[w];
        __threadfence_block();
#pragma unroll
        for (int s = 0; s < S; ++s)
          hi[w * S + s] = (unsigned)(*(volatile unsigned long long*)&sh[w][lane + 32 * s] >> 32);
        __threadfence_block();
        const unsigned v2 = *(volatile unsigned int*)&s_ver[w];
        if (!__all_sync(FULL_MASK, v1 == v2 && (v1 & 1u) == 0u)) {   // torn snapshot: skip it
#pragma unroll
          for (int s = 0; s < S; ++s) hi[w * S + s] = 0u;
        }
      }
      unsigned prefix = 0u;
      for (int bit = 31; bit >= 0; --bit) {
        const unsigned cand = prefix | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < NC * S; ++e) cnt += hi[e] >= cand ? 1 : 0;
        if (__reduce_add_sync(FULL_MASK, cnt) >= out.K) prefix = cand;
      }
      if (lane == 0) {
        unsigned nt = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : 0u;   // positive floats only
        if (gthr) {
          if (nt) atomicMax(gthr + plane, nt);
          nt = max(nt, *(volatile unsigned int*)(gthr + plane));
        }
        if (nt) atomicMax(&s_thr, nt);
      }
      __nanosleep(1500);
    }
  } else {
    // address of row y (ybase <= y < yend) in the ring
    auto row_ptr = [&](int y) -> const float* {
      const int rel = y - ybase;
      return ring + (size_t)((rel >> sr_shift) & ns_mask) * stage_floats + (rel & (SR - 1)) * W;
    };
    bool dirty = false;
    unsigned ver = 0u;

    // full NMS of one 128-column row segment from shared memory; survivors go into L
    auto slow_row = [&](int ycur, int xb, bool in, const float4& c, float t) {
      float4 vl = make_float4(ninf, ninf, ninf, ninf), vc = vl, vr = vl;
      const bool hasl = in && (xb >= 4), hasr = in && (xb + 4 < W);
#pragma unroll
      for (int d = -R; d <= R; ++d) {
        const int yy = ycur + d;
        if (yy < 0 || yy >= H) continue;
        const float* rr = row_ptr(yy);
        if (in) vc = max4(vc, *reinterpret_cast<const float4*>(rr + xb));
        if (R > 0) {
          if (hasl) vl = max4(vl, *reinterpret_cast<const float4*>(rr + xb - 4));
          if (hasr) vr = max4(vr, *reinterpret_cast<const float4*>(rr + xb + 4));
        }
      }
      const float a[12] = {vl.x, vl.y, vl.z, vl.w, vc.x, vc.y, vc.z, vc.w, vr.x, vr.y, vr.z, vr.w};
      const float cc[4] = {c.x, c.y, c.z, c.w};
      unsigned long long key[4];
      unsigned any = 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float hm = a[4 + q];
#pragma unroll
        for (int d = 1; d <= R; ++d) hm = fmaxf(hm, fmaxf(a[4 + q - d], a[4 + q + d]));
        const float v = cc[q];
        const bool cand = in && (hm == v) && (v > 0.0f) && (v >= t);
        key[q] = cand ? make_sel_key(v, (uint32_t)(ycur * W + xb + q)) : 0ull;
        any |= cand ? 1u : 0u;
      }
      if (!__any_sync(FULL_MASK, any != 0u)) return;
      bool grew = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unsigned mm = __ballot_sync(FULL_MASK, key[q] > mink);
        while (mm) {
          const int src = __ffs(mm) - 1;
          mm &= mm - 1;
          const unsigned long long kk = __shfl_sync(FULL_MASK, key[q], src);
          if (kk > mink) {
            L.insert(kk, lane);
            mink = L.min_key();
            grew = true;
          }
        }
      }
      dirty |= grew;
      if (grew && lane == 0 && mink != 0ull) {
        const unsigned nt = __float_as_uint(sel_key_value(mink));
        if (nt > __float_as_uint(t)) atomicMax(&s_thr, nt);
      }
    };

    for (int i = 0; i < nst; ++i) {
      if (i == 0) mbar_wait(full0, 0);
      if (i + 1 < nst) mbar_wait(full0 + 8u * ((i + 1) & ns_mask), ((i + 1) >> ns_shift) & 1);
      const int s0 = ybase + (i << sr_shift);
      const int ra = max(Y0, s0), rb = min(Y1, s0 + SR);
      // item = two consecutive rows x one 128-column window; rotate so no warp is always heavier
      const int items = ((rb - ra + 1) >> 1) * ncw;
      for (int it = (warp + 5 * i) & (NC - 1); it < items; it += NC) {
        const int rp = (it * inv_ncw) >> 16;
        const int xb = ((it - rp * ncw) << 7) + 4 * lane;
        const int ya = ra + 2 * rp;
        const bool in = xb < W;                     // W % 4 == 0: a strip is all in or all out
        const bool two = ya + 1 < rb;               // both rows lie in this stage: contiguous
        const float* rowa = row_ptr(ya) + xb;
        float4 c0 = make_float4(ninf, ninf, ninf, ninf), c1 = c0;
        if (in) c0 = *reinterpret_cast<const float4*>(rowa);
        if (in && two) c1 = *reinterpret_cast<const float4*>(rowa + W);
        const float t = __uint_as_float(*(volatile unsigned int*)&s_thr);
        const float m0 = fmaxf(fmaxf(c0.x, c0.y), fmaxf(c0.z, c0.w));
        const float m1 = fmaxf(fmaxf(c1.x, c1.y), fmaxf(c1.z, c1.w));
        const unsigned hit = __ballot_sync(FULL_MASK, (m0 > 0.0f) && (m0 >= t)) ? 1u : 0u;
        const unsigned hit1 = __ballot_sync(FULL_MASK, (m1 > 0.0f) && (m1 >= t)) ? 1u : 0u;
        if (hit) slow_row(ya, xb, in, c0, t);
        if (hit1) slow_row(ya + 1, xb, in, c1, t);
      }
      if (dirty) {
        // publish the list for the selector warp (seqlock: odd = writing)
        if (lane == 0) *(volatile unsigned int*)&s_ver[warp] = ver + 1u;
        __syncwarp();
        __threadfence_block();
#pragma unroll
        for (int s = 0; s < S; ++s) *(volatile unsigned long long*)&sh[warp][lane + 32 * s] = L.k[s];
        __threadfence_block();
        __syncwarp();
        ver += 2u;
        if (lane == 0) *(volatile unsigned int*)&s_ver[warp] = ver;
        dirty = false;
      }
      if (i >= 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * ((i - 1) & ns_mask));
      }
    }
  }

  // ---- merge the consumer warps of this CTA (every changed list was published at stage end)
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < NC; ++w) {
    for (int e = 0; e < 32 * S; ++e) {
      const unsigned long long kk = sh[w][e];
      if (kk == 0ull || kk <= mink) break;  // lists are sorted descending
      L.insert(kk, lane);
      mink = L.min_key();
    }
  }
  if (splits == 1) {
    topk_finalize<S>(L, plane_ptr, plane, H, W, R, out, lane);
  } else {
    unsigned long long* dst = ws_keys + ((size_t)plane * splits + split) * (32 * S);
#pragma unroll
    for (int s = 0; s < S; ++s) dst[lane + 32 * s] = L.k[s];
  }
}

// stage geometry of the streaming kernel; false = use the v1 kernel
static bool topk_stream_geometry(int planes, int H, int W, int R, int* sr_shift, int* ns_shift,
                                 int* splits, size_t* smem, int* nc) {
  if (W % 4 != 0 || W < 4) return false;
  int sr = 4;                                   // >= R for every supported kernel size
  while (sr < 32 && (size_t)sr * 2 * W * 4 <= 12288) sr *= 2;
  if (const char* e = getenv("BRTPE_TOPK_SR")) sr = atoi(e);
  const size_t stage = (size_t)sr * W * 4;
  int ns;
  if (stage * 8 <= 168 * 1024) ns = 8;
  else if (stage * 4 <= 168 * 1024) ns = 4;
  else return false;                            // very wide maps: v1
  if (const char* e = getenv("BRTPE_TOPK_NS")) ns = atoi(e);
  *nc = 8;
  if (const char* e = getenv("BRTPE_TOPK_NC")) *nc = atoi(e);
  if ((ns != 4 && ns != 8) || (*nc != 8 && *nc != 16) || stage * ns > 200 * 1024 || sr < 4 ||
      (sr & (sr - 1)))
    return false;
  int s = ceil_div(4 * num_sms(), planes);
  if (const char* e = getenv("BRTPE_TOPK_SPLITS")) s = atoi(e);
  // every split should own at least four stages of rows
  const int max_splits = H / (4 * sr) > 0 ? H / (4 * sr) : 1;
  if (s > max_splits) s = max_splits;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  int sh = 0;
  while ((1 << sh) < sr) ++sh;
  *sr_shift = sh;
  *ns_shift = ns == 8 ? 3 : 2;
  *splits = s;
  *smem = stage * ns;
  (void)R;
  return true;
}

template <int R, int S, int NC>
static int launch_topk_stream_nc(const float* det, int planes, int H, int W, int sr_shift,
                                 int ns_shift, int splits, size_t smem, unsigned long long* ws,
                                 unsigned int* gthr, const TopkOut& o, cudaStream_t st) {
  // per call (a few hundred ns): the attribute is per device, and callers may use several
  BRTPE_CUDA(cudaFuncSetAttribute(nms_topk_stream_kernel<R, S, NC>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (gthr) BRTPE_CUDA(cudaMemsetAsync(gthr, 0, (size_t)planes * sizeof(unsigned int), st));
  dim3 grid(planes, splits);
  nms_topk_stream_kernel<R, S, NC><<<grid, (NC + 2) * 32, smem, st>>>(det, H, W, splits, sr_shift,
                                                                     ns_shift, ws, gthr, o);
  BRTPE_LAUNCH_CHECK();
  if (splits > 1) {
    const int wpb = 4;
    topk_merge_kernel<S><<<ceil_div(planes, wpb), wpb * 32, 0, st>>>(det, planes, H, W, R, splits,
                                                                     ws, o);
    BRTPE_LAUNCH_CHECK();
  }
  return BRTPE_OK;
}

template <int R, int S>
static int launch_topk_stream(const float* det, int planes, int H, int W, int sr_shift, int ns_shift,
                              int splits, size_t smem, int nc, unsigned long long* ws,
                              unsigned int* gthr, const TopkOut& o, cudaStream_t st) {
  if (nc == 16)
    return launch_topk_stream_nc<R, S, 16>(det, planes, H, W, sr_shift, ns_shift, splits, smem, ws,
                                           gthr, o, st);
  return launch_topk_stream_nc<R, S, 8>(det, planes, H, W, sr_shift, ns_shift, splits, smem, ws, gthr,
                                        o, st);
}

__global__ void nms_kernel(const float* __restrict__ det, float* __restrict__ out, int planes,
                           int H, int W, int R) {
  const size_t total = (size_t)planes * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int hw = H * W;
    const size_t p = i / hw;
    const int idx = (int)(i - p * hw);
    out[i] = nms_value_direct(det + p * hw, H, W, R, idx);
  }
}

template <int R, int S>
static int launch_topk(const float* det, int planes, int H, int W, bool vec, int band_h, int nbands,
                       int ncg, int splits, unsigned long long* ws, const TopkOut& o,
                       cudaStream_t st) {
  dim3 grid(planes, splits);
  if (vec)
    nms_topk_kernel<R, S, true><<<grid, TOPK_THREADS, 0, st>>>(det, H, W, band_h, nbands, ncg,
                                                              splits, ws, o);
  else
    nms_topk_kernel<R, S, false><<<grid, TOPK_THREADS, 0, st>>>(det, H, W, band_h, nbands, ncg,
                                                               splits, ws, o);
  BRTPE_LAUNCH_CHECK();
  if (splits > 1) {
    const int wpb = 4;
    topk_merge_kernel<S><<<ceil_div(planes, wpb), wpb * 32, 0, st>>>(det, planes, H, W, R, splits,
                                                                     ws, o);
    BRTPE_LAUNCH_CHECK();
  }
  return BRTPE_OK;
}

static void topk_geometry(int planes, int H, int W, int* band_h, int* nbands, int* ncg,
                          int* splits) {
  *band_h = 32;
  *nbands = ceil_div(H, *band_h);
  *ncg = ceil_div(W, TOPK_OUT_COLS);
  const int items = (*nbands) * (*ncg);
  // aim for >= 2 CTAs per SM across the chip; each split should keep >= TOPK_WARPS/2 items
  int want = ceil_div(4 * num_sms(), planes);
  int max_splits = items / 4 > 0 ? items / 4 : 1;
  int s = want < 1 ? 1 : want;
  if (s > max_splits) s = max_splits;
  if (s > 64) s = 64;
  *splits = s;
}

}  // namespace brtpe

using namespace brtpe;

extern "C" int brtpe_nms(const float* det, float* out, int planes, int H, int W, int ksize,
                         int padding, void* stream) {
  BRTPE_CHECK_ARG(det && out && planes > 0 && H > 0 && W > 0, "brtpe_nms: bad tensor arguments");
  BRTPE_CHECK_ARG(ksize >= 1 && (ksize & 1) && 2 * padding == ksize - 1,
                  "brtpe_nms: needs odd ksize with 2*padding == ksize-1 (got k=%d p=%d); the "
                  "reference's eq(maxpool(det), det) is only shape-valid in that case",
                  ksize, padding);
  const size_t total = (size_t)planes * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  nms_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(det, out, planes, H, W, padding);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

extern "C" size_t brtpe_topk_workspace_bytes(int N, int J, int H, int W, int K) {
  if (N <= 0 || J <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
  // worst case splits = 64, S = 2; + one threshold word per plane (streaming kernel)
  return (size_t)N * J * 64 * 64 * sizeof(unsigned long long) + align_up((size_t)N * J * 4, 256);
}

extern "C" int brtpe_nms_topk_gather(const float* det, const float* tag, int N, int J, int Jt,
                                     int H, int W, int T, int K, int ksize, int padding,
                                     float* val_k, int32_t* ind_k, int64_t* loc_k, float* tag_k,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  BRTPE_CHECK_ARG(det && val_k && ind_k, "brtpe_nms_topk_gather: null det/val_k/ind_k");
  BRTPE_CHECK_ARG(N > 0 && J > 0 && H > 0 && W > 0, "brtpe_nms_topk_gather: bad shape");
  BRTPE_CHECK_ARG(K >= 1 && K <= BRTPE_MAX_TOPK, "brtpe_nms_topk_gather: K=%d outside [1,%d]", K,
                  BRTPE_MAX_TOPK);
  BRTPE_CHECK_ARG((long long)H * W >= K,
                  "brtpe_nms_topk_gather: map has %lld positions < K=%d (torch.topk would raise)",
                  (long long)H * W, K);
  BRTPE_CHECK_ARG((long long)H * W < (1ll << 31), "brtpe_nms_topk_gather: map too large");
  BRTPE_CHECK_ARG(ksize >= 1 && ksize <= 7 && (ksize & 1) && 2 * padding == ksize - 1,
                  "brtpe_nms_topk_gather: needs odd ksize <= 7 with 2*padding == ksize-1 (k=%d p=%d)",
                  ksize, padding);
  if (tag) {
    BRTPE_CHECK_ARG(tag_k != nullptr, "brtpe_nms_topk_gather: tag given but tag_k is null");
    BRTPE_CHECK_ARG(Jt == J || Jt == 1, "brtpe_nms_topk_gather: Jt must be J or 1");
    BRTPE_CHECK_ARG(T >= 1, "brtpe_nms_topk_gather: T must be >= 1");
  }
  const int S = (K <= 32) ? 1 : 2;
  const int R = padding;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* ws = reinterpret_cast<unsigned long long*>(workspace);
  TopkOut o{tag, val_k, ind_k, loc_k, tag_k, J, Jt, T, K};
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(det) & 15) == 0);
  {
    int sr_shift, ns_shift, ssplits, nc;
    size_t smem;
    // Many small planes (config 5: 17 408 planes of 320 x 320): a streaming CTA lives ~20 us, its
    // prologue / list merge shows, and the register-window kernel is faster (3.98 vs 4.50 ms for the
    // 1024-image batch; the bench's 544 planes of 640 x 640: 0.54 vs 0.345 ms the other way round).
    // BRTPE_TOPK_STREAM=1 forces the streaming kernel, BRTPE_TOPK_V1=1 the first one.
    const bool small_planes = (size_t)H * W * 4 <= 512 * 1024 && (long long)N * J >= 8ll * num_sms() &&
                              !getenv("BRTPE_TOPK_STREAM");
    if (vec && !getenv("BRTPE_TOPK_V1") && !small_planes &&
        topk_stream_geometry(N * J, H, W, R, &sr_shift, &ns_shift, &ssplits, &smem, &nc)) {
      const size_t keys = (size_t)N * J * ssplits * 32 * S * sizeof(unsigned long long);
      const size_t sneed = align_up(keys, 256) + (size_t)N * J * sizeof(unsigned int);
      if (sneed > workspace_bytes || !workspace) {
        set_error("brtpe_nms_topk_gather: workspace %zu < %zu", workspace_bytes, sneed);
        return BRTPE_EWORKSPACE;
      }
      unsigned int* gthr = ssplits > 1
          ? reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + align_up(keys, 256))
          : nullptr;
#define BRTPE_TOPKS_CASE(RR)                                                                     \
  case RR:                                                                                       \
    return (S == 1) ? launch_topk_stream<RR, 1>(det, N * J, H, W, sr_shift, ns_shift, ssplits,   \
                                                smem, nc, ws, gthr, o, st)                       \
                    : launch_topk_stream<RR, 2>(det, N * J, H, W, sr_shift, ns_shift, ssplits,   \
                                                smem, nc, ws, gthr, o, st);
      switch (R) {
        BRTPE_TOPKS_CASE(0)
        BRTPE_TOPKS_CASE(1)
        BRTPE_TOPKS_CASE(2)
        BRTPE_TOPKS_CASE(3)
      }
#undef BRTPE_TOPKS_CASE
    }
  }
  int band_h, nbands, ncg, splits;
  topk_geometry(N * J, H, W, &band_h, &nbands, &ncg, &splits);
  const size_t need = (splits > 1) ? (size_t)N * J * splits * 32 * S * sizeof(unsigned long long) : 0;
  if (need > workspace_bytes || (need && !workspace)) {
    set_error("brtpe_nms_topk_gather: workspace %zu < %zu", workspace_bytes, need);
    return BRTPE_EWORKSPACE;
  }
#define BRTPE_TOPK_CASE(RR)                                                                    \
  case RR:                                                                                     \
    return (S == 1) ? launch_topk<RR, 1>(det, N * J, H, W, vec, band_h, nbands, ncg, splits, ws, o, st) \
                    : launch_topk<RR, 2>(det, N * J, H, W, vec, band_h, nbands, ncg, splits, ws, o, st);
  switch (R) {
    BRTPE_TOPK_CASE(0)
    BRTPE_TOPK_CASE(1)
    BRTPE_TOPK_CASE(2)
    BRTPE_TOPK_CASE(3)
  }
#undef BRTPE_TOPK_CASE
  set_error("brtpe_nms_topk_gather: unsupported padding %d", padding);
  return BRTPE_EINVAL;
}
