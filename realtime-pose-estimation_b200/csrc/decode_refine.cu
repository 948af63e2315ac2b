// HeatmapParser.refine (rtpe/third_party/group.py:202-264) for every person of every image.
//
// The reference copies the full det + tag maps to the host once PER PERSON
// (group.py:278-279) and then makes 17 full-map numpy passes per person.  Only the joints a
// person is missing (keypoints[j,2] == 0) can change (group.py:256-262), so this version
//   1. refine_prep_kernel  : per person, the float32 mean tag of its detected joints
//                            (numpy reduction order, Appendix A.9);
//   2. refine_scan_kernel  : ONE streaming pass over each (image, joint) plane evaluating
//                            argmax(det - rint(||tag - mean_p||)) for all persons p missing
//                            that joint.  first-maximum semantics of np.argmax are kept by
//                            reducing 64-bit keys order(score) << 32 | ~index with max().
//                            A shared monotone lower bound prunes pixels whose det value
//                            cannot beat the current best (score <= det), so tags are only
//                            touched near real candidates;
//   3. refine_fill_kernel  : +0.5 / quarter-pixel offsets and the in-place fill.
#include "common.cuh"
#include "umma_ptx.cuh"

#include <stdlib.h>

namespace brtpe {

constexpr int MAXT = BRTPE_MAX_TAG_DIMS;
constexpr int RF_THREADS = 256;
constexpr int RF_Q = 8;          // persons evaluated per pass over a plane
constexpr int RF_LIST_CAP = 1024;

struct RefineArgs {
  const float* det;   // (N,J,H,W)
  const float* tag;   // (N,Jt,H,W,T)
  float* ans;         // (N,Pmax,J,3+T)
  const int32_t* count;
  float* prev;        // ws (N,Pmax,MAXT)
  int32_t* valid;     // ws (N,Pmax)
  unsigned long long* best;  // ws (N,Pmax,J)
  int N, J, Jt, H, W, T, Pmax;
};

__device__ __forceinline__ float pairwise_sum_local(const float* a, int n) {
  if (n < 8) {
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, a[i]);
    return s;
  }
  float r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = a[q];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = __fadd_rn(r[q], a[i + q]);
  }
  float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) s = __fadd_rn(s, a[i]);
  return s;
}

// one thread per (image, person)
__global__ void refine_prep_kernel(RefineArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N * a.Pmax) return;
  const int n = i / a.Pmax, p = i - n * a.Pmax;
  if (p >= a.count[n]) {
    a.valid[i] = 0;
    return;
  }
  const int J = a.J, T = a.T, width = 3 + T;
  const size_t HW = (size_t)a.H * a.W;
  const float* person = a.ans + (size_t)i * J * width;
  float vals[BRTPE_MAX_JOINTS];
  float mean[MAXT];
  int cnt = 0;
  for (int t = 0; t < T; ++t) {
    cnt = 0;
    float s = 0.0f;
    for (int j = 0; j < J; ++j) {
      const float* kp = person + j * width;
      if (kp[2] > 0.0f) {
        const int x = (int)kp[0], y = (int)kp[1];  // astype(np.int32): truncation
        const int jt = (a.Jt == J) ? j : 0;
        const float tv = a.tag[(((size_t)n * a.Jt + jt) * HW + (size_t)y * a.W + x) * T + t];
        if (T == 1) vals[cnt] = tv;
        else s = (cnt == 0) ? tv : __fadd_rn(s, tv);
        ++cnt;
      }
    }
    if (cnt == 0) break;
    if (T == 1) s = pairwise_sum_local(vals, cnt);
    mean[t] = __fdiv_rn(s, (float)cnt);
  }
  a.valid[i] = cnt > 0;
  for (int t = 0; t < T; ++t) a.prev[(size_t)i * MAXT + t] = (cnt > 0) ? mean[t] : 0.0f;
  for (int j = 0; j < J; ++j) a.best[(size_t)i * J + j] = 0ull;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(FULL_MASK, v, o);
    v = other > v ? other : v;
  }
  return v;
}

template <int T>
__global__ void __launch_bounds__(RF_THREADS)
refine_scan_kernel(RefineArgs a, int splits) {
  __shared__ int plist[RF_LIST_CAP];
  __shared__ int pcount;
  __shared__ float sprev[RF_Q][MAXT];
  __shared__ unsigned int slb[RF_Q];        // order keys of per-person lower bounds
  __shared__ unsigned int slbmin;
  __shared__ unsigned long long sbest[RF_Q];

  const int plane = blockIdx.x;
  const int split = blockIdx.y;
  const int n = plane / a.J, j = plane - n * a.J;
  const int width = 3 + a.T;
  const int P = min(a.count[n], a.Pmax);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- persons of this image that miss joint j (warp 0 compacts in person order)
  if (tid == 0) pcount = 0;
  __syncthreads();
  if (warp == 0) {
    int cnt = 0;
    for (int base = 0; base < P; base += 32) {
      const int p = base + lane;
      bool miss = false;
      if (p < P) {
        const size_t pi = (size_t)n * a.Pmax + p;
        miss = a.valid[pi] && (a.ans[(pi * a.J + j) * width + 2] == 0.0f);
      }
      const unsigned m = __ballot_sync(FULL_MASK, miss);
      if (miss) {
        const int slot = cnt + __popc(m & ((1u << lane) - 1u));
        if (slot < RF_LIST_CAP) plist[slot] = p;
      }
      cnt += __popc(m);
    }
    if (lane == 0) pcount = min(cnt, RF_LIST_CAP);
  }
  __syncthreads();
  const int np = pcount;
  if (np == 0) return;

  const int HW = a.H * a.W;
  const int jt = (a.Jt == a.J) ? j : 0;
  const float* __restrict__ dplane = a.det + (size_t)plane * HW;
  const float* __restrict__ tplane = a.tag + ((size_t)n * a.Jt + jt) * (size_t)HW * T;
  const int ngroups = (HW + 3) >> 2;
  const int g0 = (int)(((long long)ngroups * split) / splits);
  const int g1 = (int)(((long long)ngroups * (split + 1)) / splits);
  const bool vec = ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(dplane) & 15) == 0);
  const float ninf = __int_as_float(0xff800000);

  for (int q0 = 0; q0 < np; q0 += RF_Q) {
    const int nq = min(RF_Q, np - q0);
    __syncthreads();
    if (tid < RF_Q) {
      slb[tid] = 0u;
      sbest[tid] = 0ull;
      if (tid < nq) {
        const size_t pi = (size_t)n * a.Pmax + plist[q0 + tid];
        for (int t = 0; t < T; ++t) sprev[tid][t] = a.prev[pi * MAXT + t];
      }
    }
    if (tid == 0) slbmin = 0u;
    __syncthreads();

    float bval[RF_Q];
    int bidx[RF_Q];
    float prev[RF_Q][T];
#pragma unroll
    for (int q = 0; q < RF_Q; ++q) {
      bval[q] = ninf;
      bidx[q] = -1;
#pragma unroll
      for (int t = 0; t < T; ++t) prev[q][t] = (q < nq) ? sprev[q][t] : 0.0f;
    }

    for (int g = g0 + tid; g < g1; g += RF_THREADS) {
      const int idx0 = g << 2;
      float d[4];
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dplane + idx0));
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = (idx0 + e < HW) ? __ldg(dplane + idx0 + e) : ninf;
      }
      const float dmax = fmaxf(fmaxf(d[0], d[1]), fmaxf(d[2], d[3]));
      // score <= det, so a group whose det values are all below every person's lower
      // bound cannot contain a (first) maximum.
      if (float_order_key(dmax) < *(volatile unsigned int*)&slbmin) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (idx0 + e >= HW) continue;
        float tv[T];
#pragma unroll
        for (int t = 0; t < T; ++t) tv[t] = __ldg(tplane + (size_t)(idx0 + e) * T + t);
#pragma unroll
        for (int q = 0; q < RF_Q; ++q) {
          if (q < nq && d[e] > bval[q]) {
            float s = 0.0f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const float df = __fsub_rn(tv[t], prev[q][t]);
              const float sq = __fmul_rn(df, df);
              s = (t == 0) ? sq : __fadd_rn(s, sq);
            }
            const float score = __fsub_rn(d[e], rintf(__fsqrt_rn(s)));
            if (score > bval[q]) {
              bval[q] = score;
              bidx[q] = idx0 + e;
              const unsigned ok = float_order_key(score);
              if (ok > slb[q]) {
                atomicMax(&slb[q], ok);
                unsigned mn = 0xffffffffu;
                for (int qq = 0; qq < nq; ++qq) mn = min(mn, slb[qq]);
                atomicMax(&slbmin, mn);
              }
            }
          }
        }
      }
    }

    // ---- reduce: thread -> warp -> CTA -> global (64-bit max keeps the first maximum)
#pragma unroll
    for (int q = 0; q < RF_Q; ++q) {
      unsigned long long key = (bidx[q] >= 0) ? make_sel_key(bval[q], (uint32_t)bidx[q]) : 0ull;
      key = warp_max_u64(key);
      if (lane == 0 && q < nq && key != 0ull) atomicMax(&sbest[q], key);
    }
    __syncthreads();
    if (tid < nq && sbest[tid] != 0ull) {
      const size_t pi = (size_t)n * a.Pmax + plist[q0 + tid];
      atomicMax(&a.best[pi * a.J + j], sbest[tid]);
    }
  }
}

// ---- streaming version (default when the planes are 16-byte aligned) -------------------
// det and tag of the CTA's pixel range are pulled through a ring of shared-memory stages by 1-D
// bulk async copies (cp.async.bulk + mbarrier complete_tx) issued by a producer warp; full/empty
// mbarriers, no CTA-wide barrier in the loop: NS-1 stages (tens of KB per CTA) are in flight
// independent of registers and occupancy -- the v1 kernel above had one 16-byte load per thread
// in flight and ran at 30 % of the HBM peak.  All persons missing the joint (up to RF_PMAX per
// pass) are evaluated against the stage while it sits in shared memory, so the maps are read
// from DRAM once whatever the number of persons.  A thread owns 4 pixel pairs per stage; per
// person it reduces them to one 64-bit key (order(score) << 32 | ~index: np.argmax's first
// maximum) before touching the shared per-person best, and it evaluates a pixel only if
// det >= the person's current lower bound, since score <= det.
constexpr int RF_PMAX = 32;
constexpr int RF_CH = 2048;                       // pixels per stage
constexpr int RF_CONSUMERS = 16;
constexpr int RFS_THREADS = (RF_CONSUMERS + 1) * 32;
constexpr int RF_PAIRS = RF_CH / 2 / (RF_CONSUMERS * 32);   // pixel pairs per thread per stage
constexpr unsigned RF_LB0 = 0x00800000u;          // order key of -FLT_MAX: -inf padding never passes

template <int T>
__global__ void __launch_bounds__(RFS_THREADS)
refine_stream_kernel(RefineArgs a, int splits, int NS) {
  extern __shared__ __align__(128) unsigned char rf_ring_raw[];
  __shared__ int plist[RF_LIST_CAP];
  __shared__ int pcount;
  __shared__ float sprev[RF_PMAX][MAXT];
  __shared__ unsigned int slb[RF_PMAX];      // order keys of the per-person lower bounds
  __shared__ unsigned int slbmin;
  __shared__ unsigned long long sbest[RF_PMAX];
  __shared__ __align__(8) unsigned long long full_bar[4];
  __shared__ __align__(8) unsigned long long empty_bar[4];

  float* ring = reinterpret_cast<float*>(rf_ring_raw);
  const int plane = blockIdx.x;
  const int split = blockIdx.y;
  const int n = plane / a.J, j = plane - n * a.J;
  const int width = 3 + a.T;
  const int P = min(a.count[n], a.Pmax);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- persons of this image that miss joint j (warp 0 compacts in person order)
  if (tid == 0) pcount = 0;
  __syncthreads();
  if (warp == 0) {
    int cnt = 0;
    for (int base = 0; base < P; base += 32) {
      const int p = base + lane;
      bool miss = false;
      if (p < P) {
        const size_t pi = (size_t)n * a.Pmax + p;
        miss = a.valid[pi] && (a.ans[(pi * a.J + j) * width + 2] == 0.0f);
      }
      const unsigned m = __ballot_sync(FULL_MASK, miss);
      if (miss) {
        const int slot = cnt + __popc(m & ((1u << lane) - 1u));
        if (slot < RF_LIST_CAP) plist[slot] = p;
      }
      cnt += __popc(m);
    }
    if (lane == 0) {
      pcount = min(cnt, RF_LIST_CAP);
      for (int s = 0; s < NS; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), RF_CONSUMERS);
      }
      fence_mbar_init();
    }
  }
  __syncthreads();
  const int np = pcount;
  if (np == 0) return;

  const int HW = a.H * a.W;
  const int jt = (a.Jt == a.J) ? j : 0;
  const float* __restrict__ dplane = a.det + (size_t)plane * HW;
  const float* __restrict__ tplane = a.tag + ((size_t)n * a.Jt + jt) * (size_t)HW * T;
  const int ngroups = HW >> 2;
  const int g0 = (int)(((long long)ngroups * split) / splits);
  const int g1 = (int)(((long long)ngroups * (split + 1)) / splits);
  const int px0 = g0 << 2, px1 = g1 << 2;
  if (px1 <= px0) return;
  const int nst = (px1 - px0 + RF_CH - 1) / RF_CH;
  const int passes = (np + RF_PMAX - 1) / RF_PMAX;
  const int total = nst * passes;
  constexpr int STAGE_FLOATS = RF_CH * (1 + T);

  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t ring0 = smem_u32(ring);
  if (warp == RF_CONSUMERS) {
    // ---- producer
    if (lane == 0) {
      int slot = 0, st = 0;
      uint32_t ephase = 1;                       // first wait on a fresh barrier passes
      for (int k = 0; k < total; ++k) {
        mbar_wait(empty0 + 8u * slot, ephase);
        const int p0 = px0 + st * RF_CH;
        const uint32_t cnt = (uint32_t)min(RF_CH, px1 - p0);
        const uint32_t bar = full0 + 8u * slot;
        const uint32_t dst = ring0 + (uint32_t)slot * (STAGE_FLOATS * 4u);
        mbar_expect_tx(bar, cnt * 4u * (1 + T));
        bulk_load_1d(dst, dplane + p0, cnt * 4u, bar);
        bulk_load_1d(dst + RF_CH * 4u, tplane + (size_t)p0 * T, cnt * 4u * T, bar);
        if (++st == nst) st = 0;
        if (++slot == NS) { slot = 0; ephase ^= 1u; }
      }
    }
    return;
  }

  // ---- consumers (named barrier 1 at pass boundaries)
  const float ninf = __int_as_float(0xff800000);
  int slot = 0, st = 0, q0 = 0, nq = 0;
  uint32_t phase = 0;
  for (int k = 0; k < total; ++k) {
    if (st == 0) {
      // new pass: flush the previous person batch, load the next
      if (k > 0) {
        named_bar_sync(1, RF_CONSUMERS * 32);
        if (tid < nq && sbest[tid] != 0ull) {
          const size_t pi = (size_t)n * a.Pmax + plist[q0 + tid];
          atomicMax(&a.best[pi * a.J + j], sbest[tid]);
        }
        named_bar_sync(1, RF_CONSUMERS * 32);
        q0 += RF_PMAX;
      }
      nq = min(RF_PMAX, np - q0);
      if (tid < RF_PMAX) {
        slb[tid] = RF_LB0;
        sbest[tid] = 0ull;
        if (tid < nq) {
          const size_t pi = (size_t)n * a.Pmax + plist[q0 + tid];
#pragma unroll
          for (int t = 0; t < T; ++t) sprev[tid][t] = a.prev[pi * MAXT + t];
        }
      }
      if (tid == 0) slbmin = RF_LB0;
      named_bar_sync(1, RF_CONSUMERS * 32);
    }
    mbar_wait(full0 + 8u * slot, phase);
    const float* sd = ring + (size_t)slot * STAGE_FLOATS;
    const float* stg = sd + RF_CH;
    const int p0 = px0 + st * RF_CH;
    const int npairs = min(RF_CH, px1 - p0) >> 1;

    float d[2 * RF_PAIRS];
    float dmax = ninf;
#pragma unroll
    for (int m = 0; m < RF_PAIRS; ++m) {
      const int pp = tid + m * (RF_CONSUMERS * 32);
      float2 v = make_float2(ninf, ninf);
      if (pp < npairs) v = *reinterpret_cast<const float2*>(sd + 2 * pp);
      d[2 * m] = v.x; d[2 * m + 1] = v.y;
      dmax = fmaxf(dmax, fmaxf(v.x, v.y));
    }
    const unsigned kmax = float_order_key(dmax);
    // warp-uniform from here on: one vote per (warp, person), one shared atomic per record
    const unsigned wkmax = __reduce_max_sync(FULL_MASK, kmax);
    if (wkmax >= *(volatile unsigned int*)&slbmin) {
      float tv[2 * RF_PAIRS][T];
      unsigned kd[2 * RF_PAIRS];
#pragma unroll
      for (int m = 0; m < RF_PAIRS; ++m) {
        const int pp = tid + m * (RF_CONSUMERS * 32);
        kd[2 * m] = float_order_key(d[2 * m]);
        kd[2 * m + 1] = float_order_key(d[2 * m + 1]);
        if (pp < npairs) {
          if (T == 2) {
            const float4 w = *reinterpret_cast<const float4*>(stg + 4 * pp);
            tv[2 * m][0] = w.x; tv[2 * m][T - 1] = w.y; tv[2 * m + 1][0] = w.z; tv[2 * m + 1][T - 1] = w.w;
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
              for (int t = 0; t < T; ++t) tv[2 * m + e][t] = stg[(2 * pp + e) * T + t];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int t = 0; t < T; ++t) tv[2 * m + e][t] = 0.0f;
        }
      }
      // Scalar tags (T == 1; the sparse planted maps of config 5): the range of the warp's 128 tags bounds
      // every score of the chunk for a person from above -- score = det - rint(|tag - mean|) <= max det -
      // rint(distance of the mean to the range), with the SAME float operations (each monotone), so a chunk
      // whose bound is below the person's best so far cannot hold a winner or a tie.  Persons whose tag does
      // not occur in a chunk (almost all chunks of a sparse map) then cost ~10 instructions instead of four
      // score evaluations per lane.  Not for T >= 2: on the bench's dense two-dimensional tags the bound
      // rarely prunes and cost 10 % (profiles/r01l_decode.md).
      float wtmin = 0.0f, wtmax = 0.0f, wdmax = 0.0f;
      if constexpr (T == 1) {
        float tmn = tv[0][0], tmx = tv[0][0];
#pragma unroll
        for (int e = 1; e < 2 * RF_PAIRS; ++e) {
          tmn = fminf(tmn, tv[e][0]);
          tmx = fmaxf(tmx, tv[e][0]);
        }
        wtmin = float_from_order_key(__reduce_min_sync(FULL_MASK, float_order_key(tmn)));
        wtmax = float_from_order_key(__reduce_max_sync(FULL_MASK, float_order_key(tmx)));
        wdmax = float_from_order_key(wkmax);
      }
      for (int q = 0; q < nq; ++q) {
        const unsigned lb = *(volatile unsigned int*)&slb[q];
        if (wkmax < lb) continue;
        if constexpr (T == 1) {
          const float pv0 = sprev[q][0];
          const float tb = pv0 < wtmin ? wtmin : (pv0 > wtmax ? wtmax : pv0);
          const float df = __fsub_rn(tb, pv0);
          const float bound = __fsub_rn(wdmax, rintf(__fsqrt_rn(__fmul_rn(df, df))));
          if (float_order_key(bound) < lb) continue;
        }
        float pv[T];
#pragma unroll
        for (int t = 0; t < T; ++t) pv[t] = sprev[q][t];
        unsigned long long best = 0ull;
#pragma unroll
        for (int e = 0; e < 2 * RF_PAIRS; ++e) {
          if (kd[e] >= lb) {
            float s = 0.0f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const float df = __fsub_rn(tv[e][t], pv[t]);
              const float sq = __fmul_rn(df, df);
              s = (t == 0) ? sq : __fadd_rn(s, sq);
            }
            const float score = __fsub_rn(d[e], rintf(__fsqrt_rn(s)));
            const int idx = p0 + 2 * (tid + (e >> 1) * (RF_CONSUMERS * 32)) + (e & 1);
            const unsigned long long key = ((unsigned long long)float_order_key(score) << 32) |
                                           (unsigned long long)(0xffffffffu - (uint32_t)idx);
            best = key > best ? key : best;
          }
        }
        const unsigned kb = (unsigned)(best >> 32);
        const unsigned wb = __reduce_max_sync(FULL_MASK, kb);
        if (wb < lb || wb == 0u) continue;
        // smallest index among the lanes holding the warp's best score
        const unsigned wlo = __reduce_max_sync(FULL_MASK, kb == wb ? (unsigned)best : 0u);
        unsigned old = 0xffffffffu;
        if (lane == 0) {
          atomicMax(&sbest[q], ((unsigned long long)wb << 32) | wlo);
          if (wb > lb) old = atomicMax(&slb[q], wb);
        }
        old = __shfl_sync(FULL_MASK, old, 0);
        if (old <= *(volatile unsigned int*)&slbmin) {
          const unsigned mine = lane < nq ? *(volatile unsigned int*)&slb[lane] : 0xffffffffu;
          const unsigned mn = __reduce_min_sync(FULL_MASK, mine);
          if (lane == 0) atomicMax(&slbmin, mn);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8u * slot);
    if (++st == nst) st = 0;
    if (++slot == NS) { slot = 0; phase ^= 1u; }
  }
  named_bar_sync(1, RF_CONSUMERS * 32);
  if (tid < nq && sbest[tid] != 0ull) {
    const size_t pi = (size_t)n * a.Pmax + plist[q0 + tid];
    atomicMax(&a.best[pi * a.J + j], sbest[tid]);
  }
}

// one thread per (image, person, joint)
__global__ void refine_fill_kernel(RefineArgs a) {
  const size_t total = (size_t)a.N * a.Pmax * a.J;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % a.J);
  const size_t pi = i / a.J;
  const int p = (int)(pi % a.Pmax);
  const int n = (int)(pi / a.Pmax);
  if (p >= a.count[n] || !a.valid[pi]) return;
  float* kp = a.ans + i * (3 + a.T);
  if (kp[2] != 0.0f) return;
  const unsigned long long key = a.best[i];
  if (key == 0ull) return;
  const int idx = (int)sel_key_index(key);
  const int H = a.H, W = a.W;
  const int y = idx / W, x = idx - y * W;
  const float* plane = a.det + ((size_t)n * a.J + j) * H * W;
  const float val = plane[idx];
  double fx = (double)x + 0.5, fy = (double)y + 0.5;
  if (plane[(size_t)y * W + min(x + 1, W - 1)] > plane[(size_t)y * W + max(x - 1, 0)]) fx += 0.25;
  else fx -= 0.25;
  if (plane[(size_t)min(y + 1, H - 1) * W + x] > plane[(size_t)max(0, y - 1) * W + x]) fy += 0.25;
  else fy -= 0.25;
  if (val > 0.0f) {
    kp[0] = (float)fx;
    kp[1] = (float)fy;
    kp[2] = val;
  }
}

}  // namespace brtpe

using namespace brtpe;

extern "C" size_t brtpe_refine_workspace_bytes(int N, int J, int T, int Pmax) {
  if (N <= 0 || J <= 0 || T <= 0 || Pmax <= 0) return 0;
  size_t b = 0;
  b += align_up((size_t)N * Pmax * MAXT * sizeof(float), 256);
  b += align_up((size_t)N * Pmax * sizeof(int32_t), 256);
  b += align_up((size_t)N * Pmax * J * sizeof(unsigned long long), 256);
  return b;
}

extern "C" int brtpe_refine(const float* det, const float* tag, float* ans, const int32_t* count,
                            int N, int J, int Jt, int H, int W, int T, int Pmax, void* workspace,
                            size_t workspace_bytes, void* stream) {
  BRTPE_CHECK_ARG(det && tag && ans && count, "brtpe_refine: null argument");
  BRTPE_CHECK_ARG(N > 0 && H > 0 && W > 0 && Pmax > 0, "brtpe_refine: bad sizes");
  BRTPE_CHECK_ARG(J >= 1 && J <= BRTPE_MAX_JOINTS, "brtpe_refine: J=%d outside [1,%d]", J,
                  BRTPE_MAX_JOINTS);
  BRTPE_CHECK_ARG(Jt == J || Jt == 1, "brtpe_refine: Jt must be J or 1");
  BRTPE_CHECK_ARG(T >= 1 && T <= MAXT, "brtpe_refine: T=%d outside [1,%d]", T, MAXT);
  BRTPE_CHECK_ARG((long long)H * W < (1ll << 31), "brtpe_refine: map too large");
  const size_t need = brtpe_refine_workspace_bytes(N, J, T, Pmax);
  if (!workspace || workspace_bytes < need) {
    set_error("brtpe_refine: workspace %zu < %zu", workspace_bytes, need);
    return BRTPE_EWORKSPACE;
  }
  RefineArgs a;
  a.det = det; a.tag = tag; a.ans = ans; a.count = count;
  char* w = reinterpret_cast<char*>(workspace);
  a.prev = reinterpret_cast<float*>(w);
  w += align_up((size_t)N * Pmax * MAXT * sizeof(float), 256);
  a.valid = reinterpret_cast<int32_t*>(w);
  w += align_up((size_t)N * Pmax * sizeof(int32_t), 256);
  a.best = reinterpret_cast<unsigned long long*>(w);
  a.N = N; a.J = J; a.Jt = Jt; a.H = H; a.W = W; a.T = T; a.Pmax = Pmax;
  cudaStream_t st = (cudaStream_t)stream;

  refine_prep_kernel<<<ceil_div(N * Pmax, 128), 128, 0, st>>>(a);
  BRTPE_LAUNCH_CHECK();

  const int planes = N * J;
  const int HW = H * W;
  const bool stream_ok = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(det) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(tag) & 15) == 0) && !getenv("BRTPE_REFINE_V1");
  if (stream_ok) {
    const int NS = (T <= 2) ? 4 : 3;
    const size_t smem = (size_t)NS * RF_CH * (1 + T) * sizeof(float);
    int splits = ceil_div(8 * num_sms(), planes);
    const int max_splits = HW / (RF_CH * 4) > 0 ? HW / (RF_CH * 4) : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    dim3 grid(planes, splits);
#define BRTPE_RF_CASE(TT)                                                                         \
  case TT: {                                                                                      \
    BRTPE_CUDA(cudaFuncSetAttribute(refine_stream_kernel<TT>,                                     \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));    \
    refine_stream_kernel<TT><<<grid, RFS_THREADS, smem, st>>>(a, splits, NS);                      \
  } break;
    switch (T) {
      BRTPE_RF_CASE(1)
      BRTPE_RF_CASE(2)
      BRTPE_RF_CASE(3)
      BRTPE_RF_CASE(4)
    }
#undef BRTPE_RF_CASE
  } else {
    int splits = ceil_div(4 * num_sms(), planes);
    const int ngroups = (H * W + 3) / 4;
    const int max_splits = ngroups / (RF_THREADS * 4) > 0 ? ngroups / (RF_THREADS * 4) : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    dim3 grid(planes, splits);
    switch (T) {
      case 1: refine_scan_kernel<1><<<grid, RF_THREADS, 0, st>>>(a, splits); break;
      case 2: refine_scan_kernel<2><<<grid, RF_THREADS, 0, st>>>(a, splits); break;
      case 3: refine_scan_kernel<3><<<grid, RF_THREADS, 0, st>>>(a, splits); break;
      case 4: refine_scan_kernel<4><<<grid, RF_THREADS, 0, st>>>(a, splits); break;
    }
  }
  BRTPE_LAUNCH_CHECK();

  const size_t total = (size_t)N * Pmax * J;
  refine_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}
