// Associative-embedding grouping: HeatmapParser.match / match_by_tag / py_max_match
// (rtpe/third_party/group.py:19-97, :140-142), adjust (:181-200) and the score mean (:272).
//
// The reference runs this on the host: numpy + the pure-Python `munkres` package, 17
// sequential Hungarian rounds per image.  Here one warp owns one image: the float64 cost
// matrix (<= 32x32) lives in shared memory, lane j owns column j / row j, and every scan of
// the Kuhn-Munkres steps is a warp ballot.  All order-defining details of munkres 1.1.x are
// kept (see oracle/munkres_ref.py): first-zero starring in step 2, the "last uncovered zero
// in cyclic column order of the first row that has one, search restarted at the previous
// hit" rule of step 4, first-row / first-column star and prime look-ups, and step 6's
// "+= minval then -= minval" in float64 without contraction.
//
// The Python dict semantics of match_by_tag are reproduced literally: persons are kept in
// creation order, looked up by their float32 key tag[0] (equal keys collapse, a re-used key
// resets that person's tag list), only the first max_num_people persons are matching
// candidates, and the person list itself is unbounded (capacity Pmax, J*K always fits).
// Tag means follow numpy's float32 reduction order.
#include "common.cuh"

namespace brtpe {

constexpr int GRP_WARPS = 4;
constexpr int GRP_N = 32;          // max matrix dimension (K <= 32)
constexpr int GRP_LD = 33;         // padded row stride of the cost matrix
constexpr int MAXT = BRTPE_MAX_TAG_DIMS;

struct GroupSmem {
  double C[GRP_N * GRP_LD];
  float cval[GRP_N];
  float cx[GRP_N];
  float cy[GRP_N];
  float ctag[GRP_N][MAXT];
  float mean[GRP_N][MAXT];
};

struct GroupArgs {
  const float* val_k;     // (N,J,K)
  const int32_t* ind_k;   // (N,J,K)
  const float* tag_k;     // (N,J,K,T)
  float* ans;             // (N,Pmax,J,3+T)
  int32_t* count;         // (N)
  int32_t* overflow;      // (1)
  float* keys;            // ws (N,Pmax)
  int32_t* tl_count;      // ws (N,Pmax)
  float* tl;              // ws (N,Pmax,J+1,T)
  int N, J, K, T, W, Pmax;
  int max_people;
  double det_thr, tag_thr;
  int use_det_val, ignore_too_much, start_rule;
};

// numpy float32 add.reduce order for a strided run of n < 128 values (Appendix A.9 (i)/(ii)).
__device__ __forceinline__ float pairwise_sum_f32(const float* a, int stride, int n) {
  if (n < 8) {
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, a[i * stride]);
    return s;
  }
  float r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = a[q * stride];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = __fadd_rn(r[q], a[(i + q) * stride]);
  }
  float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) s = __fadd_rn(s, a[i * stride]);
  return s;
}

// np.mean(list of (T,) f32 vectors, axis=0)[t]
__device__ __forceinline__ float mean_tag_component(const float* list, int n, int T, int t) {
  float s;
  if (T == 1) {
    s = pairwise_sum_f32(list, 1, n);
  } else {
    s = list[t];
    for (int r = 1; r < n; ++r) s = __fadd_rn(s, list[r * T + t]);
  }
  return __fdiv_rn(s, (float)n);
}

// float64 tag distance exactly like np.linalg.norm(diff, ord=2, axis=2): sqrt(sum d*d).
__device__ __forceinline__ double tag_dist_f64(const float* a, const float* b, int T) {
  double s = 0.0;
  for (int t = 0; t < T; ++t) {
    double d = __dsub_rn((double)a[t], (double)b[t]);
    double sq = __dmul_rn(d, d);
    s = (t == 0) ? sq : __dadd_rn(s, sq);
  }
  return __dsqrt_rn(s);
}

__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double other = __shfl_xor_sync(FULL_MASK, v, o);
    v = (other < v) ? other : v;
  }
  return v;
}

// Kuhn-Munkres on the n x n matrix in sm.C; returns (per lane i < n) the starred column of
// row i.  All control flow is warp-uniform.
__device__ int munkres_warp(GroupSmem& sm, int n, int lane, int start_rule) {
  double* C = sm.C;
  const bool act = lane < n;
  int star = -1, prime = -1;
  unsigned rowcov = 0u, colcov = 0u;
  const unsigned nmask = (n >= 32) ? FULL_MASK : ((1u << n) - 1u);

  // step 1: subtract row minima (lane = row)
  if (act) {
    double mn = C[lane * GRP_LD];
    for (int j = 1; j < n; ++j) {
      double v = C[lane * GRP_LD + j];
      if (v < mn) mn = v;
    }
    for (int j = 0; j < n; ++j) C[lane * GRP_LD + j] = __dsub_rn(C[lane * GRP_LD + j], mn);
  }
  __syncwarp();

  // step 2: star the first free zero of each row (lane = column)
  for (int i = 0; i < n; ++i) {
    bool z = act && (C[i * GRP_LD + lane] == 0.0) && !((colcov >> lane) & 1u);
    unsigned m = __ballot_sync(FULL_MASK, z);
    if (m) {
      int col = __ffs(m) - 1;
      if (lane == i) star = col;
      colcov |= 1u << col;
    }
  }
  colcov = 0u;

  // zero mask of every row (lane i holds row i: bit j <=> C[i][j] == 0), rebuilt only when step 6
  // changes the matrix: step 4 then finds the next uncovered zero with ONE vote instead of a
  // scan of up to n rows with one shared-memory read and one vote each (same visiting order)
  unsigned zrow = 0u;
  auto rebuild_zero_masks = [&]() {
    for (int i = 0; i < n; ++i) {
      const unsigned m = __ballot_sync(FULL_MASK, act && (C[i * GRP_LD + lane] == 0.0));
      if (lane == i) zrow = m;
    }
  };
  rebuild_zero_masks();

  int step = 3;
  int z0r = 0, z0c = 0;
  int guard = 0;
  while (true) {
    if (++guard > 100000) break;  // never hit for finite costs; avoids a hang on NaN input
    if (step == 3) {
      unsigned m = 0u;
      // cover every column holding a star
      unsigned mine = (act && star >= 0) ? (1u << star) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine |= __shfl_xor_sync(FULL_MASK, mine, o);
      m = mine;
      colcov |= m;
      if (__popc(colcov & nmask) >= n) break;
      step = 4;
    } else if (step == 4) {
      int row = 0, col = 0;
      while (true) {
        const int i0 = start_rule ? 0 : row;
        const int j0 = start_rule ? 0 : col;
        int fr = -1, fc = -1;
        {
          // rows with an uncovered zero; the first one in cyclic order starting at i0
          const unsigned rows = __ballot_sync(
              FULL_MASK, act && !((rowcov >> lane) & 1u) && (zrow & ~colcov) != 0u);
          if (rows) {
            const unsigned from = rows & ~((1u << i0) - 1u);
            fr = __ffs(from ? from : rows) - 1;
            const unsigned m = __shfl_sync(FULL_MASK, zrow, fr) & ~colcov;
            const unsigned low = m & ((1u << j0) - 1u);  // columns visited last in cyclic order
            fc = low ? (31 - __clz(low)) : (31 - __clz(m));
          }
        }
        if (fr < 0) {
          step = 6;
          break;
        }
        row = fr;
        col = fc;
        if (lane == row) prime = col;
        int sc = __shfl_sync(FULL_MASK, star, row);
        if (sc >= 0) {
          col = sc;
          rowcov |= 1u << row;
          colcov &= ~(1u << col);
        } else {
          z0r = row;
          z0c = col;
          step = 5;
          break;
        }
      }
    } else if (step == 5) {
      // alternating path from Z0: star in column -> prime in that row -> ...
      unsigned pathrows = 1u << z0r;
      int c = z0c;
      for (int it = 0; it < 2 * GRP_N; ++it) {
        unsigned m = __ballot_sync(FULL_MASK, act && star == c);
        if (!m) break;
        int sr = __ffs(m) - 1;
        pathrows |= 1u << sr;
        c = __shfl_sync(FULL_MASK, prime, sr);
      }
      if ((pathrows >> lane) & 1u) star = prime;  // primes on the path become stars
      rowcov = 0u;
      colcov = 0u;
      prime = -1;
      step = 3;
    } else {  // step 6
      double mn = 9.2233720368547758e18;  // float(sys.maxsize)
      if (act && !((colcov >> lane) & 1u)) {
        for (int i = 0; i < n; ++i)
          if (!((rowcov >> i) & 1u)) {
            double v = C[i * GRP_LD + lane];
            if (mn > v) mn = v;
          }
      }
      mn = warp_min_f64(mn);
      if (act) {
        const bool cc = (colcov >> lane) & 1u;
        for (int i = 0; i < n; ++i) {
          double v = C[i * GRP_LD + lane];
          if ((rowcov >> i) & 1u) v = __dadd_rn(v, mn);
          if (!cc) v = __dsub_rn(v, mn);
          C[i * GRP_LD + lane] = v;
        }
      }
      __syncwarp();
      rebuild_zero_masks();
      step = 4;
    }
  }
  return star;
}

__global__ void __launch_bounds__(GRP_WARPS * 32)
group_ae_kernel(GroupArgs a) {
  __shared__ GroupSmem smem[GRP_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int img = blockIdx.x * GRP_WARPS + warp;
  if (img >= a.N) return;
  GroupSmem& sm = smem[warp];
  const int J = a.J, K = a.K, T = a.T, Pmax = a.Pmax;
  const int width = 3 + T;
  float* ans = a.ans + (size_t)img * Pmax * J * width;
  float* keys = a.keys + (size_t)img * Pmax;
  int32_t* tlc = a.tl_count + (size_t)img * Pmax;
  const int TLCAP = J + 1;
  float* tl = a.tl + (size_t)img * Pmax * TLCAP * T;
  int P = 0;
  bool overflow = false;

  // find the person whose key equals `key` (first match), -1 if none
  auto find_person = [&](float key) -> int {
    for (int base = 0; base < P; base += 32) {
      int p = base + lane;
      bool eq = (p < P) && (keys[p] == key);
      unsigned m = __ballot_sync(FULL_MASK, eq);
      if (m) return base + __ffs(m) - 1;
    }
    return -1;
  };
  // joint_dict.setdefault(key, zeros)[idx] = joint ; tag_dict[key] = [tag]   (candidate r)
  auto new_or_reset = [&](int r, int idx) {
    const float key = sm.ctag[r][0];
    int p = find_person(key);
    if (p < 0) {
      if (P >= Pmax) {
        overflow = true;
        return;
      }
      p = P++;
      for (int e = lane; e < J * width; e += 32) ans[(size_t)p * J * width + e] = 0.0f;
      if (lane == 0) keys[p] = key;
    }
    __syncwarp();
    float* row = ans + ((size_t)p * J + idx) * width;
    if (lane == 0) {
      row[0] = sm.cx[r];
      row[1] = sm.cy[r];
      row[2] = sm.cval[r];
      tlc[p] = 1;
    }
    if (lane < T) {
      row[3 + lane] = sm.ctag[r][lane];
      tl[(size_t)p * TLCAP * T + lane] = sm.ctag[r][lane];
    }
    __syncwarp();
  };

  for (int idx = 0; idx < J; ++idx) {
    // ---- candidates of this joint above the detection threshold, order preserved
    float v = 0.0f;
    int ind = 0;
    if (lane < K) {
      v = a.val_k[((size_t)img * J + idx) * K + lane];
      ind = a.ind_k[((size_t)img * J + idx) * K + lane];
    }
    const bool keep = (lane < K) && ((double)v > a.det_thr);
    const unsigned km = __ballot_sync(FULL_MASK, keep);
    const int n_add = __popc(km);
    if (n_add == 0) continue;
    __syncwarp();
    if (keep) {
      const int r = __popc(km & ((1u << lane) - 1u));
      sm.cval[r] = v;
      sm.cx[r] = (float)(ind % a.W);
      sm.cy[r] = (float)(ind / a.W);
      for (int t = 0; t < T; ++t)
        sm.ctag[r][t] = a.tag_k[(((size_t)img * J + idx) * K + lane) * T + t];
    }
    __syncwarp();

    if (idx == 0 || P == 0) {
      for (int r = 0; r < n_add; ++r) new_or_reset(r, idx);
      continue;
    }

    const int n_grp = min(P, a.max_people);
    if (lane < n_grp) {
      const int cnt = tlc[lane];
      const float* list = tl + (size_t)lane * TLCAP * T;
      for (int t = 0; t < T; ++t) sm.mean[lane][t] = mean_tag_component(list, cnt, T, t);
    }
    __syncwarp();
    if (a.ignore_too_much && n_grp == a.max_people) continue;

    // ---- cost matrix (lane = column)
    const int n = max(n_add, n_grp);
    if (lane < n) {
      for (int i = 0; i < n; ++i) {
        double c;
        if (i < n_add) {
          if (lane < n_grp) {
            double d = tag_dist_f64(sm.ctag[i], sm.mean[lane], T);
            c = a.use_det_val ? __dsub_rn(__dmul_rn(rint(d), 100.0), (double)sm.cval[i]) : d;
          } else {
            c = 1e10;
          }
        } else {
          c = 0.0;
        }
        sm.C[i * GRP_LD + lane] = c;
      }
    }
    __syncwarp();
    const int star = munkres_warp(sm, n, lane, a.start_rule);

    // ---- apply the pairs in row order (group.py:81-94)
    for (int r = 0; r < n_add; ++r) {
      const int c = __shfl_sync(FULL_MASK, star, r);
      bool matched = false;
      if (c >= 0 && c < n_grp) {
        const double d = tag_dist_f64(sm.ctag[r], sm.mean[c], T);
        matched = d < a.tag_thr;
      }
      if (matched) {
        const int p = c;  // grouped_keys[c] is the c-th person in creation order
        float* row = ans + ((size_t)p * J + idx) * width;
        int cnt = tlc[p];
        __syncwarp();
        if (lane == 0) {
          row[0] = sm.cx[r];
          row[1] = sm.cy[r];
          row[2] = sm.cval[r];
          tlc[p] = cnt + 1;
        }
        if (lane < T) {
          row[3 + lane] = sm.ctag[r][lane];
          if (cnt < TLCAP) tl[((size_t)p * TLCAP + cnt) * T + lane] = sm.ctag[r][lane];
        }
        __syncwarp();
      } else {
        new_or_reset(r, idx);
      }
    }
  }
  if (lane == 0) {
    a.count[img] = P;
    if (overflow) atomicExch(a.overflow, 1);
  }
}

// adjust (group.py:181-200): one thread per (image, person, joint)
__global__ void adjust_kernel(float* __restrict__ ans, const int32_t* __restrict__ count,
                              const float* __restrict__ det, int N, int J, int H, int W, int T,
                              int Pmax) {
  const int width = 3 + T;
  const size_t total = (size_t)N * Pmax * J;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % J);
    const size_t np = i / J;
    const int p = (int)(np % Pmax);
    const int n = (int)(np / Pmax);
    if (p >= count[n]) continue;
    float* row = ans + i * width;
    if (!(row[2] > 0.0f)) continue;
    float col = row[0], rw = row[1];
    const int r = (int)rw, c = (int)col;  // int() truncation
    const float* plane = det + ((size_t)n * J + j) * H * W;
    const float right = plane[(size_t)r * W + min(c + 1, W - 1)];
    const float left = plane[(size_t)r * W + max(c - 1, 0)];
    col = (right > left) ? __fadd_rn(col, 0.25f) : __fsub_rn(col, 0.25f);
    const float down = plane[(size_t)min(r + 1, H - 1) * W + c];
    const float up = plane[(size_t)max(0, r - 1) * W + c];
    rw = (down > up) ? __fadd_rn(rw, 0.25f) : __fsub_rn(rw, 0.25f);
    row[0] = __fadd_rn(col, 0.5f);
    row[1] = __fadd_rn(rw, 0.5f);
  }
}

__global__ void scores_kernel(const float* __restrict__ ans, const int32_t* __restrict__ count,
                              float* __restrict__ scores, int N, int J, int T, int Pmax) {
  const int width = 3 + T;
  const size_t total = (size_t)N * Pmax;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % Pmax);
    const int n = (int)(i / Pmax);
    if (p >= count[n]) {
      scores[i] = 0.0f;
      continue;
    }
    const float* base = ans + i * J * width + 2;
    scores[i] = __fdiv_rn(pairwise_sum_f32(base, width, J), (float)J);
  }
}

}  // namespace brtpe

using namespace brtpe;

extern "C" size_t brtpe_group_workspace_bytes(int N, int J, int K, int T, int Pmax) {
  if (N <= 0 || J <= 0 || K <= 0 || T <= 0 || Pmax <= 0) return 0;
  size_t b = 0;
  b += align_up((size_t)N * Pmax * sizeof(float), 256);                 // keys
  b += align_up((size_t)N * Pmax * sizeof(int32_t), 256);               // tag-list counts
  b += align_up((size_t)N * Pmax * (J + 1) * T * sizeof(float), 256);   // tag lists
  return b;
}

extern "C" int brtpe_group_ae(const float* val_k, const int32_t* ind_k, const float* tag_k, int N,
                              int W, int T, const brtpe_decode_params* prm, float* ans,
                              int32_t* count, int32_t* overflow, int Pmax, void* workspace,
                              size_t workspace_bytes, void* stream) {
  BRTPE_CHECK_ARG(val_k && ind_k && tag_k && prm && ans && count && overflow,
                  "brtpe_group_ae: null argument");
  const int J = prm->num_joints, K = prm->max_num_people;
  BRTPE_CHECK_ARG(N > 0 && W > 0 && Pmax > 0, "brtpe_group_ae: bad sizes");
  BRTPE_CHECK_ARG(J >= 1 && J <= BRTPE_MAX_JOINTS, "brtpe_group_ae: num_joints=%d outside [1,%d]",
                  J, BRTPE_MAX_JOINTS);
  BRTPE_CHECK_ARG(K >= 1 && K <= BRTPE_MAX_GROUP_K,
                  "brtpe_group_ae: max_num_people=%d outside [1,%d] (one warp lane per candidate)",
                  K, BRTPE_MAX_GROUP_K);
  BRTPE_CHECK_ARG(T >= 1 && T <= BRTPE_MAX_TAG_DIMS, "brtpe_group_ae: T=%d outside [1,%d]", T,
                  BRTPE_MAX_TAG_DIMS);
  const size_t need = brtpe_group_workspace_bytes(N, J, K, T, Pmax);
  if (!workspace || workspace_bytes < need) {
    set_error("brtpe_group_ae: workspace %zu < %zu", workspace_bytes, need);
    return BRTPE_EWORKSPACE;
  }
  GroupArgs a;
  a.val_k = val_k; a.ind_k = ind_k; a.tag_k = tag_k; a.ans = ans; a.count = count;
  a.overflow = overflow;
  char* w = reinterpret_cast<char*>(workspace);
  a.keys = reinterpret_cast<float*>(w);
  w += align_up((size_t)N * Pmax * sizeof(float), 256);
  a.tl_count = reinterpret_cast<int32_t*>(w);
  w += align_up((size_t)N * Pmax * sizeof(int32_t), 256);
  a.tl = reinterpret_cast<float*>(w);
  a.N = N; a.J = J; a.K = K; a.T = T; a.W = W; a.Pmax = Pmax;
  a.max_people = K;
  a.det_thr = prm->detection_threshold;
  a.tag_thr = prm->tag_threshold;
  a.use_det_val = prm->use_detection_val;
  a.ignore_too_much = prm->ignore_too_much;
  a.start_rule = prm->munkres_start_rule;
  cudaStream_t st = (cudaStream_t)stream;
  BRTPE_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int32_t), st));
  group_ae_kernel<<<ceil_div(N, GRP_WARPS), GRP_WARPS * 32, 0, st>>>(a);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

extern "C" int brtpe_adjust(float* ans, const int32_t* count, const float* det, int N, int J,
                            int H, int W, int T, int Pmax, void* stream) {
  BRTPE_CHECK_ARG(ans && count && det && N > 0 && J > 0 && H > 0 && W > 0 && T >= 1 && Pmax > 0,
                  "brtpe_adjust: bad arguments");
  const size_t total = (size_t)N * Pmax * J;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  adjust_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ans, count, det, N, J, H, W, T, Pmax);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}

extern "C" int brtpe_scores(const float* ans, const int32_t* count, float* scores, int N, int J,
                            int T, int Pmax, void* stream) {
  BRTPE_CHECK_ARG(ans && count && scores && N > 0 && J > 0 && J < 128 && T >= 1 && Pmax > 0,
                  "brtpe_scores: bad arguments");
  const size_t total = (size_t)N * Pmax;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  scores_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ans, count, scores, N, J, T, Pmax);
  BRTPE_LAUNCH_CHECK();
  return BRTPE_OK;
}
