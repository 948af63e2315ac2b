// C ABI of the convolution engine + the execution plan (a recorded launch list that is
// replayed natively or as a CUDA graph).  One plan = one PoseHigherResolutionNet.forward for
// a fixed (batch chunk, H, W); the host (Python) records it once and replays it per chunk.
#include "conv_common.cuh"

#include <algorithm>
#include <vector>

namespace brtpe {

enum OpKind { OP_CONV = 0, OP_STEM = 1, OP_FUSE = 2, OP_TONCHW = 3, OP_IM2COL = 4, OP_AUX = 5 };
int aux_launch(int kind, const void* in0, const void* in1, const void* in2, void* out,
               const int32_t* ip, cudaStream_t st);   // student_ops.cu
constexpr int PLAN_MAX_LANES = 8;

struct Op {
  int kind;
  // conv
  brtpe_conv_desc d;
  const void* in;
  const void* w;
  const float* bias;
  const void* res;
  void* out;
  UmmaConvPrepared* umma;  // non-null: tcgen05 per-tap path
  HaloConvPrepared* halo;  // non-null: tcgen05 halo-tile path
  HaloConvPrepared* chain; // non-null: this conv and the next conv of its lane (op `chain_with`) run as ONE
                           // launch (halo chain mode: conv1 -> conv2 of a BasicBlock, L2-blocked)
  int chain_with;
  bool nop;                // the second conv of a chain: done by the first one's launch
  const void* add_ptrs[3]; // HRNet fuse addends of the epilogue (d.n_add of them)
  void* out2;              // second output (d.out2_ld > 0)
  // stem / tonchw / fuse
  int i[16];
  const void* terms[4];
  int shifts[4], lds[4];
  const float* stem_w;
  // scheduling: capture stream index + ops (indices) that must have completed before this one
  int lane;
  std::vector<int> deps;
};

static int choose_engine(const brtpe_conv_desc* d) {
  if (d->engine == BRTPE_ENGINE_FFMA) return BRTPE_ENGINE_FFMA;
  const bool halo_ok = halo_conv_supported(d);
  if (d->engine == BRTPE_ENGINE_UMMA_HALO) {
    if (!halo_ok) {
      set_error("conv: halo tcgen05 engine requested but the layer is not a 3x3/s1 bf16 conv");
      return -1;
    }
    return BRTPE_ENGINE_UMMA_HALO;
  }
  const char* why = nullptr;
  const bool ok = umma_conv_supported(d, &why);
  if (d->engine == BRTPE_ENGINE_AUTO && halo_ok) {
    // The halo engine's 8 x 16 pixel tiles waste MMA rows on small maps (20 x 20 -> 24 x 32: 48 %);
    // wide layers there are weight-operand bound anyway, so the per-tap engine with its free tile
    // geometry (20 x 6 pixels) wins: 3x3 384 -> 384 @ 20x20, 64 images: 0.075 -> 0.054 ms
    // (profiles/r01q_conv_384.md).  BRTPE_HALO_MIN_UTIL (percent, default 60; 0 = always halo).
    static int min_util = -1;
    if (min_util < 0) {
      const char* e = getenv("BRTPE_HALO_MIN_UTIL");
      min_util = e ? atoi(e) : 60;
    }
    const long covered = (long)((d->Hm + 15) / 16 * 16) * ((d->Wm + 7) / 8 * 8);
    const bool wasteful = 100L * d->Hm * d->Wm < (long)min_util * covered;
    if (d->in_stride == 2) {
      // stride 2 (profiles/r02_halo_s2.md): the per-tap engine loads every input pixel 2.25 times; the
      // parity-plane halo tile loads it once.  Measured at 64 images: 48 -> 96 @80x80 0.064 -> 0.049 ms,
      // 48 -> 48 @80x80 0.052 -> 0.044 (both then at 70-75 % of the HBM roofline: the input is a
      // 160 x 160 map), 64 -> 64 @160x160 +7 %, 256 -> 96 +2 %; wide layers on small maps lose
      // (96 -> 192 @40x40 0.047 -> 0.065 ms: 72 KB activation stages leave two weight stages).
      // BRTPE_HALO_S2: 0 = never, 1 = by this rule (default), 2 = whenever the layer is supported.
      static int s2 = -1;
      if (s2 < 0) {
        const char* e = getenv("BRTPE_HALO_S2");
        s2 = e ? atoi(e) : 1;
      }
      const bool wins = d->Cout_store <= 96 && (d->Cout_store <= 64 || d->Hm >= 64);
      if (s2 == 2 || (s2 == 1 && wins && !(wasteful && ok))) return BRTPE_ENGINE_UMMA_HALO;
    } else if (!(wasteful && ok && d->Cin >= 256 && d->Cout_store >= 256)) {
      return BRTPE_ENGINE_UMMA_HALO;
    }
  }
  if (d->engine == BRTPE_ENGINE_UMMA) {
    if (!ok) {
      set_error("conv: tcgen05 engine requested but unsupported: %s", why ? why : "?");
      return -1;
    }
    return BRTPE_ENGINE_UMMA;
  }
  return ok ? BRTPE_ENGINE_UMMA : BRTPE_ENGINE_FFMA;
}

static int run_op(const Op& op, cudaStream_t st) {
  switch (op.kind) {
    case OP_CONV:
      if (op.nop) return BRTPE_OK;
      if (op.chain) return BRTPE_EINVAL;     // (run through run_plan_op, which knows the partner op)
      if (op.halo) return halo_conv_launch(op.halo, op.bias, op.res, op.out, st, op.add_ptrs, op.out2);
      if (op.umma) return umma_conv_launch(op.umma, op.bias, op.res, op.out, st, op.add_ptrs, op.out2);
      if (op.d.n_add > 0 || op.d.out2_ld > 0) {
        set_error("conv: fuse addends need a tcgen05 engine");
        return BRTPE_EINVAL;
      }
      return conv_ffma_launch(&op.d, op.in, op.w, op.bias, op.res, op.out, st);
    case OP_STEM:
      return stem_conv1_launch(op.in, op.i[0], op.i[1], op.i[2], op.i[3], op.stem_w, op.bias,
                               op.i[4], op.out, op.i[5], st);
    case OP_FUSE:
      return fuse_sum_launch(op.i[0], op.i[1], op.terms, op.shifts, op.lds, op.i[2], op.i[3],
                             op.i[4], op.i[5], op.out, op.i[6], op.i[7], st);
    case OP_TONCHW:
      return nhwc_to_nchw_launch(op.i[0], op.in, op.i[1], op.i[2], op.i[3], op.i[4], op.i[5],
                                 op.i[6], op.out, op.i[7], st);
    case OP_IM2COL:
      return stem_im2col_launch(op.in, op.i[0], op.i[1], op.i[2], op.i[3], op.out, st);
    case OP_AUX:
      return aux_launch(op.i[15], op.terms[0], op.terms[1], op.terms[2], op.out, op.i, st);
  }
  return BRTPE_EINVAL;
}

}  // namespace brtpe

using namespace brtpe;

struct brtpe_plan {
  std::vector<Op> ops;
  bool chains_done = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t lanes[PLAN_MAX_LANES] = {nullptr};
  ~brtpe_plan() {
    for (auto& l : lanes)
      if (l) cudaStreamDestroy(l);
    for (auto& op : ops) {
      if (op.umma) umma_conv_release(op.umma);
      if (op.halo) halo_conv_release(op.halo);
      if (op.chain) halo_conv_release(op.chain);
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
};

// One op of a plan (a chained pair runs through its first op).
static int run_plan_op(const brtpe_plan* pl, size_t i, cudaStream_t st) {
  const Op& op = pl->ops[i];
  if (op.kind == OP_CONV && op.chain && !op.nop) {
    const Op& o2 = pl->ops[op.chain_with];
    return halo_chain_launch(op.chain, op.bias, o2.bias, o2.res, o2.out, st);
  }
  return run_op(op, st);
}

// Chain pass (once, before the first run / capture): conv1 -> conv2 (+ residual) of a BasicBlock whose
// tensors do not fit the L2 become one launch of the halo engine's chain mode.  Op j = the next op of
// op i's lane; it must read what i writes, and everything else it waits for must precede i (the pair
// runs at i's position).  BRTPE_CHAIN: 0 off (default: measured slower, profiles/r02_chain.md), 1 tensors of
// >= 48 MB only, 2 every chainable pair.
static void fuse_chains(brtpe_plan* pl) {
  if (pl->chains_done) return;
  pl->chains_done = true;
  const char* e = getenv("BRTPE_CHAIN");       // read per plan (tests switch it between plans)
  const int mode = e ? atoi(e) : 0;
  if (mode <= 0) return;
  const int n = (int)pl->ops.size();
  for (int i = 0; i < n; ++i) {
    Op& a = pl->ops[i];
    if (a.kind != OP_CONV || !a.halo || a.nop || a.chain || a.res || a.d.n_add || a.d.out2_ld) continue;
    int j = -1;
    for (int k = i + 1; k < n; ++k)
      if (pl->ops[k].lane == a.lane) { j = k; break; }
    if (j < 0) continue;
    Op& b = pl->ops[j];
    if (b.kind != OP_CONV || !b.halo || b.in != a.out || !b.res || b.out == a.out || b.out == a.in) continue;
    const double tensor_bytes = (double)a.d.N * a.d.Hm * a.d.Wm * a.d.out_ld * 2.0;
    if (mode == 1 && tensor_bytes < 48e6) continue;
    bool ok = true;
    for (int dep : b.deps)
      if (dep >= i) ok = false;               // an op between i and j (other lane) that j waits for
    if (!ok) continue;
    HaloConvPrepared* c = halo_chain_prepare(&a.d, a.in, a.w, a.out, &b.d, b.w, b.out);
    if (!c) continue;                         // not a chainable pair: two launches as before
    a.chain = c;
    a.chain_with = j;
    b.nop = true;
    for (int dep : b.deps)
      if (std::find(a.deps.begin(), a.deps.end(), dep) == a.deps.end()) a.deps.push_back(dep);
  }
}

extern "C" int brtpe_conv_chain_run(const brtpe_conv_desc* d0, const void* in, const void* w0,
                                    const float* bias0, void* mid, const brtpe_conv_desc* d1,
                                    const void* w1, const float* bias1, const void* residual,
                                    void* out, void* stream) {
  int rc = conv_validate(d0);
  if (rc) return rc;
  rc = conv_validate(d1);
  if (rc) return rc;
  BRTPE_CHECK_ARG(in && w0 && mid && w1 && residual && out, "brtpe_conv_chain_run: null tensor");
  BRTPE_CHECK_ARG(halo_conv_supported(d0) && halo_conv_supported(d1),
                  "brtpe_conv_chain_run: both layers must be 3x3 / stride-1 bf16 convs of the halo engine");
  HaloConvPrepared* c = halo_chain_prepare(d0, in, w0, mid, d1, w1, out);
  if (!c) return BRTPE_EINVAL;
  rc = halo_chain_launch(c, bias0, bias1, residual, out, (cudaStream_t)stream);
  // the prepared state owns device counters the launch uses: keep it until the stream has drained
  if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = BRTPE_ECUDA;
  halo_conv_release(c);
  return rc;
}

extern "C" int brtpe_conv_run(const brtpe_conv_desc* d, const void* in, const void* weights,
                              const float* bias, const void* residual, void* out, void* stream) {
  int rc = conv_validate(d);
  if (rc) return rc;
  BRTPE_CHECK_ARG(in && weights && out, "brtpe_conv_run: null tensor");
  const int eng = choose_engine(d);
  if (eng < 0) return BRTPE_EINVAL;
  if (eng == BRTPE_ENGINE_UMMA_HALO) {
    HaloConvPrepared* p = halo_conv_prepare(d, in, weights, out);
    if (!p) return BRTPE_ECUDA;
    rc = halo_conv_launch(p, bias, residual, out, (cudaStream_t)stream);
    halo_conv_release(p);
    return rc;
  }
  if (eng == BRTPE_ENGINE_UMMA) {
    UmmaConvPrepared* p = umma_conv_prepare(d, in, weights);
    if (!p) return BRTPE_ECUDA;
    rc = umma_conv_launch(p, bias, residual, out, (cudaStream_t)stream);
    umma_conv_release(p);
    return rc;
  }
  return conv_ffma_launch(d, in, weights, bias, residual, out, (cudaStream_t)stream);
}

extern "C" int brtpe_conv_run_fused(const brtpe_conv_desc* d, const void* in, const void* weights,
                                    const float* bias, const void* residual, void* out,
                                    const void* const* add_ptrs, void* out2, void* stream) {
  int rc = conv_validate(d);
  if (rc) return rc;
  BRTPE_CHECK_ARG(in && weights && out, "brtpe_conv_run_fused: null tensor");
  BRTPE_CHECK_ARG(d->n_add >= 0 && d->n_add <= 3 && (d->n_add == 0 || add_ptrs),
                  "brtpe_conv_run_fused: bad addends");
  const int eng = choose_engine(d);
  if (eng < 0) return BRTPE_EINVAL;
  if (eng == BRTPE_ENGINE_UMMA_HALO) {
    HaloConvPrepared* p = halo_conv_prepare(d, in, weights, out);
    if (!p) return BRTPE_ECUDA;
    rc = halo_conv_launch(p, bias, residual, out, (cudaStream_t)stream, add_ptrs, out2);
    halo_conv_release(p);
    return rc;
  }
  if (eng == BRTPE_ENGINE_UMMA) {
    UmmaConvPrepared* p = umma_conv_prepare(d, in, weights);
    if (!p) return BRTPE_ECUDA;
    rc = umma_conv_launch(p, bias, residual, out, (cudaStream_t)stream, add_ptrs, out2);
    umma_conv_release(p);
    return rc;
  }
  set_error("brtpe_conv_run_fused: fuse addends need a tcgen05 engine");
  return BRTPE_EINVAL;
}

extern "C" int brtpe_plan_set_conv_fuse(brtpe_plan* pl, const void* const* add_ptrs, void* out2) {
  BRTPE_CHECK_ARG(pl && !pl->ops.empty() && pl->ops.back().kind == OP_CONV,
                  "brtpe_plan_set_conv_fuse: the last op is not a conv");
  Op& op = pl->ops.back();
  BRTPE_CHECK_ARG(op.d.n_add == 0 || add_ptrs, "brtpe_plan_set_conv_fuse: null addends");
  for (int k = 0; k < op.d.n_add; ++k) {
    BRTPE_CHECK_ARG(add_ptrs[k], "brtpe_plan_set_conv_fuse: addend %d is null", k);
    op.add_ptrs[k] = add_ptrs[k];
  }
  BRTPE_CHECK_ARG((op.d.out2_ld > 0) == (out2 != nullptr),
                  "brtpe_plan_set_conv_fuse: out2 must be given exactly when out2_ld > 0");
  op.out2 = out2;
  return BRTPE_OK;
}

extern "C" int brtpe_conv_select_engine(const brtpe_conv_desc* d) {
  int rc = conv_validate(d);
  if (rc) return rc;
  return choose_engine(d);
}

extern "C" int brtpe_stem_conv1(const void* img, int img_is_half, int N, int H, int W,
                                const float* w, const float* bias, int Cout, void* out,
                                int out_dtype, void* stream) {
  return stem_conv1_launch(img, img_is_half, N, H, W, w, bias, Cout, out, out_dtype,
                           (cudaStream_t)stream);
}

extern "C" int brtpe_fuse_sum(int dtype, int nterms, const void* const* terms,
                              const int32_t* shifts, const int32_t* term_ld, int N, int H, int W,
                              int C, void* out, int out_ld, int relu, void* stream) {
  return fuse_sum_launch(dtype, nterms, terms, shifts, term_ld, N, H, W, C, out, out_ld, relu,
                         (cudaStream_t)stream);
}

extern "C" int brtpe_nhwc_to_nchw(int dtype, const void* src, int N, int H, int W, int C, int ld,
                                  int coff, void* dst, int dst_is_half, void* stream) {
  return nhwc_to_nchw_launch(dtype, src, N, H, W, C, ld, coff, dst, dst_is_half,
                             (cudaStream_t)stream);
}

extern "C" brtpe_plan* brtpe_plan_create(void) { return new brtpe_plan(); }
extern "C" void brtpe_plan_destroy(brtpe_plan* p) { delete p; }

extern "C" int brtpe_plan_add_conv(brtpe_plan* pl, const brtpe_conv_desc* d, const void* in,
                                   const void* weights, const float* bias, const void* residual,
                                   void* out) {
  BRTPE_CHECK_ARG(pl, "brtpe_plan_add_conv: null plan");
  int rc = conv_validate(d);
  if (rc) return rc;
  BRTPE_CHECK_ARG(in && weights && out, "brtpe_plan_add_conv: null tensor");
  const int eng = choose_engine(d);
  if (eng < 0) return BRTPE_EINVAL;
  Op op{};
  op.kind = OP_CONV;
  op.d = *d;
  op.in = in; op.w = weights; op.bias = bias; op.res = residual; op.out = out;
  op.umma = nullptr;
  op.halo = nullptr;
  op.chain = nullptr;
  op.chain_with = -1;
  op.nop = false;
  if (eng == BRTPE_ENGINE_UMMA_HALO) {
    op.halo = halo_conv_prepare(d, in, weights, out);
    if (!op.halo) return BRTPE_ECUDA;
  } else if (eng == BRTPE_ENGINE_UMMA) {
    op.umma = umma_conv_prepare(d, in, weights);
    if (!op.umma) return BRTPE_ECUDA;
  }
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_add_stem(brtpe_plan* pl, const void* img, int img_is_half, int N, int H,
                                   int W, const float* w, const float* bias, int Cout, void* out,
                                   int out_dtype) {
  BRTPE_CHECK_ARG(pl && img && w && out, "brtpe_plan_add_stem: null argument");
  Op op{};
  op.kind = OP_STEM;
  op.in = img; op.stem_w = w; op.bias = bias; op.out = out;
  op.i[0] = img_is_half; op.i[1] = N; op.i[2] = H; op.i[3] = W; op.i[4] = Cout; op.i[5] = out_dtype;
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_add_fuse(brtpe_plan* pl, int dtype, int nterms, const void* const* terms,
                                   const int32_t* shifts, const int32_t* term_ld, int N, int H,
                                   int W, int C, void* out, int out_ld, int relu) {
  BRTPE_CHECK_ARG(pl && terms && shifts && term_ld && out && nterms >= 1 && nterms <= 4,
                  "brtpe_plan_add_fuse: bad arguments");
  Op op{};
  op.kind = OP_FUSE;
  for (int k = 0; k < nterms; ++k) {
    op.terms[k] = terms[k];
    op.shifts[k] = shifts[k];
    op.lds[k] = term_ld[k];
  }
  op.out = out;
  op.i[0] = dtype; op.i[1] = nterms; op.i[2] = N; op.i[3] = H; op.i[4] = W; op.i[5] = C;
  op.i[6] = out_ld; op.i[7] = relu;
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_add_nhwc_to_nchw(brtpe_plan* pl, int dtype, const void* src, int N, int H,
                                           int W, int C, int ld, int coff, void* dst,
                                           int dst_is_half) {
  BRTPE_CHECK_ARG(pl && src && dst, "brtpe_plan_add_nhwc_to_nchw: null argument");
  Op op{};
  op.kind = OP_TONCHW;
  op.in = src; op.out = dst;
  op.i[0] = dtype; op.i[1] = N; op.i[2] = H; op.i[3] = W; op.i[4] = C; op.i[5] = ld; op.i[6] = coff;
  op.i[7] = dst_is_half;
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_add_stem_im2col(brtpe_plan* pl, const void* img, int img_is_half, int N,
                                          int H, int W, void* out) {
  BRTPE_CHECK_ARG(pl && img && out, "brtpe_plan_add_stem_im2col: null argument");
  Op op{};
  op.kind = OP_IM2COL;
  op.in = img; op.out = out;
  op.i[0] = img_is_half; op.i[1] = N; op.i[2] = H; op.i[3] = W;
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_add_aux(brtpe_plan* pl, int kind, const void* in0, const void* in1,
                                  const void* in2, void* out, const int32_t* iparams, int nparams) {
  BRTPE_CHECK_ARG(pl && in0 && out && iparams && nparams >= 4 && nparams <= 12,
                  "brtpe_plan_add_aux: bad arguments");
  Op op{};
  op.kind = OP_AUX;
  op.terms[0] = in0; op.terms[1] = in1; op.terms[2] = in2;
  op.out = out;
  for (int i = 0; i < nparams; ++i) op.i[i] = iparams[i];
  op.i[15] = kind;
  pl->ops.push_back(op);
  return BRTPE_OK;
}

extern "C" int brtpe_plan_set_sched(brtpe_plan* pl, int lane, const int32_t* deps, int ndeps) {
  BRTPE_CHECK_ARG(pl && !pl->ops.empty(), "brtpe_plan_set_sched: no op to annotate");
  BRTPE_CHECK_ARG(lane >= 0 && lane < PLAN_MAX_LANES, "brtpe_plan_set_sched: lane %d outside [0,%d)",
                  lane, PLAN_MAX_LANES);
  BRTPE_CHECK_ARG(ndeps == 0 || deps, "brtpe_plan_set_sched: null deps");
  BRTPE_CHECK_ARG(!pl->exec, "brtpe_plan_set_sched: plan already captured");
  Op& op = pl->ops.back();
  const int self = (int)pl->ops.size() - 1;
  op.lane = lane;
  op.deps.clear();
  for (int k = 0; k < ndeps; ++k) {
    BRTPE_CHECK_ARG(deps[k] >= 0 && deps[k] < self, "brtpe_plan_set_sched: dependency %d of op %d is not an earlier op",
                    deps[k], self);
    op.deps.push_back(deps[k]);
  }
  return BRTPE_OK;
}

extern "C" int brtpe_stem_im2col(const void* img, int img_is_half, int N, int H, int W, void* out,
                                 void* stream) {
  return stem_im2col_launch(img, img_is_half, N, H, W, out, (cudaStream_t)stream);
}

extern "C" int brtpe_plan_num_ops(const brtpe_plan* pl) { return pl ? (int)pl->ops.size() : 0; }

extern "C" double brtpe_plan_conv_flops(const brtpe_plan* pl) {
  double f = 0;
  if (pl)
    for (auto& op : pl->ops)
      if (op.kind == OP_CONV) f += conv_flops(&op.d);
  return f;
}

extern "C" int brtpe_plan_run(brtpe_plan* pl, void* stream) {
  BRTPE_CHECK_ARG(pl, "brtpe_plan_run: null plan");
  fuse_chains(pl);
  static int sync_each = -1;               // BRTPE_PLAN_SYNC=1 (debug): synchronise after every op and name
  if (sync_each < 0) sync_each = getenv("BRTPE_PLAN_SYNC") ? atoi(getenv("BRTPE_PLAN_SYNC")) : 0;   // the one that faults
  for (size_t i = 0; i < pl->ops.size(); ++i) {
    int rc = run_plan_op(pl, i, (cudaStream_t)stream);
    if (rc) return rc;
    if (sync_each) {
      cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
      if (e != cudaSuccess) {
        const Op& op = pl->ops[i];
        set_error("plan op %zu (kind %d, engine %s, N %d %dx%d Cin %d Cout %d taps %d stride %d dtype %d res %d) failed: %s",
                  i, op.kind, op.halo ? (op.halo ? "halo" : "") : (op.umma ? "umma" : "other"), op.d.N, op.d.Hm,
                  op.d.Wm, op.d.Cin, op.d.Cout, op.d.ntaps, op.d.in_stride, op.d.dtype, op.res != nullptr,
                  cudaGetErrorString(e));
        return BRTPE_ECUDA;
      }
    }
  }
  return BRTPE_OK;
}

// Captures the op list into a CUDA graph.  Ops are issued in recording order, each on the
// capture stream of its lane; an op that depends on an op of another lane waits on an event
// recorded right after that op, so independent lanes (the HRNet branches) become parallel
// branches of the graph.  Lane 0 is the origin of the capture; every other lane forks from
// it at its first op and joins it again at the end.
static int capture_graph(brtpe_plan* pl) {
  fuse_chains(pl);
  const int n = (int)pl->ops.size();
  int nl = 1;
  for (auto& op : pl->ops) nl = std::max(nl, op.lane + 1);
  for (int l = 0; l < nl; ++l)
    if (!pl->lanes[l]) BRTPE_CUDA(cudaStreamCreateWithFlags(&pl->lanes[l], cudaStreamNonBlocking));
  std::vector<char> need_event(n, 0);
  for (int i = 0; i < n; ++i)
    for (int d : pl->ops[i].deps)
      if (d >= 0 && d < i && pl->ops[d].lane != pl->ops[i].lane) need_event[d] = 1;
  std::vector<cudaEvent_t> ev(n, nullptr);
  std::vector<cudaEvent_t> extra;
  auto new_event = [&](cudaEvent_t* e) {
    cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    if (r == cudaSuccess) extra.push_back(*e);
    return r;
  };
  int rc = BRTPE_OK;
  cudaError_t ce = cudaStreamBeginCapture(pl->lanes[0], cudaStreamCaptureModeThreadLocal);
  if (ce != cudaSuccess) {
    set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(ce));
    return BRTPE_ECUDA;
  }
  std::vector<char> joined(nl, 0);
  joined[0] = 1;
  cudaEvent_t fork = nullptr;
  for (int i = 0; i < n && !rc; ++i) {
    const Op& op = pl->ops[i];
    cudaStream_t st = pl->lanes[op.lane];
    bool waited = false;
    for (int d : op.deps) {
      if (d < 0 || d >= i || pl->ops[d].lane == op.lane) continue;
      if (cudaStreamWaitEvent(st, ev[d], 0) != cudaSuccess) rc = BRTPE_ECUDA;
      waited = true;
    }
    if (!joined[op.lane]) {
      if (!waited) {   // no cross-lane dependency yet: fork from the start of the capture
        if (!fork) {
          // an event recorded on lane 0 before anything else would be ideal; recording it
          // now (after earlier lane-0 ops) is still correct, only less parallel
          if (new_event(&fork) != cudaSuccess || cudaEventRecord(fork, pl->lanes[0]) != cudaSuccess)
            rc = BRTPE_ECUDA;
        }
        if (!rc && cudaStreamWaitEvent(st, fork, 0) != cudaSuccess) rc = BRTPE_ECUDA;
      }
      joined[op.lane] = 1;
    }
    if (rc) break;
    rc = run_plan_op(pl, (size_t)i, st);
    if (!rc && need_event[i]) {
      if (new_event(&ev[i]) != cudaSuccess || cudaEventRecord(ev[i], st) != cudaSuccess) rc = BRTPE_ECUDA;
    }
  }
  for (int l = 1; l < nl && !rc; ++l) {
    if (!joined[l]) continue;
    cudaEvent_t e;
    if (new_event(&e) != cudaSuccess || cudaEventRecord(e, pl->lanes[l]) != cudaSuccess ||
        cudaStreamWaitEvent(pl->lanes[0], e, 0) != cudaSuccess)
      rc = BRTPE_ECUDA;
  }
  cudaGraph_t g = nullptr;
  ce = cudaStreamEndCapture(pl->lanes[0], &g);
  for (auto e : extra) cudaEventDestroy(e);
  if (rc) {
    if (g) cudaGraphDestroy(g);
    if (rc == BRTPE_ECUDA) set_error("plan capture: a CUDA stream/event call failed: %s",
                                     cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  if (ce != cudaSuccess) {
    set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
    return BRTPE_ECUDA;
  }
  pl->graph = g;
  BRTPE_CUDA(cudaGraphInstantiate(&pl->exec, g, 0));
  return BRTPE_OK;
}

extern "C" int brtpe_plan_graph_launch(brtpe_plan* pl, void* stream) {
  BRTPE_CHECK_ARG(pl, "brtpe_plan_graph_launch: null plan");
  if (!pl->exec) {
    // captured on private streams: the caller's stream may be the legacy default stream,
    // which cannot be captured; the instantiated graph is then launched into `stream`.
    int rc = capture_graph(pl);
    if (rc && pdl_enabled()) {
      // programmatic (PDL) edges could not be captured / instantiated here: plain edges instead
      cudaGetLastError();
      pdl_set(false);
      rc = capture_graph(pl);
    }
    if (rc) return rc;
  }
  BRTPE_CUDA(cudaGraphLaunch(pl->exec, (cudaStream_t)stream));
  return BRTPE_OK;
}

extern "C" int brtpe_plan_profile(brtpe_plan* pl, void* stream, float* ms_out, int32_t* kinds_out,
                                  double* flops_out) {
  BRTPE_CHECK_ARG(pl && ms_out, "brtpe_plan_profile: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  fuse_chains(pl);
  const size_t n = pl->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) BRTPE_CUDA(cudaEventCreate(&e));
  int rc = BRTPE_OK;
  BRTPE_CUDA(cudaEventRecord(ev[0], st));
  for (size_t i = 0; i < n && !rc; ++i) {
    rc = run_plan_op(pl, i, st);
    cudaEventRecord(ev[i + 1], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) {
    set_error("plan_profile: %s", cudaGetErrorString(e));
    rc = BRTPE_ECUDA;
  }
  for (size_t i = 0; i < n && !rc; ++i) {
    cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
    const Op& op = pl->ops[i];
    // kinds: 0 per-tap tcgen05, 1 CUDA cores, 3 halo tcgen05, 2 other, 4 second conv of a chained pair
    // (no launch of its own: its FLOPs are reported with the first conv)
    if (kinds_out) kinds_out[i] = op.kind == OP_CONV ? (op.nop ? 4 : (op.halo ? 3 : (op.umma ? 0 : 1))) : 2;
    if (flops_out) {
      double f = (op.kind == OP_CONV && !op.nop) ? conv_flops(&op.d) : 0.0;
      if (op.kind == OP_CONV && op.chain) f += conv_flops(&pl->ops[op.chain_with].d);
      flops_out[i] = f;
    }
  }
  for (auto& evt : ev) cudaEventDestroy(evt);
  return rc;
}
