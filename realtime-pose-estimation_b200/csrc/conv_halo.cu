// Dispatcher of the 3x3/s1 tcgen05 halo-tile convolution: the kernel lives in
// conv_halo_impl.cuh and is compiled twice -- here as the single-CTA kernel (cta_group::1) and in
// conv_halo_pair.cu as the CTA-pair kernel (cta_group::2).  (A kernel that contains cta_group::2
// instructions can only be launched with an even cluster size, so the two modes cannot share one
// entry point.)
#define HL_CG 1
#define HL_NAME(x) x##_cg1
#include "conv_halo_impl.cuh"

namespace brtpe {

long long* g_halo_prof = nullptr;
int g_halo_prof_ctas = 0;

HaloConvPrepared* halo_conv_prepare_cg2(const brtpe_conv_desc* d, const void* in, const void* weights,
                                        void* out);
int halo_conv_launch_cg2(const HaloConvPrepared* P, const float* bias, const void* residual, void* out,
                         cudaStream_t st, const void* const* add_ptrs, void* out2);

bool halo_conv_supported(const brtpe_conv_desc* d) {
  if ((d->dtype != BRTPE_DT_BF16 && d->dtype != BRTPE_DT_BF16X2) || d->ntaps != 9 || d->out_scale != 1)
    return false;
  if (conv_is_split(d) && (!epi_split_ok(d) || d->in_ld / 2 < d->in_coff + d->Cin)) return false;
  if (d->out_oy || d->out_ox || d->Hout != d->Hm || d->Wout != d->Wm) return false;
  if (d->in_stride == 1) {
    if (d->Hm != d->Hin || d->Wm != d->Win) return false;
  } else if (d->in_stride == 2) {
    // 3x3 / stride 2 / pad 1 on an even-sized map: four pixel-parity planes per tile; needs the
    // chunk-aligned fast epilogue (one tile per CTA, shared by the two epilogue groups)
    if ((d->Hin | d->Win) & 1 || d->Hm * 2 != d->Hin || d->Wm * 2 != d->Win) return false;
    if (!epi_fast_ok(d) || (d->Cout_store + 15) / 16 * 16 > 256) return false;
  } else {
    return false;
  }
  for (int t = 0; t < 9; ++t)
    if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1) return false;
  if (d->Cin % 16 || d->in_ld % 8 || d->in_coff % 8 || d->out_ld % 8 || d->out_coff % 8 ||
      d->res_ld % 8 || d->res_coff % 8)
    return false;
  if ((d->Cout_store + 15) / 16 * 16 > 1024) return false;
  return halo_encode_fn() != nullptr;
}


// Pair mode (each SM reads and receives only half of the weight operand) pays off where the
// weight operand is a large share of the shared-memory traffic: Cout tiles of >= 96 channels.
// Measured in round 1 (profiles/r01e_halo_pair.md): 96ch +10 %, 48ch -25 % (the leader's single MMA
// thread then issued for both SMs).  BRTPE_HALO_CG=1 / 2 forces a mode.
HaloConvPrepared* halo_conv_prepare(const brtpe_conv_desc* d, const void* in, const void* weights,
                                    void* out) {
  if (!halo_conv_supported(d)) {
    set_error("halo conv: unsupported layer");
    return nullptr;
  }
  int cg = (d->Cout_store >= 96) ? 2 : 1;
  // Round 2, with two MMA-issuing warps (the single issuing thread was what made the pair mode 25 % slower
  // on narrow layers in round 1): stride-1 layers of fewer than 96 channels on large batches of tiles gain too
  // -- 64 images: 48 -> 48 @160^2 (+ residual) 0.0939 -> 0.0891 ms, @320^2 0.366 -> 0.353, 64 -> 64 @160^2
  // 0.138 -> 0.134, 256 -> 48 @160^2 0.400 -> 0.342; the forward 28.0 -> 27.5 ms per 64.  Stride 2: neutral;
  // 2 images @160^2: 0.0110 -> 0.0128 ms, so only from eight tiles per SM on.  BRTPE_HALO_CG_NARROW=0 = round-1 rule.
  {
    static int narrow = -1;
    if (narrow < 0) narrow = getenv("BRTPE_HALO_CG_NARROW") ? atoi(getenv("BRTPE_HALO_CG_NARROW")) : 1;
    const long tiles = (long)((d->Wm + 7) / 8) * ((d->Hm + 15) / 16) * d->N;
    if (cg == 1 && narrow && d->in_stride == 1 && tiles >= 8L * num_sms()) cg = 2;
  }
  if (getenv("BRTPE_HALO_CG")) {
    const int v = atoi(getenv("BRTPE_HALO_CG"));
    if (v == 1 || v == 2) cg = v;
  }
  HaloConvPrepared* P = nullptr;
  if (cg == 2) P = halo_conv_prepare_cg2(d, in, weights, out);
  if (!P) P = halo_conv_prepare_cg1(d, in, weights, out);     // layers the pair mode does not take
  return P;
}

void halo_conv_release(HaloConvPrepared* p) { delete p; }

int halo_conv_launch(const HaloConvPrepared* P, const float* bias, const void* residual, void* out,
                     cudaStream_t st, const void* const* add_ptrs, void* out2) {
  return P->p.cg == 2 ? halo_conv_launch_cg2(P, bias, residual, out, st, add_ptrs, out2)
                      : halo_conv_launch_cg1(P, bias, residual, out, st, add_ptrs, out2);
}

void halo_set_prof(long long* buf, int max_ctas) {
  g_halo_prof = buf;
  g_halo_prof_ctas = buf ? max_ctas : 0;
}

}  // namespace brtpe

extern "C" int brtpe_debug_halo_prof(void* buf, int max_ctas) {
  brtpe::halo_set_prof(reinterpret_cast<long long*>(buf), max_ctas);
  return BRTPE_OK;
}
BRTPE_MBAR_DEBUG_EXPORT(brtpe_debug_mbar_halo1)
