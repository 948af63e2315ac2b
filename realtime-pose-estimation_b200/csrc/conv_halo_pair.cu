// CTA-pair (tcgen05 cta_group::2) instantiation of the 3x3/s1 halo-tile convolution kernel; see
// conv_halo_impl.cuh and the dispatcher in conv_halo.cu.
#define HL_CG 2
#define HL_NAME(x) x##_cg2
#include "conv_halo_impl.cuh"
BRTPE_MBAR_DEBUG_EXPORT(brtpe_debug_mbar_halo2)
