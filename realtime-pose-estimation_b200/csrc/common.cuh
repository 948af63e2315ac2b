// Shared helpers for libbrtpe.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/brtpe.h"

namespace brtpe {

void set_error(const char* fmt, ...);

#define BRTPE_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::brtpe::set_error(__VA_ARGS__);        \
      return BRTPE_EINVAL;                    \
    }                                         \
  } while (0)

#define BRTPE_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::brtpe::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
                         __FILE__, __LINE__);                                         \
      return BRTPE_ECUDA;                                                             \
    }                                                                                 \
  } while (0)

#define BRTPE_LAUNCH_CHECK()                                                          \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::brtpe::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),  \
                         __FILE__, __LINE__);                                         \
      return BRTPE_ECUDA;                                                             \
    }                                                                                 \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int num_sms();

// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its
// predecessor in the stream (or graph lane) is still running; its prologue (barrier init, TMEM
// allocation, weight / bias loads -- nothing the predecessor writes) then overlaps the
// predecessor's tail, and pdl_wait() blocks until every prerequisite grid has completed and its
// memory operations are visible.  Every thread that reads or writes activation buffers calls
// pdl_wait() first.  Opt-in with BRTPE_PDL=1 (measured neutral to slightly slower on the 64-forward
// plan: with one 200 KB CTA per SM the successor's CTAs cannot become resident before ours leave);
// the plan executor switches it off again when a capture with programmatic edges fails.
bool pdl_enabled();
void pdl_set(bool on);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// launch with an optional cluster dimension and the PDL attribute
template <typename... KA, typename... A>
static inline cudaError_t launch_ex(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, int cluster_x, bool pdl, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 0) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}
#endif

constexpr unsigned FULL_MASK = 0xffffffffu;

// Monotonic map float -> uint32 (a < b  <=>  key(a) < key(b)), -0.0 folded onto +0.0.
__device__ __forceinline__ uint32_t float_order_key(float v) {
  v = v + 0.0f;  // -0.0 -> +0.0
  uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
// 64-bit selection key: larger value first, then smaller index first.
__device__ __forceinline__ unsigned long long make_sel_key(float v, uint32_t idx) {
  return ((unsigned long long)float_order_key(v) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ uint32_t sel_key_index(unsigned long long k) {
  return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}
__device__ __forceinline__ float sel_key_value(unsigned long long k) {
  return float_from_order_key((uint32_t)(k >> 32));
}

}  // namespace brtpe
