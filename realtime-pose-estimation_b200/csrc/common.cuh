// Shared helpers for libbrtpe.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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
