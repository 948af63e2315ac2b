// Experiment (not product code): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair,
// both operands in shared memory) for several N, against cta_group::1 on the same data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_umma_pair exp_umma_pair.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

struct Params { long long* cycles; int N, reps, pair, sbo; };

template <int PAIR>
__global__ void __launch_bounds__(128) rate_kernel(const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // 208 x 128 B
  uint8_t* sb = smem + 32768;               // 256 x 128 B
  uint64_t* bars = (uint64_t*)(smem + 65536);
  uint32_t* tslot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // bf16-ish
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(256u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;
  const uint32_t M = PAIR ? 256u : 128u;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((M >> 4) << 24);
  if (threadIdx.x == 0 && rank == 0) {
    const uint64_t ad = make_desc(smem_u32(sa), (uint32_t)p.sbo);
    const uint64_t bd = make_desc(smem_u32(sb), 1024);
    long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r)
      for (int k = 0; k < 4; ++k) {
        if (PAIR)
          asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(1u) : "memory");
      }
    if (PAIR)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "h"((uint16_t)1) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *p.cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

int main() {
  CK(cudaSetDevice(0));
  const int smem = 65536 + 2048;
  CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 8));
  for (int pair : {0, 1}) {
    for (int grid : {2, 148}) {
      for (int N : {48, 64, 96, 128, 192, 256}) {
        for (int sbo : {1024, 1280}) {
          Params p; p.cycles = dcyc; p.N = N; p.reps = 512; p.pair = pair; p.sbo = sbo;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = pair ? 1 : 0;
          if (pair) CK(cudaLaunchKernelEx(&cfg, rate_kernel<1>, p)); else CK(cudaLaunchKernelEx(&cfg, rate_kernel<0>, p));
          CK(cudaDeviceSynchronize());
          long long cyc = 0;
          CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
          printf("cta_group::%d grid=%3d N=%3d sbo=%4d: %.1f cycles/MMA (M=%d: ideal %d)\n", pair + 1, grid, N, sbo,
                 (double)cyc / (512 * 4), pair ? 256 : 128, N / 2);
        }
      }
    }
  }
  return 0;
}
