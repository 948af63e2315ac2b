// Experiment (not product code): cycle cost of the synchronisation primitives the warp-specialised
// convolution kernels are built from, measured in one CTA (optionally with spinning bystander
// warps), so that the pipeline skeleton can be budgeted.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_sync_costs exp_sync_costs.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

struct Params {
  CUtensorMap tmap;      // 64 x rows bf16, box 64 x 96 (12 KB)
  long long* out;        // [16]
  int reps, N, spinners;
};

// warps: 0 = measuring warp, 1 = partner (ping-pong tests), 2.. = bystanders spinning on a
// barrier that never completes until the end (like idle epilogue warps)
__global__ void __launch_bounds__(320) cost_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // 128 x 128 B
  uint8_t* sb = smem + 16384;               // 256 x 128 B
  uint8_t* ring = smem + 65536;             // 8 stages x 12 KB
  uint64_t* bars = (uint64_t*)(smem + 65536 + 8 * 12288);
  uint32_t* tslot = (uint32_t*)(bars + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;
  const int R = p.reps;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t stop_bar = smem_u32(&bars[63]);

  if (warp >= 2) {
    if (warp - 2 < p.spinners) mbar_wait(stop_bar, 0);      // bystander: spin until the end
  } else if (warp == 1) {
    // partner for the ping-pong test: wait ping[i&7] -> arrive pong[i&7]
    for (int i = 0; i < R; ++i) {
      mbar_wait(smem_u32(&bars[16 + (i & 7)]), (i >> 3) & 1);
      if (elect_one()) mbar_arrive(smem_u32(&bars[24 + (i & 7)]));
      __syncwarp();
    }
  } else {
    long long t0, t1;
    // 0: elect + syncwarp
    t0 = clock64();
    uint32_t sink = 0;
    for (int i = 0; i < R; ++i) { if (elect_one()) sink += i; __syncwarp(); }
    t1 = clock64();
    if (lane == 0) p.out[0] = (t1 - t0) / R + (sink == 0xffffffffu);
    // 1: arrive + try_wait on an already completed phase (same warp)
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) mbar_arrive(smem_u32(&bars[0]));
      __syncwarp();
      mbar_wait(smem_u32(&bars[0]), i & 1);
    }
    t1 = clock64();
    if (lane == 0) p.out[1] = (t1 - t0) / R;
    // 2: tcgen05.commit (nothing in flight) + wait for its arrival
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) umma_commit(smem_u32(&bars[1]));
      __syncwarp();
      mbar_wait(smem_u32(&bars[1]), i & 1);
    }
    t1 = clock64();
    if (lane == 0) p.out[2] = (t1 - t0) / R;
    // 3: one MMA (N) + commit + wait (MMA latency)
    const uint64_t ad = make_desc(smem_u32(sa)), bd = make_desc(smem_u32(sb));
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) { umma_f16(tmem, ad, bd, idesc, 0u); umma_commit(smem_u32(&bars[2])); }
      __syncwarp();
      mbar_wait(smem_u32(&bars[2]), i & 1);
    }
    t1 = clock64();
    if (lane == 0) p.out[3] = (t1 - t0) / R;
    // 4: 4 MMAs + commit per step, NOT waiting (issue cost incl. back-pressure), final wait
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) {
        for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        umma_commit(smem_u32(&bars[3 + (i & 7)]));
      }
      __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) p.out[4] = (t1 - t0) / R;
    for (int s = 0; s < 8; ++s) mbar_wait(smem_u32(&bars[3 + s]), ((R - 1 - s) >> 3) & 1);  // drain (R % 8 == 0)
    // 5: TMA issue of a 12 KB box + wait (load latency, L2 hit after the first)
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars[12]), 12288);
        tma_load_2d(smem_u32(ring), &p.tmap, smem_u32(&bars[12]), 0, 0);
      }
      __syncwarp();
      mbar_wait(smem_u32(&bars[12]), i & 1);
    }
    t1 = clock64();
    if (lane == 0) p.out[5] = (t1 - t0) / R;
    // 6: TMA issue throughput: 8-deep ring, wait for the slot 8 loads ago
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      const int s = i & 7;
      if (i >= 8) mbar_wait(smem_u32(&bars[32 + s]), ((i >> 3) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars[32 + s]), 12288);
        tma_load_2d(smem_u32(ring + s * 12288), &p.tmap, smem_u32(&bars[32 + s]), 0, (i * 96) & 1023);
      }
      __syncwarp();
    }
    for (int s = 0; s < 8; ++s) mbar_wait(smem_u32(&bars[32 + s]), ((R - 1 - s) >> 3) & 1);
    t1 = clock64();
    if (lane == 0) p.out[6] = (t1 - t0) / R;
    // 7: ping-pong with the partner warp through two mbarriers (hand-off latency x2)
    t0 = clock64();
    for (int i = 0; i < R; ++i) {
      if (elect_one()) mbar_arrive(smem_u32(&bars[16 + (i & 7)]));
      __syncwarp();
      mbar_wait(smem_u32(&bars[24 + (i & 7)]), (i >> 3) & 1);
    }
    t1 = clock64();
    if (lane == 0) p.out[7] = (t1 - t0) / R;
    // 8: single-thread variant of 4 (lane 0 only, no elect / syncwarp)
    if (lane == 0) {
      t0 = clock64();
      for (int i = 0; i < R; ++i) {
        for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        umma_commit(smem_u32(&bars[40 + (i & 7)]));
      }
      t1 = clock64();
      p.out[8] = (t1 - t0) / R;
    }
    __syncwarp();
    for (int s = 0; s < 8; ++s) mbar_wait(smem_u32(&bars[40 + s]), ((R - 1 - s) >> 3) & 1);
    if (elect_one()) mbar_arrive(stop_bar);
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  const int smem = 65536 + 8 * 12288 + 1024 + 1024;
  CK(cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  __nv_bfloat16* d; long long* dout;
  CK(cudaMalloc(&d, 2048 * 64 * 2)); CK(cudaMemset(d, 0, 2048 * 64 * 2));
  CK(cudaMalloc(&dout, 16 * 8));
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)f;
  Params p;
  cuuint64_t dim[2] = {64, 2048}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 96}; cuuint32_t es[2] = {1, 1};
  if (enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  p.out = dout; p.reps = 256;
  const char* names[9] = {"elect+syncwarp", "arrive+try_wait (ready)", "commit(empty)+wait", "1 MMA+commit+wait (latency)",
                          "4 MMA+commit issue (warp, elect)", "TMA 12KB issue+wait (latency)", "TMA 12KB ring-8 throughput",
                          "mbarrier ping-pong (2 hand-offs)", "4 MMA+commit issue (one thread)"};
  for (int spinners : {0, 4, 8}) {
    for (int N : {48, 96, 192}) {
      p.N = N; p.spinners = spinners;
      CK(cudaMemset(dout, 0, 16 * 8));
      cost_kernel<<<1, 320, smem>>>(p);
      CK(cudaDeviceSynchronize());
      long long h[16];
      CK(cudaMemcpy(h, dout, 16 * 8, cudaMemcpyDeviceToHost));
      printf("spinning bystander warps=%d N=%d\n", spinners, N);
      for (int i = 0; i < 9; ++i) printf("   %-36s %6lld cycles/step\n", names[i], h[i]);
    }
  }
  return 0;
}
