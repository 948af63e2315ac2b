// Experiment (not product code): tcgen05.mma issue rate for the exact instruction sequences of
// conv_halo_kernel's 48-channel layers (N = 48, K = 48 = 3 x K16 per tap, nine taps through shifted
// descriptors, two pixel tiles = two accumulators per work item), against the plain back-to-back
// loop that profiles/r01_exp_umma_shift.md measured at 46.1 cycles per MMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_umma_seq exp_umma_seq.cu
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
    if (!done && ++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void umma(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma64(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t dlo(uint32_t addr) { return ((addr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ constexpr uint32_t dhi(uint32_t sbo) { return (sbo >> 4) | (1u << 14) | (2u << 29); }

struct Params { int mode, N, items, k16, dstride, tps, flags; long long* cycles; };

constexpr int A_TILE = 23552;

template <int MODE, int K16>
__global__ void __launch_bounds__(320) seq_kernel(const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                        // two halo tiles
  uint8_t* sb = smem + 2 * A_TILE + 1024;    // nine taps x N rows x 128 B
  sb = (uint8_t*)(((uintptr_t)sb + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + 200 * 1024);
  uint32_t* tslot = (uint32_t*)(bars + 16);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // finite bf16 data
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), 1);
    mbar_init(smem_u32(&bars[5]), 1);
    for (int i = 8; i < 12; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[5])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;
  const int N = p.N;
  constexpr int k16 = K16;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const bool dualw = (p.flags & 8) != 0;
  if (warp == 1 || (warp == 2 && dualw)) {
    const int w = warp - 1, step = dualw ? 2 : 1;
    const uint32_t a0 = dlo(smem_u32(sa)), a1 = a0 + (A_TILE >> 4);
    const uint32_t b0 = dlo(smem_u32(sb));
    const uint32_t btap = (uint32_t)((N * 128) >> 4);
    constexpr uint32_t AH = dhi(1280), BH = dhi(1024), AH0 = dhi(1024);
    long long t0 = clock64();
    int nmma = 0;
#pragma unroll 1
    for (int it = w; it < p.items; it += step) {
      const uint32_t d0 = tmem + (uint32_t)((it & 1) * 2 * p.dstride), d1 = d0 + (uint32_t)p.dstride;
      if (p.flags & 4) {   // the product's per-item waits (already satisfied) + fence
        mbar_wait(smem_u32(&bars[5]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        mbar_wait(smem_u32(&bars[5]), 0);
      }
      if (elect_one()) {
      if (MODE == 0) {            // same operands, one accumulator (the r01 micro-benchmark)
        for (int r = 0; r < 54 / k16; ++r)
          _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(d0, a0 + 2 * k, AH0, b0 + 2 * k, BH, idesc, (r | k) ? 1u : 0u); ++nmma; }
      } else if (MODE == 1) {     // nine shifted taps, one accumulator, two passes
        for (int pass = 0; pass < 2; ++pass)
          _Pragma("unroll") for (int t = 0; t < 9; ++t) {
            const uint32_t off = (uint32_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
            _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(d0, a0 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (pass | t | k) ? 1u : 0u); ++nmma; }
          }
      } else if (MODE == 2) {     // the kernel's order: per tap, tile 0 then tile 1
        _Pragma("unroll") for (int t = 0; t < 9; ++t) {
          const uint32_t off = (uint32_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
          _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(d0, a0 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (t | k) ? 1u : 0u); ++nmma; }
          _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(d1, a1 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (t | k) ? 1u : 0u); ++nmma; }
        }
      } else if (MODE == 3) {     // tile-major: all taps of tile 0, then all taps of tile 1
        for (int tile = 0; tile < 2; ++tile)
          _Pragma("unroll") for (int t = 0; t < 9; ++t) {
            const uint32_t off = (uint32_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
            _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(tile ? d1 : d0, (tile ? a1 : a0) + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (t | k) ? 1u : 0u); ++nmma; }
          }
      } else if (MODE == 4) {     // alternate accumulators on EVERY instruction
        _Pragma("unroll") for (int t = 0; t < 9; ++t) {
          const uint32_t off = (uint32_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
          _Pragma("unroll") for (int k = 0; k < k16; ++k) {
            umma(d0, a0 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (t | k) ? 1u : 0u);
            umma(d1, a1 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, idesc, (t | k) ? 1u : 0u);
            nmma += 2;
          }
        }
      } else if (MODE == 6) {     // conv_halo_kernel's issue loop as written (runtime k16 / tps, unroll 1)
        const int rk16 = p.k16, tps = p.tps;
        int kh = 0, kw = 0;
#pragma unroll 1
        for (int t = 0; t < tps; ++t) {
          const uint32_t a_off = (uint32_t)(((kh * 10 + kw) * 128) >> 4);
          const uint32_t bt = b0 + (uint32_t)t * btap;
          const uint32_t first = (uint32_t)t;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < rk16) umma(d0, a0 + a_off + 2u * k, AH, bt + 2u * k, BH, idesc, (first | (uint32_t)k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < rk16) umma(d1, a1 + a_off + 2u * k, AH, bt + 2u * k, BH, idesc, (first | (uint32_t)k) ? 1u : 0u);
          if (++kw == 3) { kw = 0; ++kh; }
        }
      } else if (MODE == 7) {     // same order, 64-bit descriptors advanced by 64-bit adds, unrolled
        const uint64_t A0 = ((uint64_t)AH << 32) | a0, A1 = ((uint64_t)AH << 32) | a1, B0 = ((uint64_t)BH << 32) | b0;
        _Pragma("unroll") for (int t = 0; t < 9; ++t) {
          const uint64_t off = (uint64_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
          _Pragma("unroll") for (int k = 0; k < k16; ++k) umma64(d0, A0 + off + 2 * k, B0 + (uint64_t)t * btap + 2 * k, idesc, (t | k) ? 1u : 0u);
          _Pragma("unroll") for (int k = 0; k < k16; ++k) umma64(d1, A1 + off + 2 * k, B0 + (uint64_t)t * btap + 2 * k, idesc, (t | k) ? 1u : 0u);
        }
      } else if (MODE == 5) {     // one M=128 x N=2*48 instruction pattern stand-in: same A, B twice as wide
        const uint32_t id2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((128u >> 4) << 24);
        _Pragma("unroll") for (int t = 0; t < 9; ++t) {
          const uint32_t off = (uint32_t)((((t / 3) * 10 + t % 3) * 128) >> 4);
          _Pragma("unroll") for (int k = 0; k < k16; ++k) { umma(d0, a0 + off + 2 * k, AH, b0 + t * btap + 2 * k, BH, id2, (t | k) ? 1u : 0u); ++nmma; }
        }
      }
      if (p.flags & 1) umma_commit(smem_u32(&bars[2]));
      umma_commit(smem_u32(&bars[8 + (it & 3)]));
      }
      __syncwarp();
      // the hand-off of the real kernel is asynchronous (double-buffered accumulators): item it + 1
      // is issued while item it drains; only item it - 1 is waited for
      if (it >= step) mbar_wait(smem_u32(&bars[8 + ((it - step) & 3)]), (uint32_t)(((it - step) >> 2) & 1));
    }
    { const int last = p.items - 1 - ((p.items - 1 - w) % step); mbar_wait(smem_u32(&bars[8 + (last & 3)]), (uint32_t)((last >> 2) & 1)); }
    long long t1 = clock64();
    const int per_item = (MODE == 6 || MODE == 7) ? 18 * K16 : (MODE == 0) ? (54 / K16) * K16 : (MODE == 5 ? 9 * K16 : 18 * K16);
    if (p.cycles && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && warp == 1) { p.cycles[0] = t1 - t0; p.cycles[1] = (long long)per_item * p.items; }
  }
  if (warp >= 3 && (p.flags & 2)) {
    // bystander warps polling an mbarrier that only completes when the issuing warp is done
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[3])), "r"(0u) : "memory");
    }
  }
  if (warp == 1 && (threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[3])) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE, int K16>
static void run(int grid, int N, int dstride, long long* dcyc, const char* name, int flags = 0, int threads = 128) {
  const int smem = 202 * 1024;
  CK(cudaFuncSetAttribute(seq_kernel<MODE, K16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Params p{MODE, N, 64, K16, dstride, 9, flags, dcyc};
  seq_kernel<MODE, K16><<<grid, threads, smem>>>(p);
  CK(cudaDeviceSynchronize());
  long long c[2];
  CK(cudaMemcpy(c, dcyc, 16, cudaMemcpyDeviceToHost));
  printf("grid=%3d N=%2d k16=%d flags %d threads %d mode %d (%s): %.1f cycles/MMA over %lld MMAs\n", grid, N, K16,
         flags, threads, MODE, name, (double)c[0] / (double)c[1], c[1]);
}

int main() {
  CK(cudaSetDevice(0));
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 16));
  for (int grid : {1, 148}) {
    for (int N : {48, 96}) {
      for (int ds : {N}) {
        run<0, 4>(grid, N, ds, dcyc, "same operands, one accumulator");
        run<0, 3>(grid, N, ds, dcyc, "same operands, one accumulator");
        run<1, 3>(grid, N, ds, dcyc, "9 shifted taps, one accumulator");
        run<2, 3>(grid, N, ds, dcyc, "kernel order: per tap tile0(k16) tile1(k16)");
        run<3, 3>(grid, N, ds, dcyc, "tile-major: 9 taps tile0, 9 taps tile1");
        run<4, 3>(grid, N, ds, dcyc, "alternate accumulators every MMA");
        run<2, 4>(grid, N, ds, dcyc, "kernel order, K = 64");
        run<6, 3>(grid, N, ds, dcyc, "PRODUCT LOOP as written (runtime k16, unroll 1)");
        run<6, 4>(grid, N, ds, dcyc, "PRODUCT LOOP as written, K = 64");
        run<7, 3>(grid, N, ds, dcyc, "64-bit descriptors, unrolled");
        run<2, 3>(grid, N, ds, dcyc, "kernel order + second commit per item", 1);
        run<2, 3>(grid, N, ds, dcyc, "kernel order, 320 threads idle", 0, 320);
        run<2, 3>(grid, N, ds, dcyc, "kernel order + 8 polling bystander warps", 2, 320);
        run<2, 3>(grid, N, ds, dcyc, "kernel order + second commit + bystanders", 3, 320);
        run<2, 3>(grid, N, ds, dcyc, "kernel order + 2 satisfied mbarrier waits per item", 4, 128);
        run<2, 3>(grid, N, ds, dcyc, "kernel order + waits + second commit + bystanders", 7, 320);
        run<2, 3>(grid, N, ds, dcyc, "TWO issuing warps (alternate items)", 8, 128);
        run<2, 3>(grid, N, ds, dcyc, "TWO issuing warps + waits + second commit + bystanders", 15, 320);
        run<2, 4>(grid, N, ds, dcyc, "TWO issuing warps + waits + second commit + bystanders, K = 64", 15, 320);
      }
    }
  }
  return 0;
}
