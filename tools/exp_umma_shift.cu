// Experiment (not product code): how does tcgen05.mma address a K-major SWIZZLE_128B operand
// whose start address is NOT aligned to the 1024-byte swizzle repeat, and whose SBO is not a
// multiple of 1024?  This decides whether a 3x3 convolution can load one halo tile and reach
// its 9 taps purely through descriptor offsets.  Also times back-to-back MMAs for several N.
//
// Method: B = identity (N = 64, K = 64), so D[m][n] = A[m][n] as the tensor core read it.
// Pass 1 fills the TMA-loaded tile with the ROW id, pass 2 with the K id; D then reveals the
// (row, k) each (m, n) was fetched from.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_umma_shift exp_umma_shift.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROWS = 208;   // tile rows loaded by TMA (>= 128 + max shift + sbo slack)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2u << 61;
  return d;
}

struct Params {
  CUtensorMap tmap_a, tmap_b;
  float* out;          // [128][64]
  int shift_rows, sbo_bytes, base_off, N, reps;
  long long* cycles;
};

__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // ROWS x 128 B
  uint8_t* sb = smem + 32768;               // 256 x 128 B
  uint64_t* bars = (uint64_t*)(smem + 32768 + 32768);
  uint32_t* tslot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), ROWS * 128 + 256 * 128);
    tma_load_2d(smem_u32(sa), &p.tmap_a, smem_u32(&bars[0]), 0, 0);
    tma_load_2d(smem_u32(sb), &p.tmap_b, smem_u32(&bars[0]), 0, 0);
    mbar_wait(smem_u32(&bars[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t ad = make_desc(smem_u32(sa) + p.shift_rows * 128, p.sbo_bytes, p.base_off);
    const uint64_t bd = make_desc(smem_u32(sb), 1024, 0);
    long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r)
      for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, (r | k) ? 1u : 0u);
    umma_commit(smem_u32(&bars[1]));
    mbar_wait(smem_u32(&bars[1]), 0);
    long long t1 = clock64();
    if (p.cycles && blockIdx.x == 0) *p.cycles = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (p.out) {
    for (int c = 0; c < 64; c += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  return (PFN_cuTensorMapEncodeTiled_v12000)f;
}
static void make_map(CUtensorMap* m, void* ptr, int rows, int box_rows) {
  cuuint64_t dim[2] = {64, (cuuint64_t)rows};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  CK(cudaSetDevice(0));
  const int smem = 32768 + 32768 + 1024 + 1024;
  CK(cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<__nv_bfloat16> hrow(ROWS * 64), hk(ROWS * 64), hb(256 * 64);
  for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 64; ++k) { hrow[r * 64 + k] = __float2bfloat16((float)r); hk[r * 64 + k] = __float2bfloat16((float)k); }
  for (int n = 0; n < 256; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16((n % 64) == k ? 1.0f : 0.0f);
  __nv_bfloat16 *drow, *dk, *db; float* dout; long long* dcyc;
  CK(cudaMalloc(&drow, hrow.size() * 2)); CK(cudaMalloc(&dk, hk.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2));
  CK(cudaMalloc(&dout, 128 * 64 * 4)); CK(cudaMalloc(&dcyc, 8));
  CK(cudaMemcpy(drow, hrow.data(), hrow.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));

  struct V { int shift, sbo, bo; } variants[] = {
      {0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 3}, {8, 1024, 0},
      {0, 1280, 0}, {1, 1280, 0}, {1, 1280, 1}, {11, 1280, 0}, {11, 1280, 3}, {22, 1280, 0}, {0, 1664, 0}, {5, 1664, 0}, {5, 1664, 5}};
  for (auto v : variants) {
    std::vector<float> got_r(128 * 64), got_k(128 * 64);
    for (int pass = 0; pass < 2; ++pass) {
      Params p;
      make_map(&p.tmap_a, pass == 0 ? drow : dk, ROWS, ROWS);
      make_map(&p.tmap_b, db, 256, 256);
      p.out = dout; p.shift_rows = v.shift; p.sbo_bytes = v.sbo; p.base_off = v.bo; p.N = 64; p.reps = 1; p.cycles = nullptr;
      CK(cudaMemset(dout, 0, 128 * 64 * 4));
      exp_kernel<<<1, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(pass == 0 ? got_r.data() : got_k.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    }
    // hypothesis: row(m) = shift + (m/8)*(sbo/128) + m%8, k = n   (swizzle purely address based)
    int bad = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      const int er = v.shift + (m / 8) * (v.sbo / 128) + (m % 8);
      if ((int)got_r[m * 64 + n] != er || (int)got_k[m * 64 + n] != n) { if (first_bad < 0) first_bad = m * 64 + n; ++bad; }
    }
    printf("variant shift=%2d sbo=%4d base_off=%d : %s (%d mismatches)\n", v.shift, v.sbo, v.bo, bad ? "DIFFERENT" : "address-based OK", bad);
    if (bad) {
      for (int m = 0; m < 18; ++m) {
        printf("   m=%3d rows:", m);
        for (int c = 0; c < 8; ++c) printf(" %3d", (int)got_r[m * 64 + c * 8]);
        printf("  | k0 of chunk:");
        for (int c = 0; c < 8; ++c) printf(" %2d", (int)got_k[m * 64 + c * 8]);
        printf("\n");
      }
    }
  }
  // ---- MMA issue-rate timing: aligned vs halo-style (shifted start, SBO = 1280) A operand,
  // one CTA per SM (grid 1) and two co-resident CTAs per SM (grid 296)
  struct R { int shift, sbo; } rv[] = {{0, 1024}, {1, 1024}, {0, 1280}, {11, 1280}, {21, 1280}};
  for (int grid : {1, 296}) {
    for (int N : {48, 96, 192, 256}) {
      for (auto r : rv) {
        Params p;
        make_map(&p.tmap_a, drow, ROWS, ROWS);
        make_map(&p.tmap_b, db, 256, 256);
        p.out = nullptr; p.shift_rows = r.shift; p.sbo_bytes = r.sbo; p.base_off = 0; p.N = N; p.reps = 512; p.cycles = dcyc;
        exp_kernel<<<grid, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("grid=%3d N=%3d shift=%2d sbo=%4d: %.1f cycles/MMA (ideal %d)\n", grid, N, r.shift, r.sbo,
               (double)cyc / (512 * 4), N / 2);
      }
    }
  }
  return 0;
}
