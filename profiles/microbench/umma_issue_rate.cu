// Micro-benchmark 2: true tcgen05.mma (kind::tf32, cta_group::1, SMEM x SMEM operands, K-major
// SWIZZLE_NONE) dispatch interval with ZERO per-MMA issue arithmetic: 32 MMAs fully unrolled with
// identical, loop-invariant descriptors; varies M (64/128), N, and 1 vs 4 TMEM accumulators.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
#define MMA(D, ACC)                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                      \
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(D), "l"(da), "l"(db), \
               "r"(idesc), "r"(ACC) : "memory")

template <int NACC>
__global__ void __launch_bounds__(32, 1) bench(int M, int N, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  float* f = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < 16384; i += 32) f[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = smem_desc(smem_u32(sm), 4096, 128);
    const uint64_t db = smem_desc(smem_u32(sm) + 32768, 4096, 128);
    const uint32_t d0 = tm, d1 = tm + (NACC > 1 ? 128 : 0), d2 = tm + (NACC > 1 ? 256 : 0), d3 = tm + (NACC > 1 ? 384 : 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) { MMA(d0, 1u); MMA(d1, 1u); MMA(d2, 1u); MMA(d3, 1u); }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int reps = 32;   // 32 * 32 = 1024 MMAs
  printf("M N n_acc issue_cyc/mma total_cyc/mma  (tensor floor = max(M,128)*N/256 for tf32 K=8)\n");
  for (int M : {128, 64})
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 4}) {
        if (nacc == 4 && N > 128) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          if (nacc == 1) bench<1><<<1, 32, 65536 + 1024>>>(M, N, reps, d);
          else bench<4><<<1, 32, 65536 + 1024>>>(M, N, reps, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%3d %3d %5d %13.1f %13.1f\n", M, N, nacc, (double)h[0] / (reps * 32), (double)h[1] / (reps * 32));
      }
  return 0;
}
