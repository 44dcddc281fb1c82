// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) as a function of N and of the
// number of independent TMEM accumulators the MMAs are spread over (1 = one dependent chain).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_latency umma_latency.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int N, int n_acc, int iters, int a_shift16, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  float* f = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) f[i] = 1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t da = smem_desc(smem_u32(sm) + 16 * a_shift16, 4096, 128);       // A: 128 rows x 2 chunks
    const uint64_t db = smem_desc(smem_u32(sm) + 32768, 4096, 128);                // B: N rows x 2 chunks
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) mma(tm + (uint32_t)((i % n_acc) * N), da, db, idesc, i >= n_acc);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int iters = 512;
  printf("N n_acc a_shift issue_cyc/mma total_cyc/mma\n");
  for (int N : {32, 64, 128, 256})
    for (int n_acc : {1, 2, 4, 8})
      for (int sh : {0, 3}) {
        if (n_acc * N > 512) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          bench<<<1, 128, 65536 + 1024>>>(N, n_acc, iters, sh, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%3d %5d %7d %13.1f %13.1f\n", N, n_acc, sh, (double)h[0] / iters, (double)h[1] / iters);
      }
  return 0;
}
