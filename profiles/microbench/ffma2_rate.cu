// Does the packed fp32 pipe (FFMA2 / FADD2, PTX fma.rn.f32x2 / add.rn.f32x2) on sm_100a deliver more
// FMAs per issue slot than scalar FFMA?  Each thread runs long chains of 8 independent accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float acc[16];
  unsigned long long a2[8];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) a2[i] = ((unsigned long long)__float_as_uint(acc[2 * i]) << 32) | __float_as_uint(acc[2 * i + 1]);
  const unsigned long long s2 = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma1(acc[i], s, s);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i] = fma2(a2[i], s2, s2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i] = add2(a2[i], s2);
    }
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)(a2[i] >> 32)) + __uint_as_float((unsigned)a2[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[3] = {"scalar FFMA  (16 per iter)", "FFMA2        (8 per iter) ", "FADD2        (8 per iter) "};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f);
      if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters, 1.0001f);
      if (mode == 2) k<2><<<148 * 8, 256>>>(d, iters, 1.0001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = 148.0 * 8 * 256 * (double)iters * 16;   // fp32 lane-operations (an x2 op counts 2)
    printf("%s  %.3f ms  %.1f Tlane-op/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", names[mode], ms, lane_ops / ms * 1e-9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
