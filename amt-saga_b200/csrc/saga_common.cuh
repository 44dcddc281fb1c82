// Shared host/device helpers for the saga_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/saga_b200.h"

namespace saga {

// thread-local error text behind saga_last_error_string()
char* err_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
// run-time options (api.cu): SAGA_OPT("SAGA_X") is getenv("SAGA_X") read once at start-up, overridable through
// saga_set_option; the id lookup happens once per call site
int opt_id(const char* name);
const char* opt_value(int id);
#define SAGA_OPT(name) ([]() -> const char* { static const int id_ = saga::opt_id(name); return saga::opt_value(id_); }())

#define SAGA_CUDA_OK(expr)                                                          \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return saga::set_error(SAGA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define SAGA_LAUNCH_CHECK()                                                         \
  do {                                                                              \
    saga::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess)                                                          \
      return saga::set_error(SAGA_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
  } while (0)

// Index map of np.pad(mode='reflect'): period 2(L-1), edge sample not repeated,
// keeps bouncing when the pad exceeds the signal (oracle/spectral.py:reflect_index).
__host__ __device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t len) {
  if (len <= 1) return 0;
  const int64_t period = 2 * (len - 1);
  int64_t m = i % period;
  if (m < 0) m += period;
  return m < len ? m : period - m;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// max over non-negative floats through their bit pattern (monotone for x >= 0)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

}  // namespace saga
