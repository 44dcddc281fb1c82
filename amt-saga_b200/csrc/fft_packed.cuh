// Packed-fp32 FFT building blocks shared by the ring kernels (stft_ring.cu: K1, istft_ring.cu: K4).
// Everything is written with the sm_100 packed intrinsics (__ffma2_rn / __fadd2_rn / __fmul2_rn): ptxas folds
// component swaps (.LO_HI), per-half negations (.NP), scalar broadcasts (.F32) and uniform-register constants
// into the operand modifiers of FFMA2 / FADD2 / FMUL2, so a complex multiplication is two issue slots, a
// multiplication by -i is free, and there are no MOVs between butterflies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "fft_radix.cuh"

namespace saga {
namespace ring {

// exp(-2 pi i k / 32), k = 1..7, as the two operand pairs of a two-instruction complex multiplication:
// A = (c, -s), B = (s, c).  They travel as kernel parameters (constant bank -> uniform registers), where FMUL2 /
// FFMA2 take them with swap / negation modifiers; literal constants would be rebuilt in general registers
// (MOV / HFMA2 / FADD) in front of every use.
struct Rot32 {
  float2 A[8], B[8];
};

typedef float2 f2;
#define F2(a, b) make_float2((a), (b))

__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, F2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
// d * (w.x + i w.y): FMUL2 (broadcast d.x) + FFMA2 (broadcast d.y, swapped / half-negated w)
__device__ __forceinline__ f2 cmulp(f2 d, f2 w) { return fma2(F2(d.y, d.y), F2(-w.y, w.x), mul2(F2(d.x, d.x), w)); }

// (neg ? -d : d) * exp(-2 pi i idx / 32), idx in [0, 16): two packed instructions on uniform-register
// constants; the trivial angles become operand modifiers of the consuming packed add
__device__ __forceinline__ f2 mulw(f2 d, int idx, bool neg, const Rot32& w) {
  if (idx == 0) return neg ? F2(-d.x, -d.y) : d;
  if (idx == 8) return neg ? F2(-d.y, d.x) : F2(d.y, -d.x);
  const f2 dx = neg ? F2(-d.x, -d.x) : F2(d.x, d.x), dy = neg ? F2(-d.y, -d.y) : F2(d.y, d.y);
  if (idx < 8) return fma2(dy, w.B[idx], mul2(dx, w.A[idx]));                 // dx (c, -s) + dy (s, c)
  const f2 A = w.A[16 - idx], B = w.B[16 - idx];                             // c = -c', s = s'
  return fma2(dy, F2(-A.y, -A.x), mul2(dx, F2(-B.y, -B.x)));                  // dx (-c', -s') + dy (s', -c')
}

// radix-2 DIF stages half = HALF0 .. 1 on 32 register-resident points; X[bitrev(i)] == v[i] afterwards
template <int HALF0>
__device__ __forceinline__ void fft32_tail(f2 (&v)[32], const Rot32& w) {
#pragma unroll
  for (int half = HALF0; half >= 1; half >>= 1) {
#pragma unroll
    for (int b = 0; b < 32; b += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; ++k) {
        const f2 a = v[b + k], c = v[b + k + half];
        v[b + k] = add2(a, c);
        v[b + k + half] = mulw(sub2(a, c), k * (16 / half), false, w);
      }
    }
  }
}

__device__ __forceinline__ float fast_sqrt(float x) {   // sqrt.approx: ~1 ulp, exact 0 -> 0
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}


inline void fill_rot32(Rot32& w) {
  for (int k = 0; k < 8; ++k) {
    const double th = 2.0 * 3.14159265358979323846 * k / 32.0;
    w.A[k] = make_float2((float)std::cos(th), (float)-std::sin(th));
    w.B[k] = make_float2((float)std::sin(th), (float)std::cos(th));
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

}  // namespace ring
}  // namespace saga
