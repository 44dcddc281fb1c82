// Register-resident small FFTs (radix 2..32) used by the warp-per-frame
// Stockham passes of the STFT / iSTFT kernels.  Everything is fully unrolled so
// the data stays in registers and every twiddle is an immediate.
#pragma once
#include <cuda_runtime.h>

namespace saga {

// Complex add / subtract as ONE packed instruction (FADD2 on sm_100a, PTX add/sub.f32x2): the fp32 pipe
// delivers the same 128 lane-ops/clk/SM either way (profiles/microbench/ffma2_rate.cu), but the packed
// form takes one issue slot instead of two and the FFT kernels are issue-bound.  Same IEEE results.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
// componentwise product (a.x*b.x, a.y*b.y) in one FMUL2
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// cos(2*pi*i/32), i = 0..8 (float64 values rounded once to fp32)
__host__ __device__ constexpr float cos32_tab(int i) {
  switch (i) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960218f;
    case 6: return 0.38268343236508978f;
    case 7: return 0.19509032201612825f;
    default: return 0.0f;
  }
}
// cos / sin of 2*pi*i/32 for i = 0..16
__host__ __device__ constexpr float cos32(int i) { return i <= 8 ? cos32_tab(i) : -cos32_tab(16 - i); }
__host__ __device__ constexpr float sin32(int i) { return i <= 8 ? cos32_tab(8 - i) : cos32_tab(i - 8); }

// d * exp(-2*pi*i * idx/32), idx in [0,16), trivial angles special-cased
__device__ __forceinline__ float2 mul_w32(float2 d, int idx) {
  if (idx == 0) return d;
  if (idx == 8) return make_float2(d.y, -d.x);
  if (idx == 4) {
    const float h = 0.70710678118654757f;
    return make_float2((d.x + d.y) * h, (d.y - d.x) * h);
  }
  if (idx == 12) {
    const float h = 0.70710678118654757f;
    return make_float2((d.y - d.x) * h, -(d.x + d.y) * h);
  }
  const float c = cos32(idx), s = sin32(idx);
  return make_float2(d.x * c + d.y * s, d.y * c - d.x * s);
}

__host__ __device__ constexpr int bitrev(int i, int R) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) {
    r = (r << 1) | (i & 1);
    i >>= 1;
  }
  return r;
}

// In-place forward DFT of v[0..R) (decimation in frequency, radix 2).
// Result is in bit-reversed order:  X[bitrev(i, R)] == v[i].
template <int R>
__device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
  static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
#pragma unroll
  for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int b = 0; b < R; b += 2 * half) {
#pragma unroll
      for (int k = 0; k < half; ++k) {
        const float2 a = v[b + k], c = v[b + k + half];
        v[b + k] = cadd(a, c);
        v[b + k + half] = mul_w32(csub(a, c), k * (16 / half));
      }
    }
  }
}

}  // namespace saga
