// K3: generative-subtractive chain + amplitude_to_db epilogue.
// Replaces /root/reference/util_audio.py:221-259 (subtract), :170-174 (ref_mag)
// and :176-180 (D = librosa.amplitude_to_db(mag, ref=ref_mag)).
//
// Semantics kept from the reference (numpy float32):
//   mag_sub *= ref/ref_sub          -> p = fl32(g * fl32(ref / gref))
//   mag_sub *= overkill_factor      -> p = fl32(p * ok)
//   mag -= pad(mag_sub); relu       -> w = max(fl32(w - p), 0)       (no FMA contraction)
//   the mag setter resets ref_mag   -> the next step's ref is the max of the updated window
//
// Layout is frame-major (a guess spanning columns [off, off+Tg) is ONE contiguous
// slab of Tg*pitch floats), so every pass is a flat, 16-byte vectorised stream.
// One CTA owns one window for the whole chain: per-frame maxima live in shared
// memory, so a step only touches the guess's column range, and the window max
// needed by the next step / by the dB floor is a block reduction, not a re-read.
#include <algorithm>

#include "saga_common.cuh"

namespace saga {

constexpr int SUB_THREADS = 512;
constexpr int SUB_WARPS = SUB_THREADS / 32;

struct SubArgs {
  float* win_mag;
  const int64_t* win_offsets;
  int64_t win_stride;
  const float* guess_mag;
  const int64_t* guess_offsets;
  int64_t guess_stride;
  const int32_t* guess_frames;
  int guess_frames_all;
  const int32_t* offset_frames;
  const float* overkill;
  const float* guess_ref;
  const float* ref_init;
  const float* frame_max_in;
  int64_t frame_max_stride;
  int flags;
  float* D_out;
  float* ref_out;
  float* vmax_scratch;
  int n_steps, n_bins, n_frames;
  int64_t frame_pitch;
  float amin, top_db;
};

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < SUB_WARPS; ++i) r = fmaxf(r, red[i]);
  return r;
}

// 10 log10(max(amin^2, v^2)) - ref_db through the hardware log2 (MUFU.LG2: |error| < 2^-21 on the log2,
// i.e. < 2e-6 dB -- the parity bar is 0.01 dB): 6 instructions per element instead of libm's ~30, which
// had made the dB pass instruction-bound (ncu r1 v6: 38 instructions per element, 66 % issue-active).
// The argument is >= amin^2 > 0 and finite, so the .ftz form never sees a denormal.
__device__ __forceinline__ float db_abs(float v, float amin2) {      // 10 log10(max(amin^2, v^2))
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaxf(amin2, v * v)));
  return __fmul_rn(l2, 3.01029995663981195f);          // 10 log10(2) * log2(x); _rn: no FMA contraction with the
                                                       // subtraction below, or x == ref would not give exactly 0
}
// the reference level goes through the same function, so an element equal to the reference (the window
// maximum, or everything in an all-zero window) maps to exactly 0 dB as it does in numpy
__device__ __forceinline__ float db_of(float v, float amin2, float ref_db) { return __fsub_rn(db_abs(v, amin2), ref_db); }

// max of row[0..n_bins) cooperatively by one warp
template <bool VEC>
__device__ __forceinline__ float row_max(const float* row, int n_bins, int lane) {
  float m = 0.f;
  if (VEC) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const int n4 = n_bins >> 2;
    for (int i = lane; i < n4; i += 32) {
      const float4 v = r4[i];
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int k = (n4 << 2) + lane; k < n_bins; k += 32) m = fmaxf(m, row[k]);
  } else {
    for (int k = lane; k < n_bins; k += 32) m = fmaxf(m, row[k]);
  }
  return warp_max(m);
}

template <bool VEC>
__global__ void __launch_bounds__(SUB_THREADS) subtract_chain_kernel(const SubArgs a) {
  extern __shared__ float fmax_s[];  // [n_frames] per-frame maxima
  __shared__ float red[SUB_WARPS];
  const int w = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t P = a.frame_pitch;
  const int T = a.n_frames, B = a.n_bins;
  float* win = a.win_mag + (a.win_offsets ? a.win_offsets[w] : (int64_t)w * a.win_stride);
  const bool relu = (a.flags & SAGA_SUB_RELU) != 0;
  const bool normalize = (a.flags & SAGA_SUB_NORMALIZE) != 0;

  // ---- per-frame maxima of the incoming window --------------------------------
  if (a.frame_max_in) {
    for (int t = threadIdx.x; t < T; t += SUB_THREADS) fmax_s[t] = a.frame_max_in[(int64_t)w * a.frame_max_stride + t];
  } else {
    for (int t = warp; t < T; t += SUB_WARPS) {
      const float m = row_max<VEC>(win + (int64_t)t * P, B, lane);
      if (lane == 0) fmax_s[t] = m;
    }
  }
  __syncthreads();

  for (int j = 0; j < a.n_steps; ++j) {
    const int64_t idx = (int64_t)w * a.n_steps + j;
    const float* g = a.guess_mag + (a.guess_offsets ? a.guess_offsets[idx] : idx * a.guess_stride);
    const int Tg_full = a.guess_frames ? a.guess_frames[idx] : a.guess_frames_all;
    const int off = a.offset_frames[idx];
    // ---- ref of the current window (reference: self.ref_mag, util_audio.py:239) ----
    float ref;
    if (j == 0 && a.ref_init && a.ref_init[w] >= 0.f) {
      ref = a.ref_init[w];
    } else {
      float m = 0.f;
      for (int t = threadIdx.x; t < T; t += SUB_THREADS) m = fmaxf(m, fmax_s[t]);
      ref = block_max(m, red);
    }
    // ---- ref of the subtrahend (its own ref_mag: max over ALL its frames) ----
    float gref;
    if (a.guess_ref) {
      gref = a.guess_ref[idx];
    } else {
      float m = 0.f;
      for (int t = warp; t < Tg_full; t += SUB_WARPS) m = fmaxf(m, row_max<VEC>(g + (int64_t)t * P, B, lane));
      gref = block_max(m, red);
    }
    const float scale = normalize ? __fdiv_rn(ref, gref) : 1.0f;
    const float ok = a.overkill ? a.overkill[idx] : 1.0f;
    int Tg = Tg_full;
    if (off < 0 || off >= T) Tg = 0;
    else if (off + Tg > T) Tg = T - off;

    for (int tg = warp; tg < Tg; tg += SUB_WARPS) {
      float* wr = win + (int64_t)(off + tg) * P;
      const float* gr = g + (int64_t)tg * P;
      float m = 0.f;
      if (VEC) {
        float4* w4 = reinterpret_cast<float4*>(wr);
        const float4* g4 = reinterpret_cast<const float4*>(gr);
        const int n4 = B >> 2;
#pragma unroll 3
        for (int i = lane; i < n4; i += 32) {
          float4 x = w4[i];
          const float4 y = __ldg(g4 + i);
          x.x = __fsub_rn(x.x, __fmul_rn(__fmul_rn(y.x, scale), ok));
          x.y = __fsub_rn(x.y, __fmul_rn(__fmul_rn(y.y, scale), ok));
          x.z = __fsub_rn(x.z, __fmul_rn(__fmul_rn(y.z, scale), ok));
          x.w = __fsub_rn(x.w, __fmul_rn(__fmul_rn(y.w, scale), ok));
          if (relu) {
            x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
          }
          m = fmaxf(fmaxf(m, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
          w4[i] = x;
        }
        for (int k = (n4 << 2) + lane; k < B; k += 32) {
          float x = __fsub_rn(wr[k], __fmul_rn(__fmul_rn(__ldg(gr + k), scale), ok));
          if (relu) x = fmaxf(x, 0.f);
          m = fmaxf(m, x);
          wr[k] = x;
        }
      } else {
        for (int k = lane; k < B; k += 32) {
          float x = __fsub_rn(wr[k], __fmul_rn(__fmul_rn(__ldg(gr + k), scale), ok));
          if (relu) x = fmaxf(x, 0.f);
          m = fmaxf(m, x);
          wr[k] = x;
        }
      }
      m = warp_max(m);
      if (lane == 0) fmax_s[off + tg] = m;
    }
    __syncthreads();
  }

  // ---- final window max (== ref_mag after the chain) and dB epilogue -------------
  float m = 0.f;
  for (int t = threadIdx.x; t < T; t += SUB_THREADS) m = fmaxf(m, fmax_s[t]);
  const float vmax = block_max(m, red);
  if (threadIdx.x == 0) {
    if (a.ref_out) a.ref_out[w] = vmax;
    if (a.vmax_scratch) a.vmax_scratch[w] = vmax;
  }
}

// dB epilogue as its own flat, perfectly balanced pass (a window per CTA leaves the last
// wave nearly empty: 600 windows over 296 resident CTAs = 2.03 waves).
//   D = 10 log10(max(amin^2, w^2)) - 10 log10(max(amin^2, ref^2)),  floor at max(D) - top_db,
// and max(D) == 0 because ref is the window's own max.
template <bool VEC>
__global__ void __launch_bounds__(256) window_db_kernel(const float* __restrict__ win_mag,
                                                         const int64_t* __restrict__ win_offsets,
                                                         int64_t win_stride, float* __restrict__ D_out,
                                                         const float* __restrict__ vmax_w, int n_bins, int n_frames,
                                                         int64_t P, float amin, float top_db, int chunks_per_window) {
  const int w = blockIdx.x / chunks_per_window, chunk = blockIdx.x % chunks_per_window;
  const int64_t base = win_offsets ? win_offsets[w] : (int64_t)w * win_stride;
  const float* win = win_mag + base;
  float* D = D_out + base;
  const float amin2 = amin * amin;
  const float vmax = vmax_w[w];
  const float ref_db = db_abs(vmax, amin2);
  const float floor_db = (top_db >= 0.f) ? (0.0f - top_db) : -INFINITY;
  const int rows = (n_frames + chunks_per_window - 1) / chunks_per_window;
  const int t0 = chunk * rows, t1 = min(n_frames, t0 + rows);
  if (t0 >= t1) return;
  if (VEC) {
    const int Pq = (int)(P >> 2);
    const float4* w4 = reinterpret_cast<const float4*>(win + (int64_t)t0 * P);
    float4* d4 = reinterpret_cast<float4*>(D + (int64_t)t0 * P);
    const int n4 = (t1 - t0) * Pq;
    int c4 = threadIdx.x % Pq;            // float4 column of element i, advanced incrementally (no modulo per element)
    const int cstep = 256 % Pq;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 x = __ldcs(w4 + i);
      float4 d;
      d.x = fmaxf(db_of(x.x, amin2, ref_db), floor_db);
      d.y = fmaxf(db_of(x.y, amin2, ref_db), floor_db);
      d.z = fmaxf(db_of(x.z, amin2, ref_db), floor_db);
      d.w = fmaxf(db_of(x.w, amin2, ref_db), floor_db);
      const int c = c4 << 2;                // keep the padding columns at zero
      c4 += cstep;
      if (c4 >= Pq) c4 -= Pq;
      if (c + 3 >= n_bins) {
        if (c >= n_bins) d.x = 0.f;
        if (c + 1 >= n_bins) d.y = 0.f;
        if (c + 2 >= n_bins) d.z = 0.f;
        d.w = 0.f;
      }
      __stcs(d4 + i, d);
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = t0 + warp; t < t1; t += 8)
      for (int k = lane; k < n_bins; k += 32)
        D[(int64_t)t * P + k] = fmaxf(db_of(win[(int64_t)t * P + k], amin2, ref_db), floor_db);
  }
}

// ---- stand-alone amplitude_to_db -------------------------------------------------
__global__ void clip_max_kernel(const float* mag, float* out, int n_bins, int n_frames,
                                int64_t P, int64_t clip_stride, int rows_per_cta) {
  const int clip = blockIdx.y;
  const int t0 = blockIdx.x * rows_per_cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = 0.f;
  for (int t = t0 + warp; t < min(n_frames, t0 + rows_per_cta); t += nw) {
    const float* row = mag + clip * clip_stride + (int64_t)t * P;
    for (int k = lane; k < n_bins; k += 32) m = fmaxf(m, row[k]);
  }
  m = warp_max(m);
  if (lane == 0) atomic_max_nonneg(out + clip, m);
}

__global__ void db_kernel(const float* mag, float* D, const float* ref, const float* vmax,
                          int n_bins, int n_frames, int64_t P, int64_t clip_stride, float amin,
                          float top_db, int rows_per_cta) {
  const int clip = blockIdx.y;
  const int t0 = blockIdx.x * rows_per_cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float amin2 = amin * amin;
  const float mx = vmax[clip];
  const float r = (ref && ref[clip] >= 0.f) ? ref[clip] : mx;
  const float ref_db = db_abs(r, amin2);
  const float floor_db = (top_db >= 0.f) ? (db_of(mx, amin2, ref_db) - top_db) : -INFINITY;
  for (int t = t0 + warp; t < min(n_frames, t0 + rows_per_cta); t += nw) {
    const float* row = mag + clip * clip_stride + (int64_t)t * P;
    float* drow = D + clip * clip_stride + (int64_t)t * P;
    for (int k = lane; k < n_bins; k += 32) drow[k] = fmaxf(db_of(row[k], amin2, ref_db), floor_db);
    for (int64_t k = n_bins + lane; k < P; k += 32) drow[k] = 0.f;
  }
}

}  // namespace saga

using namespace saga;

extern "C" int saga_subtract_db_exec(float* win_mag, const int64_t* win_offsets, int64_t win_stride,
                                     const float* guess_mag, const int64_t* guess_offsets,
                                     int64_t guess_stride, const int32_t* guess_frames,
                                     int guess_frames_all, const int32_t* offset_frames,
                                     const float* overkill, const float* guess_ref,
                                     const float* ref_init, const float* frame_max_in,
                                     int64_t frame_max_stride, int flags, float* D_out, float* ref_out,
                                     int n_windows, int n_steps, int n_bins, int n_frames,
                                     int64_t frame_pitch, float amin, float top_db, void* stream) {
  if (!win_mag) return set_error(SAGA_ERR_INVALID, "subtract_db_exec: null window pointer");
  if (n_steps > 0 && (!guess_mag || !offset_frames))
    return set_error(SAGA_ERR_INVALID, "subtract_db_exec: guesses/offsets missing for n_steps=%d", n_steps);
  if (n_windows < 0 || n_steps < 0 || n_bins < 1 || n_frames < 0 || frame_pitch < n_bins)
    return set_error(SAGA_ERR_INVALID, "subtract_db_exec: bad shape");
  if (n_windows == 0 || n_frames == 0) return SAGA_OK;
  SubArgs a;
  a.win_mag = win_mag; a.win_offsets = win_offsets; a.win_stride = win_stride;
  a.guess_mag = guess_mag; a.guess_offsets = guess_offsets; a.guess_stride = guess_stride;
  a.guess_frames = guess_frames; a.guess_frames_all = guess_frames_all;
  a.offset_frames = offset_frames; a.overkill = overkill; a.guess_ref = guess_ref;
  a.ref_init = ref_init; a.frame_max_in = frame_max_in; a.frame_max_stride = frame_max_stride; a.flags = flags; a.D_out = D_out; a.ref_out = ref_out;
  a.n_steps = n_steps; a.n_bins = n_bins; a.n_frames = n_frames; a.frame_pitch = frame_pitch;
  a.amin = amin; a.top_db = top_db;
  // 16-byte vector path needs aligned bases and pitch; explicit offsets are checked by the caller
  // contract (multiples of 4 elements) -- fall back to the scalar path when unsure.
  bool vec = (frame_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(win_mag) & 15) == 0) &&
             (!guess_mag || (reinterpret_cast<uintptr_t>(guess_mag) & 15) == 0) &&
             (!D_out || (reinterpret_cast<uintptr_t>(D_out) & 15) == 0) &&
             (((!win_offsets || (flags & SAGA_SUB_OFFSETS_ALIGNED)) && (win_offsets || win_stride % 4 == 0))) &&
             (((!guess_offsets || (flags & SAGA_SUB_OFFSETS_ALIGNED)) && (guess_offsets || guess_stride % 4 == 0)));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * (size_t)n_frames;
  if (smem > 160 * 1024) return set_error(SAGA_ERR_UNSUPPORTED, "subtract_db_exec: n_frames too large");
  // the dB pass needs every window's final max: ref_out doubles as that buffer when the caller
  // provides it, otherwise a stream-ordered scratch allocation is used
  float* vmax = ref_out;
  if (D_out && !vmax) SAGA_CUDA_OK(cudaMallocAsync(&vmax, sizeof(float) * n_windows, st));
  a.vmax_scratch = (vmax != ref_out) ? vmax : nullptr;
  if (vec) {
    if (smem > 40 * 1024)
      SAGA_CUDA_OK(cudaFuncSetAttribute(subtract_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    subtract_chain_kernel<true><<<n_windows, SUB_THREADS, smem, st>>>(a);
  } else {
    if (smem > 40 * 1024)
      SAGA_CUDA_OK(cudaFuncSetAttribute(subtract_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    subtract_chain_kernel<false><<<n_windows, SUB_THREADS, smem, st>>>(a);
  }
  SAGA_LAUNCH_CHECK();
  if (D_out) {
    // ~8 CTAs of 256 threads per SM in flight, whole chunks of rows per CTA
    int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(n_frames, (148 * 16 + n_windows - 1) / n_windows));
    const int64_t blocks = (int64_t)n_windows * chunks;
    if (vec)
      window_db_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(win_mag, win_offsets, win_stride, D_out, vmax, n_bins,
                                                              n_frames, frame_pitch, amin, top_db, chunks);
    else
      window_db_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(win_mag, win_offsets, win_stride, D_out, vmax, n_bins,
                                                               n_frames, frame_pitch, amin, top_db, chunks);
    SAGA_LAUNCH_CHECK();
    if (vmax != ref_out) SAGA_CUDA_OK(cudaFreeAsync(vmax, st));
  }
  return SAGA_OK;
}

extern "C" int saga_amplitude_to_db_exec(const float* mag, float* D_out, const float* ref, int n_clips,
                                         int n_bins, int n_frames, int64_t frame_pitch,
                                         int64_t clip_stride, float amin, float top_db, void* stream) {
  if (!mag || !D_out) return set_error(SAGA_ERR_INVALID, "amplitude_to_db_exec: null argument");
  if (n_clips <= 0 || n_frames <= 0) return SAGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* vmax = nullptr;
  SAGA_CUDA_OK(cudaMallocAsync(&vmax, sizeof(float) * n_clips, st));
  SAGA_CUDA_OK(cudaMemsetAsync(vmax, 0, sizeof(float) * n_clips, st));
  const int rows = 32;
  dim3 grid((n_frames + rows - 1) / rows, n_clips);
  clip_max_kernel<<<grid, 256, 0, st>>>(mag, vmax, n_bins, n_frames, frame_pitch, clip_stride, rows);
  SAGA_LAUNCH_CHECK();
  db_kernel<<<grid, 256, 0, st>>>(mag, D_out, ref, vmax, n_bins, n_frames, frame_pitch, clip_stride,
                                   amin, top_db, rows);
  SAGA_LAUNCH_CHECK();
  SAGA_CUDA_OK(cudaFreeAsync(vmax, st));
  return SAGA_OK;
}
