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
#include <cstdlib>

#include <cooperative_groups.h>

#include "saga_common.cuh"

namespace cg = cooperative_groups;

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

// ---------------------------------------------------------------------------
// Multi-step chains (n_steps >= 2, e.g. BASELINE cfg4: 16 guesses per window): a CLUSTER of 4 CTAs owns one
// window (row t belongs to CTA t % 4, so any guess slab spreads evenly), 1 CTA per SM.  Only 37 windows
// (78 MB) are then in flight on the chip, so a window stays L2-resident across its steps and the fused dB
// epilogue reads it from L2 as well: DRAM sees each window once in, once out, plus the guesses -- the
// algorithmic traffic of SURVEY 8(d) -- where one CTA per window with ~440 windows in flight re-reads and
// re-writes every slab from HBM (cfg4, 4096 windows x 16 guesses: 22.1 ms; this kernel: 17.0 ms; an L2
// prefetch of the next step's guess rows made it slower, 18.2 ms).  The per-step window maximum (the reference's ref_mag) is a
// DSMEM exchange + cluster barrier; arithmetic and operation order per element are those of
// subtract_chain_kernel (bit-exact against the numpy float32 replay).
// ---------------------------------------------------------------------------
constexpr int SUBC_CTAS = 4;
constexpr int SUBC_THREADS = 512;      // 128 registers per thread: 16 outstanding 16-byte loads per lane
constexpr int SUBC_UNR = 8;            // float4 per lane kept in flight per operand
constexpr int SUBC_WARPS = SUBC_THREADS / 32;

__device__ __forceinline__ float block_max_c(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < SUBC_WARPS; ++i) r = fmaxf(r, red[i]);
  return r;
}

// cluster-wide max of one value per CTA: write it into slot [parity][rank] of every CTA, barrier, read back
__device__ __forceinline__ float cluster_max(cg::cluster_group& cluster, float v, float (*xch)[SUBC_CTAS], int parity,
                                             int rank) {
  if (threadIdx.x < SUBC_CTAS) {
    float* remote = cluster.map_shared_rank(&xch[parity][rank], threadIdx.x);
    *remote = v;
  }
  cluster.sync();
  float r = xch[parity][0];
#pragma unroll
  for (int i = 1; i < SUBC_CTAS; ++i) r = fmaxf(r, xch[parity][i]);
  return r;
}

__global__ void __cluster_dims__(SUBC_CTAS, 1, 1) __launch_bounds__(SUBC_THREADS, 1)
subtract_chain_cluster_kernel(const SubArgs a) {
  extern __shared__ float fmax_s[];          // maxima of the rows this CTA owns: row t -> fmax_s[t / SUBC_CTAS]
  __shared__ float red[SUBC_WARPS];
  __shared__ float xch[2][SUBC_CTAS];
  __shared__ float xch_g[2][SUBC_CTAS];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int w = blockIdx.x / SUBC_CTAS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t P = a.frame_pitch;
  const int T = a.n_frames, B = a.n_bins;
  const int n_own = (T - rank + SUBC_CTAS - 1) / SUBC_CTAS;           // rows rank, rank + 4, ...
  float* win = a.win_mag + (a.win_offsets ? a.win_offsets[w] : (int64_t)w * a.win_stride);
  const bool relu = (a.flags & SAGA_SUB_RELU) != 0;
  const bool normalize = (a.flags & SAGA_SUB_NORMALIZE) != 0;

  if (a.frame_max_in) {
    for (int i = threadIdx.x; i < n_own; i += SUBC_THREADS)
      fmax_s[i] = a.frame_max_in[(int64_t)w * a.frame_max_stride + rank + SUBC_CTAS * i];
  } else {
    for (int i = warp; i < n_own; i += SUBC_WARPS) {
      const float m = row_max<true>(win + (int64_t)(rank + SUBC_CTAS * i) * P, B, lane);
      if (lane == 0) fmax_s[i] = m;
    }
  }
  cluster.sync();      // every CTA of the cluster is resident (its shared memory is addressable) from here on

  int parity = 0;
  for (int j = 0; j < a.n_steps; ++j) {
    const int64_t idx = (int64_t)w * a.n_steps + j;
    const float* g = a.guess_mag + (a.guess_offsets ? a.guess_offsets[idx] : idx * a.guess_stride);
    const int Tg_full = a.guess_frames ? a.guess_frames[idx] : a.guess_frames_all;
    const int off = a.offset_frames[idx];
    // ---- ref of the current window and of the subtrahend (util_audio.py:239) ----
    float ref, gref;
    const bool need_ref = !(j == 0 && a.ref_init && a.ref_init[w] >= 0.f);
    if (need_ref || !a.guess_ref) {
      float m = 0.f, mg = 0.f;
      if (need_ref)
        for (int i = threadIdx.x; i < n_own; i += SUBC_THREADS) m = fmaxf(m, fmax_s[i]);
      if (!a.guess_ref)
        for (int tg = rank + SUBC_CTAS * warp; tg < Tg_full; tg += SUBC_CTAS * SUBC_WARPS)
          mg = fmaxf(mg, row_max<true>(g + (int64_t)tg * P, B, lane));
      m = block_max_c(m, red);
      mg = block_max_c(mg, red);
      if (threadIdx.x < SUBC_CTAS) *cluster.map_shared_rank(&xch_g[parity][rank], threadIdx.x) = mg;
      ref = cluster_max(cluster, m, xch, parity, rank);
      gref = xch_g[parity][0];
#pragma unroll
      for (int i = 1; i < SUBC_CTAS; ++i) gref = fmaxf(gref, xch_g[parity][i]);
      parity ^= 1;
    }
    if (!need_ref) ref = a.ref_init[w];
    if (a.guess_ref) gref = a.guess_ref[idx];
    const float scale = normalize ? __fdiv_rn(ref, gref) : 1.0f;
    const float ok = a.overkill ? a.overkill[idx] : 1.0f;
    int Tg = Tg_full;
    if (off < 0 || off >= T) Tg = 0;
    else if (off + Tg > T) Tg = T - off;
    // guess rows whose window row (off + tg) this CTA owns
    const int tg0 = (((rank - off) % SUBC_CTAS) + SUBC_CTAS) % SUBC_CTAS;
    for (int tg = tg0 + SUBC_CTAS * warp; tg < Tg; tg += SUBC_CTAS * SUBC_WARPS) {
      float4* w4 = reinterpret_cast<float4*>(win + (int64_t)(off + tg) * P);
      const float4* g4 = reinterpret_cast<const float4*>(g + (int64_t)tg * P);
      float* wr = reinterpret_cast<float*>(w4);
      const float* gr = reinterpret_cast<const float*>(g4);
      float m = 0.f;
      const int n4 = B >> 2;
      // the CTA is alone on its SM: all loads of a row chunk are issued before the first use
      for (int i0 = lane; i0 < n4; i0 += 32 * SUBC_UNR) {
        float4 x[SUBC_UNR], y[SUBC_UNR];
#pragma unroll
        for (int u = 0; u < SUBC_UNR; ++u) {
          const int i = i0 + 32 * u;
          if (i < n4) {
            x[u] = w4[i];
            y[u] = __ldcs(g4 + i);            // guesses are read once: do not let them evict the windows
          }
        }
#pragma unroll
        for (int u = 0; u < SUBC_UNR; ++u) {
          const int i = i0 + 32 * u;
          if (i < n4) {
            float4 v = x[u];
            v.x = __fsub_rn(v.x, __fmul_rn(__fmul_rn(y[u].x, scale), ok));
            v.y = __fsub_rn(v.y, __fmul_rn(__fmul_rn(y[u].y, scale), ok));
            v.z = __fsub_rn(v.z, __fmul_rn(__fmul_rn(y[u].z, scale), ok));
            v.w = __fsub_rn(v.w, __fmul_rn(__fmul_rn(y[u].w, scale), ok));
            if (relu) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            w4[i] = v;
          }
        }
      }
      for (int k = (n4 << 2) + lane; k < B; k += 32) {
        float x = __fsub_rn(wr[k], __fmul_rn(__fmul_rn(__ldg(gr + k), scale), ok));
        if (relu) x = fmaxf(x, 0.f);
        m = fmaxf(m, x);
        wr[k] = x;
      }
      m = warp_max(m);
      if (lane == 0) fmax_s[(off + tg) / SUBC_CTAS] = m;
    }
    __syncthreads();
  }

  // ---- final window max (== ref_mag after the chain) and the dB image of the owned rows ----
  float m = 0.f;
  for (int i = threadIdx.x; i < n_own; i += SUBC_THREADS) m = fmaxf(m, fmax_s[i]);
  m = block_max_c(m, red);
  const float vmax = cluster_max(cluster, m, xch, parity, rank);
  if (rank == 0 && threadIdx.x == 0) {
    if (a.ref_out) a.ref_out[w] = vmax;
    if (a.vmax_scratch) a.vmax_scratch[w] = vmax;
  }
  if (a.D_out) {
    float* D = a.D_out + (a.win_offsets ? a.win_offsets[w] : (int64_t)w * a.win_stride);
    const float amin2 = a.amin * a.amin;
    const float ref_db = db_abs(vmax, amin2);
    const float floor_db = (a.top_db >= 0.f) ? (0.0f - a.top_db) : -INFINITY;
    const int Pq = (int)(P >> 2);
    for (int i = warp; i < n_own; i += SUBC_WARPS) {
      const int64_t t = rank + SUBC_CTAS * i;
      const float4* w4 = reinterpret_cast<const float4*>(win + t * P);
      float4* d4 = reinterpret_cast<float4*>(D + t * P);
      for (int q0 = lane; q0 < Pq; q0 += 32 * SUBC_UNR) {
        float4 x[SUBC_UNR];
#pragma unroll
        for (int u = 0; u < SUBC_UNR; ++u)
          if (q0 + 32 * u < Pq) x[u] = w4[q0 + 32 * u];
#pragma unroll
        for (int u = 0; u < SUBC_UNR; ++u) {
          const int q = q0 + 32 * u;
          if (q < Pq) {
            float4 d;
            d.x = fmaxf(db_of(x[u].x, amin2, ref_db), floor_db);
            d.y = fmaxf(db_of(x[u].y, amin2, ref_db), floor_db);
            d.z = fmaxf(db_of(x[u].z, amin2, ref_db), floor_db);
            d.w = fmaxf(db_of(x[u].w, amin2, ref_db), floor_db);
            const int c = q << 2;                 // keep the padding columns at zero
            if (c + 3 >= B) {
              if (c >= B) d.x = 0.f;
              if (c + 1 >= B) d.y = 0.f;
              if (c + 2 >= B) d.z = 0.f;
              if (c + 3 >= B) d.w = 0.f;
            }
            __stcs(d4 + q, d);
          }
        }
      }
    }
  }
  cluster.sync();      // no CTA of the cluster may exit while a peer can still address its shared memory
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

// Single-step subtraction as a flat pass (the producer loop's common case: one guessed note per window, K1 having
// left the per-frame maxima and the guess maximum behind).  With one step there is no sequential dependence inside
// a window, so the slab is cut into `chunks` row ranges per window -- 2400 small CTAs instead of 600 large ones
// (1.35 waves) -- and every thread keeps SF_UNR window + SF_UNR guess 16-byte loads in flight.  Element arithmetic
// is that of subtract_chain_kernel (bit-identical); the window's new maximum is assembled with atomicMax on the
// bit patterns (ReLU output and magnitudes are >= 0), the untouched frames contributing their K1 maxima.
constexpr int SF_THREADS = 256;
constexpr int SF_UNR = 6;
__global__ void __launch_bounds__(SF_THREADS) subtract_single_flat_kernel(const SubArgs a, float* __restrict__ vmax_out,
                                                                          int chunks) {
  __shared__ float red[SUB_WARPS];
  const int w = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int64_t P = a.frame_pitch;
  const int T = a.n_frames, B = a.n_bins;
  float* win = a.win_mag + (a.win_offsets ? a.win_offsets[w] : (int64_t)w * a.win_stride);
  const float* g = a.guess_mag + (a.guess_offsets ? a.guess_offsets[w] : (int64_t)w * a.guess_stride);
  const float* fmax_in = a.frame_max_in + (int64_t)w * a.frame_max_stride;
  const int Tg_full = a.guess_frames ? a.guess_frames[w] : a.guess_frames_all;
  const int off = a.offset_frames[w];
  const bool normalize = (a.flags & SAGA_SUB_NORMALIZE) != 0;
  float ref = 0.f;
  if (a.ref_init && a.ref_init[w] >= 0.f) {
    ref = a.ref_init[w];
  } else if (normalize) {
    float m = 0.f;
    for (int t = threadIdx.x; t < T; t += SF_THREADS) m = fmaxf(m, fmax_in[t]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    ref = red[0];
#pragma unroll
    for (int i = 1; i < SF_THREADS / 32; ++i) ref = fmaxf(ref, red[i]);
    __syncthreads();
  }
  const float scale = normalize ? __fdiv_rn(ref, a.guess_ref[w]) : 1.0f;
  const float ok = a.overkill ? a.overkill[w] : 1.0f;
  int Tg = Tg_full;
  if (off < 0 || off >= T) Tg = 0;
  else if (off + Tg > T) Tg = T - off;

  float m = 0.f;
  // this CTA's share of the frames the subtraction does not touch
  {
    const int per = (T + chunks - 1) / chunks;
    for (int t = chunk * per + threadIdx.x; t < min(T, (chunk + 1) * per); t += SF_THREADS)
      if (t < off || t >= off + Tg) m = fmaxf(m, fmax_in[t]);
  }
  // this CTA's rows of the slab, as one flat run of 16-byte vectors (rows are whole pitches, hence contiguous)
  const int rows = (Tg + chunks - 1) / chunks;
  const int r0 = chunk * rows, r1 = min(Tg, r0 + rows);
  if (r0 < r1) {
    const int Pq = (int)(P >> 2);
    float4* w4 = reinterpret_cast<float4*>(win + (int64_t)(off + r0) * P);
    const float4* g4 = reinterpret_cast<const float4*>(g + (int64_t)r0 * P);
    const int n4 = (r1 - r0) * Pq;
    const int s1 = SF_THREADS % Pq, sk = (SF_THREADS * SF_UNR) % Pq;
    int c0 = threadIdx.x % Pq;
    for (int i0 = threadIdx.x; i0 < n4; i0 += SF_THREADS * SF_UNR) {
      float4 x[SF_UNR], y[SF_UNR];
#pragma unroll
      for (int u = 0; u < SF_UNR; ++u)
        if (i0 + u * SF_THREADS < n4) {
          x[u] = w4[i0 + u * SF_THREADS];
          y[u] = __ldcs(g4 + i0 + u * SF_THREADS);
        }
      int cu = c0;
#pragma unroll
      for (int u = 0; u < SF_UNR; ++u) {
        const int i = i0 + u * SF_THREADS;
        if (i < n4) {
          float4 v;
          v.x = fmaxf(__fsub_rn(x[u].x, __fmul_rn(__fmul_rn(y[u].x, scale), ok)), 0.f);
          v.y = fmaxf(__fsub_rn(x[u].y, __fmul_rn(__fmul_rn(y[u].y, scale), ok)), 0.f);
          v.z = fmaxf(__fsub_rn(x[u].z, __fmul_rn(__fmul_rn(y[u].z, scale), ok)), 0.f);
          v.w = fmaxf(__fsub_rn(x[u].w, __fmul_rn(__fmul_rn(y[u].w, scale), ok)), 0.f);
          const int c = cu << 2;
          if (c + 3 >= B) {                  // padding columns: neither changed nor counted
            if (c >= B) v.x = x[u].x;
            if (c + 1 >= B) v.y = x[u].y;
            if (c + 2 >= B) v.z = x[u].z;
            v.w = x[u].w;
            m = fmaxf(m, fmaxf(c < B ? v.x : 0.f, fmaxf(c + 1 < B ? v.y : 0.f, c + 2 < B ? v.z : 0.f)));
          } else {
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
          }
          w4[i] = v;
        }
        cu += s1;
        if (cu >= Pq) cu -= Pq;
      }
      c0 += sk;
      if (c0 >= Pq) c0 -= Pq;
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = red[0];
#pragma unroll
    for (int i = 1; i < SF_THREADS / 32; ++i) r = fmaxf(r, red[i]);
    atomic_max_nonneg(vmax_out + w, r);
  }
}

// The same pass with the bytes in flight coming from depth instead of occupancy: every thread keeps DBL_UNR 16-byte
// loads outstanding.  Written to run beside the tensor-core CQT contraction (128 threads x 80 registers is what fits
// next to its 768-thread CTA) -- that co-residency turned out to be worth nothing (the two kernels together take as
// long as one after the other, profiles/microbench/db_umma_overlap_b200.txt), but the deep-load shape itself is
// 20 % faster than the shallow kernel on an empty GPU, so it is the default.  Same arithmetic, same results.
template <int DBL_THREADS, int DBL_UNR>
__global__ void __launch_bounds__(DBL_THREADS) window_db_lean_kernel(const float* __restrict__ win_mag,
                                                                      const int64_t* __restrict__ win_offsets,
                                                                      int64_t win_stride, float* __restrict__ D_out,
                                                                      const float* __restrict__ vmax_w, int n_bins,
                                                                      int n_frames, int64_t P, float amin, float top_db,
                                                                      int chunks_per_window) {
  const int w = blockIdx.x / chunks_per_window, chunk = blockIdx.x % chunks_per_window;
  const int64_t base = win_offsets ? win_offsets[w] : (int64_t)w * win_stride;
  const float amin2 = amin * amin;
  const float ref_db = db_abs(vmax_w[w], amin2);
  const float floor_db = (top_db >= 0.f) ? (0.0f - top_db) : -INFINITY;
  const int rows = (n_frames + chunks_per_window - 1) / chunks_per_window;
  const int t0 = chunk * rows, t1 = min(n_frames, t0 + rows);
  if (t0 >= t1) return;
  const int Pq = (int)(P >> 2);
  const float4* w4 = reinterpret_cast<const float4*>(win_mag + base + (int64_t)t0 * P);
  float4* d4 = reinterpret_cast<float4*>(D_out + base + (int64_t)t0 * P);
  const int n4 = (t1 - t0) * Pq;
  const int s1 = DBL_THREADS % Pq, sk = (DBL_THREADS * DBL_UNR) % Pq;
  int c0 = threadIdx.x % Pq;               // float4 column of the batch's first element, advanced incrementally
  for (int i0 = threadIdx.x; i0 < n4; i0 += DBL_THREADS * DBL_UNR) {
    float4 x[DBL_UNR];
#pragma unroll
    for (int u = 0; u < DBL_UNR; ++u)
      if (i0 + u * DBL_THREADS < n4) x[u] = __ldcs(w4 + i0 + u * DBL_THREADS);
    int cu = c0;
#pragma unroll
    for (int u = 0; u < DBL_UNR; ++u) {
      const int i = i0 + u * DBL_THREADS;
      if (i < n4) {
        float4 d;
        d.x = fmaxf(db_of(x[u].x, amin2, ref_db), floor_db);
        d.y = fmaxf(db_of(x[u].y, amin2, ref_db), floor_db);
        d.z = fmaxf(db_of(x[u].z, amin2, ref_db), floor_db);
        d.w = fmaxf(db_of(x[u].w, amin2, ref_db), floor_db);
        const int c = cu << 2;               // keep the padding columns at zero
        if (c + 3 >= n_bins) {
          if (c >= n_bins) d.x = 0.f;
          if (c + 1 >= n_bins) d.y = 0.f;
          if (c + 2 >= n_bins) d.z = 0.f;
          d.w = 0.f;
        }
        __stcs(d4 + i, d);
      }
      cu += s1;
      if (cu >= Pq) cu -= Pq;
    }
    c0 += sk;
    if (c0 >= Pq) c0 -= Pq;
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
  // phase flags (as SAGA_CQT_SKIP_*): let a caller put the chain and the dB pass on different streams
  const bool skip_db = (flags & SAGA_SUB_SKIP_DB) != 0, only_db = (flags & SAGA_SUB_ONLY_DB) != 0;
  if (skip_db && only_db) return set_error(SAGA_ERR_INVALID, "subtract_db_exec: SKIP_DB and ONLY_DB together");
  if (only_db && (!D_out || !ref_out))
    return set_error(SAGA_ERR_INVALID, "subtract_db_exec: ONLY_DB needs D_out and the ref_out an earlier SKIP_DB call filled");
  if (skip_db) D_out = nullptr;
  flags &= 0xFF;
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
  int min_steps = 2;
  if (const char* e = SAGA_OPT("SAGA_SUB_CLUSTER_MIN_STEPS")) min_steps = atoi(e);       // tuning aid
  const bool clustered = vec && n_steps >= min_steps && n_steps >= 1 && !SAGA_OPT("SAGA_SUB_NO_CLUSTER") && !skip_db && !only_db;
  if (clustered) {
    // >= 120 KB of dynamic shared memory per CTA keeps it alone on its SM: 37 windows in flight, L2-resident
    const size_t csmem = std::max<size_t>(sizeof(float) * ((size_t)n_frames / SUBC_CTAS + 2), 120 * 1024);
    SAGA_CUDA_OK(cudaFuncSetAttribute(subtract_chain_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
    subtract_chain_cluster_kernel<<<(unsigned)n_windows * SUBC_CTAS, SUBC_THREADS, csmem, st>>>(a);
    SAGA_LAUNCH_CHECK();
    if (D_out && vmax != ref_out) SAGA_CUDA_OK(cudaFreeAsync(vmax, st));
    return SAGA_OK;           // the dB image was written by the same kernel
  }
  // one step, ReLU, K1's by-products at hand: the flat single-step kernel (values >= 0 make the atomic max valid)
  const bool flat = vec && !only_db && n_steps == 1 && (flags & SAGA_SUB_RELU) && frame_max_in && guess_ref && vmax &&
                    !SAGA_OPT("SAGA_SUB_NO_FLAT");
  if (only_db) {
    // the chain ran in an earlier call (SAGA_SUB_SKIP_DB) and left every window's final max in ref_out
  } else if (flat) {
    SAGA_CUDA_OK(cudaMemsetAsync(vmax, 0, sizeof(float) * n_windows, st));
    int fchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, guess_frames_all), (148 * 16 + n_windows - 1) / n_windows));
    if (const char* e = SAGA_OPT("SAGA_SUB_FLAT_CHUNKS")) fchunks = std::max(1, std::min(std::max(1, guess_frames_all), atoi(e)));   // tuning aid
    subtract_single_flat_kernel<<<(unsigned)((int64_t)n_windows * fchunks), SF_THREADS, 0, st>>>(a, vmax, fchunks);
    SAGA_LAUNCH_CHECK();
  } else if (vec) {
    if (smem > 40 * 1024)
      SAGA_CUDA_OK(cudaFuncSetAttribute(subtract_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    subtract_chain_kernel<true><<<n_windows, SUB_THREADS, smem, st>>>(a);
    SAGA_LAUNCH_CHECK();
  } else {
    if (smem > 40 * 1024)
      SAGA_CUDA_OK(cudaFuncSetAttribute(subtract_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    subtract_chain_kernel<false><<<n_windows, SUB_THREADS, smem, st>>>(a);
    SAGA_LAUNCH_CHECK();
  }
  if (D_out) {
    // ~8 CTAs of 256 threads per SM in flight, whole chunks of rows per CTA
    int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(n_frames, (148 * 16 + n_windows - 1) / n_windows));
    // default: 256 threads x 12 outstanding 16-byte loads, ~32 CTAs of whole row chunks per SM slot (0.54 -> 0.42 ms for
    // 600 windows = 6.05 TB/s, profiles/microbench/db_lean_sweep_b200.txt); SAGA_DB_LEAN=0 restores the shallow kernel
    const char* lean = SAGA_OPT("SAGA_DB_LEAN");
    const int lv = lean ? atoi(lean) : 4;
    if (vec && lv > 0) {
      // one deep batch per thread: a CTA takes as many whole rows as 256 threads x 12 vectors cover (11 rows of 1028
      // floats -> 47 chunks of a 516-frame window; the step measures 2.908 ms at 48-64 chunks, 2.950 at 8, 2.97 at 96)
      const int rows_per_cta = (int)std::max<int64_t>(1, (256 * 12) / std::max<int64_t>(1, frame_pitch >> 2));
      chunks = std::max(1, (n_frames + rows_per_cta - 1) / rows_per_cta);
    }
    if (const char* e = SAGA_OPT("SAGA_DB_CHUNKS")) chunks = std::max(1, std::min(n_frames, atoi(e)));
    const int64_t blocks = (int64_t)n_windows * chunks, lblocks = blocks;
#define SAGA_DBL(T, U) window_db_lean_kernel<T, U><<<(unsigned)lblocks, T, 0, st>>>(win_mag, win_offsets, win_stride, D_out, vmax, n_bins, n_frames, frame_pitch, amin, top_db, chunks)
    if (vec && lv == 1) SAGA_DBL(128, 12);
    else if (vec && lv == 2) SAGA_DBL(128, 16);
    else if (vec && lv == 3) SAGA_DBL(256, 8);
    else if (vec && lv >= 4) SAGA_DBL(256, 12);
#undef SAGA_DBL
    else if (vec)
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
