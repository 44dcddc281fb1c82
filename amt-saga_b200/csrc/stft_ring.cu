// K1, ring form: batched windowed real FFT -> magnitude (/ unit phasor / complex) for n_fft 2048, hop 512
// (librosa.stft + magphase as reached from /root/reference/util_audio.py:127-128, :147, :173).
//
// Why a second kernel.  The first-generation kernel (stft.cu) gives a CTA ten frames, stages their span, meets at a
// barrier and lets every warp transform one frame: all warps of an SM are in the same phase at the same time, so
// the three resources a frame needs about equally -- issue slots, the fp32 pipe, the shared-memory / L1 data path --
// are used one after the other (ncu r1: issue 48 %, fp32 36 %, LSU 39 % busy, sum > 100 %).  Here
//   * one persistent CTA per SM: 16 consumer warps + 1 producer warp, NO CTA barrier after start-up;
//   * the producer streams the audio through a shared-memory RING of hop-sized blocks with 1-D bulk copies
//     (cp.async.bulk -> mbarrier complete_tx; clip edges are mirrored by the producer's own lanes), each sample
//     crosses L2 -> SM once although four frames read it;
//   * every consumer warp runs its own frame loop (wait full -> load 32 points per lane -> release -> FFT ->
//     store), so warps drift apart and the pipes overlap;
//   * the butterflies are written with the packed fp32 intrinsics (__ffma2_rn / __fadd2_rn / __fmul2_rn): ptxas
//     folds component swaps, per-half negations, scalar broadcasts and uniform-register constants into the
//     operand modifiers of FFMA2 / FADD2 / FMUL2, so a complex multiplication is TWO issue slots, a
//     multiplication by -i is free, and there are no MOVs between butterflies;
//   * the periodic Hann window satisfies w[n + N/2] = 1 - w[n]: the window product and the first radix-2 stage
//     fuse into two adds and two FMAs per point pair with HALF the window table;
//   * window / inter-pass twiddle / split twiddle tables live in shared memory in per-lane rows read with
//     conflict-free 16-byte loads (32 LDS.128 per frame and lane instead of 79 eight-byte global loads that
//     missed the 6 KB of L1 the old kernel left over).
// Work decomposition: 1024 complex points z[n] = (x[2n], x[2n+1]) = 32 lanes x 32 registers, two radix-32
// passes with one transpose through a per-warp exchange buffer, last pass register-resident, bins k / M-k paired
// by warp shuffles for the real-FFT split (same mathematics as stft.cu; parity tests run both).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fft_packed.cuh"
#include "saga_common.cuh"
#include "stft_plan.cuh"

namespace saga {

namespace ring {

// NW consumer warps (one frame in flight each) + 1 producer warp.  Registers are allocated to warps in groups of
// four, so 19 + 1 = 20 warps at <= 102 registers and 15 + 1 = 16 warps at <= 128 are the two shapes that waste none.
// S ring slots of one hop = 512 samples each (NW + 3 blocks are in use, the rest is the producer's lookahead).
constexpr int HOP = 512, NFFT = 2048, M = 1024;
constexpr int EXW = 32 * 34;              // float2 per warp exchange buffer (row pitch 34: 16-byte rows, conflict-free)
// table image, float2 units, per-lane rows whose pitch is = 4 (mod 32) words: conflict-free LDS.128
constexpr int T_NWIN = 0;                 // [32][18]: -(w[2n], w[2n+1]), n = lane + 32 r, r < 16
constexpr int T_TW0 = T_NWIN + 32 * 18;   // [32][34]: exp(-2 pi i lane rp / 1024), rp < 32
constexpr int T_TWN = T_TW0 + 32 * 34;    // [32][18]: 1/2 (-i) exp(-2 pi i k / 2048), k = lane + 32 i, i <= 16
constexpr int T_TOTAL = T_TWN + 32 * 18;  // 2240 float2 = 17920 bytes
constexpr size_t SMEM_TABLES = (size_t)T_TOTAL * 8;
template <int NW, int S>
struct Shape {
  static constexpr int THREADS = (NW + 1) * 32;
  static constexpr size_t SMEM_RING = (size_t)S * HOP * 4;
  static constexpr size_t SMEM_EXCH = (size_t)NW * EXW * 8;
  static constexpr size_t SMEM_BYTES = SMEM_TABLES + SMEM_RING + SMEM_EXCH + 2 * S * 8 + 16;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};
constexpr uint32_t SPIN_LIMIT = 1u << 26;

struct Args {
  StftArgs s;
  Rot32 w;
  const float2* tables;
  int R;                 // frames per run (item); a run needs R + 3 hop blocks
  int runs_per_clip;
  int n_items;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded: a pipeline bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}

// `issued`: number of ring blocks the producer has handed to the copy engine / filled.  A consumer may test a
// slot's mbarrier by PARITY only while it is at most one phase away from it; a warp that runs a ring lap ahead of
// the producer would otherwise read "phase u complete" off the parity of phase u-2.  Block g has been issued =>
// the slot's previous block was filled, read and released => the barrier is in (or past) phase u: unambiguous.
__device__ __forceinline__ void publish_issued(uint32_t* cnt, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(cnt)), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_issued(const uint32_t* cnt, uint32_t g) {
  uint32_t v, spins = 0;
  while (true) {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(cnt)) : "memory");
    if (v > g) return;
    __nanosleep(100);               // rare (start-up, run boundaries): do not steal issue slots from working warps
    if (++spins > (SPIN_LIMIT >> 4)) __trap();
  }
}

struct RunInfo {
  const float* x;      // clip base
  int64_t len;         // clip samples
  int clip, t0, nF;    // first frame of the run, frames of the run that exist (0 .. R)
};

__device__ __forceinline__ RunInfo run_info(const Args& a, int item) {
  RunInfo r;
  r.clip = item / a.runs_per_clip;
  const int run = item - r.clip * a.runs_per_clip;
  r.t0 = run * a.R;
  r.len = a.s.clip_lens[r.clip];
  r.x = a.s.wav + a.s.clip_offsets[r.clip];
  int64_t T = 0;
  if (r.len > 0) T = a.s.center ? 1 + r.len / HOP : (r.len >= NFFT ? 1 + (r.len - NFFT) / HOP : 0);
  const int64_t left = T - r.t0;
  r.nF = left <= 0 ? 0 : (left < a.R ? (int)left : a.R);
  return r;
}

// ------------------------------------------------------------------ producer warp
template <int S>
__device__ __forceinline__ void producer(const Args& a, float* ring_base, uint64_t* full, uint64_t* empty,
                                         uint32_t* issued, int lane) {
  const int G = gridDim.x;
  const int64_t trim = a.s.center ? NFFT / 2 : 0;
  int slot = 0;
  uint32_t use = 0;       // how often this slot position has wrapped
  uint32_t g = 0;         // CTA-wide block counter
  for (int item = blockIdx.x; item < a.n_items; item += G) {
    const RunInfo r = run_info(a, item);
    for (int b = 0; b < a.R + 3; ++b) {
      if (use > 0) mbar_wait(empty + slot, (use - 1) & 1);
      float* dst = ring_base + slot * HOP;
      const bool needed = b < r.nF + 3 && r.nF > 0;
      if (!needed) {
        if (lane == 0) mbar_arrive(full + slot, 1);
      } else {
        const int64_t s0 = (int64_t)(r.t0 + b) * HOP - trim;
        const float* src = r.x + s0;
        if (s0 >= 0 && s0 + HOP <= r.len && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          if (lane == 0) {
            mbar_arrive_expect_tx(full + slot, HOP * 4);
            bulk_g2s(smem_u32(dst), src, HOP * 4, full + slot);
          }
        } else {
          // clip edge (np.pad reflect: edge sample not repeated, keeps bouncing on tiny clips) or unaligned clip.
          // All 16 loads of a lane are issued before the first store: one memory round trip per edge block --
          // the ring stands still while the producer is in here.
          if (r.len > NFFT / 2) {            // a single reflection always lands inside the clip
            float vals[HOP / 32];
#pragma unroll
            for (int u = 0; u < HOP / 32; ++u) {
              int64_t sx = s0 + lane + 32 * u;
              if (sx < 0) sx = -sx;
              else if (sx >= r.len) sx = 2 * (r.len - 1) - sx;
              vals[u] = __ldg(r.x + sx);
            }
#pragma unroll
            for (int u = 0; u < HOP / 32; ++u) dst[lane + 32 * u] = vals[u];
          } else {
#pragma unroll 1
            for (int u = 0; u < HOP / 32; ++u) dst[lane + 32 * u] = __ldg(r.x + reflect_index(s0 + lane + 32 * u, r.len));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(full + slot, 1);
        }
      }
      ++g;
      if (lane == 0) publish_issued(issued, g);
      if (++slot == S) { slot = 0; ++use; }
    }
  }
}

// ------------------------------------------------------------------ one frame: ring -> |X| row
template <bool EXTRA>
__device__ __forceinline__ void transform_frame(const Args& a, const float* const (&blk)[4], const float4* nwin4,
                                                const float4* tw04, const float4* twN4, const float2* twN_row,
                                                float2* ex_st, const float4* ex_ld, int lane, int clip, int t,
                                                uint64_t* const (&rel)[4], const uint32_t (&rel_cnt)[4]) {
  f2 v[32];
  // point n = lane + 32 r is the sample pair (2n, 2n+1); 2n = 2 lane + 64 r lies in hop block r / 8
#pragma unroll
  for (int r = 0; r < 32; ++r)
    v[r] = *reinterpret_cast<const f2*>(blk[r >> 3] + 2 * lane + 64 * (r & 7));
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) mbar_arrive(rel[j], rel_cnt[j]);   // the four hop blocks may be refilled
  }

  // ---- pass 1: window product fused with the first radix-2 stage (w[n + 512] = 1 - w[n]; table holds -w) ----
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 w4 = nwin4[q];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = 2 * q + h;
      const f2 nw = h ? F2(w4.z, w4.w) : F2(w4.x, w4.y);
      const f2 xa = v[r], xc = v[r + 16];
      const f2 nd = sub2(xc, xa), sm = add2(xa, xc);
      v[r] = fma2(nw, nd, xc);                 //   w xa + (1 - w) xc
      const f2 tneg = fma2(nw, sm, xc);        // -(w xa - (1 - w) xc)
      v[r + 16] = mulw(tneg, r, true, a.w);
    }
  }
  fft32_tail<8>(v, a.w);
  // inter-pass twiddles exp(-2 pi i lane rp / 1024) and the transpose through the exchange buffer
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 t4 = tw04[j];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rp = 2 * j + h;
      f2 o = v[bitrev(rp, 32)];
      if (rp > 0) o = cmulp(o, h ? F2(t4.z, t4.w) : F2(t4.x, t4.y));
      ex_st[rp * 34] = o;
    }
  }
  __syncwarp();
  // ---- pass 2: lane = k1, registers = l; afterwards v[bitrev(k2)] = Z[lane + 32 k2] ----
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 q4 = ex_ld[j];
    v[2 * j] = F2(q4.x, q4.y);
    v[2 * j + 1] = F2(q4.z, q4.w);
  }
  __syncwarp();                                // exchange buffer free for the next frame
  fft32_tail<16>(v, a.w);

  // ---- real-FFT split: bin k = lane + 32 i pairs with M - k, held by lane (32 - lane) & 31 in register 31 - i
  //   X[k] = p/2 + u,  conj X[M-k] = p/2 - u,   p = Z[k] + conj Z[M-k],  u = 1/2 (-i) W^k (Z[k] - conj Z[M-k])
  const int64_t row = (int64_t)clip * a.s.out_clip_stride + (int64_t)t * a.s.frame_pitch;
  float* mag_up = a.s.mag_out + row + lane;
  float* mag_dn = a.s.mag_out + row + (M - lane);
  const int partner = (32 - lane) & 31;
  const f2 half2 = F2(0.5f, 0.5f);
  float vmax = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 w4 = twN4[j];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * j + h;
      const f2 zk = v[bitrev(i, 32)];
      const f2 give = v[bitrev(31 - i, 32)];
      f2 zm;
      zm.x = __shfl_sync(0xffffffffu, give.x, partner);
      zm.y = __shfl_sync(0xffffffffu, give.y, partner);
      if (lane == 0) zm = (i == 0) ? v[0] : v[bitrev(32 - i, 32)];   // lane 0 pairs with itself: M - 32 i = 32 (32 - i)
      const f2 p = add2(zk, F2(zm.x, -zm.y));
      const f2 q = add2(zk, F2(-zm.x, zm.y));
      const f2 u = cmulp(q, h ? F2(w4.z, w4.w) : F2(w4.x, w4.y));
      const f2 Xk = fma2(p, half2, u);
      const f2 Xm = fma2(p, half2, F2(-u.x, -u.y));     // conj X[M - k]
      const float mk = fast_sqrt(fmaf(Xk.x, Xk.x, Xk.y * Xk.y));
      const float mm = fast_sqrt(fmaf(Xm.x, Xm.x, Xm.y * Xm.y));
      vmax = fmaxf(vmax, fmaxf(mk, mm));
      mag_up[32 * i] = mk;
      mag_dn[-32 * i] = mm;
      if (EXTRA) {
        const int k = lane + 32 * i;
        if (a.s.cplx_out) {
          float2* c = a.s.cplx_out + row;
          c[k] = Xk;
          c[M - k] = F2(Xm.x, -Xm.y);
        }
        if (a.s.phase_out) {
          float2* ph = a.s.phase_out + row;
          ph[k] = mk > 0.f ? F2(Xk.x / mk, Xk.y / mk) : F2(1.f, 0.f);
          ph[M - k] = mm > 0.f ? F2(Xm.x / mm, -Xm.y / mm) : F2(1.f, 0.f);
        }
      }
    }
  }
  if (lane == 0) {
    // k = M/2 = 512 pairs with itself: X[512] = conj Z[512]
    const f2 z = v[bitrev(16, 32)];
    const f2 p = add2(z, F2(z.x, -z.y));
    const f2 q = add2(z, F2(-z.x, z.y));
    const f2 Xk = fma2(p, half2, cmulp(q, twN_row[16]));
    const float mk = fast_sqrt(fmaf(Xk.x, Xk.x, Xk.y * Xk.y));
    vmax = fmaxf(vmax, mk);
    a.s.mag_out[row + M / 2] = mk;
    if (EXTRA) {
      if (a.s.cplx_out) a.s.cplx_out[row + M / 2] = Xk;
      if (a.s.phase_out) a.s.phase_out[row + M / 2] = mk > 0.f ? F2(Xk.x / mk, Xk.y / mk) : F2(1.f, 0.f);
    }
  }
  // padding columns [M+1, frame_pitch) are defined as zero
  for (int64_t k = M + 1 + lane; k < a.s.frame_pitch; k += 32) {
    a.s.mag_out[row + k] = 0.f;
    if (EXTRA) {
      if (a.s.cplx_out) a.s.cplx_out[row + k] = F2(0.f, 0.f);
      if (a.s.phase_out) a.s.phase_out[row + k] = F2(0.f, 0.f);
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0) {
    if (a.s.frame_max_out) a.s.frame_max_out[(int64_t)clip * a.s.max_frames + t] = vmax;
    if (a.s.clip_max_out) atomic_max_nonneg(a.s.clip_max_out + clip, vmax);
  }
}

template <int NW, int S, bool EXTRA>
__global__ void __launch_bounds__((NW + 1) * 32, 1) stft_ring_kernel(const Args a) {
  constexpr int THREADS = Shape<NW, S>::THREADS;
  constexpr size_t SMEM_RING = Shape<NW, S>::SMEM_RING, SMEM_EXCH = Shape<NW, S>::SMEM_EXCH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tab = reinterpret_cast<float2*>(smem_raw);
  float* ring_base = reinterpret_cast<float*>(smem_raw + SMEM_TABLES);
  float2* exch = reinterpret_cast<float2*>(smem_raw + SMEM_TABLES + SMEM_RING);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + SMEM_TABLES + SMEM_RING + SMEM_EXCH);
  uint64_t* empty = full + S;
  uint32_t* issued = reinterpret_cast<uint32_t*>(empty + S);
  uint32_t* next_job = issued + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < T_TOTAL / 2; i += THREADS)
    reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(a.tables) + i);
  if (threadIdx.x == 0) {
    *issued = 0;
    *next_job = 0;
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);     // the producer's arrive (+ the bulk copy's bytes)
      mbar_init(empty + s, 4);    // four frames read every hop block (run ends make up the difference)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();                // the only CTA-wide barrier

  if (warp == NW) {
    producer<S>(a, ring_base, full, empty, issued, lane);
    return;
  }

  const float4* nwin4 = reinterpret_cast<const float4*>(tab + T_NWIN + lane * 18);
  const float4* tw04 = reinterpret_cast<const float4*>(tab + T_TW0 + lane * 34);
  const float2* twN_row = tab + T_TWN + lane * 18;
  const float4* twN4 = reinterpret_cast<const float4*>(twN_row);
  float2* ex = exch + warp * EXW;
  float2* ex_st = ex + lane;
  const float4* ex_ld = reinterpret_cast<const float4*>(ex + lane * 34);

  const int G = gridDim.x, R = a.R;
  // Job J of this CTA = frame f of its k-th item.  Warps draw jobs from a shared counter: jobs START in order (the
  // frames in flight are NW consecutive ones, so the ring only has to cover NW + 3 blocks plus the producer's
  // lookahead) and a scheduler that hosts three consumer warps instead of four simply takes more jobs.
  const uint32_t n_jobs = (uint32_t)((a.n_items - (int)blockIdx.x + G - 1) / G) * (uint32_t)R;
  int cur_k = -1;
  RunInfo r;
  while (true) {
    uint32_t J = 0;
    if (lane == 0) J = atomicAdd(next_job, 1u);
    J = __shfl_sync(0xffffffffu, J, 0);
    if (J >= n_jobs) break;
    const int k = (int)(J / (uint32_t)R), f = (int)J - k * R;
    if (k != cur_k) {
      cur_k = k;
      r = run_info(a, (int)blockIdx.x + k * G);
    }
    const int g0 = k * (R + 3) + f;               // first of the frame's four hop blocks (CTA-wide block counter)
    const float* blk[4];
    uint64_t* rel[4];
    uint32_t rel_cnt[4];
    wait_issued(issued, (uint32_t)(g0 + 3));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = g0 + j;
      const int use = g / S, slot = g - use * S;
      mbar_wait(full + slot, use & 1);
      blk[j] = ring_base + slot * HOP;
      rel[j] = empty + slot;
      // block b = f + j of the run is read by frames max(b-3,0) .. min(b,R-1); its first reader arrives for the missing ones
      const int b = f + j;
      uint32_t c = 1;
      if (f == 0 && j < 3) c += 3 - j;
      if (b >= R && j == 3) c += b - R + 1;
      rel_cnt[j] = c;
    }
    if (f < r.nF) {
      transform_frame<EXTRA>(a, blk, nwin4, tw04, twN4, twN_row, ex_st, ex_ld, lane, r.clip, r.t0 + f, rel, rel_cnt);
    } else {
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mbar_arrive(rel[j], rel_cnt[j]);
      }
    }
  }
}

}  // namespace ring

bool stft_ring_supported(const saga_stft_plan* p) {
  return p->n_fft == ring::NFFT && p->hop == ring::HOP && p->default_window && p->d_ring_tables != nullptr;
}

int stft_ring_build_tables(saga_stft_plan* p) {
  using namespace ring;
  p->d_ring_tables = nullptr;
  if (!(p->n_fft == NFFT && p->hop == HOP && p->default_window)) return SAGA_OK;
  const double PI = 3.14159265358979323846;
  std::vector<float2> t(T_TOTAL, make_float2(0.f, 0.f));
  for (int lane = 0; lane < 32; ++lane) {
    for (int r = 0; r < 16; ++r) {
      const int n = lane + 32 * r;
      const double w0 = 0.5 - 0.5 * std::cos(2.0 * PI * (2 * n) / NFFT), w1 = 0.5 - 0.5 * std::cos(2.0 * PI * (2 * n + 1) / NFFT);
      t[T_NWIN + lane * 18 + r] = make_float2((float)-w0, (float)-w1);
    }
    for (int rp = 0; rp < 32; ++rp) {
      const double ang = -2.0 * PI * (double)((lane * rp) % M) / (double)M;
      t[T_TW0 + lane * 34 + rp] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int i = 0; i <= 16; ++i) {
      const int k = lane + 32 * i;
      const double th = 2.0 * PI * k / (double)NFFT;
      t[T_TWN + lane * 18 + i] = make_float2((float)(-0.5 * std::sin(th)), (float)(-0.5 * std::cos(th)));
    }
  }
  SAGA_CUDA_OK(cudaMalloc(&p->d_ring_tables, sizeof(float2) * T_TOTAL));
  SAGA_CUDA_OK(cudaMemcpy(p->d_ring_tables, t.data(), sizeof(float2) * T_TOTAL, cudaMemcpyHostToDevice));
  return SAGA_OK;
}

template <int NW, int S>
static int launch_ring_shape(const saga_stft_plan* p, const StftArgs& s, int n_clips, int64_t max_frames, cudaStream_t st) {
  using namespace ring;
  constexpr int THREADS = Shape<NW, S>::THREADS;
  constexpr size_t SMEM_BYTES = Shape<NW, S>::SMEM_BYTES;
  const bool extra = s.phase_out || s.cplx_out;
  auto kern = extra ? stft_ring_kernel<NW, S, true> : stft_ring_kernel<NW, S, false>;
  int dev = 0, n_sm = 0;
  SAGA_CUDA_OK(cudaGetDevice(&dev));
  SAGA_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SAGA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  // runs: enough items to balance the persistent CTAs (>= ~24 per CTA when the batch allows), 8 .. 48 frames each
  const int64_t T = max_frames;
  const int64_t min_runs = (T + 47) / 48, max_runs = (T + 7) / 8;
  int64_t want = ((int64_t)n_sm * 24 + n_clips - 1) / n_clips;
  int64_t runs = std::min(std::max(want, min_runs), std::max(max_runs, min_runs));
  Args a;
  a.s = s;
  fill_rot32(a.w);
  a.tables = p->d_ring_tables;
  a.R = (int)((T + runs - 1) / runs);
  if (a.R < 4) a.R = 4;
  a.runs_per_clip = (int)((T + a.R - 1) / a.R);
  const int64_t items = (int64_t)n_clips * a.runs_per_clip;
  if (items > 0x3fffffffLL) return set_error(SAGA_ERR_INVALID, "stft: batch too large");
  a.n_items = (int)items;
  const int grid = (int)std::min<int64_t>(n_sm, items);
  kern<<<grid, THREADS, SMEM_BYTES, st>>>(a);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

int launch_stft_ring(const saga_stft_plan* p, const StftArgs& s, int n_clips, int64_t max_frames, cudaStream_t st) {
  // SAGA_STFT_RING_SHAPE=15: 15 + 1 warps at <= 128 registers (tuning aid); default 19 + 1 warps at <= 102
  const char* shape_opt = SAGA_OPT("SAGA_STFT_RING_SHAPE");
  const int shape = shape_opt ? atoi(shape_opt) : 19;
  if (shape == 15) return launch_ring_shape<15, 24>(p, s, n_clips, max_frames, st);
  return launch_ring_shape<19, 20>(p, s, n_clips, max_frames, st);
}

}  // namespace saga
