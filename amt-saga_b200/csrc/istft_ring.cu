// K4, ring form: inverse STFT (librosa.istft as reached from /root/reference/util_audio.py:92-104: the waveform a
// window gets back from mag * ph after a subtraction) for n_fft 2048 / hop 512 / periodic Hann.
//
// The first-generation kernel (stft.cu: istft_kernel) gives a CTA FO output hops, inverse-transforms every frame that
// overlaps them (1.5-2x halo recompute) into shared memory, meets at a barrier and gathers.  Here every frame is
// inverse-transformed exactly ONCE:
//   * persistent CTAs, NW autonomous warps, jobs (= frames, in frame order) drawn from a shared counter, no CTA
//     barrier after start-up;
//   * a warp loads its frame's half spectrum straight from global memory (mag * phase or complex; every byte is
//     read once, there is nothing to stage), rebuilds conj Z[k] with the packed-fp32 formulation of fft_packed.cuh
//     (bins k / M-k are produced by the same lane and cross to their owner by warp shuffles), runs the two
//     radix-32 passes of the forward ring kernel, multiplies by window / M;
//   * overlap-add through a 4-block shared-memory ring with an IN-ORDER COMMIT: the warp that holds frame t adds its
//     four hop-sized pieces after frame t-1 has committed (ascending frame order = the reference's accumulation
//     order, deterministic, no atomics), which completes block t: that block is normalised by the window
//     sum-of-squares and stored while it is still in registers.
// A run (item) is a range of output blocks of one clip; it re-transforms the 3 frames before its first block
// (their earlier blocks belong to the previous run) -- 3 extra frames per ~48.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fft_packed.cuh"
#include "saga_common.cuh"
#include "stft_plan.cuh"

namespace saga {

namespace iring {

using namespace ring;

constexpr int NW = 16;                    // warps per CTA (one frame in flight each), <= 128 registers
constexpr int THREADS = NW * 32;
constexpr int HOP = 512, NFFT = 2048, M = 1024;
constexpr int EXW = 32 * 34;              // float2 of the exchange layout (8704 bytes)
constexpr int PITCH = M + 4;              // frame pitch (floats / float2) the staging path needs: 1028
constexpr int WBUF = 12352;               // bytes per warp: staging [phasors or complex: 8224][magnitudes: 4112], later the exchange buffer
// table image (float2 units), per-lane rows read with conflict-free 16-byte loads
constexpr int T_TW0 = 0;                  // [32][34]: exp(-2 pi i lane rp / 1024)
constexpr int T_WQ = T_TW0 + 32 * 34;     // [32][18]: 1/2 (cos, sin)(2 pi k / 2048), k = lane + 32 i, i <= 16
constexpr int T_WIN = T_WQ + 32 * 18;     // [32][34]: (w[2n], -w[2n+1]) / M, n = lane + 32 k2
constexpr int T_INV = T_WIN + 32 * 34;    // [32][10]: 1 / sum_j w^2[r + j hop] for samples (2p, 2p+1), p = lane + 32 q, q < 8
constexpr int T_TOTAL = T_INV + 32 * 10;
constexpr size_t SMEM_TABLES = (size_t)T_TOTAL * 8;
constexpr size_t SMEM_EXCH = (size_t)NW * WBUF;
constexpr size_t SMEM_ACC = (size_t)4 * HOP * 4;
constexpr size_t SMEM_BYTES = SMEM_TABLES + SMEM_EXCH + SMEM_ACC + 16 + NW * 8;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

struct Args {
  const float2* cplx_in;
  const float* mag_in;
  const float2* phase_in;
  float* wav_out;
  const float* wsq;          // [NFFT] window^2 (edge blocks)
  const float2* tables;
  Rot32 w;
  int64_t frame_pitch, in_clip_stride, wav_clip_stride;
  int center, T;             // frames per clip
  int RB;                    // output blocks per run
  int runs_per_clip;
  int n_items;
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > SPIN_LIMIT) __trap();
  }
}

// what a job is: frame t of a run; derived from the job number alone, so a warp can look at its NEXT job early
struct Job {
  int clip, t, b0, b1;
  bool real, emit, first;
};

__global__ void __launch_bounds__(THREADS, 1) istft_ring_kernel(const Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tab = reinterpret_cast<float2*>(smem_raw);
  unsigned char* exch = smem_raw + SMEM_TABLES;
  float* acc = reinterpret_cast<float*>(smem_raw + SMEM_TABLES + SMEM_EXCH);
  uint32_t* committed = reinterpret_cast<uint32_t*>(smem_raw + SMEM_TABLES + SMEM_EXCH + SMEM_ACC);
  uint32_t* next_job = committed + 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(next_job + 3);          // one staging barrier per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < T_TOTAL / 2; i += THREADS)
    reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(a.tables) + i);
  if (threadIdx.x == 0) {
    *committed = 0;
    *next_job = 0;
    for (int w = 0; w < NW; ++w) mbar_init(bars + w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();                // the only CTA-wide barrier

  const float4* tw04 = reinterpret_cast<const float4*>(tab + T_TW0 + lane * 34);
  const float4* wq4 = reinterpret_cast<const float4*>(tab + T_WQ + lane * 18);
  const float4* win4 = reinterpret_cast<const float4*>(tab + T_WIN + lane * 34);
  const float4* inv4 = reinterpret_cast<const float4*>(tab + T_INV + lane * 10);
  unsigned char* wbuf = exch + (size_t)warp * WBUF;
  const float2* st_ph = reinterpret_cast<const float2*>(wbuf);           // staged phasors / complex bins
  const float* st_mag = reinterpret_cast<const float*>(wbuf + PITCH * 8); // staged magnitudes
  float2* ex = reinterpret_cast<float2*>(wbuf);
  float2* ex_st = ex + lane;
  const float4* ex_ld = reinterpret_cast<const float4*>(ex + lane * 34);
  float2* acc2 = reinterpret_cast<float2*>(acc) + lane;       // block slot s, pair p = lane + 32 q: acc2[s * 256 + 32 q]
  uint64_t* bar = bars + warp;
  uint32_t bar_phase = 0;

  const int G = gridDim.x, RB = a.RB, JPI = RB + 3;           // jobs per item
  const int T = a.T;
  const int nB = T + 3;                                       // hop blocks of the padded signal
  const int keep_lo = a.center ? 2 : 0, keep_hi = a.center ? T : T + 2;   // blocks that exist in the output
  const int64_t trim = a.center ? NFFT / 2 : 0;
  const uint32_t n_jobs = (uint32_t)((a.n_items - (int)blockIdx.x + G - 1) / G) * (uint32_t)JPI;
  const int partner = (32 - lane) & 31;
  const f2 halfpm = F2(0.5f, -0.5f);

  auto decode = [&](uint32_t J) {
    Job j;
    const int k_item = (int)(J / (uint32_t)JPI), f = (int)J - k_item * JPI;
    const int item = (int)blockIdx.x + k_item * G;
    j.clip = item / a.runs_per_clip;
    j.b0 = (item - j.clip * a.runs_per_clip) * RB;             // first block this run emits
    j.b1 = min(j.b0 + RB, nB);
    j.t = j.b0 - 3 + f;                                        // frame of this job; it completes block t
    j.real = j.t >= 0 && j.t < T && j.t < j.b1;                // contributes samples
    j.emit = j.t >= j.b0 && j.t < j.b1 && j.t >= keep_lo && j.t <= keep_hi;
    j.first = j.t == max(j.b0 - 3, 0);                         // first frame of the run: its pieces start their blocks
    return j;
  };
  auto draw = [&]() {
    uint32_t J = 0;
    if (lane == 0) J = atomicAdd(next_job, 1u);
    return __shfl_sync(0xffffffffu, J, 0);
  };
  // the frame's half spectrum goes straight from global memory into this warp's buffer (one DRAM round trip,
  // no registers held while it is in flight)
  auto stage = [&](const Job& j) {
    if (!j.real || lane != 0) return;
    const int64_t row = (int64_t)j.clip * a.in_clip_stride + (int64_t)j.t * a.frame_pitch;
    if (a.cplx_in) {
      mbar_arrive_expect_tx(bar, PITCH * 8);
      bulk_g2s(smem_u32(wbuf), a.cplx_in + row, PITCH * 8, bar);
    } else {
      mbar_arrive_expect_tx(bar, PITCH * 12);
      bulk_g2s(smem_u32(wbuf), a.phase_in + row, PITCH * 8, bar);
      bulk_g2s(smem_u32(wbuf + PITCH * 8), a.mag_in + row, PITCH * 4, bar);
    }
  };

  uint32_t J = draw();
  Job job;
  if (J < n_jobs) {
    job = decode(J);
    stage(job);
  }
  while (J < n_jobs) {
    const int clip = job.clip, t = job.t;
    const bool real = job.real, emit = job.emit, first = job.first;
    uint32_t Jn = n_jobs;
    Job jobn;
    f2 v[32];
    if (real) {
      mbar_wait(bar, bar_phase);
      bar_phase ^= 1;
      // ---- half spectrum -> conj Z[k] (k = lane + 32 i) and H = conj(conj Z[M - k]) for the partner lane ----
      f2 give[16];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w4 = wq4[j4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * j4 + h;
          const int ku = lane + 32 * i, kd = M - lane - 32 * i;
          f2 ak = st_ph[ku], am = st_ph[kd];
          if (!a.cplx_in) {
            const float mk = st_mag[ku], mm = st_mag[kd];
            ak = mul2(F2(mk, mk), ak);
            am = mul2(F2(mm, mm), am);
          }
          if (i == 0 && lane == 0) { ak.y = 0.f; am.y = 0.f; }      // irfft ignores Im of DC and Nyquist
          const f2 p = add2(ak, F2(am.x, -am.y));                   // ak + conj am
          const f2 q = add2(ak, F2(-am.x, am.y));                   // ak - conj am
          const f2 wq = h ? F2(w4.z, w4.w) : F2(w4.x, w4.y);        // 1/2 (c, s)
          // G = (O.y, O.x),  O = 1/2 q (c + i s)
          const f2 Gv = fma2(F2(q.y, q.y), F2(wq.x, -wq.y), mul2(F2(q.x, q.x), F2(wq.y, wq.x)));
          v[i] = fma2(p, halfpm, F2(-Gv.x, -Gv.y));                 // conj Z[k]    = (p.x/2 - G.x, -p.y/2 - G.y)
          give[i] = fma2(p, halfpm, Gv);                            // H, conj Z[M-k] = conj H
        }
      }
      // bins M - k cross to their owner: lane (32 - lane) & 31, register 31 - i; lane 0 pairs with itself
      // (M - 32 i = 32 (32 - i)) and owns the self-paired bin M/2
      f2 x512 = st_ph[M / 2];
      if (!a.cplx_in) { const float m5 = st_mag[M / 2]; x512 = mul2(F2(m5, m5), x512); }
#pragma unroll
      for (int r = 16; r < 32; ++r) {
        const int i = 31 - r;
        f2 h2;
        h2.x = __shfl_sync(0xffffffffu, give[i].x, partner);
        h2.y = __shfl_sync(0xffffffffu, give[i].y, partner);
        if (lane == 0) h2 = (r == 16) ? F2(x512.x, -x512.y) : give[32 - r];   // conj Z[512] = X[512]
        v[r] = h2;
      }
      __syncwarp();               // every lane has read its staged bins: the buffer becomes the exchange buffer
      // ---- pass 1: first radix-2 stage takes the received values conjugated, then the usual network ----
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const f2 x0 = v[r], c = F2(v[r + 16].x, -v[r + 16].y);
        v[r] = add2(x0, c);
        v[r + 16] = mulw(sub2(x0, c), r, false, a.w);
      }
      fft32_tail<8>(v, a.w);
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
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 q4 = ex_ld[j];
        v[2 * j] = F2(q4.x, q4.y);
        v[2 * j + 1] = F2(q4.z, q4.w);
      }
      __syncwarp();               // buffer free: the next frame's spectrum can start to arrive
    }
    // look at the next job now, so that its spectrum streams in behind the second pass and the commit
    Jn = draw();
    if (Jn < n_jobs) {
      jobn = decode(Jn);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads / writes of the buffer before the async writes
      stage(jobn);
    }
    if (real) fft32_tail<16>(v, a.w);      // v[bitrev(k2)] = y[lane + 32 k2];  x[2n] = Re y[n] / M,  x[2n+1] = -Im y[n] / M

    // ---- in-order commit: wait for frame t - 1, overlap-add, emit block t ----
    {
      uint32_t spins = 0;
      while (ld_acquire(committed) != J) {
        __nanosleep(20);
        if (++spins > SPIN_LIMIT) __trap();
      }
    }
    f2 fin[8];
    if (real) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 w4 = win4[j];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k2 = 2 * j + h, piece = k2 >> 3, q = k2 & 7;
          const f2 o = mul2(v[bitrev(k2, 32)], h ? F2(w4.z, w4.w) : F2(w4.x, w4.y));
          f2* dst = acc2 + ((t + piece) & 3) * (HOP / 2) + 32 * q;
          f2 s;
          if (piece == 3) s = o;
          else s = first ? o : add2(*dst, o);
          if (piece == 0) fin[q] = s;
          else *dst = s;
        }
      }
    } else if (emit) {
#pragma unroll
      for (int q = 0; q < 8; ++q) fin[q] = acc2[(t & 3) * (HOP / 2) + 32 * q];
    }
    __syncwarp();
    if (lane == 0) st_release(committed, J + 1);       // the next frame may add; the finished block is in registers
    if (emit) {
      float* y = a.wav_out + (int64_t)clip * a.wav_clip_stride + ((int64_t)t * HOP - trim);
      if (t >= 3 && t <= T - 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 n4 = inv4[q4];
          fin[2 * q4] = mul2(fin[2 * q4], F2(n4.x, n4.y));
          fin[2 * q4 + 1] = mul2(fin[2 * q4 + 1], F2(n4.z, n4.w));
        }
      } else {
        // clip edge: fewer than four frames overlap this block
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = 2 * (lane + 32 * q);
          float s0 = 0.f, s1 = 0.f;
          for (int j = 3; j >= 0; --j)
            if (t - j >= 0 && t - j < T) { s0 += __ldg(a.wsq + r + j * HOP); s1 += __ldg(a.wsq + r + 1 + j * HOP); }
          if (s0 > 1.17549435e-38f) fin[q].x /= s0;
          if (s1 > 1.17549435e-38f) fin[q].y /= s1;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<f2*>(y + 2 * (lane + 32 * q)) = fin[q];
    }
    J = Jn;
    job = jobn;
  }
}

}  // namespace iring

bool istft_ring_supported(const saga_stft_plan* p) {
  return p->n_fft == iring::NFFT && p->hop == iring::HOP && p->default_window && p->d_iring_tables != nullptr;
}

int istft_ring_build_tables(saga_stft_plan* p) {
  using namespace iring;
  p->d_iring_tables = nullptr;
  p->d_wsq = nullptr;
  if (!(p->n_fft == NFFT && p->hop == HOP && p->default_window)) return SAGA_OK;
  const double PI = 3.14159265358979323846;
  std::vector<float> w(NFFT), wsq(NFFT);
  for (int n = 0; n < NFFT; ++n) {
    w[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / NFFT));
    wsq[n] = w[n] * w[n];
  }
  std::vector<float2> t(T_TOTAL, make_float2(0.f, 0.f));
  for (int lane = 0; lane < 32; ++lane) {
    for (int rp = 0; rp < 32; ++rp) {
      const double ang = -2.0 * PI * (double)((lane * rp) % M) / (double)M;
      t[T_TW0 + lane * 34 + rp] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int i = 0; i <= 16; ++i) {
      const double th = 2.0 * PI * (lane + 32 * i) / (double)NFFT;
      t[T_WQ + lane * 18 + i] = make_float2((float)(0.5 * std::cos(th)), (float)(0.5 * std::sin(th)));
    }
    for (int k2 = 0; k2 < 32; ++k2) {
      const int n = lane + 32 * k2;
      t[T_WIN + lane * 34 + k2] = make_float2(w[2 * n] / (float)M, -w[2 * n + 1] / (float)M);
    }
    for (int q = 0; q < 8; ++q) {
      const int r = 2 * (lane + 32 * q);
      float s0 = 0.f, s1 = 0.f;
      for (int j = 3; j >= 0; --j) { s0 += wsq[r + j * HOP]; s1 += wsq[r + 1 + j * HOP]; }
      t[T_INV + lane * 10 + q] = make_float2(s0 > 1.17549435e-38f ? 1.0f / s0 : 1.0f, s1 > 1.17549435e-38f ? 1.0f / s1 : 1.0f);
    }
  }
  SAGA_CUDA_OK(cudaMalloc(&p->d_iring_tables, sizeof(float2) * T_TOTAL));
  SAGA_CUDA_OK(cudaMemcpy(p->d_iring_tables, t.data(), sizeof(float2) * T_TOTAL, cudaMemcpyHostToDevice));
  SAGA_CUDA_OK(cudaMalloc(&p->d_wsq, sizeof(float) * NFFT));
  SAGA_CUDA_OK(cudaMemcpy(p->d_wsq, wsq.data(), sizeof(float) * NFFT, cudaMemcpyHostToDevice));
  return SAGA_OK;
}

int launch_istft_ring(const saga_stft_plan* p, const void* cplx_in, const float* mag_in, const void* phase_in,
                      int n_clips, int n_frames, int64_t frame_pitch, int64_t in_clip_stride, float* wav_out,
                      int64_t wav_clip_stride, cudaStream_t st) {
  using namespace iring;
  int dev = 0, n_sm = 0;
  SAGA_CUDA_OK(cudaGetDevice(&dev));
  SAGA_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  SAGA_CUDA_OK(cudaFuncSetAttribute(istft_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  Args a;
  a.cplx_in = (const float2*)cplx_in;
  a.mag_in = mag_in;
  a.phase_in = (const float2*)phase_in;
  a.wav_out = wav_out;
  a.wsq = p->d_wsq;
  a.tables = p->d_iring_tables;
  fill_rot32(a.w);
  a.frame_pitch = frame_pitch;
  a.in_clip_stride = in_clip_stride;
  a.wav_clip_stride = wav_clip_stride;
  a.center = p->center;
  a.T = n_frames;
  const int64_t nB = (int64_t)n_frames + 3;
  // a run re-transforms the 3 frames before its first block: longer runs = less halo (48 blocks: 6.3 %), as long as
  // the job list still balances over the CTAs
  int64_t run_cap = 48;
  if (const char* e = SAGA_OPT("SAGA_ISTFT_RING_RUN")) run_cap = std::max(8, atoi(e));
  const int64_t min_runs = (nB + run_cap - 1) / run_cap, max_runs = (nB + 7) / 8;
  const int64_t want = ((int64_t)n_sm * 24 + n_clips - 1) / n_clips;
  const int64_t runs = std::min(std::max(want, min_runs), std::max(max_runs, min_runs));
  a.RB = (int)((nB + runs - 1) / runs);
  if (a.RB < 4) a.RB = 4;
  a.runs_per_clip = (int)((nB + a.RB - 1) / a.RB);
  const int64_t items = (int64_t)n_clips * a.runs_per_clip;
  if (items > 0x3fffffffLL) return set_error(SAGA_ERR_INVALID, "istft: batch too large");
  a.n_items = (int)items;
  const int grid = (int)std::min<int64_t>(n_sm, items);
  istft_ring_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(a);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

}  // namespace saga
