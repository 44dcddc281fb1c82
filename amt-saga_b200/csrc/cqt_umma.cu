// K2 contraction on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
//   C_o[t, :] = sum_n y_o[reflect(t*hop + n - n_fft/2)] * G_o[n, :]       (cqt.cu header)
//
// is a GEMM whose A operand is the strided-frame (Hankel) view of the decimated
// signal.  The Hankel matrix is never materialised: write n = q*hop + 4g + j
// (q < Q = n_fft/hop, g < hop/4, j < 4) and view the signal as rows of `hop`
// samples, Y[s][.] = y[(t0+s)*hop - n_fft/2 + .].  Then A[r, n] = Y[r + q][4g + j].
// Shared memory holds the tile's (128 + Q - 1) signal rows ONCE, as "planes"
//     plane g : row s -> 16 bytes = Y[s][4g .. 4g+3]           (row pitch 16 B)
// which is exactly the canonical K-major SWIZZLE_NONE UMMA layout with
// SBO = 128 B (8-row core matrices back to back) and LBO = plane pitch, so the
// operand for shift q is the SAME plane with the descriptor start address
// advanced by q rows (16*q bytes).  For hop == 4 there is one plane and the two
// 16-byte K-chunks of an MMA are consecutive shifts: LBO = 16 B.
// Every sample is fetched from L2 once per tile instead of n_fft/hop times.
//
// fp32 accuracy on TF32 tensor cores: operands are split x = hi + lo (both
// TF32, round-to-nearest) and three MMAs hi*hi + lo*hi + hi*lo accumulate in
// fp32 TMEM (n_split = 3); n_split = 1 keeps only hi*hi.
//
// Warp roles per persistent CTA (one per SM): 8 producer warps (global -> split
// -> planes), 1 MMA issuer (single elected thread), 4 epilogue warps
// (tcgen05.ld -> |re + i im| -> global).  mbarrier pipelines: smem full/empty per
// A stage, TMEM full/empty per accumulator buffer.  The bank G_o (packed for the
// B descriptor on the host) stays resident in SMEM while the CTA works through
// items of one octave (items are ordered octave-major).
#include <algorithm>
#include <cstring>
#include <vector>

#include "cqt_plan.cuh"
#include "saga_common.cuh"

namespace saga {

constexpr int UM_TILE_M = 128;
constexpr int UM_PRODUCER_WARPS = 8;
constexpr int UM_EPI_WARPS = 4;
constexpr int UM_THREADS = 32 * (1 + UM_EPI_WARPS + UM_PRODUCER_WARPS);   // warp0 = MMA, 1-4 epilogue, 5-12 producers
constexpr int UM_STAGES = 2;
constexpr int UM_PLANES_PER_STAGE = 8;      // 8 planes x 4 samples = 32 k-values per shift
constexpr int UM_MAX_OCT = 12;
constexpr uint32_t UM_SPIN_LIMIT = 1u << 27;

struct UmmaOct {
  const float* sig;
  const int64_t* sig_offsets;
  int64_t sig_stride;
  const float* b_hi;      // packed [n_fft/4][npad][4]
  const float* b_lo;
  int level, hop, n_fft, ncol, npad, first_bin;
  int planes, n_stages, Q, rows, rows_pad;
  int64_t item_begin;
};

struct UmmaArgs {
  UmmaOct oct[UM_MAX_OCT];
  int n_oct, n_clips, tiles_per_clip, early_factor, n_bins, n_split;
  int64_t total_items;
  const int64_t* clip_lens;
  const int32_t* clip_frames;
  float* mag_out;
  float2* cplx_out;
  int64_t frame_pitch, out_clip_stride;
  uint32_t b_region_bytes;   // one of hi / lo
  uint32_t a_region_bytes;   // one of hi / lo, per stage
  uint32_t tmem_cols;        // allocation (power of two >= 2*npad_max)
  uint32_t acc_stride;       // columns between the two accumulator buffers
  int* error_flag;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// bounded wait: a pipeline bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > UM_SPIN_LIMIT) {
      if (error_flag) atomicExch(error_flag, 1);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct ItemInfo {
  int o, clip, t0, T;
  int64_t len;
};

__device__ __forceinline__ ItemInfo decode_item(const UmmaArgs& a, int64_t item) {
  ItemInfo it;
  int o = 0;
  while (o + 1 < a.n_oct && item >= a.oct[o + 1].item_begin) ++o;
  const int64_t local = item - a.oct[o].item_begin;
  it.o = o;
  it.clip = (int)(local / a.tiles_per_clip);
  it.t0 = (int)(local % a.tiles_per_clip) * UM_TILE_M;
  it.T = a.clip_frames[it.clip];
  int64_t len = a.clip_lens[it.clip];
  if (a.early_factor > 1) len = (len + a.early_factor - 1) / a.early_factor;
  for (int s = 0; s < a.oct[o].level; ++s) len = (len + 1) >> 1;
  it.len = len;
  return it;
}

__global__ void __launch_bounds__(UM_THREADS, 1) cqt_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* b_hi_s = smem_raw;
  uint8_t* b_lo_s = b_hi_s + a.b_region_bytes;
  uint8_t* a_base = b_lo_s + a.b_region_bytes;                  // stages: [hi | lo] x UM_STAGES
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + 2 * UM_STAGES * a.a_region_bytes);
  uint64_t* full = bars;                    // [UM_STAGES]
  uint64_t* empty = bars + UM_STAGES;       // [UM_STAGES]
  uint64_t* tfull = bars + 2 * UM_STAGES;   // [2]
  uint64_t* tempty = bars + 2 * UM_STAGES + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * UM_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < UM_STAGES; ++s) {
      mbar_init(&full[s], UM_PRODUCER_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], UM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int64_t G = gridDim.x;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    uint32_t it_stage = 0, it_acc = 0;
    for (int64_t item = blockIdx.x; item < a.total_items; item += G) {
      const ItemInfo inf = decode_item(a, item);
      if (inf.t0 >= inf.T) continue;
      const UmmaOct& oc = a.oct[inf.o];
      // instruction descriptor: D=f32, A=B=tf32, K-major both, N = npad, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(oc.npad >> 3) << 17) |
                             ((uint32_t)(UM_TILE_M >> 4) << 24);
      const uint32_t acc = it_acc & 1;
      mbar_wait(&tempty[acc], ((it_acc >> 1) & 1) ^ 1, a.error_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * a.acc_stride;
      const uint32_t plane_bytes = (uint32_t)oc.rows_pad * 16u;
      const uint32_t b_chunk_bytes = (uint32_t)oc.npad * 16u;
      uint32_t first = 1;
      for (int st = 0; st < oc.n_stages; ++st, ++it_stage) {
        const uint32_t s = it_stage % UM_STAGES;
        mbar_wait(&full[s], (it_stage / UM_STAGES) & 1, a.error_flag);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t ah = smem_u32(a_base + (2 * s) * a.a_region_bytes);
          const uint32_t al = smem_u32(a_base + (2 * s + 1) * a.a_region_bytes);
          const uint32_t bh = smem_u32(b_hi_s), bl = smem_u32(b_lo_s);
          const int g0 = st * UM_PLANES_PER_STAGE;
          const int np = min(UM_PLANES_PER_STAGE, oc.planes - g0);
          if (oc.planes >= 2) {
            for (int q = 0; q < oc.Q; ++q) {
              for (int p = 0; p < np; p += 2) {
                const uint32_t a_off = (uint32_t)p * plane_bytes + (uint32_t)q * 16u;
                const uint32_t ck = (uint32_t)(q * oc.planes + g0 + p);
                const uint64_t dah = smem_desc(ah + a_off, plane_bytes, 128);
                const uint64_t dbh = smem_desc(bh + ck * b_chunk_bytes, b_chunk_bytes, 128);
                tc_mma_tf32(d_tmem, dah, dbh, idesc, first ? 0u : 1u);
                first = 0;
                if (a.n_split == 3) {
                  const uint64_t dal = smem_desc(al + a_off, plane_bytes, 128);
                  const uint64_t dbl = smem_desc(bl + ck * b_chunk_bytes, b_chunk_bytes, 128);
                  tc_mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                  tc_mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                }
              }
            }
          } else {
            // hop == 4: one plane; the two K-chunks of an MMA are shifts q and q+1
            for (int q = 0; q < oc.Q; q += 2) {
              const uint32_t a_off = (uint32_t)q * 16u;
              const uint64_t dah = smem_desc(ah + a_off, 16, 128);
              const uint64_t dbh = smem_desc(bh + (uint32_t)q * b_chunk_bytes, b_chunk_bytes, 128);
              tc_mma_tf32(d_tmem, dah, dbh, idesc, first ? 0u : 1u);
              first = 0;
              if (a.n_split == 3) {
                const uint64_t dal = smem_desc(al + a_off, 16, 128);
                const uint64_t dbl = smem_desc(bl + (uint32_t)q * b_chunk_bytes, b_chunk_bytes, 128);
                tc_mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                tc_mma_tf32(d_tmem, dah, dbl, idesc, 1u);
              }
            }
          }
          tc_commit(&empty[s]);                       // smem stage reusable once these MMAs retire
          if (st == oc.n_stages - 1) tc_commit(&tfull[acc]);   // accumulator complete
        }
        __syncwarp();
      }
      ++it_acc;
    }
  } else if (warp <= UM_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                 // a warp may only touch TMEM lanes [32*(warp%4), +32)
    uint32_t it_acc = 0;
    for (int64_t item = blockIdx.x; item < a.total_items; item += G) {
      const ItemInfo inf = decode_item(a, item);
      if (inf.t0 >= inf.T) continue;
      const UmmaOct& oc = a.oct[inf.o];
      const uint32_t acc = it_acc & 1;
      mbar_wait(&tfull[acc], (it_acc >> 1) & 1, a.error_flag);
      tc_fence_after();
      const int t = inf.t0 + ew * 32 + lane;
      const int64_t row = (int64_t)inf.clip * a.out_clip_stride + (int64_t)t * a.frame_pitch;
      for (int c0 = 0; c0 < oc.npad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + acc * a.acc_stride + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (t < inf.T) {
#pragma unroll
          for (int f = 0; f < 16; f += 2) {
            const int col = c0 + f;
            const int bin = oc.first_bin + (col >> 1);
            if (col < oc.ncol && bin >= 0 && bin < a.n_bins) {
              const float re = __uint_as_float(v[f]), im = __uint_as_float(v[f + 1]);
              a.mag_out[row + bin] = sqrtf(re * re + im * im);
              if (a.cplx_out) a.cplx_out[row + bin] = make_float2(re, im);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      ++it_acc;
    }
  } else {
    // =========================== producers ===========================
    const int ptid = threadIdx.x - 32 * (1 + UM_EPI_WARPS);
    constexpr int PT = 32 * UM_PRODUCER_WARPS;
    uint32_t it_stage = 0;
    int cur_oct = -1;
    for (int64_t item = blockIdx.x; item < a.total_items; item += G) {
      const ItemInfo inf = decode_item(a, item);
      if (inf.t0 >= inf.T) continue;
      const UmmaOct& oc = a.oct[inf.o];
      if (inf.o != cur_oct) {
        // new bank: every MMA that reads the old one must have retired
        if (it_stage > 0) {
          const uint32_t last = it_stage - 1;
          mbar_wait(&empty[last % UM_STAGES], (last / UM_STAGES) & 1, a.error_flag);
        }
        const int n16 = (oc.n_fft / 4) * oc.npad;     // 16-byte units per split
        const float4* gh = reinterpret_cast<const float4*>(oc.b_hi);
        const float4* gl = reinterpret_cast<const float4*>(oc.b_lo);
        float4* sh = reinterpret_cast<float4*>(b_hi_s);
        float4* sl = reinterpret_cast<float4*>(b_lo_s);
        for (int i = ptid; i < n16; i += PT) {
          sh[i] = __ldg(gh + i);
          if (a.n_split == 3) sl[i] = __ldg(gl + i);
        }
        cur_oct = inf.o;
      }
      const float* y = oc.sig + (oc.sig_offsets ? oc.sig_offsets[inf.clip] : (int64_t)inf.clip * oc.sig_stride);
      const bool base_al = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
      const int64_t origin = (int64_t)inf.t0 * oc.hop - (oc.n_fft >> 1);   // sample index of Y[0][0]
      for (int st = 0; st < oc.n_stages; ++st, ++it_stage) {
        const uint32_t s = it_stage % UM_STAGES;
        mbar_wait(&empty[s], ((it_stage / UM_STAGES) & 1) ^ 1, a.error_flag);
        float4* dh = reinterpret_cast<float4*>(a_base + (2 * s) * a.a_region_bytes);
        float4* dl = reinterpret_cast<float4*>(a_base + (2 * s + 1) * a.a_region_bytes);
        const int g0 = st * UM_PLANES_PER_STAGE;
        const int np = min(UM_PLANES_PER_STAGE, oc.planes - g0);
        const int total = np * oc.rows;
        for (int e = ptid; e < total; e += PT) {
          const int g = e % np, srow = e / np;
          const int64_t idx = origin + (int64_t)srow * oc.hop + 4 * (g0 + g);
          float4 x;
          if (base_al && idx >= 0 && idx + 3 < inf.len) {
            x = __ldg(reinterpret_cast<const float4*>(y + idx));
          } else {
            x.x = __ldg(y + reflect_index(idx, inf.len));
            x.y = __ldg(y + reflect_index(idx + 1, inf.len));
            x.z = __ldg(y + reflect_index(idx + 2, inf.len));
            x.w = __ldg(y + reflect_index(idx + 3, inf.len));
          }
          float4 h;
          h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
          const int dst = g * oc.rows_pad + srow;
          dh[dst] = h;
          if (a.n_split == 3) {
            float4 l;
            l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
            dl[dst] = l;
          }
        }
        fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
  }
}

__global__ void cqt_zero_pad_kernel(float* mag, float2* cplx, const int32_t* clip_frames, int n_bins,
                                    int64_t pitch, int64_t clip_stride) {
  const int clip = blockIdx.y;
  const int T = clip_frames[clip];
  for (int t = blockIdx.x * blockDim.y + threadIdx.y; t < T; t += gridDim.x * blockDim.y)
    for (int64_t k = n_bins + threadIdx.x; k < pitch; k += blockDim.x) {
      mag[clip * clip_stride + t * pitch + k] = 0.f;
      if (cplx) cplx[clip * clip_stride + t * pitch + k] = make_float2(0.f, 0.f);
    }
}

// ------------------------------------------------------------------ host side
struct OctPack {
  float* d_hi = nullptr;
  float* d_lo = nullptr;
  int npad = 0;
};

}  // namespace saga

struct CqtUmmaState {
  std::vector<saga::OctPack> packs;
  bool supported = false;
  uint32_t b_region_bytes = 0, a_region_bytes = 0, tmem_cols = 0, acc_stride = 0;
  size_t smem_bytes = 0;
  int* d_error = nullptr;
  int num_sms = 0;
  int n_split = 3;
};

namespace saga {

static float tf32_rna_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0x1000u;
  u &= ~0x1FFFu;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

void cqt_umma_plan_init(saga_cqt_plan* p) {
  CqtUmmaState* st = new CqtUmmaState();
  p->umma = st;
  if ((int)p->oct.size() > UM_MAX_OCT) return;
  uint32_t bmax = 0, amax = 0;
  int npad_max = 0;
  for (auto& o : p->oct) {
    const int ncol = 2 * o.n_filters;
    const int npad = (ncol + 15) & ~15;
    if (npad > 256) return;
    if (o.hop < 4 || (o.hop % 4) != 0 || (o.n_fft % o.hop) != 0 || (o.n_fft % 8) != 0) return;
    const int planes = o.hop / 4;
    if (planes >= 2 && (std::min(planes, UM_PLANES_PER_STAGE) % 2) != 0) return;
    const int Q = o.n_fft / o.hop;
    if (planes == 1 && (Q % 2) != 0) return;
    const int rows = UM_TILE_M + Q - 1;
    const int rows_pad = rows | 1;
    bmax = std::max<uint32_t>(bmax, (uint32_t)o.n_fft * npad * 4u);
    amax = std::max<uint32_t>(amax, (uint32_t)std::min(planes, UM_PLANES_PER_STAGE) * rows_pad * 16u);
    npad_max = std::max(npad_max, npad);
  }
  amax = (amax + 127u) & ~127u;
  bmax = (bmax + 127u) & ~127u;
  const size_t smem = 2ull * bmax + 2ull * UM_STAGES * amax + 256;
  if (smem > 225 * 1024) return;
  uint32_t cols = 32;
  while (cols < 2u * npad_max) cols <<= 1;
  if (cols > 512) return;
  st->b_region_bytes = bmax;
  st->a_region_bytes = amax;
  st->tmem_cols = cols;
  st->acc_stride = cols / 2;
  st->smem_bytes = smem;
  // pack each bank for the B descriptor: [n_fft/4][npad][4], split into TF32 hi / lo
  for (auto& o : p->oct) {
    OctPack pk;
    const int ncol = 2 * o.n_filters;
    pk.npad = (ncol + 15) & ~15;
    const size_t n = (size_t)o.n_fft * pk.npad;
    std::vector<float> hi(n, 0.f), lo(n, 0.f);
    for (int k = 0; k < o.n_fft; ++k)
      for (int c = 0; c < ncol; ++c) {
        const float b = o.bank_host[(size_t)k * ncol + c];
        const float h = tf32_rna_host(b);
        const size_t dst = ((size_t)(k / 4) * pk.npad + c) * 4 + (k % 4);
        hi[dst] = h;
        lo[dst] = tf32_rna_host(b - h);
      }
    if (cudaMalloc(&pk.d_hi, n * 4) != cudaSuccess || cudaMalloc(&pk.d_lo, n * 4) != cudaSuccess) return;
    cudaMemcpy(pk.d_hi, hi.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(pk.d_lo, lo.data(), n * 4, cudaMemcpyHostToDevice);
    st->packs.push_back(pk);
  }
  if (cudaMalloc(&st->d_error, sizeof(int)) != cudaSuccess) return;
  cudaMemset(st->d_error, 0, sizeof(int));
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaFuncSetAttribute(cqt_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  st->supported = true;
}

void cqt_umma_plan_free(saga_cqt_plan* p) {
  if (!p->umma) return;
  for (auto& pk : p->umma->packs) {
    cudaFree(pk.d_hi);
    cudaFree(pk.d_lo);
  }
  cudaFree(p->umma->d_error);
  delete p->umma;
  p->umma = nullptr;
}

int cqt_umma_exec(const saga_cqt_plan* p, const CqtLevels& lv, int n_clips, int64_t max_len, int64_t T_max,
                  float* mag_out, float2* cplx_out, int64_t frame_pitch, int64_t out_clip_stride,
                  int n_split, cudaStream_t stream) {
  const CqtUmmaState* st = p->umma;
  if (!st || !st->supported) return set_error(SAGA_ERR_UNSUPPORTED, "cqt: plan does not fit the tcgen05 path");
  UmmaArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n_oct = (int)p->oct.size();
  a.n_clips = n_clips;
  a.tiles_per_clip = (int)((T_max + UM_TILE_M - 1) / UM_TILE_M);
  a.early_factor = p->early_factor;
  a.n_bins = p->n_bins;
  a.n_split = (n_split == 1) ? 1 : 3;
  a.clip_lens = lv.clip_lens;
  a.clip_frames = lv.clip_frames;
  a.mag_out = mag_out;
  a.cplx_out = cplx_out;
  a.frame_pitch = frame_pitch;
  a.out_clip_stride = out_clip_stride;
  a.b_region_bytes = st->b_region_bytes;
  a.a_region_bytes = st->a_region_bytes;
  a.tmem_cols = st->tmem_cols;
  a.acc_stride = st->acc_stride;
  a.error_flag = st->d_error;
  const int64_t per_oct = (int64_t)n_clips * a.tiles_per_clip;
  for (int i = 0; i < a.n_oct; ++i) {
    const CqtOctaveDev& o = p->oct[i];
    UmmaOct& u = a.oct[i];
    const bool raw = (o.level == 0 && p->early_factor == 1);
    u.sig = raw ? lv.wav : lv.lvl[o.level];
    u.sig_offsets = raw ? lv.clip_offsets : nullptr;
    u.sig_stride = raw ? 0 : lv.pitch[o.level];
    u.b_hi = st->packs[i].d_hi;
    u.b_lo = st->packs[i].d_lo;
    u.level = o.level;
    u.hop = o.hop;
    u.n_fft = o.n_fft;
    u.ncol = 2 * o.n_filters;
    u.npad = st->packs[i].npad;
    u.first_bin = o.first_bin;
    u.planes = o.hop / 4;
    u.n_stages = (u.planes + UM_PLANES_PER_STAGE - 1) / UM_PLANES_PER_STAGE;
    u.Q = o.n_fft / o.hop;
    u.rows = UM_TILE_M + u.Q - 1;
    u.rows_pad = u.rows | 1;
    u.item_begin = per_oct * i;
  }
  a.total_items = per_oct * a.n_oct;
  if (a.total_items <= 0) return SAGA_OK;
  if (frame_pitch > p->n_bins) {
    dim3 grid(8, n_clips), block(32, 8);
    cqt_zero_pad_kernel<<<grid, block, 0, stream>>>(mag_out, cplx_out, lv.clip_frames, p->n_bins, frame_pitch,
                                                    out_clip_stride);
    SAGA_LAUNCH_CHECK();
  }
  const int grid = (int)std::min<int64_t>(a.total_items, st->num_sms > 0 ? st->num_sms : 148);
  cqt_umma_kernel<<<grid, UM_THREADS, st->smem_bytes, stream>>>(a);
  SAGA_LAUNCH_CHECK();
  return SAGA_OK;
}

}  // namespace saga
