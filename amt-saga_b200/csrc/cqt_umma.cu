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
// The level signals arrive REFLECT-PADDED (cqt.cu: cqt_pad_kernel), so every tile of
// every clip is a plain strided copy: no bounds tests, no index mirroring, always
// 16-byte aligned.
//
// fp32 accuracy on TF32 tensor cores: operands are split x = hi + lo and
// hi*hi + hi*lo + lo*hi is accumulated in fp32 TMEM (n_split = 3).  The bank is
// packed [B_hi | B_lo] along N (2*ncol columns), so per K-slice ONE MMA with
// N = n_main (A_hi x [B_hi|B_lo]) plus one with N = n_lo (A_lo x B_hi..., into the
// SAME accumulator columns) do the work of three; the epilogue adds column c and
// column ncol + c.  kind::tf32 ignores the low 13 mantissa bits of an operand, so
// the raw fp32 samples ARE A_hi; only lo = tf32(x - trunc13(x)) is computed.
//
// Measured (profiles/microbench): an SS-mode tcgen05.mma is paced by the shared-memory
// bytes it reads (A 4 KB + B 32*N bytes at 128 B/clk: 44 cycles for N<=32, 48 @ N=64),
// so the contraction is bound by one thread-issued MMA pair (~88 cycles) per 8 k-values.
//
// Warp roles per persistent CTA (one per SM):
//   4 MMA issuer warps  K-slices dealt round robin, one TMEM column group each
//   4 epilogue warps    tcgen05.ld, sum of the column groups, |re + i im| -> global
//   8 loader warps      cp.async 16 B (L2 -> plane layout), completion signalled to an
//                       mbarrier by cp.async.mbarrier.arrive -- loaders never wait for data
//   8 converter warps   lo plane from the landed hi plane, proxy fence, arrive `full`
// mbarrier pipelines: raw / full / empty per A stage, full / empty per TMEM buffer, one
// for the bank.  The bank G_o stays resident in SMEM while the CTA works through tiles of
// one octave (tiles are ordered octave-major) and is replaced by ONE bulk async copy.
// Partial tiles with <= tail_max valid frames are left to cqt_tail_kernel (below):
// an MMA costs the same for 5 valid rows as for 128.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "cqt_plan.cuh"
#include "saga_common.cuh"
#include "umma_ptx.cuh"

namespace saga {

constexpr int UM_TILE_M = 128;
constexpr int UM_MMA_WARPS = 4;             // issuer warps: K-slices round-robin, one TMEM column group each
constexpr int UM_EPI_WARPS = 4;
constexpr int UM_LOAD_WARPS = 8;
constexpr int UM_CONV_WARPS = 8;
constexpr int UM_THREADS = 32 * (UM_MMA_WARPS + UM_EPI_WARPS + UM_LOAD_WARPS + UM_CONV_WARPS);
constexpr int UM_TAIL_WARPS = 4;            // warps per CTA of cqt_tail_kernel
constexpr int UM_TAIL_MAX = 16;             // largest partial tile (frames) left to cqt_tail_kernel
constexpr int UM_MAX_STAGES = 8;            // barrier array capacity; the actual count is a plan parameter
constexpr int UM_MAX_OCT = 12;
constexpr int UM_MAX_COLS = 128;
constexpr int UM_PROF_SLOTS = 24;
constexpr uint32_t UM_SMEM_LIMIT = 232448 - 1024;   // 227 KB opt-in maximum, minus slack

struct UmmaOct {
  const float* sig;       // padded level signal, element 0 of a clip = sample -n_fft/2
  int64_t sig_stride;     // floats per clip
  const float* b_pack;    // packed [n_fft/4][n_main][4]: rows [0,ncol) = TF32 hi, [ncol,2ncol) = lo, rest 0
  const float* bank_f32;  // [n_fft][ncol] fp32 bank of THIS octave (tail warps)
  uint32_t bank_bytes;
  int hop, n_fft, ncol, first_bin;
  int planes, n_stages, Q, rows, rows_pad;
  int n_main, n_lo;       // MMA N of the main and of the correction MMA
};

struct UmmaArgs {
  UmmaOct oct[UM_MAX_OCT];
  int n_oct, n_clips, n_bins, n_split;
  uint32_t tiles_per_clip, per_oct, total_items;
  const int32_t* clip_frames;
  int uniform_T;             // > 0: every clip has this many frames (no per-item loads)
  int tail_max;              // tiles with <= tail_max valid frames are skipped (cqt_tail_kernel does them)
  float* mag_out;
  float2* cplx_out;
  int64_t frame_pitch, out_clip_stride;
  uint32_t b_region_bytes;   // resident bank
  uint32_t a_region_bytes;   // one of hi / lo, per stage
  uint32_t tmem_cols;        // allocation (power of two)
  uint32_t acc_stride;       // columns between the two accumulator buffers
  uint32_t grp_stride;       // columns between issuer-warp column groups
  int* error_flag;
  int stages, pps;           // pipeline shape: SMEM stages, planes per stage (2, 4 or 8)
  int shared_bank;           // every octave's bank is a per-column multiple of octave 0's (librosa re-uses one
                             // basis): ONE resident bank, tiles interleave octaves, the epilogue applies col_scale
  const float* col_scale;    // [n_oct][UM_MAX_COLS]
  int debug;                 // SAGA_UMMA_DEBUG: 1 = no MMAs, 2 = no loads / conversion, 4 = no epilogue work,
                             // 16 = per-role cycle breakdown into `prof`
  long long* prof;           // [grid][UM_PROF_SLOTS]
};


// One (octave, clip, 128-frame tile) work item.  A CTA walks items blockIdx.x, +G, +2G, ...; the walk is
// kept incremental (two divisions at the start, a handful of adds per step) because every role of the CTA
// decodes every item and a division chain on one warp is ~500 cycles of exposed latency.
//   shared bank : item = local * n_oct + o   (octave-minor: consecutive tiles mix load-heavy and light octaves)
//   else        : item = o * per_oct + local (octave-major: the resident bank changes n_oct times per CTA)
// local = clip * tiles_per_clip + tile
struct Tile {
  int o, clip, t0, valid;
};
struct TileIter {
  uint32_t item, o, local, clip, tile;
  uint32_t dO, dL, dC, dT;
  __device__ __forceinline__ void init(const UmmaArgs& a, uint32_t first, uint32_t G) {
    item = first;
    if (a.shared_bank) {
      local = first / (uint32_t)a.n_oct;
      o = first - local * (uint32_t)a.n_oct;
      dL = G / (uint32_t)a.n_oct;
      dO = G - dL * (uint32_t)a.n_oct;
    } else {
      o = first / a.per_oct;
      local = first - o * a.per_oct;
      dL = G;
      dO = 0;
    }
    clip = local / a.tiles_per_clip;
    tile = local - clip * a.tiles_per_clip;
    dC = dL / a.tiles_per_clip;
    dT = dL - dC * a.tiles_per_clip;
  }
  __device__ __forceinline__ bool done(const UmmaArgs& a) const { return item >= a.total_items; }
  __device__ __forceinline__ void step(const UmmaArgs& a, uint32_t G) {
    item += G;
    local += dL;
    clip += dC;
    tile += dT;
    if (a.shared_bank) {
      o += dO;
      if (o >= (uint32_t)a.n_oct) {
        o -= (uint32_t)a.n_oct;
        ++local;
        ++tile;
      }
    }
    if (tile >= a.tiles_per_clip) {
      tile -= a.tiles_per_clip;
      ++clip;
    }
    if (!a.shared_bank && local >= a.per_oct && item < a.total_items) {   // crossed into the next octave(s)
      o = item / a.per_oct;
      local = item - o * a.per_oct;
      clip = local / a.tiles_per_clip;
      tile = local - clip * a.tiles_per_clip;
    }
  }
  // false: nothing to do for this item (beyond the clip's frames, or a short tail left to cqt_tail_kernel)
  __device__ __forceinline__ bool get(const UmmaArgs& a, Tile& t) const {
    t.o = (int)o;
    t.clip = (int)clip;
    t.t0 = (int)tile * UM_TILE_M;
    const int T = a.uniform_T > 0 ? a.uniform_T : a.clip_frames[clip];
    t.valid = min(T - t.t0, UM_TILE_M);
    return t.valid > a.tail_max;
  }
};

__global__ void __launch_bounds__(UM_THREADS, 1) cqt_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* b_s = smem_raw;
  uint8_t* a_base = b_s + a.b_region_bytes;                     // stages: [hi | lo] x S
  const uint32_t S = (uint32_t)a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + 2 * S * a.a_region_bytes);
  uint64_t* raw = bars;                             // [S] hi plane landed (cp.async completion)
  uint64_t* full = bars + UM_MAX_STAGES;            // [S] lo plane written, stage ready for the tensor core
  uint64_t* empty = bars + 2 * UM_MAX_STAGES;       // [S] MMAs reading the stage retired
  uint64_t* tfull = bars + 3 * UM_MAX_STAGES;       // [2]
  uint64_t* tempty = bars + 3 * UM_MAX_STAGES + 2;  // [2]
  uint64_t* bankbar = bars + 3 * UM_MAX_STAGES + 4; // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 3 * UM_MAX_STAGES + 5);

  // broadcast so the compiler knows the role index is warp-uniform (keeps MMA operands in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&raw[s], 32 * UM_LOAD_WARPS);
      mbar_init(&full[s], UM_CONV_WARPS);
      mbar_init(&empty[s], UM_MMA_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], UM_MMA_WARPS);
      mbar_init(&tempty[s], UM_EPI_WARPS);
    }
    mbar_init(bankbar, 1);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  const uint32_t G = gridDim.x;
  const bool split = (a.n_split == 3);

  if (warp < UM_MMA_WARPS) {
    // =========================== MMA issuers ===========================
    uint32_t k = 0, it_acc = 0, n_bank = 0;
    int cur_oct = -1;
    PROF_DECL;
    TileIter ti;
    for (ti.init(a, blockIdx.x, G); !ti.done(a); ti.step(a, G)) {
      Tile tl;
      if (!ti.get(a, tl)) continue;
      const UmmaOct& oc = a.oct[tl.o];
      PROF_T(0);
      const int bank_id = a.shared_bank ? 0 : tl.o;
      if (bank_id != cur_oct) {       // the loaders replace the resident bank at an octave switch
        mbar_wait_warp(bankbar, n_bank & 1, a.error_flag);
        ++n_bank;
        cur_oct = bank_id;
        PROF_T(1);
      }
      // instruction descriptors: D=f32, A=B=tf32, K-major both, M = 128
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UM_TILE_M >> 4) << 24);
      const uint32_t idesc_lo = idesc_base | ((uint32_t)(oc.n_lo >> 3) << 17);
      const uint32_t idesc_main = split ? (idesc_base | ((uint32_t)(oc.n_main >> 3) << 17)) : idesc_lo;
      const uint32_t acc = it_acc & 1;
      mbar_wait_warp(&tempty[acc], ((it_acc >> 1) & 1) ^ 1, a.error_flag);
      tc_fence_after();
      PROF_T(2);
      const uint32_t d_tmem = tmem_base + acc * a.acc_stride + (uint32_t)warp * a.grp_stride;
      const uint32_t plane16 = (uint32_t)oc.rows_pad;          // plane pitch in 16-byte units
      const uint32_t bchunk16 = (uint32_t)oc.n_main;           // K-chunk pitch of the bank in 16-byte units
      const uint64_t db0 = smem_desc(smem_u32(b_s), bchunk16 * 16u, 128);
      uint32_t accum = 0;
      int rr = warp;      // next K-slice (within the current stage) owned by this warp
      for (int st = 0; st < oc.n_stages; ++st, ++k) {
        const uint32_t s = k % S;
        // the converters observed the cp.async data (mbarrier), wrote the lo plane and executed the
        // generic->async proxy fence before their release-arrive on `full`
        mbar_wait_warp(&full[s], (k / S) & 1, a.error_flag);
        tc_fence_after();
        PROF_T(3);
        if (!(a.debug & 1)) {
          const uint32_t ah = smem_u32(a_base + (2 * s) * a.a_region_bytes);
          const uint32_t al = smem_u32(a_base + (2 * s + 1) * a.a_region_bytes);
          if (oc.planes >= 2) {
            const uint64_t dah0 = smem_desc(ah, plane16 * 16u, 128);
            const uint64_t dal0 = smem_desc(al, plane16 * 16u, 128);
            const int g0 = st * a.pps;
            const int np = min(a.pps, oc.planes - g0);
            const int np2_log = 30 - __clz(np);                     // log2(plane pairs per shift): np = 2, 4, 8 -> 0, 1, 2
            const int n_slices = oc.Q << np2_log;
            const uint32_t bq = (uint32_t)oc.planes * bchunk16;
            if (split && rr == warp && (n_slices % UM_MMA_WARPS) == 0) {
              // this warp's slices sl = warp + 4 i: shift q advances by 4 >> np2_log per step, the plane
              // pair is fixed  ->  both descriptors are arithmetic progressions
              const uint32_t q0 = (uint32_t)(warp >> np2_log);
              const uint32_t pp = (uint32_t)(warp & ((1 << np2_log) - 1)) * 2u;
              const uint32_t dq = (uint32_t)(UM_MMA_WARPS >> np2_log);   // needs pairs-per-shift <= UM_MMA_WARPS
              const uint32_t a_off = pp * plane16 + q0;
              const uint64_t db = db0 + (uint64_t)(q0 * bq + ((uint32_t)g0 + pp) * bchunk16);
              tc_mma_split_run(d_tmem, dah0 + a_off, dal0 + a_off, db, idesc_main, idesc_lo, accum, dq, dq * bq,
                               n_slices / UM_MMA_WARPS);
            } else {
              int sl = rr;
              for (; sl < n_slices; sl += UM_MMA_WARPS) {
                const uint32_t q = (uint32_t)(sl >> np2_log);
                const uint32_t pp = (uint32_t)(sl & ((1 << np2_log) - 1)) * 2u;
                const uint32_t a_off = pp * plane16 + q;
                const uint64_t db = db0 + (uint64_t)(q * bq + ((uint32_t)g0 + pp) * bchunk16);
                tc_mma_tf32(d_tmem, dah0 + a_off, db, idesc_main, accum);
                accum = 1;
                if (split) tc_mma_tf32(d_tmem, dal0 + a_off, db, idesc_lo, 1u);
              }
              rr = sl - n_slices;
            }
          } else {
            // hop == 4: one plane; the two K-chunks of an MMA are shifts q and q+1 (LBO = 16 B)
            const uint64_t dah0 = smem_desc(ah, 16, 128);
            const uint64_t dal0 = smem_desc(al, 16, 128);
            const int n_slices = oc.Q >> 1;
            if (split && rr == warp && (n_slices % UM_MMA_WARPS) == 0) {
              const uint32_t q0 = 2u * (uint32_t)warp;
              tc_mma_split_run(d_tmem, dah0 + q0, dal0 + q0, db0 + (uint64_t)(q0 * bchunk16), idesc_main, idesc_lo,
                               accum, 2u * UM_MMA_WARPS, 2u * UM_MMA_WARPS * bchunk16, n_slices / UM_MMA_WARPS);
            } else {
              int sl = rr;
              for (; sl < n_slices; sl += UM_MMA_WARPS) {
                const uint32_t q = 2u * (uint32_t)sl;
                const uint64_t db = db0 + (uint64_t)(q * bchunk16);
                tc_mma_tf32(d_tmem, dah0 + q, db, idesc_main, accum);
                accum = 1;
                if (split) tc_mma_tf32(d_tmem, dal0 + q, db, idesc_lo, 1u);
              }
              rr = sl - n_slices;
            }
          }
        }
        PROF_T(4);
        tc_commit(&empty[s]);                                   // smem stage reusable once these MMAs retire
        if (st == oc.n_stages - 1) tc_commit(&tfull[acc]);      // accumulator complete
        __syncwarp();
        PROF_T(5);
      }
      ++it_acc;
    }
    if (warp == 0 && lane == 0) PROF_OUT(0, 6);
  } else if (warp < UM_MMA_WARPS + UM_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                 // a warp may only touch TMEM lanes [32*(warp%4), +32)
    uint32_t it_acc = 0;
    PROF_DECL;
    TileIter ti;
    for (ti.init(a, blockIdx.x, G); !ti.done(a); ti.step(a, G)) {
      Tile tl;
      if (!ti.get(a, tl)) continue;
      const UmmaOct& oc = a.oct[tl.o];
      const uint32_t acc = it_acc & 1;
      mbar_wait(&tfull[acc], (it_acc >> 1) & 1, a.error_flag);
      tc_fence_after();
      PROF_T(0);
      const int f = ew * 32 + lane;                                        // frame within the tile
      const int64_t row = (int64_t)tl.clip * a.out_clip_stride + (int64_t)(tl.t0 + f) * a.frame_pitch;
      const uint32_t tbase = tmem_base + acc * a.acc_stride + ((uint32_t)(ew * 32) << 16);
      const bool vec_ok = ((a.frame_pitch | a.out_clip_stride | (int64_t)oc.first_bin) & 3) == 0 && !a.cplx_out &&
                          (reinterpret_cast<uintptr_t>(a.mag_out) & 15) == 0;
      for (int c0 = 0; c0 < oc.ncol && !(a.debug & 4); c0 += 8) {
        float sum[8];
        uint32_t v[UM_MMA_WARPS][8], u[UM_MMA_WARPS][8];
#pragma unroll
        for (int g = 0; g < UM_MMA_WARPS; ++g) {
          tmem_ld8(tbase + (uint32_t)g * a.grp_stride + (uint32_t)c0, v[g]);
          if (split) tmem_ld8(tbase + (uint32_t)g * a.grp_stride + (uint32_t)(oc.ncol + c0), u[g]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float h = 0.f, l = 0.f;
#pragma unroll
          for (int g = 0; g < UM_MMA_WARPS; ++g) {
            h += __uint_as_float(v[g][i]);
            if (split) l += __uint_as_float(u[g][i]);
          }
          sum[i] = h + l;      // the correction columns are ~2^-11 of the main ones: add them last
        }
        if (f < tl.valid) {
          const int bin0 = oc.first_bin + (c0 >> 1);
          if (a.shared_bank) {
            const float4* sc = reinterpret_cast<const float4*>(a.col_scale + tl.o * UM_MAX_COLS + c0);
            const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1);
            sum[0] *= s0.x; sum[1] *= s0.y; sum[2] *= s0.z; sum[3] *= s0.w;
            sum[4] *= s1.x; sum[5] *= s1.y; sum[6] *= s1.z; sum[7] *= s1.w;
          }
          float m[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) m[i] = sqrtf(sum[2 * i] * sum[2 * i] + sum[2 * i + 1] * sum[2 * i + 1]);
          if (vec_ok && bin0 >= 0 && bin0 + 4 <= a.n_bins) {
            *reinterpret_cast<float4*>(a.mag_out + row + bin0) = make_float4(m[0], m[1], m[2], m[3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int bin = bin0 + i;
              if (bin >= 0 && bin < a.n_bins) {
                a.mag_out[row + bin] = m[i];
                if (a.cplx_out) a.cplx_out[row + bin] = make_float2(sum[2 * i], sum[2 * i + 1]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      ++it_acc;
      PROF_T(1);
    }
    if (ew == 0 && lane == 0) PROF_OUT(6, 2);
  } else if (warp < UM_MMA_WARPS + UM_EPI_WARPS + UM_LOAD_WARPS) {
    // =========================== loaders ===========================
    const int ltid = threadIdx.x - 32 * (UM_MMA_WARPS + UM_EPI_WARPS);
    constexpr int LT = 32 * UM_LOAD_WARPS;
    uint32_t k = 0;
    int cur_oct = -1;
    PROF_DECL;
    TileIter ti;
    for (ti.init(a, blockIdx.x, G); !ti.done(a); ti.step(a, G)) {
      Tile tl;
      if (!ti.get(a, tl)) continue;
      const UmmaOct& oc = a.oct[tl.o];
      PROF_T(0);
      const int bank_id = a.shared_bank ? 0 : tl.o;
      if (bank_id != cur_oct) {
        const UmmaOct& ob = a.oct[bank_id];
        if (ltid == 0) {
          // new bank: every MMA that reads the old one must have retired (= all stages in flight consumed)
          const uint32_t back = min(k, S);
          for (uint32_t j = 1; j <= back; ++j) mbar_wait(&empty[(k - j) % S], ((k - j) / S) & 1, a.error_flag);
          mbar_arrive_expect_tx(bankbar, ob.bank_bytes);
          const uint32_t dst = smem_u32(b_s);
          for (uint32_t off = 0; off < ob.bank_bytes; off += 32768u)
            bulk_g2s(dst + off, reinterpret_cast<const uint8_t*>(ob.b_pack) + off, min(32768u, ob.bank_bytes - off),
                     bankbar);
        }
        cur_oct = bank_id;
        PROF_T(1);
      }
      // rows that valid frames read: frame r uses rows r .. r+Q-1; the rest of a partial tile stays stale
      // (MMA rows are independent and frames >= T are never stored)
      const int rows_used = min(oc.rows, tl.valid + oc.Q - 1);
      const float* src_tile = oc.sig + (int64_t)tl.clip * oc.sig_stride + (int64_t)tl.t0 * oc.hop;
      {
        // pull the NEXT tile of this CTA from HBM into L2 now: its cp.asyncs are issued a few stages from
        // now and then pay L2 latency instead of DRAM latency (the stage ring is too short to hide the latter)
        TileIter nx = ti;
        nx.step(a, G);
        Tile tn;
        if (!nx.done(a) && nx.get(a, tn)) {
          const UmmaOct& on = a.oct[tn.o];
          const char* p0 = reinterpret_cast<const char*>(on.sig + (int64_t)tn.clip * on.sig_stride + (int64_t)tn.t0 * on.hop);
          const int nbytes = min(on.rows, tn.valid + on.Q - 1) * on.hop * 4;
          for (int off = ltid * 128; off < nbytes; off += LT * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + off));
        }
      }
      for (int st = 0; st < oc.n_stages; ++st, ++k) {
        const uint32_t s = k % S;
        mbar_wait(&empty[s], ((k / S) & 1) ^ 1, a.error_flag);
        PROF_T(2);
        const int g0 = st * a.pps;
        const int np = min(a.pps, oc.planes - g0);
        const int lg = 31 - __clz(np);
        const int total = (a.debug & 2) ? 0 : np * rows_used;
        const uint32_t dst0 = smem_u32(a_base + (2 * s) * a.a_region_bytes);
        const float* src0 = src_tile + 4 * g0;
        for (int e = ltid; e < total; e += LT) {
          const int g = e & (np - 1), srow = e >> lg;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16u * (uint32_t)(g * oc.rows_pad + srow)),
                       "l"(src0 + (srow * oc.hop + 4 * g))
                       : "memory");
        }
        cp_async_arrive_noinc(&raw[s]);
        PROF_T(3);
      }
    }
    if (ltid == 0) PROF_OUT(8, 4);
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    // =========================== converters ===========================
    const int ctid = threadIdx.x - 32 * (UM_MMA_WARPS + UM_EPI_WARPS + UM_LOAD_WARPS);
    constexpr int CT = 32 * UM_CONV_WARPS;
    uint32_t k = 0;
    PROF_DECL;
    TileIter ti;
    for (ti.init(a, blockIdx.x, G); !ti.done(a); ti.step(a, G)) {
      Tile tl;
      if (!ti.get(a, tl)) continue;
      const UmmaOct& oc = a.oct[tl.o];
      const int rows_used = min(oc.rows, tl.valid + oc.Q - 1);
      PROF_T(0);
      for (int st = 0; st < oc.n_stages; ++st, ++k) {
        const uint32_t s = k % S;
        mbar_wait(&raw[s], (k / S) & 1, a.error_flag);
        PROF_T(1);
        float4* dh = reinterpret_cast<float4*>(a_base + (2 * s) * a.a_region_bytes);
        float4* dl = reinterpret_cast<float4*>(a_base + (2 * s + 1) * a.a_region_bytes);
        const int g0 = st * a.pps;
        const int np = min(a.pps, oc.planes - g0);
        const int lg = 31 - __clz(np);
        const int total = (a.debug & 2) ? 0 : np * rows_used;
        // a stage holds at most 8 planes x ~129 rows = 5 elements per thread: all loads first, then the
        // splits, then the stores (one shared-memory round trip instead of five in a row)
        constexpr int CV = 5;
        float4 x[CV];
        int d[CV];
        for (int e0 = ctid; e0 < total; e0 += CV * CT) {
#pragma unroll
          for (int i = 0; i < CV; ++i) {
            const int e = e0 + i * CT;
            d[i] = (e & (np - 1)) * oc.rows_pad + (e >> lg);
            if (e < total) x[i] = dh[d[i]];
          }
#pragma unroll
          for (int i = 0; i < CV; ++i) {
            if (e0 + i * CT < total) {
              if (split) {
                // kind::tf32 reads only the top 19 bits of each operand, so the raw fp32 samples already ARE
                // the hi operand (hi = trunc13(x)); lo = x - hi is exact in fp32 and is itself truncated by
                // the hardware to its top 11 mantissa bits: |error| <= 2^-21 |x|, the size of the dropped
                // lo*lo term -- no conversion instruction needed
                float4 l;
                l.x = x[i].x - __uint_as_float(__float_as_uint(x[i].x) & 0xFFFFE000u);
                l.y = x[i].y - __uint_as_float(__float_as_uint(x[i].y) & 0xFFFFE000u);
                l.z = x[i].z - __uint_as_float(__float_as_uint(x[i].z) & 0xFFFFE000u);
                l.w = x[i].w - __uint_as_float(__float_as_uint(x[i].w) & 0xFFFFE000u);
                dl[d[i]] = l;
              } else {
                // single pass: round to nearest instead of the hardware's truncation
                float4 h;
                h.x = to_tf32(x[i].x); h.y = to_tf32(x[i].y); h.z = to_tf32(x[i].z); h.w = to_tf32(x[i].w);
                dh[d[i]] = h;
              }
            }
          }
        }
        fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        PROF_T(2);
      }
    }
    if (ctid == 0) PROF_OUT(12, 3);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
  }
}

// A clip's last tile with <= tail_max valid frames would cost a full 128-row MMA sequence, so the
// tensor-core kernel skips it and this kernel contracts it on the CUDA cores: ONE WARP per (octave, clip),
// lane = output column, fp32 bank rows coalesced, signal values as 16-byte broadcast loads from the padded
// level buffers.  No shared memory and <= 80 registers: launched on the plan's side stream it co-resides
// with the persistent tensor-core kernel (which leaves ~10 k registers per SM unused) and hides under it.
template <int NF>
__device__ __forceinline__ void tail_item(const UmmaArgs& a, const UmmaOct& oc, int o, uint32_t clip, int rem, int t0,
                                          int lane) {
  const float* y = oc.sig + (int64_t)clip * oc.sig_stride + (int64_t)t0 * oc.hop;
  // frame f reads y + f*hop + n; frames >= rem re-read the last valid one (branch-free loads, results unused)
  int foff[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) foff[f] = min(f, rem - 1) * oc.hop;
  for (int c0 = 0; c0 < oc.ncol; c0 += 32) {
    const int col = c0 + lane;
    const bool live = col < oc.ncol;
    // shared-bank plans: every octave reads octave 0's bank (one 48 KB table for all warps of the SM) and
    // applies its per-column scale at the end, like the tensor-core epilogue
    const float* bk = (a.shared_bank ? a.oct[0].bank_f32 : oc.bank_f32) + (live ? col : 0);
    const float cs = (a.shared_bank && live) ? __ldg(a.col_scale + o * UM_MAX_COLS + col) : 1.f;
    float acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.f;
#pragma unroll 1
    for (int n = 0; n < oc.n_fft; n += 4) {
      const float g0 = __ldg(bk + (int64_t)(n + 0) * oc.ncol), g1 = __ldg(bk + (int64_t)(n + 1) * oc.ncol);
      const float g2 = __ldg(bk + (int64_t)(n + 2) * oc.ncol), g3 = __ldg(bk + (int64_t)(n + 3) * oc.ncol);
      float4 v[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) v[f] = __ldg(reinterpret_cast<const float4*>(y + foff[f] + n));
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[f] = fmaf(g3, v[f].w, fmaf(g2, v[f].z, fmaf(g1, v[f].y, fmaf(g0, v[f].x, acc[f]))));
    }
    // lanes (2j, 2j+1) hold re / im of filter j
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      acc[f] *= cs;
      const float im = __shfl_down_sync(0xffffffffu, acc[f], 1);
      const int bin = oc.first_bin + (col >> 1);
      if (f < rem && live && !(lane & 1) && bin >= 0 && bin < a.n_bins) {
        const int64_t row = (int64_t)clip * a.out_clip_stride + (int64_t)(t0 + f) * a.frame_pitch;
        a.mag_out[row + bin] = sqrtf(acc[f] * acc[f] + im * im);
        if (a.cplx_out) a.cplx_out[row + bin] = make_float2(acc[f], im);
      }
    }
  }
}

__global__ void __launch_bounds__(32 * UM_TAIL_WARPS, 6) cqt_tail_kernel(const __grid_constant__ UmmaArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t idx = blockIdx.x * UM_TAIL_WARPS + (threadIdx.x >> 5);
  if (idx >= (uint32_t)a.n_oct * (uint32_t)a.n_clips) return;
  const uint32_t clip = idx / (uint32_t)a.n_oct;
  const int o = (int)(idx - clip * (uint32_t)a.n_oct);
  const UmmaOct& oc = a.oct[o];
  const int T = a.uniform_T > 0 ? a.uniform_T : a.clip_frames[clip];
  const int rem = T % UM_TILE_M;
  if (rem == 0 || rem > a.tail_max) return;
  const int t0 = T - rem;
  if (rem <= 4) {
    tail_item<4>(a, oc, o, clip, rem, t0, lane);
  } else {
    for (int f0 = 0; f0 < rem; f0 += 8) tail_item<8>(a, oc, o, clip, min(rem - f0, 8), t0 + f0, lane);   // 8 frames per sweep
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
  float* d_pack = nullptr;
  int n_main = 0, n_lo = 0;
  uint32_t bytes = 0;
};

}  // namespace saga

struct CqtUmmaState {
  std::vector<saga::OctPack> packs;
  bool supported = false;
  uint32_t b_region_bytes = 0, a_region_bytes = 0, tmem_cols = 0, acc_stride = 0, grp_stride = 0;
  size_t smem_bytes = 0;
  int* d_error = nullptr;
  long long* d_prof = nullptr;
  float* d_col_scale = nullptr;
  // tail kernel runs on a side stream, concurrently with the persistent kernel: ONE side stream and event pair per
  // CALLER stream (created on first use, reused afterwards), so two streams driving the same cached plan neither
  // serialise on a plan-owned stream nor pay two cudaEventCreate per call
  struct Fork { cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; };
  mutable std::map<cudaStream_t, Fork> forks;
  mutable std::mutex forks_mu;
  bool shared_bank = false;
  int num_sms = 0;
  int stages = 0, pps = 8;   // pipeline shape (SAGA_UMMA_CFG="stages,planes" overrides; stages 0 = as many as fit)
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
  if (const char* cfg = SAGA_OPT("SAGA_UMMA_CFG")) {
    int a_ = 0, b_ = 0;
    if (sscanf(cfg, "%d,%d", &a_, &b_) == 2 && a_ >= 0 && a_ <= UM_MAX_STAGES && a_ != 1 &&
        (b_ == 2 || b_ == 4 || b_ == 8)) {
      st->stages = a_; st->pps = b_;
    }
  }
  uint32_t bmax = 0, amax = 0;
  int nmain_max = 0;
  for (auto& o : p->oct) {
    const int ncol = 2 * o.n_filters;
    if (ncol % 8) return;                           // the epilogue reads 8 columns (4 filters) at a time
    const int n_main = (2 * ncol + 15) & ~15;       // [B_hi | B_lo]; M = 128 needs N % 16 == 0
    if (n_main > 256 || ncol > UM_MAX_COLS) return;
    if (o.hop < 4 || (o.hop % 4) != 0 || (o.n_fft % o.hop) != 0 || (o.n_fft % 8) != 0) return;
    const int planes = o.hop / 4;
    if (planes >= 2 && (std::min(planes, st->pps) % 2) != 0) return;
    const int Q = o.n_fft / o.hop;
    if (planes == 1 && (Q % 2) != 0) return;
    const int np0 = std::min(planes, st->pps);
    if ((np0 / 2) > UM_MMA_WARPS) return;                         // arithmetic-progression issue path
    const int slices = planes >= 2 ? Q * (np0 / 2) : Q / 2;      // K-slices per stage
    const int n_st = (planes + st->pps - 1) / st->pps;
    if (slices * n_st < UM_MMA_WARPS) return;                     // every issuer warp must get work in an item
    const int rows = UM_TILE_M + Q - 1;
    const int rows_pad = rows | 1;
    bmax = std::max<uint32_t>(bmax, (uint32_t)o.n_fft * n_main * 4u);
    amax = std::max<uint32_t>(amax, (uint32_t)np0 * rows_pad * 16u);
    nmain_max = std::max(nmain_max, n_main);
  }
  amax = (amax + 127u) & ~127u;
  bmax = (bmax + 127u) & ~127u;
  const uint32_t fixed = bmax + 512;              // bank + barriers
  if (fixed + 2u * 2u * amax > UM_SMEM_LIMIT) return;
  int fit = (int)((UM_SMEM_LIMIT - fixed) / (2u * amax));
  fit = std::min(fit, UM_MAX_STAGES);
  if (st->stages == 0 || st->stages > fit) st->stages = fit;
  const size_t smem = (size_t)bmax + 2ull * st->stages * amax + 512;
  const uint32_t grp = ((uint32_t)nmain_max + 31u) & ~31u;
  uint32_t cols = 32;
  while (cols < 2u * UM_MMA_WARPS * grp) cols <<= 1;   // 2 buffers x issuer groups
  if (cols > 512) return;
  st->b_region_bytes = bmax;
  st->a_region_bytes = amax;
  st->tmem_cols = cols;
  st->acc_stride = cols / 2;
  st->grp_stride = grp;
  st->smem_bytes = smem;
  // librosa analyses every octave with the SAME sparsified basis (times sqrt(2)^level and the per-bin length
  // normalisation), so each bank is normally a per-column multiple of the first one: detect that and keep
  // one bank resident for the whole kernel
  {
    const CqtOctaveDev& o0 = p->oct[0];
    const int ncol0 = 2 * o0.n_filters;
    std::vector<float> scale((size_t)p->oct.size() * UM_MAX_COLS, 0.f);
    bool shared = p->oct.size() > 1 && SAGA_OPT("SAGA_UMMA_NO_SHARED_BANK") == nullptr;
    for (size_t i = 0; shared && i < p->oct.size(); ++i) {
      const CqtOctaveDev& o = p->oct[i];
      if (o.n_fft != o0.n_fft || o.n_filters != o0.n_filters) { shared = false; break; }
      for (int c = 0; c < ncol0 && shared; ++c) {
        double num = 0, den = 0, peak = 0;
        for (int k = 0; k < o.n_fft; ++k) {
          const double b0 = o0.bank_host[(size_t)k * ncol0 + c], b = o.bank_host[(size_t)k * ncol0 + c];
          num += b * b0; den += b0 * b0; peak = std::max(peak, std::fabs(b));
        }
        const double sc = den > 0 ? num / den : 0.0;
        double dev = 0;
        for (int k = 0; k < o.n_fft; ++k)
          dev = std::max(dev, std::fabs(o.bank_host[(size_t)k * ncol0 + c] - sc * o0.bank_host[(size_t)k * ncol0 + c]));
        if (dev > 4e-7 * peak) shared = false;     // fp32 rounding of the host-built banks is ~6e-8 relative
        scale[i * UM_MAX_COLS + c] = (float)sc;
      }
    }
    st->shared_bank = shared;
    if (shared) {
      if (cudaMalloc(&st->d_col_scale, scale.size() * 4) != cudaSuccess) return;
      cudaMemcpy(st->d_col_scale, scale.data(), scale.size() * 4, cudaMemcpyHostToDevice);
    }
  }
  // pack each bank for the B descriptor: [n_fft/4][n_main][4]; rows [0,ncol) TF32 hi, [ncol,2ncol) lo
  for (auto& o : p->oct) {
    OctPack pk;
    const int ncol = 2 * o.n_filters;
    pk.n_main = (2 * ncol + 15) & ~15;
    pk.n_lo = (ncol + 15) & ~15;
    const size_t n = (size_t)o.n_fft * pk.n_main;
    pk.bytes = (uint32_t)(n * 4);
    std::vector<float> pack(n, 0.f);
    for (int k = 0; k < o.n_fft; ++k)
      for (int c = 0; c < ncol; ++c) {
        const float b = o.bank_host[(size_t)k * ncol + c];
        const float h = tf32_rna_host(b);
        const size_t chunk = (size_t)(k / 4) * pk.n_main;
        pack[(chunk + c) * 4 + (k % 4)] = h;
        pack[(chunk + ncol + c) * 4 + (k % 4)] = tf32_rna_host(b - h);
      }
    if (cudaMalloc(&pk.d_pack, n * 4) != cudaSuccess) return;
    cudaMemcpy(pk.d_pack, pack.data(), n * 4, cudaMemcpyHostToDevice);
    st->packs.push_back(pk);
  }
  if (cudaMalloc(&st->d_error, sizeof(int)) != cudaSuccess) return;
  cudaMemset(st->d_error, 0, sizeof(int));
  if (cudaMalloc(&st->d_prof, sizeof(long long) * 256 * UM_PROF_SLOTS) != cudaSuccess) return;
  cudaMemset(st->d_prof, 0, sizeof(long long) * 256 * UM_PROF_SLOTS);
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
  for (auto& pk : p->umma->packs) cudaFree(pk.d_pack);
  cudaFree(p->umma->d_error);
  cudaFree(p->umma->d_prof);
  cudaFree(p->umma->d_col_scale);
  for (auto& kv : p->umma->forks) {
    if (kv.second.ev_fork) cudaEventDestroy(kv.second.ev_fork);
    if (kv.second.ev_join) cudaEventDestroy(kv.second.ev_join);
    if (kv.second.side) cudaStreamDestroy(kv.second.side);
  }
  delete p->umma;
  p->umma = nullptr;
}

bool cqt_umma_supported(const saga_cqt_plan* p) { return p->umma && p->umma->supported; }

int cqt_umma_exec(const saga_cqt_plan* p, const CqtLevels& lv, int n_clips, int64_t max_len, int64_t T_max,
                  float* mag_out, float2* cplx_out, int64_t frame_pitch, int64_t out_clip_stride,
                  int n_split, int tail_max, cudaStream_t stream) {
  const CqtUmmaState* st = p->umma;
  if (!st || !st->supported) return set_error(SAGA_ERR_UNSUPPORTED, "cqt: plan does not fit the tcgen05 path");
  UmmaArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n_oct = (int)p->oct.size();
  a.n_clips = n_clips;
  const int64_t tpc = (T_max + UM_TILE_M - 1) / UM_TILE_M;
  const int64_t per_oct = (int64_t)n_clips * tpc;
  if (per_oct * a.n_oct >= (int64_t)1 << 31) return set_error(SAGA_ERR_UNSUPPORTED, "cqt: batch too large for one launch");
  a.tiles_per_clip = (uint32_t)tpc;
  a.per_oct = (uint32_t)per_oct;
  a.total_items = (uint32_t)(per_oct * a.n_oct);
  a.n_bins = p->n_bins;
  a.n_split = (n_split == 1) ? 1 : 3;
  a.clip_frames = lv.clip_frames;
  a.uniform_T = lv.clip_lens ? 0 : (int)T_max;
  a.tail_max = std::min(tail_max, UM_TAIL_MAX);
  a.mag_out = mag_out;
  a.cplx_out = cplx_out;
  a.frame_pitch = frame_pitch;
  a.out_clip_stride = out_clip_stride;
  a.b_region_bytes = st->b_region_bytes;
  a.a_region_bytes = st->a_region_bytes;
  a.tmem_cols = st->tmem_cols;
  a.acc_stride = st->acc_stride;
  a.grp_stride = st->grp_stride;
  a.error_flag = st->d_error;
  a.prof = st->d_prof;
  a.stages = st->stages;
  a.pps = st->pps;
  a.shared_bank = st->shared_bank ? 1 : 0;
  a.col_scale = st->d_col_scale;
  {
    const char* dbg = SAGA_OPT("SAGA_UMMA_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
  }
  for (int i = 0; i < a.n_oct; ++i) {
    const CqtOctaveDev& o = p->oct[i];
    UmmaOct& u = a.oct[i];
    // padded level buffer: sample s of clip c lives at lvl + c*pitch + pad + s
    u.sig = lv.lvl[o.level] + (lv.pad[o.level] - o.n_fft / 2);
    u.sig_stride = lv.pitch[o.level];
    u.b_pack = st->packs[i].d_pack;
    u.bank_f32 = o.bank;
    u.bank_bytes = st->packs[i].bytes;
    u.hop = o.hop;
    u.n_fft = o.n_fft;
    u.ncol = 2 * o.n_filters;
    u.n_main = st->packs[i].n_main;
    u.n_lo = st->packs[i].n_lo;
    u.first_bin = o.first_bin;
    u.planes = o.hop / 4;
    u.n_stages = (u.planes + st->pps - 1) / st->pps;
    u.Q = o.n_fft / o.hop;
    u.rows = UM_TILE_M + u.Q - 1;
    u.rows_pad = u.rows | 1;
  }
  if (a.total_items == 0) return SAGA_OK;
  if (frame_pitch > p->n_bins) {
    dim3 grid(8, n_clips), block(32, 8);
    cqt_zero_pad_kernel<<<grid, block, 0, stream>>>(mag_out, cplx_out, lv.clip_frames, p->n_bins, frame_pitch,
                                                    out_clip_stride);
    SAGA_LAUNCH_CHECK();
  }
  const int grid = (int)std::min<int64_t>(a.total_items, st->num_sms > 0 ? st->num_sms : 148);
  // short partial tiles: cqt_tail_kernel runs on the side stream, forked before and joined after the
  // persistent kernel.  The persistent kernel is ENQUEUED FIRST: its 148 CTAs each need a whole SM's shared
  // memory and 55 k registers, so tail CTAs that got there first (six fit per SM) would hold it back until
  // they retire; the other way round the tail kernel fills the ~10 k registers per SM that are left over.
  // Events are per call (a plan may be driven from several streams).
  const int rem = (int)(T_max % UM_TILE_M);
  const bool tails = a.tail_max > 0 && (lv.clip_lens || (rem > 0 && rem <= a.tail_max));
  const unsigned n_items = (unsigned)a.n_oct * (unsigned)n_clips;
  const unsigned tgrid = (n_items + UM_TAIL_WARPS - 1) / UM_TAIL_WARPS;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t side = nullptr;
  if (tails) {
    std::lock_guard<std::mutex> lock(st->forks_mu);
    auto it = st->forks.find(stream);
    if (it == st->forks.end() && st->forks.size() < 64) {
      CqtUmmaState::Fork f;
      if (cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking) == cudaSuccess &&
          cudaEventCreateWithFlags(&f.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
          cudaEventCreateWithFlags(&f.ev_join, cudaEventDisableTiming) == cudaSuccess) {
        it = st->forks.emplace(stream, f).first;
      } else {
        cudaGetLastError();
        if (f.ev_fork) cudaEventDestroy(f.ev_fork);
        if (f.side) cudaStreamDestroy(f.side);
      }
    }
    if (it != st->forks.end()) { side = it->second.side; ev_fork = it->second.ev_fork; ev_join = it->second.ev_join; }
  }
  const bool forked = tails && side != nullptr;
  if (forked) {
    cudaEventRecord(ev_fork, stream);
    cudaStreamWaitEvent(side, ev_fork, 0);
  }
  cqt_umma_kernel<<<grid, UM_THREADS, st->smem_bytes, stream>>>(a);
  SAGA_LAUNCH_CHECK();
  if (tails) {
    cqt_tail_kernel<<<tgrid, 32 * UM_TAIL_WARPS, 0, forked ? side : stream>>>(a);
    SAGA_LAUNCH_CHECK();
  }
  if (forked) {
    cudaEventRecord(ev_join, side);
    cudaStreamWaitEvent(stream, ev_join, 0);
  }
  if (a.debug & 16) {
    // profiling aid only: synchronous read-back of the per-role cycle counters, mean over CTAs
    std::vector<long long> h((size_t)grid * UM_PROF_SLOTS);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), st->d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    static const char* names[15] = {"mma.decode", "mma.wait_bank", "mma.wait_tempty", "mma.wait_full", "mma.issue",
                                    "mma.commit", "epi.wait_tfull", "epi.work", "load.decode", "load.bank",
                                    "load.wait_empty", "load.issue", "conv.decode", "conv.wait_raw", "conv.convert"};
    fprintf(stderr, "umma_prof stages=%d planes/stage=%d smem=%zu shared_bank=%d\n", st->stages, st->pps, st->smem_bytes,
            (int)st->shared_bank);
    for (int s = 0; s < 15; ++s) {
      double sum = 0, mx = 0;
      for (int c = 0; c < grid; ++c) {
        sum += (double)h[(size_t)c * UM_PROF_SLOTS + s];
        mx = std::max(mx, (double)h[(size_t)c * UM_PROF_SLOTS + s]);
      }
      fprintf(stderr, "umma_prof %-18s %10.0f cycles/CTA (max %.0f)\n", names[s], sum / grid, mx);
    }
    {
      double mx = 0, mn = 1e30;
      for (int c = 0; c < grid; ++c) {
        double t = 0;
        for (int s = 0; s < 6; ++s) t += (double)h[(size_t)c * UM_PROF_SLOTS + s];
        mx = std::max(mx, t);
        mn = std::min(mn, t);
      }
      fprintf(stderr, "umma_prof mma role total per CTA: min %.0f max %.0f cycles\n", mn, mx);
    }
  }
  return SAGA_OK;
}

}  // namespace saga
