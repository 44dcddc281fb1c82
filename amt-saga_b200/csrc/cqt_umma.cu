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
// TF32, round-to-nearest) and hi*hi + hi*lo + lo*hi is accumulated in fp32
// TMEM (n_split = 3).  The bank is packed [B_hi | B_lo] along N, so per K-slice
// ONE N=2*npad MMA (A_hi x [B_hi|B_lo]) plus one N=npad MMA (A_lo x B_hi) do the
// work of three; the epilogue adds the two column groups.  n_split = 1 keeps
// only hi*hi.  N is tiny here (24 real columns per octave), so the kernel is
// bound by how fast ONE thread can issue MMAs: descriptors are built once per
// stage and advanced by adding to their low word.
//
// Warp roles per persistent CTA (one per SM): 16 producer warps (L2 -> cp.async -> planes; lo = x - trunc13(x)), 4 MMA issuer warps (one elected thread each; K-slices dealt round
// robin, each warp accumulating into its own TMEM column group, because with
// N <= 64 an MMA retires in 16-32 cycles and a single issuing thread cannot keep
// up -- profiles/microbench/umma_latency.cu), 4 epilogue warps (tcgen05.ld, sum of
// the column groups, |re + i im| -> global).  mbarrier pipelines: smem full/empty per
// A stage, TMEM full/empty per accumulator buffer.  The bank G_o (packed for the
// B descriptor on the host) stays resident in SMEM while the CTA works through
// items of one octave (items are ordered octave-major).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cqt_plan.cuh"
#include "saga_common.cuh"

namespace saga {

constexpr int UM_TILE_M = 128;
constexpr int UM_PRODUCER_WARPS = 16;
constexpr int UM_EPI_WARPS = 4;
constexpr int UM_MMA_WARPS = 4;             // issuer warps: K-slices round-robin, one TMEM column group each
constexpr int UM_THREADS = 32 * (UM_MMA_WARPS + UM_EPI_WARPS + UM_PRODUCER_WARPS);   // 0-3 MMA, 4-7 epilogue, 8-15 producers
constexpr int UM_MAX_STAGES = 6;            // barrier array capacity; the actual count is a plan parameter
constexpr int UM_MAX_OCT = 12;
constexpr uint32_t UM_SPIN_LIMIT = 1u << 27;
constexpr int UM_PROF_SLOTS = 16;

struct UmmaOct {
  const float* sig;
  const int64_t* sig_offsets;
  int64_t sig_stride;
  const float* b_pack;    // packed [n_fft/4][2*npad][4]: rows [0,npad) = TF32 hi, [npad,2npad) = lo
  int level, hop, n_fft, ncol, npad, first_bin;
  int planes, np_log2, n_stages, Q, rows, rows_pad;
  int64_t item_begin;
  int64_t uniform_len;    // length of every clip at this octave's level (equal-length batches)
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
  uint32_t b_region_bytes;   // resident bank (hi and lo rows interleaved per K-chunk)
  uint32_t a_region_bytes;   // one of hi / lo, per stage
  uint32_t tmem_cols;        // allocation (power of two >= 2*npad_max)
  uint32_t acc_stride;       // columns between the two accumulator buffers
  uint32_t grp_stride;       // columns between issuer-warp column groups (2*npad_max)
  int* error_flag;
  int stages, pps, prefetch; // pipeline shape: SMEM stages, planes per stage (4 or 8), producer look-ahead (< stages)
  int uniform_T;             // > 0: every clip has this many frames and `uniform_len` samples (no per-item loads)
  int64_t uniform_len;
  int debug;                 // profiling bisect (SAGA_UMMA_DEBUG): 1 = no MMAs, 2 = no producer data, 4 = no epilogue work,
                             // 8 = no proxy fence, 16 = per-role cycle breakdown into `prof`
  long long* prof;           // [grid][UM_PROF_SLOTS] cycle counters (debug & 16)
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
// MMA-warp variant: returns with the warp converged, so what follows is uniform code
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int* error_flag) {
  mbar_wait(bar, parity, error_flag);
  __syncwarp();
}
// cycle breakdown (debug & 16): PROF_T(slot) adds the cycles since the previous mark to pr[slot]
#define PROF_DECL const bool prof_on = (a.debug & 16) != 0; long long pr[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt = prof_on ? clock64() : 0; const long long pt0 = pt
#define PROF_T(slot) do { if (prof_on) { const long long n_ = clock64(); pr[slot] += n_ - pt; pt = n_; } } while (0)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {   // whole warp calls, one lane issues
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xFFFFFFFF;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// Issued by ONE elected lane, but called by the whole (converged) MMA warp with warp-uniform
// operands: that keeps descriptors in uniform registers (UIADD3 + UTCHMMA per MMA) instead of a
// per-instruction R2UR waterfall -- with N = 32..64 the tensor pipe needs a new MMA every 16-32 cycles.
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xFFFFFFFF;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- multi-slice issue blocks -------------------------------------------------------------
// STEPS consecutive K-slices of one issuer warp in ONE asm block: per slice the main MMA
// (A_hi x [B_hi|B_lo], N = 2*npad) and the correction MMA (A_lo x B_hi, N = npad); then both A
// descriptors advance by `da` and the B descriptor by `db` (16-byte units, added to the low word).
// Keeping the descriptor arithmetic inside the block lets ptxas carry it in uniform registers
// instead of re-broadcasting seven vector registers in front of every MMA.
#define UM_ASM_HEAD                                   \
  "{\n\t"                                             \
  ".reg .pred e, p, t;\n\t"                           \
  ".reg .b64 ah, al, bb, dda, ddb;\n\t"               \
  "elect.sync _|e, 0xFFFFFFFF;\n\t"                   \
  "setp.ne.b32 p, %6, 0;\n\t"                         \
  "setp.eq.b32 t, 0, 0;\n\t"                          \
  "mov.b64 ah, %1;\n\t"                               \
  "mov.b64 al, %2;\n\t"                               \
  "mov.b64 bb, %3;\n\t"                               \
  "cvt.u64.u32 dda, %7;\n\t"                          \
  "cvt.u64.u32 ddb, %8;\n\t"                          \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bb, %4, p;\n\t" \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bb, %5, t;\n\t"
#define UM_ASM_NEXT                                   \
  "add.u64 ah, ah, dda;\n\t"                          \
  "add.u64 al, al, dda;\n\t"                          \
  "add.u64 bb, bb, ddb;\n\t"                          \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bb, %4, t;\n\t" \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bb, %5, t;\n\t"
#define UM_ASM_TAIL "}\n"
#define UM_ASM_ARGS                                                                                        \
  ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b), "r"(idesc_main), "r"(idesc_lo), "r"(accumulate), "r"(da), \
      "r"(db)                                                                                             \
      : "memory"

__device__ __forceinline__ void tc_mma_split_x1(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b,
                                                uint32_t idesc_main, uint32_t idesc_lo, uint32_t accumulate,
                                                uint32_t da, uint32_t db) {
  asm volatile(UM_ASM_HEAD UM_ASM_TAIL UM_ASM_ARGS);
}
__device__ __forceinline__ void tc_mma_split_x2(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b,
                                                uint32_t idesc_main, uint32_t idesc_lo, uint32_t accumulate,
                                                uint32_t da, uint32_t db) {
  asm volatile(UM_ASM_HEAD UM_ASM_NEXT UM_ASM_TAIL UM_ASM_ARGS);
}
__device__ __forceinline__ void tc_mma_split_x4(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b,
                                                uint32_t idesc_main, uint32_t idesc_lo, uint32_t accumulate,
                                                uint32_t da, uint32_t db) {
  asm volatile(UM_ASM_HEAD UM_ASM_NEXT UM_ASM_NEXT UM_ASM_NEXT UM_ASM_TAIL UM_ASM_ARGS);
}
// `steps` consecutive slices starting at (a_hi, a_lo, b)
__device__ __forceinline__ void tc_mma_split_run(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b,
                                                 uint32_t idesc_main, uint32_t idesc_lo, uint32_t& accumulate,
                                                 uint32_t da, uint32_t db, int steps) {
  while (steps >= 4) {
    tc_mma_split_x4(d_tmem, a_hi, a_lo, b, idesc_main, idesc_lo, accumulate, da, db);
    accumulate = 1;
    a_hi += 4ull * da; a_lo += 4ull * da; b += 4ull * db;
    steps -= 4;
  }
  if (steps >= 2) {
    tc_mma_split_x2(d_tmem, a_hi, a_lo, b, idesc_main, idesc_lo, accumulate, da, db);
    accumulate = 1;
    a_hi += 2ull * da; a_lo += 2ull * da; b += 2ull * db;
    steps -= 2;
  }
  if (steps >= 1) {
    tc_mma_split_x1(d_tmem, a_hi, a_lo, b, idesc_main, idesc_lo, accumulate, da, db);
    accumulate = 1;
  }
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
  const uint32_t local = (uint32_t)(item - a.oct[o].item_begin);      // < n_clips * tiles_per_clip < 2^31
  it.o = o;
  it.clip = (int)(local / (uint32_t)a.tiles_per_clip);
  it.t0 = (int)(local % (uint32_t)a.tiles_per_clip) * UM_TILE_M;
  if (a.uniform_T > 0) {          // equal-length batch: nothing to fetch or derive
    it.T = a.uniform_T;
    it.len = a.oct[o].uniform_len;
    return it;
  }
  it.T = a.clip_frames[it.clip];
  int64_t len = a.clip_lens[it.clip];
  if (a.early_factor > 1) len = (len + a.early_factor - 1) / a.early_factor;
  for (int s = 0; s < a.oct[o].level; ++s) len = (len + 1) >> 1;
  it.len = len;
  return it;
}

// walks this CTA's (item, stage) jobs in the order every role processes them
struct JobIter {
  int64_t item, G;
  int st;
  bool started;
  ItemInfo inf;
  __device__ __forceinline__ void reset(int64_t first, int64_t stride) {
    item = first - stride;
    G = stride;
    st = 0;
    started = false;
  }
  __device__ __forceinline__ bool next(const UmmaArgs& a) {
    if (started && st + 1 < a.oct[inf.o].n_stages) {
      ++st;
      return true;
    }
    for (;;) {
      item += G;
      if (item >= a.total_items) return false;
      inf = decode_item(a, item);
      if (inf.t0 < inf.T) break;
    }
    st = 0;
    started = true;
    return true;
  }
};

__global__ void __launch_bounds__(UM_THREADS, 1) cqt_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* b_s = smem_raw;
  uint8_t* a_base = b_s + a.b_region_bytes;                     // stages: [hi | lo] x UM_STAGES
  const uint32_t UM_STAGES = (uint32_t)a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + 2 * UM_STAGES * a.a_region_bytes);
  uint64_t* full = bars;                        // [stages]
  uint64_t* empty = bars + UM_MAX_STAGES;       // [stages]
  uint64_t* tfull = bars + 2 * UM_MAX_STAGES;   // [2]
  uint64_t* tempty = bars + 2 * UM_MAX_STAGES + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * UM_MAX_STAGES + 4);

  // broadcast so the compiler knows the role index is warp-uniform (keeps MMA operands in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < UM_STAGES; ++s) {
      mbar_init(&full[s], UM_PRODUCER_WARPS);
      mbar_init(&empty[s], UM_MMA_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], UM_MMA_WARPS);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  const int64_t G = gridDim.x;

  if (warp < UM_MMA_WARPS) {
    // =========================== MMA issuers ===========================
    uint32_t it_stage = 0, it_acc = 0;
    PROF_DECL;
    for (int64_t item = blockIdx.x; item < a.total_items; item += G) {
      const ItemInfo inf = decode_item(a, item);
      if (inf.t0 >= inf.T) continue;
      const UmmaOct& oc = a.oct[inf.o];
      PROF_T(0);
      // instruction descriptors: D=f32, A=B=tf32, K-major both, M = 128; N = 2*npad (hi|lo) and N = npad
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UM_TILE_M >> 4) << 24);
      const uint32_t idesc_n1 = idesc_base | ((uint32_t)(oc.npad >> 3) << 17);
      const uint32_t idesc_n2 = idesc_base | ((uint32_t)((2 * oc.npad) >> 3) << 17);
      const uint32_t acc = it_acc & 1;
      mbar_wait_warp(&tempty[acc], ((it_acc >> 1) & 1) ^ 1, a.error_flag);
      tc_fence_after();
      PROF_T(1);
      const uint32_t d_tmem = tmem_base + acc * a.acc_stride + (uint32_t)warp * a.grp_stride;
      const uint32_t plane16 = (uint32_t)oc.rows_pad;          // plane pitch in 16-byte units
      const uint32_t bchunk16 = 2u * (uint32_t)oc.npad;        // K-chunk pitch of the bank in 16-byte units
      const bool split = (a.n_split == 3);
      const uint32_t idesc_main = split ? idesc_n2 : idesc_n1;
      const uint64_t db0 = smem_desc(smem_u32(b_s), bchunk16 * 16u, 128);
      uint32_t accum = 0;
      int rr = warp;      // next K-slice (within the current stage) owned by this warp
      for (int st = 0; st < oc.n_stages; ++st, ++it_stage) {
        const uint32_t s = it_stage % UM_STAGES;
        mbar_wait_warp(&full[s], (it_stage / UM_STAGES) & 1, a.error_flag);
        tc_fence_after();
        PROF_T(2);
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
              tc_mma_split_run(d_tmem, dah0 + a_off, dal0 + a_off, db, idesc_main, idesc_n1, accum, dq, dq * bq,
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
                if (split) tc_mma_tf32(d_tmem, dal0 + a_off, db, idesc_n1, 1u);
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
              tc_mma_split_run(d_tmem, dah0 + q0, dal0 + q0, db0 + (uint64_t)(q0 * bchunk16), idesc_main, idesc_n1,
                               accum, 2u * UM_MMA_WARPS, 2u * UM_MMA_WARPS * bchunk16, n_slices / UM_MMA_WARPS);
            } else {
              int sl = rr;
              for (; sl < n_slices; sl += UM_MMA_WARPS) {
                const uint32_t q = 2u * (uint32_t)sl;
                const uint64_t db = db0 + (uint64_t)(q * bchunk16);
                tc_mma_tf32(d_tmem, dah0 + q, db, idesc_main, accum);
                accum = 1;
                if (split) tc_mma_tf32(d_tmem, dal0 + q, db, idesc_n1, 1u);
              }
              rr = sl - n_slices;
            }
          }
        }
        PROF_T(3);
        tc_commit(&empty[s]);                                   // smem stage reusable once these MMAs retire
        if (st == oc.n_stages - 1) tc_commit(&tfull[acc]);      // accumulator complete
        __syncwarp();
        PROF_T(4);
      }
      ++it_acc;
    }
    if (prof_on && warp == 0 && lane == 0)
      for (int i = 0; i < 5; ++i) a.prof[blockIdx.x * UM_PROF_SLOTS + i] = pr[i];
  } else if (warp < UM_MMA_WARPS + UM_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                 // a warp may only touch TMEM lanes [32*(warp%4), +32)
    uint32_t it_acc = 0;
    PROF_DECL;
    for (int64_t item = blockIdx.x; item < a.total_items; item += G) {
      const ItemInfo inf = decode_item(a, item);
      if (inf.t0 >= inf.T) continue;
      const UmmaOct& oc = a.oct[inf.o];
      const uint32_t acc = it_acc & 1;
      mbar_wait(&tfull[acc], (it_acc >> 1) & 1, a.error_flag);
      tc_fence_after();
      PROF_T(0);
      const int t = inf.t0 + ew * 32 + lane;
      const int64_t row = (int64_t)inf.clip * a.out_clip_stride + (int64_t)t * a.frame_pitch;
      const uint32_t tbase = tmem_base + acc * a.acc_stride + ((uint32_t)(ew * 32) << 16);
      for (int c0 = 0; c0 < oc.npad && !(a.debug & 4); c0 += 16) {
        float sum[16];
#pragma unroll
        for (int f = 0; f < 16; ++f) sum[f] = 0.f;
#pragma unroll
        for (int g = 0; g < UM_MMA_WARPS; ++g) {
          uint32_t v[16], u[16];
          tmem_ld16(tbase + (uint32_t)g * a.grp_stride + (uint32_t)c0, v);
          if (a.n_split == 3) tmem_ld16(tbase + (uint32_t)g * a.grp_stride + (uint32_t)(oc.npad + c0), u);
          tmem_ld_wait();
#pragma unroll
          for (int f = 0; f < 16; ++f) {
            sum[f] += __uint_as_float(v[f]);
            if (a.n_split == 3) sum[f] += __uint_as_float(u[f]);
          }
        }
        if (t < inf.T) {
#pragma unroll
          for (int f = 0; f < 16; f += 2) {
            const int col = c0 + f;
            const int bin = oc.first_bin + (col >> 1);
            if (col < oc.ncol && bin >= 0 && bin < a.n_bins) {
              const float re = sum[f], im = sum[f + 1];
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
      PROF_T(1);
    }
    if (prof_on && ew == 0 && lane == 0)
      for (int i = 0; i < 2; ++i) a.prof[blockIdx.x * UM_PROF_SLOTS + 5 + i] = pr[i];
  } else {
    // =========================== producers ===========================
    // Software pipeline over this CTA's (item, stage) jobs: job k's raw fp32 rows are fetched with
    // cp.async (16 B, L2 -> SMEM, no register staging) a.prefetch jobs ahead of their conversion, so
    // several stages of L2 requests are in flight per SM; each thread later converts exactly the
    // elements it fetched (in place -> TF32 hi, plus lo), so cp.async.wait_group is the only sync.
    const int ptid = threadIdx.x - 32 * (UM_MMA_WARPS + UM_EPI_WARPS);
    constexpr int PT = 32 * UM_PRODUCER_WARPS;
    JobIter is, cv;
    is.reset(blockIdx.x, G);
    cv.reset(blockIdx.x, G);
    bool have_issue = is.next(a);
    int cur_oct = -1;
    const bool split = (a.n_split == 3);
    PROF_DECL;
    for (uint32_t k = 0;; ++k) {
      // ---- issue job k -------------------------------------------------------------------
      if (have_issue) {
        const UmmaOct& oc = a.oct[is.inf.o];
        const uint32_t s = k % UM_STAGES;
        PROF_T(0);
        mbar_wait(&empty[s], ((k / UM_STAGES) & 1) ^ 1, a.error_flag);
        PROF_T(1);
        float4* raw = reinterpret_cast<float4*>(a_base + (2 * s) * a.a_region_bytes);
        const float* y = oc.sig + (oc.sig_offsets ? oc.sig_offsets[is.inf.clip] : (int64_t)is.inf.clip * oc.sig_stride);
        const bool base_al = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
        const int64_t origin = (int64_t)is.inf.t0 * oc.hop - (oc.n_fft >> 1);
        const int g0 = is.st * a.pps;
        const int np = min(a.pps, oc.planes - g0);
        const int lg = 31 - __clz(np);
        // rows that valid frames read: frame r uses rows r .. r+Q-1; the rest of a partial tile stays stale
        // (MMA rows are independent and frames >= T are never stored)
        const int rows_used = min(oc.rows, is.inf.T - is.inf.t0 + oc.Q - 1);
        const int total = np * rows_used;
        const int64_t first = origin + 4 * g0;                                  // sample index of element (row 0, plane g0)
        const int64_t last = first + (int64_t)(rows_used - 1) * oc.hop + 4 * np;  // one past the last sample touched
        const uint32_t raw_s = smem_u32(raw);
        if (a.debug & 2) {
        } else if (base_al && first >= 0 && last <= is.inf.len) {
          // interior tile: no bounds / reflection tests, 32-bit offsets
          const float* src0 = y + first;
          for (int e = ptid; e < total; e += PT) {
            const int g = e & (np - 1), srow = e >> lg;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_s + 16u * (uint32_t)(g * oc.rows_pad + srow)),
                         "l"(src0 + (srow * oc.hop + 4 * g))
                         : "memory");
          }
        } else {
          for (int e = ptid; e < total; e += PT) {
            const int g = e & (np - 1), srow = e >> lg;
            const int64_t idx = first + (int64_t)srow * oc.hop + 4 * g;
            float4* dst = raw + (g * oc.rows_pad + srow);
            if (base_al && idx >= 0 && idx + 3 < is.inf.len) {
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(y + idx) : "memory");
            } else {
              float4 x;
              x.x = __ldg(y + reflect_index(idx, is.inf.len));
              x.y = __ldg(y + reflect_index(idx + 1, is.inf.len));
              x.z = __ldg(y + reflect_index(idx + 2, is.inf.len));
              x.w = __ldg(y + reflect_index(idx + 3, is.inf.len));
              *dst = x;
            }
          }
        }
        have_issue = is.next(a);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");      // one group per k, possibly empty
      PROF_T(2);
      if (k < a.prefetch) continue;
      // ---- convert job k - a.prefetch ------------------------------------------------------
      if (!cv.next(a)) break;
      const uint32_t kc = k - a.prefetch;
      const UmmaOct& oc = a.oct[cv.inf.o];
      if (cv.inf.o != cur_oct) {
        // new bank: every MMA that reads the old one must have retired
        if (kc > 0) mbar_wait(&empty[(kc - 1) % UM_STAGES], ((kc - 1) / UM_STAGES) & 1, a.error_flag);
        const int n16 = (oc.n_fft / 4) * 2 * oc.npad;      // 16-byte units of the packed bank
        const float4* gb = reinterpret_cast<const float4*>(oc.b_pack);
        float4* sb = reinterpret_cast<float4*>(b_s);
        for (int i = ptid; i < n16; i += PT) sb[i] = __ldg(gb + i);
        cur_oct = cv.inf.o;
        PROF_T(3);
      }
      PROF_T(0);
      if (a.prefetch == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (a.prefetch == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else asm volatile("cp.async.wait_group 3;" ::: "memory");
      PROF_T(4);
      {
        const uint32_t s = kc % UM_STAGES;
        float4* dh = reinterpret_cast<float4*>(a_base + (2 * s) * a.a_region_bytes);
        float4* dl = reinterpret_cast<float4*>(a_base + (2 * s + 1) * a.a_region_bytes);
        const int g0 = cv.st * a.pps;
        const int np = min(a.pps, oc.planes - g0);
        const int lg = 31 - __clz(np);
        const int total = np * min(oc.rows, cv.inf.T - cv.inf.t0 + oc.Q - 1);
        for (int e = ptid; e < total && !(a.debug & 2); e += PT) {
          const int g = e & (np - 1), srow = e >> lg;
          const int d = g * oc.rows_pad + srow;
          const float4 x = dh[d];
          if (split) {
            // kind::tf32 reads only the top 19 bits of each operand (the low 13 mantissa bits are
            // ignored), so the raw fp32 samples already ARE the hi operand: hi = trunc13(x), and
            // lo = tf32(x - hi) is exact up to 2^-21 |x|.  Only the lo plane has to be written.
            float4 l;
            l.x = to_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
            l.y = to_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
            l.z = to_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
            l.w = to_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
            dl[d] = l;
          } else {
            // single pass: round to nearest instead of the hardware's truncation
            float4 h;
            h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
            dh[d] = h;
          }
        }
        PROF_T(5);
        if (!(a.debug & 8)) fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        PROF_T(6);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (prof_on && ptid == 0)
      for (int i = 0; i < 7; ++i) a.prof[blockIdx.x * UM_PROF_SLOTS + 7 + i] = pr[i];
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
  float* d_pack = nullptr;
  int npad = 0;
};

}  // namespace saga

struct CqtUmmaState {
  std::vector<saga::OctPack> packs;
  bool supported = false;
  uint32_t b_region_bytes = 0, a_region_bytes = 0, tmem_cols = 0, acc_stride = 0, grp_stride = 0;
  size_t smem_bytes = 0;
  int* d_error = nullptr;
  long long* d_prof = nullptr;
  int num_sms = 0;
  int n_split = 3;
  int stages = 2, pps = 8, prefetch = 1;   // pipeline shape (SAGA_UMMA_CFG="stages,planes,prefetch" overrides)
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
  if (const char* cfg = getenv("SAGA_UMMA_CFG")) {
    int a_ = 0, b_ = 0, c_ = 0;
    if (sscanf(cfg, "%d,%d,%d", &a_, &b_, &c_) == 3 && a_ >= 2 && a_ <= UM_MAX_STAGES && (b_ == 4 || b_ == 8) && c_ >= 1 &&
        c_ <= 3 && c_ < a_) {
      st->stages = a_; st->pps = b_; st->prefetch = c_;
    }
  }
  uint32_t bmax = 0, amax = 0;
  int npad_max = 0;
  for (auto& o : p->oct) {
    const int ncol = 2 * o.n_filters;
    const int npad = (ncol + 15) & ~15;
    if (npad > 128) return;   // [hi|lo] operand: N = 2*npad <= 256
    if (o.hop < 4 || (o.hop % 4) != 0 || (o.n_fft % o.hop) != 0 || (o.n_fft % 8) != 0) return;
    const int planes = o.hop / 4;
    if (planes >= 2 && (std::min(planes, st->pps) % 2) != 0) return;
    const int Q = o.n_fft / o.hop;
    if (planes == 1 && (Q % 2) != 0) return;
    const int np0 = std::min(planes, st->pps);
    const int slices = planes >= 2 ? Q * (np0 / 2) : Q / 2;      // K-slices per stage
    const int n_st = (planes + st->pps - 1) / st->pps;
    if (slices * n_st < UM_MMA_WARPS) return;                     // every issuer warp must get work in an item
    const int rows = UM_TILE_M + Q - 1;
    const int rows_pad = rows | 1;
    bmax = std::max<uint32_t>(bmax, (uint32_t)o.n_fft * 2u * npad * 4u);
    amax = std::max<uint32_t>(amax, (uint32_t)std::min(planes, st->pps) * rows_pad * 16u);
    npad_max = std::max(npad_max, npad);
  }
  amax = (amax + 127u) & ~127u;
  bmax = (bmax + 127u) & ~127u;
  const size_t smem = (size_t)bmax + 2ull * st->stages * amax + 256;
  if (smem > 225 * 1024) return;
  uint32_t cols = 32;
  while (cols < 2u * UM_MMA_WARPS * 2u * npad_max) cols <<= 1;   // 2 buffers x issuer groups x (hi|lo)
  if (cols > 512) return;
  st->b_region_bytes = bmax;
  st->a_region_bytes = amax;
  st->tmem_cols = cols;
  st->acc_stride = cols / 2;
  st->grp_stride = 2u * npad_max;
  st->smem_bytes = smem;
  // pack each bank for the B descriptor: [n_fft/4][npad][4], split into TF32 hi / lo
  for (auto& o : p->oct) {
    OctPack pk;
    const int ncol = 2 * o.n_filters;
    pk.npad = (ncol + 15) & ~15;
    const size_t n = (size_t)o.n_fft * 2 * pk.npad;
    std::vector<float> pack(n, 0.f);
    for (int k = 0; k < o.n_fft; ++k)
      for (int c = 0; c < ncol; ++c) {
        const float b = o.bank_host[(size_t)k * ncol + c];
        const float h = tf32_rna_host(b);
        const size_t chunk = (size_t)(k / 4) * 2 * pk.npad;
        pack[(chunk + c) * 4 + (k % 4)] = h;
        pack[(chunk + pk.npad + c) * 4 + (k % 4)] = tf32_rna_host(b - h);
      }
    if (cudaMalloc(&pk.d_pack, n * 4) != cudaSuccess) return;
    cudaMemcpy(pk.d_pack, pack.data(), n * 4, cudaMemcpyHostToDevice);
    st->packs.push_back(pk);
  }
  if (cudaMalloc(&st->d_error, sizeof(int)) != cudaSuccess) return;
  cudaMemset(st->d_error, 0, sizeof(int));
  if (cudaMalloc(&st->d_prof, sizeof(long long) * 256 * UM_PROF_SLOTS) != cudaSuccess) return;
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
  a.grp_stride = st->grp_stride;
  a.error_flag = st->d_error;
  a.prof = st->d_prof;
  a.stages = st->stages;
  a.pps = st->pps;
  a.prefetch = st->prefetch;
  a.uniform_T = lv.clip_lens ? 0 : (int)T_max;
  a.uniform_len = lv.max_len;
  {
    const char* dbg = getenv("SAGA_UMMA_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
  }
  const int64_t per_oct = (int64_t)n_clips * a.tiles_per_clip;
  for (int i = 0; i < a.n_oct; ++i) {
    const CqtOctaveDev& o = p->oct[i];
    UmmaOct& u = a.oct[i];
    const bool raw = (o.level == 0 && p->early_factor == 1);
    u.sig = raw ? lv.wav : lv.lvl[o.level];
    u.sig_offsets = raw ? lv.clip_offsets : nullptr;
    u.sig_stride = raw ? 0 : lv.pitch[o.level];
    u.b_pack = st->packs[i].d_pack;
    u.level = o.level;
    u.hop = o.hop;
    u.n_fft = o.n_fft;
    u.ncol = 2 * o.n_filters;
    u.npad = st->packs[i].npad;
    u.first_bin = o.first_bin;
    u.planes = o.hop / 4;
    u.np_log2 = u.planes >= 4 ? 2 : (u.planes == 2 ? 1 : 0);
    u.n_stages = (u.planes + st->pps - 1) / st->pps;
    u.Q = o.n_fft / o.hop;
    u.rows = UM_TILE_M + u.Q - 1;
    u.rows_pad = u.rows | 1;
    u.item_begin = per_oct * i;
    {
      int64_t len = lv.max_len;
      if (p->early_factor > 1) len = (len + p->early_factor - 1) / p->early_factor;
      for (int s = 0; s < o.level; ++s) len = (len + 1) >> 1;
      u.uniform_len = len;
    }
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
  if (a.debug & 16) {
    // profiling aid only: synchronous read-back of the per-role cycle counters, mean over CTAs
    std::vector<long long> h((size_t)grid * UM_PROF_SLOTS);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), st->d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    static const char* names[14] = {"mma.decode", "mma.wait_tempty", "mma.wait_full", "mma.issue", "mma.commit",
                                    "epi.wait_tfull", "epi.work", "prod.loop", "prod.wait_empty", "prod.issue",
                                    "prod.bank", "prod.wait_group", "prod.convert", "prod.fence_arrive"};
    for (int s = 0; s < 14; ++s) {
      double sum = 0;
      for (int c = 0; c < grid; ++c) sum += (double)h[(size_t)c * UM_PROF_SLOTS + s];
      fprintf(stderr, "umma_prof %-18s %10.0f cycles/CTA\n", names[s], sum / grid);
    }
  }
  return SAGA_OK;
}

}  // namespace saga
