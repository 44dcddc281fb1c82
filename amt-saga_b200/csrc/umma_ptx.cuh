// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the two tensor-core contraction kernels
// (cqt_umma.cu: resident bank + Hankel planes; cqt_umma_stream.cu: gathered rows + streamed bank).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace saga {

constexpr uint32_t UM_SPIN_LIMIT = 1u << 27;

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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
#define PROF_DECL                            \
  const bool prof_on = (a.debug & 16) != 0;  \
  long long pr[6] = {0, 0, 0, 0, 0, 0};      \
  long long pt = prof_on ? clock64() : 0
#define PROF_T(slot)                \
  do {                              \
    if (prof_on) {                  \
      const long long n_ = clock64(); \
      pr[slot] += n_ - pt;          \
      pt = n_;                      \
    }                               \
  } while (0)
#define PROF_OUT(base, n)                                                                               \
  do {                                                                                                  \
    if (prof_on)                                                                                        \
      for (int i_ = 0; i_ < (n); ++i_) a.prof[blockIdx.x * UM_PROF_SLOTS + (base) + i_] = pr[i_];       \
  } while (0)
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
// per-instruction R2UR waterfall.
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
// (A_hi x [B_hi|B_lo], N = n_main) and the correction MMA (A_lo x B_hi, N = n_lo); then both A
// descriptors advance by `da` and the B descriptor by `db` (16-byte units, added to the low word).
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace saga
