// tcgen05 / tensor-memory wrappers (inline PTX, sm_100a) for the tensor-core tier of the B=8
// transform.  Descriptor layouts follow cute/arch/mma_sm100_desc.hpp (InstrDescriptor,
// SmemDescriptor) and cute/atom/mma_traits_sm100.hpp (canonical no-swizzle layouts); they were
// checked on the hardware with profiles/microbench/tc_bringup.cu before anything was built on them.
#pragma once
#include <stdio.h>

#include <stdint.h>

#include "tma.cuh"

namespace vcfb {
namespace tc {

// ---- mbarrier pieces not in tma.cuh -------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma::smem_u32(bar)) : "memory");
}
// Wait for the phase with this parity.  try_wait suspends the warp in hardware for up to the hinted time
// instead of spinning: a polling warp takes issue slots from the warps that do the work (ncu on the first
// version of kernels_tc.cu: 83 % of the issue slots busy, half of them in wait loops).
#ifdef VCFB_TC_WATCHDOG
// Debug build: a wait that does not complete within ~0.2 s reports itself and returns, so that a dead-locked
// pipeline terminates (with garbage) and its printf buffer shows who waited for what.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < 2000; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(tma::smem_u32(bar)), "r"(parity), "r"(100000u)
        : "memory");
    if (done) return;
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0)
    printf("[watchdog] warp %d stuck on barrier at smem offset %u parity %u\n", int(threadIdx.x >> 5), tma::smem_u32(bar), parity);
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TCW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@!p bra TCW_%=;\n\t"
      "}" ::"r"(tma::smem_u32(bar)),
      "r"(parity), "r"(200000u)
      : "memory");
}
// for the latency-insensitive control warps: back off between polls
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
#ifdef VCFB_TC_WATCHDOG
  (void)ns;
  mbar_wait(bar, parity);
  return;
#endif
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(tma::smem_u32(bar)), "r"(parity), "r"(200000u)
        : "memory");
    if (!done) __nanosleep(ns);
  } while (!done);
}

// ---- tensor memory ------------------------------------------------------------------------
// One warp allocates NCOLS (power of two >= 32) columns; the base address lands in shared memory.
template <int NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t* holder) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(holder)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// lane i of the warp <-> TMEM lane 32 * (warp % 4) + i; 8 consecutive 32-bit columns
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---- descriptors --------------------------------------------------------------------------
// Shared-memory matrix descriptor, no swizzle.  K-major operand of 16-bit elements:
//   byte offset(mn, k) = (mn % 8) * 16 + (mn / 8) * SBO + (k / 8) * LBO + (k % 8) * 2
__host__ __device__ inline int off_kmajor16(int mn, int k, int sbo, int lbo) {
  return (mn % 8) * 16 + (mn / 8) * sbo + (k / 8) * lbo + (k % 8) * 2;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, int lbo, int sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 |
         (uint64_t)1 << 46;          // version 1 (Blackwell), base offset 0, layout type 0 (no swizzle)
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, shape M x N (K = 16)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(bar))
               : "memory");
}

// one lane of a converged warp (elect.sync): the tcgen05.mma operands live in uniform registers, and code the
// compiler knows to run on a single elected lane needs no per-instruction uniformity loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// named barrier among a subset of the CTA's warps
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace tc
}  // namespace vcfb
