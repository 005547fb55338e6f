// Tensor memory (TMEM, 256 KB per SM on sm_100a) used as plain per-thread scratch: the spectral kernels have no matrix
// products, so the 512 columns x 128 lanes x 32 bit that normally hold tcgen05.mma accumulators are free, and the
// register file is the scarcest resource of the two-transforms-per-thread kernels.  Only tcgen05.alloc / st / ld /
// dealloc are used.  Addressing: bits 31..16 = lane, 15..0 = column; with the 32x32b shape warp w of a CTA reaches
// lanes 32 (w % 4) .. 32 (w % 4) + 31, lane l of the warp owning TMEM lane 32 (w % 4) + l, so every thread reads back
// exactly what it wrote and no cross-thread ordering is needed (warps w and w + 4 share lanes and use different columns).
#pragma once
#include <stdint.h>

namespace adept {

// executed by ONE full warp; ncols: power of two, 32..512.  The base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// this thread's address for column `col`: lane quadrant of its warp
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int warp, int col) {
  return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
}

// 8 doubles (16 columns) per call; all 32 lanes of the warp must execute these together
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&v)[8]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r[2 * i] = (uint32_t)__double2loint(v[i]);
    r[2 * i + 1] = (uint32_t)__double2hiint(v[i]);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double (&v)[8]) {  // follow with tmem_wait_ld() before use
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

// 4 doubles (8 columns) / 2 doubles (4 columns) per call
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, double (&v)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, double (&v)[2]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 2; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

}  // namespace adept
