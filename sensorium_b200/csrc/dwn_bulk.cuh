// 1-D bulk (TMA) global -> shared staging for the streaming kernels: one elected thread issues
// cp.async.bulk copies of whole contiguous row chunks, completion is tracked by an mbarrier, and the CTA consumes the
// chunk from shared memory.  Bytes in flight per SM = stages x chunk size (not registers x occupancy), which is what
// an HBM-bound pass needs to cover the ~1 us DRAM latency at 6.5 TB/s.
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t bk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bk_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bk_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bk_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bk_mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a protocol bug traps (clean launch failure) instead of hanging the GPU
  for (uint32_t it = 0; !bk_mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 28)) __trap();
}
// bytes: multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   bk_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bk_smem_u32(bar))
               : "memory");
}
