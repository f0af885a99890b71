// fsweep_tma.cuh — the TMA bulk-copy engine and mbarriers as inline PTX (sm_90+; SASS: UBLKCP, SYNCS.*).
// Used by the streaming table kernels (tile ring, fsweep_stream.cuh) and by the thread-per-bin kernels to stage the
// feedback matrix of a block (fsweep_tpc.cuh).
#pragma once
#include <cstdint>

namespace fsweep {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
  }
}
// global -> shared, completion counted in bytes on `bar`; src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   s_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s_u32(smem_src)), "r"(bytes)
               : "memory");
}


}  // namespace fsweep
