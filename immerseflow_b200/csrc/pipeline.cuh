// pipeline.cuh — sm_100a async-copy pipeline primitives (mbarrier + cp.async.bulk, the TMA engine's
// 1-D bulk mode; SASS: UBLKCP / SYNCS).  A dedicated producer warp streams row segments of the fields
// from HBM into a ring of shared-memory stages; consumer warps pick rows up as they land.  The number
// of bytes in flight per SM is set by the ring depth, not by registers or occupancy — which is what a
// 6.5 TB/s x ~0.7 us memory system needs (~30 KB per SM in flight) and what the register-rolled v1
// kernels could not provide (ncu: 45 % / 28 % of DRAM peak at 12 / 20 resident warps).
//
// Two proxies meet in a stage.  Filled -> consumed: the bulk copy's complete_tx on the stage's "full" barrier makes its
// (async-proxy) writes visible to whoever observes the phase — nothing to add.  Consumed -> refilled: the consumers' arrive
// on the "empty" barrier does NOT order their (generic-proxy) loads before the refill; an LDS that has only been issued
// can be overtaken by the copy.  So a consumer arrives only after the values it loaded have been USED (here: after the
// row's stores, which depend on them), or after fence.proxy.async.  profiles/r2_nc2_race.md has the measurement.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion signalled on an mbarrier (complete_tx::bytes).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

}  // namespace ifx
