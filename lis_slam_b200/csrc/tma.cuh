// Minimal sm_100a bulk-copy (TMA, 1-D form) + mbarrier + cluster helpers as inline PTX.
//   cp.async.bulk.shared::cluster.global   -> SASS UBLKCP.S.G          (bytes land in shared memory without passing registers)
//   mbarrier.arrive.expect_tx / try_wait   -> SASS SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK
//   .multicast::cluster                    -> one L2 read delivered to the same offset of every CTA in the mask
// Used by k_epsc_score (epsc.cuh): descriptor rows (1600 B) are exactly the "contiguous rows staged into shared
// memory" case the copy engine is made for; the query rows of a tile are shared by all CTAs of a cluster -> multicast.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lisreg { namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barrier visible to the async proxy / the other CTAs of the cluster
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
// global -> this CTA's shared memory; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// global -> the same shared-memory offset of every CTA whose bit is set in cta_mask; each destination's barrier (same
// offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned short cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}}  // namespace lisreg::tma
