// Two-CTA ("pair") primitives for sm_100a: thread-block cluster of two CTAs sharing one tcgen05.mma.cta_group::2.
// The leader (cluster rank 0) issues the MMAs for both; every CTA feeds its own half of the operands from its own shared
// memory and owns the 128 accumulator rows that land in its own TMEM. Verified on B200 by tools/umma_pair_test.cu.
#pragma once
#include "common.cuh"

namespace cilrs {

CILRS_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster
CILRS_DEVINL uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
  return a;
}
CILRS_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
CILRS_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
CILRS_DEVINL void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

// arrive on the mbarrier at the same shared-memory offset in CTA `rank` (release at cluster scope)
CILRS_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(bar, rank)) : "memory");
}
// wait with acquire at cluster scope (the arrivals may come from the peer CTA)
CILRS_DEVINL bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
CILRS_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (globaltimer_ns() - t0 > 4000000000ull) { __trap(); }
  }
}

// TMA load into THIS CTA's shared memory whose completion is signalled on the LEADER's mbarrier (same offset, rank 0)
CILRS_DEVINL void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(mapa_u32(bar, 0)), "r"(c0), "r"(c1)
      : "memory");
}

// TMEM allocation for a CTA pair: the same warp of BOTH CTAs executes these
CILRS_DEVINL void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
CILRS_DEVINL void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
CILRS_DEVINL void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows per CTA), B's N rows split half / half between the two CTAs
CILRS_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `cta_mask` once all MMAs issued so far by this thread completed
CILRS_DEVINL void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

}  // namespace cilrs
