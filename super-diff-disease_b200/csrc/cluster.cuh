// Helpers shared by the 2-CTA (cta_group::2) conv kernels: cluster rank / sync, remote and relaxed mbarrier arrives,
// cluster-scope waits, the 2-CTA UMMA issue + multicast commit, the generic->async proxy fence.
// (The generation-2 kernel that used to live here -- TMA halo loads + in-place shared-memory transform + per-tile
// statistics publishing, 0.61 of peak on 128->128 -- was superseded by conv_tc3.cuh / conv_tc4.cuh; its measurements are
// in profiles/r1_v2_summary.md and DESIGN.md.)
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"

namespace sdd {

constexpr int kHaloRowsV2 = (kTileH + 2) * kHaloW;  // 180 rows of 128 B in one 64-channel halo box
constexpr int kC2SmemLimit = 232448;                // 227 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the barrier at the same smem offset in CTA `rank` of the pair
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// relaxed flavour: orders nothing but the barrier itself (used where only tcgen05 fences matter, so the
// arrive does not wait for the thread's outstanding global stores)
__device__ __forceinline__ void mbar_arrive_relaxed_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait_cluster(bar, parity)) return;
  printf("sdd: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread arrives on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// debug stamps go to shared memory (a global store would be dragged into the release of the next mbarrier
// arrive and perturb exactly what is being measured) and are dumped once at kernel exit
constexpr int kTraceIters = 12;
#define SDD_TRACE(role, iter, ev)                                                                    \
  do {                                                                                               \
    if (a.trace && blockIdx.x < 2 && (iter) < kTraceIters) s_trace[role][iter][ev] = clock64();      \
  } while (0)

}  // namespace sdd
