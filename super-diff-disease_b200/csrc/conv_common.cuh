// Shared by the 2-CTA (cta_group::2) tensor-core conv kernels: tile geometry, the argument block, cluster rank / sync,
// remote and relaxed mbarrier arrives, cluster-scope waits, the 2-CTA UMMA issue + multicast commit, the
// generic->async proxy fence, shared-memory sizing, and the two tiny GroupNorm helper kernels.
// (History of the superseded kernel generations 1-3 -- streamed weights, TMA halo + in-place transform, single loader
// group -- with their measurements: DESIGN.md section 3 and profiles/r1_v*_summary.md.)
#pragma once
#include "common.cuh"

namespace sdd {

#ifndef SDD_CONV_ACCS
#define SDD_CONV_ACCS 4  // TMEM accumulators per CTA pair (2 = round 1; A/B builds: -DSDD_CONV_ACCS=2)
#endif
#ifndef SDD_CONV_HALF_MATH
#define SDD_CONV_HALF_MATH false  // fused GroupNorm+SiLU transform in fp32 (true: packed fp16 math, see conv_tc4.cuh)
#endif

constexpr int kTileH = 16, kTileW = 8;  // 128 output pixels per CTA and accumulator
constexpr int kHaloW = kTileW + 2;
constexpr int kHaloRowsV2 = (kTileH + 2) * kHaloW;                              // 180 rows of 128 B per 64-channel box
constexpr int kHaloBytes = ((kTileH + 2) * kHaloW * 128 + 1023) / 1024 * 1024;  // 23552
constexpr int kC2SmemLimit = 232448;                                            // 227 KB

constexpr int kC3LoaderWarps = 8;
constexpr int kC3LoaderThreads = kC3LoaderWarps * 32;  // 256
// warps: 0 weights TMA, 1 MMA issuer, 2 TMEM alloc, 3 raw-ring TMA producer, 4-11 epilogue, 12-19 loaders
constexpr int kC3Threads = 384 + kC3LoaderThreads;
constexpr int kC3MaxStages = 6;

struct ConvTc3Args {
  const act_t* in;             // raw input, fp16 NHWC [B][H][W][Cin]
  act_t* out;                  // fp16 NHWC [B][H][W][COUT]
  BiasRef bias;
  const float* in_ab;          // [2][B][Cin] pre-halved GroupNorm+SiLU scale (plane 0) / shift (plane 1) per sample and
                               // channel, written by gn_scale_shift_kernel; nullptr: the input is used as is
  long long* out_sums;         // [B][4][2] fixed-point accumulators of the OUTPUT (zeroed by the caller), or nullptr
  int B, H, W;
  int tiles_w, tiles_per_sample, num_tiles, num_pairs;
  int stages;
  int raw_slots;               // kRaw: raw TMA slots behind the operand stages (0 = register-path loader)
  int prefetch;                // register-path loader: TMA L2 prefetch of a group's item after next (0 = off)
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// relaxed remote arrive: orders nothing but the barrier itself (used where only tcgen05 fences / proxy fences matter, so
// the arrive does not wait for the thread's outstanding global accesses)
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
// kind::f16 covers fp16 and bf16 operands (selected by the instruction descriptor)
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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

// shared memory needed for (COUT, Cin) with `stages` halo stages (operand stages + raw slots)
inline int conv_tc3_smem_bytes(int Cout, int Cin, int stages) {
  return 9 * (Cin / 64) * (Cout / 2) * 128 + stages * kHaloBytes + 1024 /*alignment*/ + 512 /*barriers*/;
}
inline int conv_tc3_stages(int Cout, int Cin) {
  int s = kC3MaxStages;
  while (s > 1 && conv_tc3_smem_bytes(Cout, Cin, s) > kC2SmemLimit - 4096 /*static smem*/) --s;
  return s;
}

// Per-(sample, channel) scale / shift of the fused GroupNorm+SiLU transform, pre-halved for silu(v) = h + h tanh(h):
// ab[0][b][c] = rstd * gamma / 2, ab[1][b][c] = (beta - mean * rstd * gamma) / 2.  One tiny launch per conv layer; the
// loaders then fetch their eight channels' values with four 16-byte loads when the sample changes instead of rebuilding
// them behind two named barriers (double-precision statistics, an LDS round trip: 7.6 % of the loaders' stall samples
// in the round-1 128->128 ncu capture).
__global__ void gn_scale_shift_kernel(const long long* __restrict__ sums, const float* __restrict__ meanrstd,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, double count,
                                      float eps, float* __restrict__ ab, int B, int Cin) {
  const int b = blockIdx.x, ch = threadIdx.x;
  if (ch >= Cin) return;
  const int g = ch / (Cin / 4);
  float mean, rstd;
  if (sums) gn_mean_rstd_from_sums(sums + ((size_t)b * 4 + g) * 2, count, eps, mean, rstd);
  else { mean = meanrstd[(b * 4 + g) * 2]; rstd = meanrstd[(b * 4 + g) * 2 + 1]; }
  const float sc = rstd * gamma[ch];
  ab[(size_t)b * Cin + ch] = 0.5f * sc;
  ab[(size_t)B * Cin + (size_t)b * Cin + ch] = 0.5f * (beta[ch] - mean * sc);
}

// Operator API only: (mean, rstd) floats from the fixed-point sums.
__global__ void gn_sums_to_meanrstd_kernel(const long long* sums, float* meanrstd, int groups_total, double count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups_total) return;
  float mean, rstd;
  gn_mean_rstd_from_sums(sums + (size_t)i * 2, count, eps, mean, rstd);
  meanrstd[2 * i] = mean;
  meanrstd[2 * i + 1] = rstd;
}

}  // namespace sdd
