// Product tensor-core 3x3 conv: GroupNorm+SiLU -> conv3x3 -> (+bias, next GroupNorm's statistics), one kernel.
//
//   out[n,h,w,co] = bias[n,co] + sum_{ky,kx,ci} f(raw[n,h+ky-1,w+kx-1,ci]) * wt[kx][ky][co][ci]
//   f(v) = silu((v - mean[n,g]) * rstd[n,g] * gamma[ci] + beta[ci])  inside the image, 0 in the padding
//
// CTA pair (cluster of 2, tcgen05 cta_group::2): UMMA M = 256 = two 16x8-pixel tiles (one per CTA),
// N = Cout, K = 9*Cin, fp32 accumulators double-buffered in TMEM.  Per CTA:
//   * its half of the weights (Cout/2 rows of every tap) is loaded ONCE by TMA and stays resident in shared
//     memory for the whole persistent kernel (128->128: 144 KB) -- only activations move per tile;
//   * per tile and 64-channel chunk, four loader warps read the (18 x 10)-pixel halo box of the RAW input with
//     coalesced 16-byte global loads, apply GroupNorm+SiLU in registers (halo overhead 1.4x, not 9x), and store
//     it as 180 rows of 128 B in the UMMA 128-byte swizzle.  The nine taps are then descriptor VIEWS of that
//     one box: start at row ky*10+kx, 1280-byte stride between 8-row groups (bit-exact, tools/exp_halo.py).
//     Measured: an in-place shared-memory transform behind a TMA load starves -- SS-mode UMMA already uses
//     ~100 of the 128 B/clk of shared-memory bandwidth -- so the operand is transformed BEFORE it is stored;
//   * one elected lane issues the 36 MMAs per chunk back to back (elect.sync keeps the descriptors in uniform
//     registers: 61 cycles per M256 x N128 x K16 MMA = the nominal rate);
//   * eight epilogue warps (4 TMEM lane quadrants x 2 column halves) drain TMEM, add the bias, accumulate the
//     output's GroupNorm partial sums and write bf16 NHWC with full-sector 32-byte stores; a separate warp
//     publishes the statistics (fence + atomic + fixed-order finalize) off the critical path.
// Barriers (arrival count): ready[s] loaders->MMA (8, on the leader) | empty[s] MMA->loaders (1, multicast
// commit) | tfull[a] MMA->epilogue (1, multicast commit) | tempty[a] epilogue->MMA (16, on the leader) |
// wbar weights landed (1+tx) | sfull/sempty[2] epilogue<->statistics publisher (8 / 1).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"

namespace sdd {

constexpr int kC2XformWarps = 8;  // transform warps (2 per SM sub-partition: one warp alone is latency-bound)
constexpr int kC2XformThreads = kC2XformWarps * 32;
// warps: 0 TMA producer, 1 MMA, 2 TMEM alloc, 3 stats publisher, 4-11 epilogue, 12.. transform
constexpr int kC2Threads = 384 + kC2XformThreads;
constexpr int kC2MaxStages = 6;
constexpr int kHaloRowsV2 = (kTileH + 2) * kHaloW;  // 180
constexpr int kHaloVecs = kHaloRowsV2 * 8;          // 1440 16-byte vectors per stage
constexpr int kVecsPerLoader = (kHaloVecs + kC2XformThreads - 1) / kC2XformThreads;  // 6 with 8 warps
constexpr int kC2SmemLimit = 232448;                // 227 KB

struct ConvTc2Args {
  const __nv_bfloat16* in;   // raw input, bf16 NHWC [B][H][W][Cin]
  __nv_bfloat16* out;
  BiasRef bias;
  const float* in_meanrstd;  // [B][4][2] of the INPUT tensor, or nullptr: input is already activated
  const float* in_gamma;     // [Cin]
  const float* in_beta;      // [Cin]
  float* partials;           // [B][tiles_per_sample][4][2]
  int* counters;             // [B]
  float* meanrstd;           // [B][4][2] of the OUTPUT tensor
  int B, H, W, Cin;
  int tiles_w, tiles_per_sample, num_tiles, num_pairs;
  int stages;
  long long* trace;  // optional [2 ctas][6 roles][64 iters][4 events] stamps of CTA pair 0 (debug)
  int dbg;           // timing experiments only (results invalid): 2 = no stores/stats, 4 = no MMA, 8 = no stats publish
  int prefetch;      // halo boxes pulled into L2 this many pipeline items ahead of their shared-memory load (0 = off)
};

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

template <int COUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC2Threads, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const ConvTc2Args a) {
  constexpr int kWSlot = (COUT / 2) * 128;  // bytes of one (tap, chunk) weight slice held by this CTA
  constexpr int kTmemCols = 2 * COUT;
  extern __shared__ uint8_t smem_raw[];
  auto gtimer = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  // debug: role row 5 of the trace holds [entry, after-setup, loop-done, exit] globaltimer ns of the first 64 CTAs
  auto gstamp = [&](int k) {
    if (a.trace && blockIdx.x < 64) a.trace[((size_t)(blockIdx.x & 1) * 6 + 5) * 256 + (blockIdx.x >> 1) * 4 + k] = gtimer();
  };
  if (threadIdx.x == 0) gstamp(0);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int kchunks = a.Cin / 64;
  const uint32_t w_bytes = 9u * kchunks * kWSlot;
  const uint32_t a_base = smem_base + w_bytes;
  const uint32_t bar_base = a_base + (uint32_t)a.stages * kHaloBytes;
  auto ready_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC2MaxStages + s); };
  auto full_bar = [&](int s) { return bar_base + 8u * (2 * kC2MaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * kC2MaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * kC2MaxStages + 2 + s); };
  const uint32_t w_bar = bar_base + 8u * (3 * kC2MaxStages + 4);
  auto sfull_bar = [&](int s) { return bar_base + 8u * (3 * kC2MaxStages + 5 + s); };
  auto sempty_bar = [&](int s) { return bar_base + 8u * (3 * kC2MaxStages + 7 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * kC2MaxStages + 9);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  __shared__ long long s_trace[5][kTraceIters][4];
  if (a.trace && blockIdx.x < 2)
    for (int i = threadIdx.x; i < 5 * kTraceIters * 4; i += kC2Threads) (&s_trace[0][0][0])[i] = 0;
  __shared__ float s_red[2][8][4];  // per-tile GroupNorm partial sums: [slot][epilogue warp][g0 s, g0 ss, g1 s, g1 ss]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(ready_bar(s), 2 * kC2XformWarps); mbar_init(empty_bar(s), 1); mbar_init(full_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 16); }
    mbar_init(w_bar, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(sfull_bar(s), 8); mbar_init(sempty_bar(s), 1); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) gstamp(1);

  // tile of this CTA in a pair-iteration; an odd tile count leaves one dummy (clamped, computed, not stored)
  auto tile_of = [&](int pair, bool& valid) {
    int t = 2 * pair + (int)rank;
    valid = t < a.num_tiles;
    return valid ? t : a.num_tiles - 1;
  };

  if (warp == 0) {
    // ===================== resident weights: this CTA's Cout/2 rows of all 9 taps, once =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, w_bytes);
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(smem_base + (uint32_t)(tap * kchunks + kc) * kWSlot, &tmB, w_bar, kc * 64,
                      (int)rank * (COUT / 2), tap);
      // then one (64 ci x 10 w x 18 h) halo box per (tile, chunk); out-of-image pixels are zero-filled by TMA.
      // Shared memory only holds a few stages (3 for 128->128), too few to cover DRAM latency, so the boxes of the
      // items kPrefetch ahead are pulled into L2 first (cp.async.bulk.prefetch.tensor): the real load then hits L2.
      const int kPrefetch = a.prefetch;
      const int total_items = ((a.num_pairs - pair0 + pair_stride - 1) / pair_stride) * kchunks;
      auto coords = [&](int j, int& n, int& h0, int& w0, int& kc) {
        bool valid;
        const int tile = tile_of(pair0 + (j / kchunks) * pair_stride, valid);
        n = tile / a.tiles_per_sample;
        const int tr = tile % a.tiles_per_sample;
        h0 = (tr / a.tiles_w) * kTileH; w0 = (tr % a.tiles_w) * kTileW; kc = j % kchunks;
      };
      for (int j = 0; j < kPrefetch && j < total_items; ++j) {
        int n, h0, w0, kc;
        coords(j, n, h0, w0, kc);
        tma_prefetch_l2_4d(&tmA, kc * 64, w0 - 1, h0 - 1, n);
      }
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < total_items; ++j) {
        int n, h0, w0, kc;
        if (j + kPrefetch < total_items) {
          coords(j + kPrefetch, n, h0, w0, kc);
          tma_prefetch_l2_4d(&tmA, kc * 64, w0 - 1, h0 - 1, n);
        }
        coords(j, n, h0, w0, kc);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), kHaloVecs * 16);
        tma_load_4d(a_base + stage * kHaloBytes, &tmA, full_bar(stage), kc * 64, w0 - 1, h0 - 1, n);
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp walks the loop, one elected lane issues) ===
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, COUT);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int it = 0;
      for (int pair = pair0; pair < a.num_pairs; pair += pair_stride, ++it) {
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0) SDD_TRACE(1, it, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * COUT);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait_cluster(ready_bar(stage), phase);
          tc_fence_after();
          if (lane == 0) SDD_TRACE(1, it, 1 + kc);
          const uint32_t sa = a_base + stage * kHaloBytes;
          if (elect_one_sync()) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const uint64_t adesc = umma_desc_sw128(sa + (ky * kHaloW + kx) * 128, kHaloW * 128);
                const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)((kx * 3 + ky) * kchunks + kc) * kWSlot);
                if (a.dbg & 4) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x UMMA_K (16 bf16 = 32 B) inside the 128-byte swizzle row
                  umma_bf16_2cta(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                 (kc | kx | ky | k) ? 1u : 0u);
              }
            umma_commit_2cta(empty_bar(stage));                         // frees the stage in both CTAs
            if (kc == kchunks - 1) umma_commit_2cta(tfull_bar(acc));    // accumulator complete -> both epilogues
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        if (lane == 0) SDD_TRACE(1, it, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp == 3) {
    // ===================== GroupNorm statistics publisher (off the epilogue's critical path) ==========
    __shared__ float s_sums[8];
    int it = 0;
    for (int pair = pair0; pair < a.num_pairs; pair += pair_stride, ++it) {
      bool valid;
      const int tile = tile_of(pair, valid);
      const int n = tile / a.tiles_per_sample, tr = tile % a.tiles_per_sample;
      const int slot = it & 1;
      mbar_wait(sfull_bar(slot), (uint32_t)((it >> 1) & 1));
      if (lane == 0) SDD_TRACE(4, it, 0);
      if (valid && !(a.dbg & (2 | 8))) {
        if (lane < 8) {  // lane = group*2 + {sum, sumsq}; fixed order over the four lane quadrants
          const int g = lane >> 1, hc = g >> 1, k = (g & 1) * 2 + (lane & 1);
          s_sums[lane] = (s_red[slot][hc * 4 + 0][k] + s_red[slot][hc * 4 + 1][k]) +
                         (s_red[slot][hc * 4 + 2][k] + s_red[slot][hc * 4 + 3][k]);
        }
        __syncwarp();
        gn_publish_and_finalize_warp(s_sums, a.partials, a.counters, a.meanrstd, n, tr, a.tiles_per_sample, 4,
                                     (float)a.H * (float)a.W * (float)(COUT / 4), kGnEps);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty_bar(slot));
      if (lane == 0) SDD_TRACE(4, it, 1);
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves =====================
    constexpr int COLS = COUT / 2;  // columns drained by this warp (two GroupNorm groups)
    const int e = warp - 4, q = e & 3, hcol = e >> 2;
    const int col0 = hcol * COLS;
    const int m = q * 32 + lane;  // accumulator row = pixel within the tile
    int acc = 0; uint32_t acc_phase = 0;
    int it = 0;
    for (int pair = pair0; pair < a.num_pairs; pair += pair_stride, ++it) {
      bool valid;
      const int tile = tile_of(pair, valid);
      const int n = tile / a.tiles_per_sample, tr = tile % a.tiles_per_sample;
      const int h = (tr / a.tiles_w) * kTileH + (m >> 3), w = (tr % a.tiles_w) * kTileW + (m & 7);
      const float* bp = bias_ptr(a.bias, n) + col0;
      __nv_bfloat16* orow = a.out + (((size_t)n * a.H + h) * a.W + w) * COUT + col0;
      const bool do_store = valid && !(a.dbg & 2);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (e == 0 && lane == 0) SDD_TRACE(3, it, 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * COUT + col0);
      float sg[2] = {0.f, 0.f}, ssg[2] = {0.f, 0.f};
      uint32_t v[2][16];
      tmem_ld_32x16(taddr, v[0]);
#pragma unroll
      for (int st = 0; st < COLS / 16; ++st) {
        tmem_ld_wait();
        if (st + 1 < COLS / 16) tmem_ld_32x16(taddr + (uint32_t)((st + 1) * 16), v[(st + 1) & 1]);
        const uint32_t* vv = v[st & 1];
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + st * 16 + j));
          f[j] = __uint_as_float(vv[j]) + b4.x;         f[j + 1] = __uint_as_float(vv[j + 1]) + b4.y;
          f[j + 2] = __uint_as_float(vv[j + 2]) + b4.z; f[j + 3] = __uint_as_float(vv[j + 3]) + b4.w;
        }
        // two GroupNorm groups per warp; 4 independent partial accumulators per statistic
        const int g = (st * 16 >= COLS / 2) ? 1 : 0;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          p0 += f[j]; p1 += f[j + 1]; p2 += f[j + 2]; p3 += f[j + 3];
          r0 = fmaf(f[j], f[j], r0); r1 = fmaf(f[j + 1], f[j + 1], r1);
          r2 = fmaf(f[j + 2], f[j + 2], r2); r3 = fmaf(f[j + 3], f[j + 3], r3);
        }
        sg[g] += (p0 + p1) + (p2 + p3);
        ssg[g] += (r0 + r1) + (r2 + r3);
        if (do_store) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
          st_global_v8(orow + st * 16, pk);  // 16 channels = 32 B = one full sector per thread
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_remote(tempty_bar(acc), 0);  // orders TMEM reads only, not the stores
      if (e == 0 && lane == 0) SDD_TRACE(3, it, 1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sg[0] += __shfl_xor_sync(0xffffffffu, sg[0], o);   ssg[0] += __shfl_xor_sync(0xffffffffu, ssg[0], o);
        sg[1] += __shfl_xor_sync(0xffffffffu, sg[1], o);   ssg[1] += __shfl_xor_sync(0xffffffffu, ssg[1], o);
      }
      const int slot = it & 1;
      if (lane == 0) {
        mbar_wait(sempty_bar(slot), (uint32_t)(((it >> 1) & 1) ^ 1));
        s_red[slot][e][0] = sg[0]; s_red[slot][e][1] = ssg[0]; s_red[slot][e][2] = sg[1]; s_red[slot][e][3] = ssg[1];
        mbar_arrive(sfull_bar(slot));
      }
      if (e == 0 && lane == 0) SDD_TRACE(3, it, 2);
    }
  } else if (warp >= 12) {
    // ===================== transform warps: GroupNorm + SiLU on the landed halo box, in place ==========
    const int tt = threadIdx.x - 384;  // 0..kC2XformThreads-1
    __shared__ __align__(16) float s_ga[128], s_gb[128];
    const bool fuse = a.in_meanrstd != nullptr;
    // vector v = tt + 128*i sits at smem row v>>3, physical 16-byte chunk v&7; (v & 7) and (row & 7) are the
    // same for every vector of a thread, hence so is its logical channel chunk (physical ^ row&7): the eight
    // scale/shift pairs live in registers and the only shared-memory traffic is one LDS.128 + one STS.128
    const int lchunk = (tt & 7) ^ ((tt >> 3) & 7);
    int stage = 0; uint32_t phase = 0;
    int cur_n = -1;
    bool w_ready = false;
    int it = 0;
    for (int pair = pair0; pair < a.num_pairs; pair += pair_stride, ++it) {
      bool valid;
      const int tile = tile_of(pair, valid);
      const int n = tile / a.tiles_per_sample, tr = tile % a.tiles_per_sample;
      const int h0 = (tr / a.tiles_w) * kTileH, w0 = (tr % a.tiles_w) * kTileW;
      if (fuse && n != cur_n) {
        named_bar_sync(2, kC2XformThreads);  // previous readers of s_ga/s_gb are done
        if (tt < a.Cin) {
          const int g = tt / (a.Cin / 4);
          const float mean = a.in_meanrstd[(n * 4 + g) * 2], rstd = a.in_meanrstd[(n * 4 + g) * 2 + 1];
          const float sc = rstd * a.in_gamma[tt];
          s_ga[tt] = sc;
          s_gb[tt] = a.in_beta[tt] - mean * sc;
        }
        named_bar_sync(2, kC2XformThreads);
        cur_n = n;
      }
      for (int kc = 0; kc < kchunks; ++kc) {
        float ga[8], gb[8];
        uint32_t okmask = 0;
        if (fuse) {
          const int coff = kc * 64 + lchunk * 8;
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const float4 x = *reinterpret_cast<const float4*>(&s_ga[coff + j]);
            const float4 y = *reinterpret_cast<const float4*>(&s_gb[coff + j]);
            ga[j] = x.x; ga[j + 1] = x.y; ga[j + 2] = x.z; ga[j + 3] = x.w;
            gb[j] = y.x; gb[j + 1] = y.y; gb[j + 2] = y.z; gb[j + 3] = y.w;
          }
#pragma unroll
          for (int i = 0; i < kVecsPerLoader; ++i) {
            const int row = (tt >> 3) + (kC2XformThreads / 8) * i;
            const int hr = row / kHaloW, wr = row - hr * kHaloW;
            const int hh = h0 - 1 + hr, ww = w0 - 1 + wr;
            if (row < kHaloRowsV2 && hh >= 0 && hh < a.H && ww >= 0 && ww < a.W) okmask |= 1u << i;
          }
        }
        mbar_wait(full_bar(stage), phase);
        if (tt == 0) SDD_TRACE(2, it, kc);
        if (fuse && !(a.dbg & 64)) {
          uint8_t* sp = smem_raw + (a_base + stage * kHaloBytes - smem_u32(smem_raw));
          uint4 r[kVecsPerLoader];
#pragma unroll
          for (int i = 0; i < kVecsPerLoader; ++i)
            if ((okmask >> i) & 1u) r[i] = *reinterpret_cast<const uint4*>(sp + (size_t)(tt + kC2XformThreads * i) * 16);
#pragma unroll
          for (int i = 0; i < kVecsPerLoader; ++i) {
            if (!((okmask >> i) & 1u)) continue;  // padding stays the exact zeros TMA wrote
            uint32_t u[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 hv = *reinterpret_cast<__nv_bfloat162*>(&u[j]);
              const float lo = silu_tanh(fmaf(__low2float(hv), ga[2 * j], gb[2 * j]));
              const float hi = silu_tanh(fmaf(__high2float(hv), ga[2 * j + 1], gb[2 * j + 1]));
              u[j] = pack_bf16x2(lo, hi);
            }
            *reinterpret_cast<uint4*>(sp + (size_t)(tt + kC2XformThreads * i) * 16) = make_uint4(u[0], u[1], u[2], u[3]);
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        }
        if (!w_ready) { mbar_wait(w_bar, 0); w_ready = true; }  // this CTA's weights have landed too
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(ready_bar(stage), 0);
        if (tt == 0) SDD_TRACE(2, it, 2 + kc);
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (a.trace && blockIdx.x < 2)
    for (int i = threadIdx.x; i < 5 * kTraceIters * 4; i += kC2Threads) {
      const int role = i / (kTraceIters * 4), rem = i % (kTraceIters * 4);
      a.trace[(((size_t)blockIdx.x * 6 + role) * 64 + rem / 4) * 4 + (rem & 3)] = (&s_trace[0][0][0])[i];
    }
  if (threadIdx.x == 0) gstamp(2);
  cluster_sync_all();  // the peer may still be reading our smem / arriving on our barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (threadIdx.x == 64) gstamp(3);
}

// shared memory needed for (COUT, Cin) with `stages` halo stages
inline int conv_tc2_smem_bytes(int Cout, int Cin, int stages) {
  return 9 * (Cin / 64) * (Cout / 2) * 128 + stages * kHaloBytes + 1024 + 512;
}
inline int conv_tc2_stages(int Cout, int Cin) {
  int s = kC2MaxStages;
  while (s > 1 && conv_tc2_smem_bytes(Cout, Cin, s) > kC2SmemLimit - 2048 /*static smem*/) --s;
  return s;
}

}  // namespace sdd
