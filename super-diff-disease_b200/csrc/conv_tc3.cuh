// Product tensor-core 3x3 conv, generation 3: GroupNorm+SiLU -> conv3x3 -> (+bias, next GroupNorm's sums), one kernel.
//
//   out[n,h,w,co] = bias[n,co] + sum_{ky,kx,ci} f(raw[n,h+ky-1,w+kx-1,ci]) * wt[kx][ky][co][ci]
//   f(v) = silu((v - mean[n,g]) * rstd[n,g] * gamma[ci] + beta[ci])  inside the image, 0 in the padding
//   (reference: GroupNorm -> SiLU -> Conv2d of ResidualBlock, /root/reference/src/models/unet.py:21-28)
//
// What v2 measured and v3 changes (profiles/r1_v3_notes.md):
//   * shared memory, not the tensor pipe, is the scarce resource: an SS-mode M256xN128xK16 UMMA reads 6 KB per CTA
//     in 64 cycles (96 of the 128 B/clk), so every extra pass over the operand tile stalls the MMAs.  v2 made three
//     passes per halo box (TMA write, LDS, STS); v3 makes ONE: the loader warps read the raw activations from
//     global memory into registers (coalesced 16-byte loads, software-pipelined one item ahead; the boxes are
//     pulled into L2 a few items earlier by cp.async.bulk.prefetch.tensor), apply GroupNorm+SiLU in registers and
//     store the activated tile straight into the UMMA 128-byte swizzle;
//   * per-tile statistics publishing (store + __threadfence + atomic counter + last-CTA finalize) cost up to 28% of
//     a launch.  v3 accumulates the output's GroupNorm sums with fire-and-forget 64-bit integer RED.ADDs of
//     fixed-point values (2^-20 units): integer addition is associative, so the result is bit-reproducible and
//     independent of batch sharding without any ordering protocol, and the consumer derives mean / rstd itself;
//   * per-tile index arithmetic and the bias-row pointer chase are hoisted out of the critical loops.
// Unchanged from v2: CTA pair (cluster of 2, tcgen05 cta_group::2), UMMA M = 256 = two 16x8-pixel tiles, N = Cout,
// K = 9*Cin; this CTA's half of the weights resident in shared memory for the whole persistent kernel; the nine taps
// are descriptor VIEWS of one (18 x 10)-pixel halo box (start row ky*10+kx, 1280-byte group stride); fp32
// accumulators double-buffered in TMEM; eight epilogue warps with full-sector 32-byte stores.
// Barriers (arrival count): ready[s] loaders->MMA (16, on the leader) | empty[s] MMA->loaders (1, multicast commit)
// | tfull[a] MMA->epilogue (1, multicast commit) | tempty[a] epilogue->MMA (16, on the leader) | wbar weights (1+tx).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "cluster.cuh"

namespace sdd {

constexpr int kC3LoaderWarps = 8;
constexpr int kC3LoaderThreads = kC3LoaderWarps * 32;  // 256
// warps: 0 weights TMA, 1 MMA issuer, 2 TMEM alloc, 3 spare, 4-11 epilogue, 12-19 loaders
constexpr int kC3Threads = 384 + kC3LoaderThreads;
constexpr int kC3MaxStages = 6;
constexpr int kC3Vecs = 6;  // 16-byte vectors per loader thread and item: 180 rows x 8 / 256 threads, rounded up

struct ConvTc3Args {
  const __nv_bfloat16* in;     // raw input, bf16 NHWC [B][H][W][Cin]
  __nv_bfloat16* out;          // bf16 NHWC [B][H][W][COUT]
  BiasRef bias;
  const long long* in_sums;    // [B][4][2] fixed-point (sum, sumsq) per GroupNorm group of the INPUT, or nullptr
  const float* in_meanrstd;    // or [B][4][2] (mean, rstd) as floats (operator API); both null: input used as is
  const float* in_gamma;       // [Cin]
  const float* in_beta;        // [Cin]
  const float* in_ab;          // generation 4+: [2][B][Cin] pre-halved GroupNorm+SiLU scale (plane 0) / shift (plane 1) per
                               // sample and channel, written by gn_scale_shift_kernel; replaces the four fields above
  long long* out_sums;         // [B][4][2] fixed-point accumulators of the OUTPUT (zeroed by the caller), or nullptr
  int B, H, W, Cin;
  int tiles_w, tiles_per_sample, num_tiles, num_pairs;
  int stages;
  int contig;                  // generation 4+: contiguous tile-pair ranges per CTA pair (0 = strided by the grid)
  int raw_slots;               // generation 5 (kRaw): raw TMA slots behind the operand stages (0 = register-path loader)
  int prefetch;                // halo boxes pulled into L2 this many items ahead (0 = off)
  int dbg;                     // timing experiments only (results invalid): 2 = no stores/stats, 4 = no MMA, 64 = no transform math
  long long* trace;
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

// stamps exist only in the kTrace instantiation (timing experiments, tools/conv_exp.py); the product kernel has none
#define SDD_TRACE3(role, iter, ev)                                                                    \
  do {                                                                                                \
    if constexpr (kTrace) {                                                                           \
      if (a.trace && blockIdx.x < 2 && (iter) < kTraceIters) s_trace[role][iter][ev] = clock64();     \
    }                                                                                                 \
  } while (0)

template <int COUT, bool kTrace = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC3Threads, 1)
conv3x3_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const ConvTc3Args a) {
  constexpr int kWSlot = (COUT / 2) * 128;  // bytes of one (tap, chunk) weight slice held by this CTA
  constexpr int kTmemCols = 2 * COUT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int kchunks = a.Cin / 64;
  const uint32_t w_bytes = 9u * kchunks * kWSlot;
  const uint32_t a_base = smem_base + w_bytes;
  const uint32_t bar_base = a_base + (uint32_t)a.stages * kHaloBytes;
  auto ready_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC3MaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kC3MaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kC3MaxStages + 2 + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * kC3MaxStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kC3MaxStages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  __shared__ long long s_trace[kTrace ? 5 : 1][kTrace ? kTraceIters : 1][4];
  if constexpr (kTrace) {
    if (a.trace && blockIdx.x < 2)
      for (int i = threadIdx.x; i < 5 * kTraceIters * 4; i += kC3Threads) (&s_trace[0][0][0])[i] = 0;
  }

  __shared__ volatile int s_progress_v;  // items the loaders have started (paces the L2 prefetcher, warp 3)
  volatile int* s_progress = &s_progress_v;
  if (threadIdx.x == 0) s_progress_v = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(ready_bar(s), 2 * kC3LoaderWarps); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 16); }
    mbar_init(w_bar, 1);
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

  // tile of this CTA in a pair-iteration; an odd tile count leaves one dummy (clamped, computed, not stored)
  auto tile_of = [&](int pair, bool& valid) {
    int t = 2 * pair + (int)rank;
    valid = t < a.num_tiles;
    return valid ? t : a.num_tiles - 1;
  };

  // Register re-partitioning by warpgroup: the kernel is launched with 96 registers per thread (61440 per CTA, and only
  // registers the CTA itself releases can be re-acquired): control warps 96 -> 40 (frees 7168), epilogue 96 -> 88
  // (frees 2048), and the two loader warpgroups (two 24-register load buffers + the transform) 96 -> 128 (takes 8192).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ===================== resident weights: this CTA's Cout/2 rows of all 9 taps, once =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, w_bytes);
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_3d(smem_base + (uint32_t)(tap * kchunks + kc) * kWSlot, &tmB, w_bar, kc * 64,
                      (int)rank * (COUT / 2), tap);
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
        if (lane == 0) SDD_TRACE3(1, it, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * COUT);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait_cluster(ready_bar(stage), phase);
          tc_fence_after();
          if (lane == 0) SDD_TRACE3(1, it, 1 + kc);
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
        if (lane == 0) SDD_TRACE3(1, it, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp == 3) {
    // ===================== L2 prefetcher: pulls the halo box of item (progress + prefetch) into L2 ==========
    // cp.async.bulk.prefetch.tensor moves no data into the SM and needs no barrier; the loaders' global loads then
    // hit L2.  Paced by the loaders' progress counter (a plain shared-memory word, polled with a short sleep).
    if (lane == 0 && a.prefetch > 0) {
      const int tiles_h = a.H / kTileH;
      const int step = 2 * pair_stride;
      const int d_n = step / a.tiles_per_sample, d_r = step - d_n * a.tiles_per_sample;
      const int d_th = d_r / a.tiles_w, d_tw = d_r - d_th * a.tiles_w;
      const int my_items = ((a.num_pairs - pair0 + pair_stride - 1) / pair_stride) * kchunks;
      bool valid;
      const int tile0 = tile_of(pair0, valid);
      int n = tile0 / a.tiles_per_sample;
      const int tr0 = tile0 - n * a.tiles_per_sample;
      int th = tr0 / a.tiles_w, tw = tr0 - th * a.tiles_w;
      int kc = 0, pair = pair0;
      for (int j = 0; j < my_items; ++j) {
        while (j >= *s_progress + a.prefetch) __nanosleep(800);
        tma_prefetch_l2_4d(&tmA, kc * 64, tw * kTileW - 1, th * kTileH - 1, n);
        if (++kc == kchunks) {
          kc = 0; pair += pair_stride;
          if (2 * pair + (int)rank < a.num_tiles) {
            tw += d_tw; th += d_th; n += d_n;
            if (tw >= a.tiles_w) { tw -= a.tiles_w; ++th; }
            if (th >= tiles_h) { th -= tiles_h; ++n; }
          }
        }
      }
    }
  }
  } else if (warp < 12) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    constexpr int COLS = COUT / 2;  // columns drained by this warp (two GroupNorm groups)
    constexpr int G = COLS / 16;    // 16-channel chunks per pixel and warp: 4 (Cout = 128) or 2 (Cout = 64)
    const int e = warp - 4, q = e & 3, hcol = e >> 2;
    const int col0 = hcol * COLS;
    const int m = q * 32 + lane;  // accumulator row = pixel within the tile
    const float* bias_row = a.bias.base + (a.bias.row_ptr ? (int64_t)(*a.bias.row_ptr) : 0) * a.bias.row_stride + col0;
    int acc = 0; uint32_t acc_phase = 0;
    int it = 0;
    for (int pair = pair0; pair < a.num_pairs; pair += pair_stride, ++it) {
      bool valid;
      const int tile = tile_of(pair, valid);
      const int n = tile / a.tiles_per_sample, tr = tile - n * a.tiles_per_sample;
      const int th = tr / a.tiles_w, tw = tr - th * a.tiles_w;
      const int h = th * kTileH + (m >> 3), w = tw * kTileW + (m & 7);
      const float* bp = bias_row + (int64_t)n * a.bias.batch_stride;
      __nv_bfloat16* orow = a.out + (((size_t)n * a.H + h) * a.W + w) * COUT + col0;
      const bool do_store = valid && !(a.dbg & 2);  // dbg 2: no stores (statistics stay), dbg 8: no statistics
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (e == 0 && lane == 0) SDD_TRACE3(3, it, 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * COUT + col0);
      float sg[2] = {0.f, 0.f}, ssg[2] = {0.f, 0.f};
      uint32_t v[2][16];
      uint32_t pk[G][8];  // this lane's pixel: G chunks of 16 channels (32 B each), packed bf16
      tmem_ld_32x16(taddr, v[0]);
#pragma unroll
      for (int st = 0; st < G; ++st) {
        float4 b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(bp + st * 16 + j * 4));
        tmem_ld_wait();
        if (st + 1 < G) tmem_ld_32x16(taddr + (uint32_t)((st + 1) * 16), v[(st + 1) & 1]);
        const uint32_t* vv = v[st & 1];
        float f[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[4 * j] = __uint_as_float(vv[4 * j]) + b4[j].x;         f[4 * j + 1] = __uint_as_float(vv[4 * j + 1]) + b4[j].y;
          f[4 * j + 2] = __uint_as_float(vv[4 * j + 2]) + b4[j].z; f[4 * j + 3] = __uint_as_float(vv[4 * j + 3]) + b4[j].w;
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
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[st][j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_remote(tempty_bar(acc), 0);  // orders TMEM reads only, not the stores
      if (do_store) {
        // G x G transpose of 32-byte chunks inside each group of G lanes (G consecutive pixels of one image row): lane j
        // of a group ends up with chunk j of all G pixels, so one store instruction writes G*32 contiguous bytes per
        // group -- 128-byte lines for Cout = 128 -- instead of 32 isolated sectors: 4x (2x) fewer L1 wavefronts.
        // Measured before the change: the stores cost 19 % of the launch through LSU contention with the loaders.
#pragma unroll
        for (int mbit = 1; mbit < G; mbit <<= 1) {
          const bool up = (lane & mbit) != 0;
#pragma unroll
          for (int c = 0; c < G; ++c) {
            if (c & mbit) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t send = up ? pk[c][j] : pk[c | mbit][j];
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, mbit);
              if (up) pk[c][j] = recv; else pk[c | mbit][j] = recv;
            }
          }
        }
        // pk[i] now holds chunk (lane % G) of pixel (lane - lane % G + i)
        __nv_bfloat16* obase = orow - (size_t)(lane & (G - 1)) * COUT + (lane & (G - 1)) * 16;
#pragma unroll
        for (int i = 0; i < G; ++i) st_global_v8(obase + (size_t)i * COUT, pk[i]);
      }
      if (e == 0 && lane == 0) SDD_TRACE3(3, it, 1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      // warp reduction of (sg0, ssg0, sg1, ssg1) in 6 shuffles: halve the value count while halving the lanes.
      // lane bit 4 selects the group it keeps, bit 3 the statistic; bits 2..0 are summed out.
      {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        const float keep_s = up16 ? sg[1] : sg[0], keep_ss = up16 ? ssg[1] : ssg[0];
        const float send_s = up16 ? sg[0] : sg[1], send_ss = up16 ? ssg[0] : ssg[1];
        const float s2 = keep_s + __shfl_xor_sync(0xffffffffu, send_s, 16);
        const float ss2 = keep_ss + __shfl_xor_sync(0xffffffffu, send_ss, 16);
        float val = (up8 ? ss2 : s2) + __shfl_xor_sync(0xffffffffu, up8 ? s2 : ss2, 8);
        val += __shfl_xor_sync(0xffffffffu, val, 4);
        val += __shfl_xor_sync(0xffffffffu, val, 2);
        val += __shfl_xor_sync(0xffffffffu, val, 1);
        if ((lane & 7) == 0 && valid && !(a.dbg & 8) && a.out_sums)
          gn_red_add(a.out_sums + ((size_t)n * 4 + hcol * 2 + (lane >> 4)) * 2 + ((lane >> 3) & 1), val);
      }
      if (e == 0 && lane == 0) SDD_TRACE3(3, it, 2);
    }
  } else {
    // ===================== loaders: global -> registers -> GroupNorm+SiLU -> swizzled shared memory ==========
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // Thread = (16-byte piece of the 64-channel chunk, halo column wr, group of 6 halo rows): its six vectors of an
    // item are one image row apart, so addresses and border masks need no tables.  240 of the 256 threads work.
    // Two register buffers: the loads of item j+1 are issued BEFORE item j is transformed, so they complete under
    // the transform and the MEMBAR inside fence.proxy.async (which waits for every outstanding load of the thread)
    // finds nothing left to wait for.
    const int tt = threadIdx.x - 384;  // 0..255
    __shared__ __align__(16) float s_ga[128], s_gb[128];
    const bool fuse = (a.in_sums != nullptr) || (a.in_meanrstd != nullptr);
    const int piece = tt & 7, col = tt >> 3;            // col 0..31
    const bool active = col < 3 * kHaloW;
    const int hg = active ? col / kHaloW : 0, wr = active ? col - hg * kHaloW : 0;
    const int hr0 = hg * 6;                             // halo rows hr0 .. hr0+5
    const long long row_bytes = (long long)a.W * a.Cin * 2;
    const uint8_t* in_bytes = reinterpret_cast<const uint8_t*>(a.in) + ((long long)wr * a.Cin + piece * 8) * 2;
    const int tiles_h = a.H / kTileH;

    // tile cursor advanced without divisions: step = 2 * pair_stride tiles
    struct Cursor { int n, th, tw; };
    const int step = 2 * pair_stride;
    const int d_n = step / a.tiles_per_sample, d_r = step - d_n * a.tiles_per_sample;
    const int d_th = d_r / a.tiles_w, d_tw = d_r - d_th * a.tiles_w;
    auto cursor_init = [&](int pair) {
      bool valid;
      const int tile = tile_of(pair, valid);
      Cursor c;
      c.n = tile / a.tiles_per_sample;
      const int tr = tile - c.n * a.tiles_per_sample;
      c.th = tr / a.tiles_w; c.tw = tr - c.th * a.tiles_w;
      return c;
    };
    auto cursor_next = [&](Cursor c, int next_pair) {  // coordinates of the tile `step` further, if it exists
      if (2 * next_pair + (int)rank >= a.num_tiles) return c;  // dummy tile of an odd count: any valid tile will do
      c.tw += d_tw; c.th += d_th; c.n += d_n;
      if (c.tw >= a.tiles_w) { c.tw -= a.tiles_w; ++c.th; }
      if (c.th >= tiles_h) { c.th -= tiles_h; ++c.n; }
      return c;
    };
    // base pointer of this thread's first vector and the mask of its in-image vectors
    auto item_src = [&](const Cursor& c, int kc, const uint8_t*& src, uint32_t& okmask) {
      const int h0 = c.th * kTileH - 1 + hr0, w0 = c.tw * kTileW - 1;
      src = in_bytes + (((long long)c.n * a.H + h0) * a.W + w0) * a.Cin * 2 + kc * 128;
      okmask = 0;
      const bool wok = active && (w0 + wr) >= 0 && (w0 + wr) < a.W;
#pragma unroll
      for (int i = 0; i < kC3Vecs; ++i)
        if (wok && (h0 + i) >= 0 && (h0 + i) < a.H) okmask |= 1u << i;
    };
    auto issue_loads = [&](uint4 (&r)[kC3Vecs], const uint8_t* src, uint32_t okmask) {
#pragma unroll
      for (int i = 0; i < kC3Vecs; ++i)
        r[i] = ((okmask >> i) & 1u) ? ldg_nc_v4(src + i * row_bytes) : make_uint4(0u, 0u, 0u, 0u);
    };

    const int num_my_pairs = (a.num_pairs - pair0 + pair_stride - 1) / pair_stride;
    const int my_items = num_my_pairs * kchunks;
    // one step of the item loop: `cur` holds item j's raw data, `nxt` receives item j+1's
    Cursor c = cursor_init(pair0);
    int pair = pair0, kc = 0, item = 0, it = 0;
    int stage = 0; uint32_t phase = 0;
    int cur_n = -1;
    bool w_ready = false;
    const uint8_t* src; uint32_t ok_c;
    uint4 ra[kC3Vecs], rb[kC3Vecs];
    if (my_items > 0) { item_src(c, 0, src, ok_c); issue_loads(ra, src, ok_c); }
    // look-ahead: coordinates, source pointer and mask of item j+1 are computed during item j-1's transform, so the
    // top of an item only has to issue six loads (the serial address arithmetic no longer sits between two items)
    auto advance = [&](const Cursor& cc, int kk, int pp, Cursor& c2, int& k2, int& p2) {
      c2 = cc; k2 = kk + 1; p2 = pp;
      if (k2 == kchunks) { k2 = 0; p2 = pp + pair_stride; c2 = cursor_next(cc, p2); }
    };
    Cursor c1; int kc1, pair1; const uint8_t* src1 = nullptr; uint32_t ok1 = 0;
    advance(c, kc, pair, c1, kc1, pair1);
    if (my_items > 1) item_src(c1, kc1, src1, ok1);

    auto process = [&](uint4 (&cur)[kC3Vecs], uint4 (&nxt)[kC3Vecs]) {
      // ---- item j+1: its loads go out first (addresses were prepared during the previous item)
      if (item + 1 < my_items) issue_loads(nxt, src1, ok1);
      // ---- GroupNorm scale / shift of this sample (pre-halved: silu(v) = h + h tanh(h), h = v/2)
      if (fuse && c.n != cur_n) {
        named_bar_sync(2, kC3LoaderThreads);  // previous readers of s_ga/s_gb are done
        if (tt < a.Cin) {
          const int g = tt / (a.Cin / 4);
          float mean, rstd;
          if (a.in_sums)
            gn_mean_rstd_from_sums(a.in_sums + ((size_t)c.n * 4 + g) * 2, (double)a.H * (double)a.W * (double)(a.Cin / 4),
                                   kGnEps, mean, rstd);
          else { mean = a.in_meanrstd[(c.n * 4 + g) * 2]; rstd = a.in_meanrstd[(c.n * 4 + g) * 2 + 1]; }
          const float sc = rstd * a.in_gamma[tt];
          s_ga[tt] = 0.5f * sc;
          s_gb[tt] = 0.5f * (a.in_beta[tt] - mean * sc);
        }
        named_bar_sync(2, kC3LoaderThreads);
        cur_n = c.n;
      }
      // scale / shift of this thread's 8 channels as packed fp32 pairs for FFMA2 (fma.rn.f32x2)
      uint64_t ga2[4], gb2[4];
      if (fuse) {
        const int coff = kc * 64 + piece * 8;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float4 x = *reinterpret_cast<const float4*>(&s_ga[coff + 2 * j]);
          const float4 y = *reinterpret_cast<const float4*>(&s_gb[coff + 2 * j]);
          ga2[j] = pack_f32x2(x.x, x.y); ga2[j + 1] = pack_f32x2(x.z, x.w);
          gb2[j] = pack_f32x2(y.x, y.y); gb2[j + 1] = pack_f32x2(y.z, y.w);
        }
      }
      // silu(gn(v)) of one bf16 pair: 2 unpack + FFMA2 + 2 MUFU.TANH + FFMA2 + 1 pack
      auto xform_pair = [&](uint32_t u, int j) -> uint32_t {
        const uint64_t h = fma_f32x2(pack_f32x2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)), ga2[j], gb2[j]);
        float hl, hh;
        unpack_f32x2(h, hl, hh);
        const uint64_t r = fma_f32x2(h, pack_f32x2(tanh_approx(hl), tanh_approx(hh)), h);
        float rl, rh;
        unpack_f32x2(r, rl, rh);
        return pack_bf16x2(rl, rh);
      };
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (tt == 0) { *s_progress = item + 1; SDD_TRACE3(2, it, kc); }
      // ---- item j+2: coordinates / pointer / mask, interleaved by the scheduler with the transform below
      Cursor c2; int kc2, pair2; const uint8_t* src2 = nullptr; uint32_t ok2 = 0;
      advance(c1, kc1, pair1, c2, kc2, pair2);
      if (item + 2 < my_items) item_src(c2, kc2, src2, ok2);
      const uint32_t dst = a_base + (uint32_t)stage * kHaloBytes;
      if (active) {
        if (fuse && ok_c == (1u << kC3Vecs) - 1u && !(a.dbg & 64)) {
          // interior tile (the common case): straight-line code, the six vectors' chains interleave freely
#pragma unroll
          for (int i = 0; i < kC3Vecs; ++i) {
            const int row = (hr0 + i) * kHaloW + wr;
            sts_v4(dst + (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4),
                   make_uint4(xform_pair(cur[i].x, 0), xform_pair(cur[i].y, 1), xform_pair(cur[i].z, 2), xform_pair(cur[i].w, 3)));
          }
        } else {
#pragma unroll
          for (int i = 0; i < kC3Vecs; ++i) {
            uint4 v = cur[i];
            if (fuse && ((ok_c >> i) & 1u) && !(a.dbg & 64))
              v = make_uint4(xform_pair(v.x, 0), xform_pair(v.y, 1), xform_pair(v.z, 2), xform_pair(v.w, 3));
            const int row = (hr0 + i) * kHaloW + wr;  // padding pixels hold the zeros they were "loaded" as
            sts_v4(dst + (uint32_t)row * 128u + (uint32_t)((piece ^ (row & 7)) << 4), v);
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      if (!w_ready) { mbar_wait(w_bar, 0); w_ready = true; }  // this CTA's weights have landed too
      __syncwarp();
      // relaxed: the proxy fence above has already completed this thread's shared-memory writes (MEMBAR.ALL.CTA) and
      // made them visible to the async proxy of THIS CTA's tensor core, which is the only reader; a release at
      // cluster scope would additionally wait ~900 cycles on the in-flight global loads of the next item
      if (lane == 0) mbar_arrive_relaxed_remote(ready_bar(stage), 0);
      if (tt == 0) SDD_TRACE3(2, it, 2 + kc);
      if (lane == 0 && kc == kchunks - 1) SDD_TRACE3((warp < 16 ? 0 : 4), it, (warp & 3));  // every loader warp's arrive
      if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      ++item;
      if (kc1 == 0) ++it;
      c = c1; kc = kc1; pair = pair1; ok_c = (item < my_items) ? ok1 : 0u;
      c1 = c2; kc1 = kc2; pair1 = pair2; src1 = src2; ok1 = ok2;
    };
    while (item < my_items) {
      process(ra, rb);
      if (item >= my_items) break;
      process(rb, ra);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kTrace) {
    if (a.trace && blockIdx.x < 2)
      for (int i = threadIdx.x; i < 5 * kTraceIters * 4; i += kC3Threads) {
        const int role = i / (kTraceIters * 4), rem = i % (kTraceIters * 4);
        a.trace[(((size_t)blockIdx.x * 6 + role) * 64 + rem / 4) * 4 + (rem & 3)] = (&s_trace[0][0][0])[i];
      }
  }
  cluster_sync_all();  // the peer may still be reading our smem / arriving on our barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// shared memory needed for (COUT, Cin) with `stages` halo stages
inline int conv_tc3_smem_bytes(int Cout, int Cin, int stages) {
  return 9 * (Cin / 64) * (Cout / 2) * 128 + stages * kHaloBytes + 1024 + 256;
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
// in the 128->128 ncu capture).
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
