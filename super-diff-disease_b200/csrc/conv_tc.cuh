// 3x3 / pad-1 convolution, Cin, Cout in {64,128}, as an implicit GEMM on the 5th-gen tensor cores.
//
//   out[n,h,w,co] = bias[n,co] + sum_{ky,kx,ci} act[n,h+ky-1,w+kx-1,ci] * wt[kx][ky][co][ci]
//
// GEMM view per output tile: M = 128 pixels (16 rows x 8 cols of the image), N = Cout, K = 9*Cin.
//   A (pixels x ci, K-major, 128-byte rows, SWIZZLE_128B): for each kx one TMA box of
//     (64 ci, 8 w, 18 h) lands as 144 rows of 128 B.  Because a tile row is exactly 8 pixels = one
//     1024-byte swizzle atom, the three ky taps are the SAME smem box viewed at +0/+1/+2 atoms:
//     3 loads feed 9 taps, and image-border padding is TMA out-of-bounds zero fill.
//   B (co x ci, K-major): the three ky taps of one kx are contiguous in wt -> one TMA box (64, Cout, 3).
//   D: fp32 accumulators in TMEM, double buffered (2 x Cout columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1.
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA
// issuer (one lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> regs -> +bias ->
// GroupNorm partial sums -> bf16 NHWC store).
#pragma once
#include "common.cuh"
#include "unet_kernels.cuh"

namespace sdd {

constexpr int kTileH = 16, kTileW = 8;            // 128 output pixels per tile
constexpr int kHaloRows = (kTileH + 2) * kTileW;  // 144 smem rows per kx load
constexpr int kABytes = kHaloRows * 128;          // 18432
constexpr int kConvThreads = 256;

// HALO experiment: one (64 ci, 10 w, 18 h) box per stage; the 9 taps are descriptor views starting at
// row (ky*10 + kx) with a 1280-byte stride between 8-row groups (start not 1024-aligned, SBO not a multiple
// of 1024) -- tests whether UMMA applies the 128B swizzle on absolute smem address bits.
constexpr int kHaloW = kTileW + 2;
constexpr int kHaloBytes = ((kTileH + 2) * kHaloW * 128 + 1023) / 1024 * 1024;  // 23552

template <int COUT, bool HALO = false>
struct ConvCfg {
  static constexpr int kBBytes = 3 * COUT * 128;
  static constexpr int kAStage = HALO ? kHaloBytes : kABytes;
  static constexpr int kStageBytes = kAStage + kBBytes;
  static constexpr int kTxBytes = (HALO ? (kTileH + 2) * kHaloW * 128 : kABytes) + kBBytes;
  static constexpr int kStages = (COUT == 64) ? (HALO ? 4 : 5) : 3;
  static constexpr int kTmemCols = 2 * COUT;  // 128 or 256: power of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct ConvTcArgs {
  __nv_bfloat16* out;
  BiasRef bias;
  float* partials;  // [B][tiles_per_sample][4][2]
  int* counters;    // [B]
  float* meanrstd;  // [B][4][2]
  int B, H, W, Cin;
  int tiles_w, tiles_per_sample, num_tiles;
};

template <int COUT, bool HALO = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvTcArgs a) {
  using Cfg = ConvCfg<COUT, HALO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = a.Cin / 64;
  const int kiters = 3 * kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int n = tile / a.tiles_per_sample;
        const int tr = tile % a.tiles_per_sample;
        const int h0 = (tr / a.tiles_w) * kTileH, w0 = (tr % a.tiles_w) * kTileW;
        for (int kx = 0; kx < 3; ++kx)
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kTxBytes);
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            tma_load_4d(sa, &tmA, full_bar(stage), kc * 64, HALO ? w0 - 1 : w0 + kx - 1, h0 - 1, n);
            tma_load_3d(sa + Cfg::kAStage, &tmB, full_bar(stage), kc * 64, 0, kx * 3);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, COUT);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * COUT);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kAStage;
          const int kx = it / kchunks;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t adesc = HALO ? umma_desc_sw128(sa + (ky * kHaloW + kx) * 128, kHaloW * 128)
                                        : umma_desc_sw128(sa + ky * (kTileW * 128));
            const uint64_t bdesc = umma_desc_sw128(sb + ky * (COUT * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x UMMA_K(16 bf16 = 32 B) inside the 128-byte swizzle row
              umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                        (it | ky | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;           // TMEM lane quadrant (== warp % 4)
    const int m = ew * 32 + lane;      // accumulator row = pixel within the tile
    __shared__ float s_red[4][8];
    __shared__ float s_sums[8];
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int n = tile / a.tiles_per_sample;
      const int tr = tile % a.tiles_per_sample;
      const int h = (tr / a.tiles_w) * kTileH + (m >> 3), w = (tr % a.tiles_w) * kTileW + (m & 7);
      const float* bp = bias_ptr(a.bias, n);
      __nv_bfloat16* orow = a.out + (((size_t)n * a.H + h) * a.W + w) * COUT;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      float gs[4] = {0.f, 0.f, 0.f, 0.f}, gss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * COUT + c0), v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __ldg(bp + c0 + j);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int g = (c0 + j) / (COUT / 4);
          gs[g] += f[j];
          gss[g] = fmaf(f[j], f[j], gss[g]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(f[j], f[j + 1]); pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
          pk.z = pack_bf16x2(f[j + 4], f[j + 5]); pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(orow + c0 + j) = pk;
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }

      // GroupNorm partial sums of this tile (fixed order: lanes by shuffle, then warps 0..3)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], o);
          gss[g] += __shfl_xor_sync(0xffffffffu, gss[g], o);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { s_red[ew][2 * g] = gs[g]; s_red[ew][2 * g + 1] = gss[g]; }
      }
      named_bar_sync(1, 128);
      if (ew == 0) {
        if (lane < 8) s_sums[lane] = (s_red[0][lane] + s_red[1][lane]) + (s_red[2][lane] + s_red[3][lane]);
        __syncwarp();
        gn_publish_and_finalize_warp(s_sums, a.partials, a.counters, a.meanrstd, n, tr, a.tiles_per_sample, 4,
                                     (float)a.H * (float)a.W * (float)(COUT / 4), kGnEps);
      }
      named_bar_sync(1, 128);  // s_red / s_sums reusable
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace sdd
