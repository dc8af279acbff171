// Product tensor-core 3x3 conv: GroupNorm+SiLU -> conv3x3 -> (+bias, next GroupNorm's sums), one kernel.
//
//   out[n,h,w,co] = bias[n,co] + sum_{ky,kx,ci} f(raw[n,h+ky-1,w+kx-1,ci]) * wt[kx][ky][co][ci]
//   f(v) = silu((v - mean[n,g]) * rstd[n,g] * gamma[ci] + beta[ci])  inside the image, 0 in the padding
//   (reference: GroupNorm -> SiLU -> Conv2d of ResidualBlock, /root/reference/src/models/unet.py:21-28)
//
// CTA pair (cluster of 2, tcgen05 cta_group::2), UMMA M = 256 = two 16x8-pixel tiles, N = Cout, K = 9*Cin, fp16 x fp16
// -> fp32 in TMEM (double-buffered accumulators); this CTA's half of the weights resident in shared memory for the whole
// persistent kernel; the nine taps are descriptor VIEWS of one (18 x 10)-pixel halo box (start row ky*10+kx, 1280-byte
// group stride under SWIZZLE_128B).  Shared memory, not the tensor pipe, is the scarce resource (an SS-mode
// M256xN128xK16 UMMA reads 6 KB per CTA in 64 cycles), so the operand tile is touched ONCE: loader warps read the raw
// activations into registers, apply GroupNorm+SiLU there and store the activated tile straight into the UMMA swizzle.
// The output's GroupNorm sums leave the epilogue as fire-and-forget 64-bit fixed-point RED.ADDs (common.cuh).
// Barriers (arrival count): ready[s] loaders->MMA (8, on the leader) | empty[s] MMA->loaders (1, multicast commit)
// | tfull[a] MMA->epilogue (1, multicast commit) | tempty[a] epilogue->MMA (16, on the leader) | wbar weights (1+tx).
// The loader warps form TWO groups that work on alternating 64-channel items (see the loader section).
#pragma once
#include "conv_common.cuh"

namespace sdd {

// -DSDD_CONV_PROF (tools/conv_prof.py, never the product build): per-role cycle accounting of CTA 0 -- where the MMA
// warp, one loader warp per group and one epilogue warp spend their clocks per item.
#ifdef SDD_CONV_PROF
__device__ unsigned long long g_conv_prof[64];
#define SDD_PROF_DECL(n) long long pf[n] = {}; long long pf_t = clock64()
#define SDD_PROF_LAP(i) { const long long t_ = clock64(); pf[i] += t_ - pf_t; pf_t = t_; }
#define SDD_PROF_FLUSH(cond, base, n) if (cond) { for (int i_ = 0; i_ < (n); ++i_) atomicAdd(&g_conv_prof[(base) + i_], (unsigned long long)pf[i_]); }
#define SDD_PROF_SINK(v) asm volatile("" ::"r"(v))
#else
#define SDD_PROF_DECL(n)
#define SDD_PROF_LAP(i)
#define SDD_PROF_FLUSH(cond, base, n)
#define SDD_PROF_SINK(v)
#endif

// kRaw (generation 5, layers whose resident weights leave >= 3 spare 23 KB slots: every layer but 128->128): the
// halo boxes are not fetched by the loader threads' own global loads but by TMA into a ring of RAW shared-memory slots,
// issued by the otherwise idle warp 3 up to `raw_slots` items ahead.  Measured on v4: every layer spends 3800-3900
// cycles per 64-channel item because a loader group can keep only ONE item of loads in flight (its single 48-register
// buffer; the MEMBAR.ALL.CTA that ptxas puts in front of fence.proxy.async waits for every outstanding load of the
// thread, so loads cannot be left in flight across a hand-off) -- ~1.2 items = 24-48 KB per SM at 1400-4400 cycles
// of latency.  With the ring 70-90 KB per SM are in flight without any registers, the loaders read their vectors with
// conflict-free LDS.128 from the same swizzled offsets they write to (the TMA box lands in the UMMA layout), no thread
// has a global load outstanding at the fence, and address arithmetic leaves the loaders entirely.  Cost: one more
// shared-memory write + read per byte (23 KB + 23 KB per item), affordable where the MMA leaves bandwidth.
// CIN is a template parameter: with the chunk count a compile-time constant every weight-slice / tap offset below is an
// immediate added to ONE base descriptor (the 40-register control warps spill anything the compiler hoists), the loaders'
// address arithmetic folds, and all chunk loops unroll.
// kHalfMath: the fused GroupNorm+SiLU transform on packed fp16 pairs (HFMA2, MUFU.TANH.F16 x2, HFMA2: 5 instructions per
// pair, no conversions) instead of fp32 (7 per pair); costs ~1e-3 of forward rel-L2 (tools/error_budget.py).
template <int COUT, int CIN, bool kRaw = false, bool kHalfMath = SDD_CONV_HALF_MATH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC3Threads, 1)
conv3x3_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const ConvTc3Args a) {
  constexpr int kWSlot = (COUT / 2) * 128;  // bytes of one (tap, chunk) weight slice held by this CTA
  // FOUR fp32 accumulators in TMEM (4 x COUT columns: 256 or all 512; the kernel owns its SM): with two, the MMA warp of
  // the 128->128 launch spent 11 % of its time waiting for the slower epilogue of the pair to release one (round-2 ncu
  // capture) although the epilogue warps themselves idle 27 % -- hand-off jitter that two more tiles of slack absorb.
  constexpr int kAccs = SDD_CONV_ACCS;
  constexpr int kTmemCols = kAccs * COUT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int kchunks = CIN / 64;
  constexpr uint32_t w_bytes = 9u * kchunks * kWSlot;
  const uint32_t a_base = smem_base + w_bytes;
  const uint32_t raw_base = a_base + (uint32_t)a.stages * kHaloBytes;
  const uint32_t bar_base = raw_base + (uint32_t)(kRaw ? a.raw_slots : 0) * kHaloBytes;
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * kC3MaxStages + 10 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (3 * kC3MaxStages + 10 + s); };
  auto ready_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC3MaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kC3MaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kC3MaxStages + 4 + s); };
  const uint32_t w_bar = bar_base + 8u * (2 * kC3MaxStages + 8);
  // CTA-lifetime values the 40-register control warps and the 88-register epilogue need inside their loops live in STATIC
  // shared memory (address = an immediate, no register to keep alive): [0] TMEM base (written by tcgen05.alloc),
  // [1] operand stages, [2] resident weights, [3] barriers, [4..6] this CTA pair's tile range.  Kept in registers,
  // ptxas spilled them and re-loaded them from LOCAL memory in front of every MMA batch and at every loop test (~240
  // cycles per reload with 213 KB of the L1 carved out as shared memory: 12 % of the MMA warp's stall samples in the 64->64
  // capture, 13 % of the epilogue's; profiles/r2_stalls_conv.md).
  __shared__ uint32_t s_cfg[8];
  const uint32_t tmem_slot = smem_u32(&s_cfg[0]);
#define SDD_CFG(i) (*reinterpret_cast<volatile uint32_t*>(&s_cfg[i]))

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // Tile pairs of this CTA pair: a CONTIGUOUS range [pair0, pair_end) (kContig) or every (gridDim.x / 2)-th
  // pair.  Contiguous ranges keep a loader group inside one sample for hundreds of items, so the per-sample GroupNorm
  // scale / shift (two named barriers, double-precision statistics, an LDS round trip: 15 % of the loaders' samples in
  // the 64->64 ncu capture, where the strided order changed sample every ~1.7 items) is rebuilt almost never.
  const int npp = gridDim.x >> 1, pidx = blockIdx.x >> 1;
  // (compile-time: only 64->64 gains from contiguous ranges; with the strided order pair_end is the
  // kernel parameter itself, which lives in the constant bank -- as a computed value ptxas spilled it and the epilogue's
  // loop test waited on the local-memory reload every tile: 17 % of its stall samples in the round-2 128->128 capture)
  constexpr bool kContig = (CIN == 64 && COUT == 64);
  const int pair0 = kContig ? (int)(((long long)pidx * a.num_pairs) / npp) : pidx;
  const int pair_end = kContig ? (int)(((long long)(pidx + 1) * a.num_pairs) / npp) : a.num_pairs;
  const int pair_stride = kContig ? 1 : npp;

  if (threadIdx.x == 0) {
    // [5] pair iterations of this CTA pair, [6] how many of them are valid tiles for THIS CTA (an odd tile count leaves
    // one dummy in the last pair): the MMA and epilogue loops count iterations and derive accumulator index / phase
    // from the counter, so that ONE loop-carried control register is live across their bodies
    const int n_it = pair_end > pair0 ? (pair_end - pair0 + pair_stride - 1) / pair_stride : 0;
    const int last_pair = pair0 + (n_it - 1) * pair_stride;
    s_cfg[1] = a_base; s_cfg[2] = smem_base; s_cfg[3] = bar_base; s_cfg[4] = (uint32_t)pair_end;
    s_cfg[5] = (uint32_t)n_it;
    s_cfg[6] = (uint32_t)((n_it > 0 && 2 * last_pair + (int)rank >= a.num_tiles) ? n_it - 1 : n_it);
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    // ready: one loader group (4 warps) per CTA and item -> 8 arrivals
    for (int s = 0; s < a.stages; ++s) { mbar_init(ready_bar(s), kC3LoaderWarps); mbar_init(empty_bar(s), 1); }
    // tempty: every epilogue warp that drains the accumulator, in both CTAs (Cout = 64: four warps per tile, see there)
    for (int s = 0; s < kAccs; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), COUT == 64 ? 8 : 16); }
    mbar_init(w_bar, 1);
    if constexpr (kRaw)
      for (int s = 0; s < a.raw_slots; ++s) { mbar_init(raw_full_bar(s), 1); mbar_init(raw_empty_bar(s), 4); }
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
  // CTA-lifetime scalars are NOT kept in registers across the role loops: at 88 (epilogue) / 40 (issuer) registers
  // ptxas spilled them and reloaded them every tile with LDL, and with 213 KB of shared memory there is no L1 left to
  // hit in (~17 % of the epilogue's stall samples in the 128->128 ncu capture, and a ~300-cycle bubble per tile in
  // front of the MMA issue).  The TMEM base is re-read from its shared-memory slot (volatile: one LDS per use) and the
  // pair stride is re-derived from %nctaid behind a volatile asm, which ptxas cannot hoist.
#define SDD_TMEM_BASE() SDD_CFG(0)
  auto pair_step = [&]() -> int {
    uint32_t g;
    asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(g));
    return kContig ? 1 : (int)(g >> 1);
  };

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
      constexpr uint32_t idesc = SDD_ACT_IDESC(256, COUT);
      int stage = 0; uint32_t phase = 0;
      static_assert((kAccs & (kAccs - 1)) == 0, "accumulator count must be a power of two");
      SDD_PROF_DECL(4);
      for (int it = 0; it < (int)SDD_CFG(5); ++it) {
        const int acc = it & (kAccs - 1);
        const uint32_t acc_phase = (uint32_t)(it / kAccs) & 1u;
        SDD_PROF_LAP(3);
        const uint32_t bars = SDD_CFG(3);  // == bar_base (the lambdas above would keep it alive across the loop)
        mbar_wait_cluster(bars + 8u * (2 * kC3MaxStages + 4 + acc), acc_phase ^ 1u);  // tempty_bar(acc)
        SDD_PROF_LAP(0);
        tc_fence_after();
        const uint32_t d_tmem = SDD_TMEM_BASE() + (uint32_t)(acc * COUT);
#pragma unroll
        for (int kc = 0; kc < kchunks; ++kc) {
          SDD_PROF_LAP(3);
          mbar_wait_cluster(bars + 8u * stage, phase);  // ready_bar(stage)
          SDD_PROF_LAP(1);
          tc_fence_after();
          if (elect_one_sync()) {
            // one base descriptor per operand; taps, chunks and K steps are immediates in 16-byte units (shared-memory
            // addresses are < 2^18, so the 14-bit start-address field never carries)
            // (`opaque` is a zero ptxas cannot see through: without it the 36 x 2 loop-invariant B descriptors are hoisted
            // out of the tile loop and, at 40 registers, spilled and re-loaded from local memory in front of every MMA)
            uint32_t opaque;
            asm volatile("mov.u32 %0, 0;" : "=r"(opaque));
            const uint64_t adesc0 = umma_desc_sw128(SDD_CFG(1) + stage * kHaloBytes, kHaloW * 128);
            const uint64_t bdesc0 = umma_desc_sw128(SDD_CFG(2) + opaque);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x UMMA_K (16 fp16 = 32 B) inside the 128-byte swizzle row
                  umma_f16_2cta(d_tmem, adesc0 + (uint64_t)((ky * kHaloW + kx) * 8 + k * 2),
                                bdesc0 + (uint64_t)((((kx * 3 + ky) * kchunks + kc) * kWSlot) / 16 + k * 2), idesc,
                                (kc | kx | ky | k) ? 1u : 0u);
              }
            umma_commit_2cta(bars + 8u * (kC3MaxStages + stage));       // empty_bar(stage): frees the stage in both CTAs
            if (kc == kchunks - 1) umma_commit_2cta(bars + 8u * (2 * kC3MaxStages + acc));  // tfull_bar(acc) -> both epilogues
          }
          __syncwarp();
          SDD_PROF_LAP(2);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
      SDD_PROF_FLUSH(blockIdx.x == 0 && lane == 0, 0, 4);
    }
  } else if (warp == 3) {
    // ===================== kRaw: TMA producer of the raw halo boxes, `raw_slots` items ahead =====================
    if constexpr (kRaw) {
      if (lane == 0) {
        const int my_items = ((pair_end - pair0 + pair_stride - 1) / pair_stride) * kchunks;
        for (int j = 0; j < my_items; ++j) {
          const int slot = j % a.raw_slots;
          if (j >= a.raw_slots) mbar_wait(raw_empty_bar(slot), (uint32_t)(((j / a.raw_slots) - 1) & 1));
          bool valid;
          const int tile = tile_of(pair0 + (j / kchunks) * pair_stride, valid);
          const int n = tile / a.tiles_per_sample, tr = tile - n * a.tiles_per_sample;
          const int th = tr / a.tiles_w, tw = tr - th * a.tiles_w;
          mbar_arrive_expect_tx(raw_full_bar(slot), (uint32_t)(kHaloRowsV2 * 128));
          tma_load_4d(raw_base + (uint32_t)slot * kHaloBytes, &tmA, raw_full_bar(slot), (j % kchunks) * 64,
                      tw * kTileW - 1, th * kTileH - 1, n);
        }
      }
    }
  }
  } else if (warp < 12) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quadrants x 2 =====================
    // Cout = 128: x 2 column halves of every tile.  Cout = 64 (kSplit): x 2 TILE PARITIES -- a warp drains all 64 columns of
    // its quadrant, of every other tile.  Per warp and tile the work is then the same 64 columns in both cases; what changes
    // for Cout = 64 is that a warp has TWO tile periods (2 x 1152..1440 MMA cycles) for its dependent chain -- accumulator
    // wake-up, two-deep tcgen05.ld pipeline, statistics shuffles, stores -- instead of one, with the chain's fixed latencies
    // paid once per 64 columns instead of once per 32.  (The instruction count per tile is unchanged: measured, 586 per
    // warp and 64 columns against 2 x 290.)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    constexpr bool kSplit = COUT == 64;
    constexpr int COLS = 64;        // columns drained by this warp
    constexpr int G = COLS / 16;    // 16-channel chunks per pixel and warp
    constexpr int kGroups = kSplit ? 4 : 2;  // GroupNorm groups inside the warp's columns (16 / 32 channels each)
    const int e = warp - 4, q = e & 3, hcol = kSplit ? 0 : (e >> 2);
    const int par = kSplit ? (e >> 2) : 0, it_step = kSplit ? 2 : 1;
    const int col0 = hcol * COLS;
    const int m = q * 32 + lane;  // accumulator row = pixel within the tile
    const float* bias_row = a.bias.base + (a.bias.row_ptr ? (int64_t)(*a.bias.row_ptr) : 0) * a.bias.row_stride + col0;
    __shared__ __align__(16) float s_bias[8][64];
    int bias_cur = -1;
    // division-free tile cursor (the two runtime-divisor divisions per tile were ~55 of the epilogue's ~450 instructions
    // per warp and tile): (n, th, tw) of this CTA's tile advance by a constant tile step with carries
    int n, th, tw;
    {
      bool v0;
      const int tile0 = tile_of(pair0 + par * pair_stride, v0);
      n = tile0 / a.tiles_per_sample;
      const int tr = tile0 - n * a.tiles_per_sample;
      th = tr / a.tiles_w; tw = tr - th * a.tiles_w;
    }
    const int tiles_h_e = a.H / kTileH;
    const int tstep = 2 * it_step * pair_step();
    const int e_dn = tstep / a.tiles_per_sample, e_dr = tstep - e_dn * a.tiles_per_sample;
    const int e_dth = e_dr / a.tiles_w, e_dtw = e_dr - e_dth * a.tiles_w;
    SDD_PROF_DECL(4);
    for (int it = par; it < (int)SDD_CFG(5); it += it_step) {
      SDD_PROF_LAP(3);
      const int acc = it & (kAccs - 1);
      const uint32_t acc_phase = (uint32_t)(it / kAccs) & 1u;
      const bool valid = it < (int)SDD_CFG(6);  // false: the dummy tile of an odd count (coordinates stay at the
                                                // previous, valid tile; nothing is stored)
      const int h = th * kTileH + (m >> 3), w = tw * kTileW + (m & 7);
      const float* bp = bias_row + (int64_t)n * a.bias.batch_stride;
      act_t* orow = a.out + (((size_t)n * a.H + h) * a.W + w) * COUT + col0;
      const bool do_store = valid;
      // This warp's COLS bias values live in shared memory and are refreshed only when the sample changes (never, when
      // all samples share one row).  As per-tile global loads they were the first use behind the accumulator wait and
      // cost 10 % of the epilogue's stall samples: with 213 KB of shared memory there is almost no L1 left to hit in.
      const int bias_key = a.bias.batch_stride ? n : 0;
      if (bias_key != bias_cur) {
        __syncwarp();
        for (int j = lane; j < COLS; j += 32) s_bias[e][j] = __ldg(bp + j);
        __syncwarp();
        bias_cur = bias_key;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      SDD_PROF_LAP(0);
      tc_fence_after();
      const uint32_t taddr = SDD_TMEM_BASE() + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * COUT + col0);
      float sg[kGroups], ssg[kGroups];
#pragma unroll
      for (int g = 0; g < kGroups; ++g) { sg[g] = 0.f; ssg[g] = 0.f; }
      uint32_t v[2][16];
      uint32_t pk[G][8];  // this lane's pixel: G chunks of 16 channels (32 B each), packed fp16
      tmem_ld_32x16(taddr, v[0]);
#pragma unroll
      for (int st = 0; st < G; ++st) {
        float4 b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b4[j] = *reinterpret_cast<const float4*>(&s_bias[e][st * 16 + j * 4]);
        tmem_ld_wait();
        if (st + 1 < G) tmem_ld_32x16(taddr + (uint32_t)((st + 1) * 16), v[(st + 1) & 1]);
        const uint32_t* vv = v[st & 1];
        // packed fp32 pairs (FADD2 / FFMA2): bias add, sum and sum of squares at two columns per instruction -- the
        // epilogue is bound by its own instruction stream on the 64-channel layers
        uint64_t f2[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f2[2 * j] = add_f32x2(pack_f32x2(__uint_as_float(vv[4 * j]), __uint_as_float(vv[4 * j + 1])),
                                pack_f32x2(b4[j].x, b4[j].y));
          f2[2 * j + 1] = add_f32x2(pack_f32x2(__uint_as_float(vv[4 * j + 2]), __uint_as_float(vv[4 * j + 3])),
                                    pack_f32x2(b4[j].z, b4[j].w));
        }
        // GroupNorm group of this 16-column chunk (Cout = 128: 32-channel groups, Cout = 64: 16-channel groups); two
        // independent pair accumulators per statistic
        const int g = kSplit ? st : (st >> 1);
        uint64_t p01 = f2[0], p23 = f2[1];
        uint64_t r01 = fma_f32x2(f2[0], f2[0], 0ull), r23 = fma_f32x2(f2[1], f2[1], 0ull);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
          p01 = add_f32x2(p01, f2[j]); p23 = add_f32x2(p23, f2[j + 1]);
          r01 = fma_f32x2(f2[j], f2[j], r01); r23 = fma_f32x2(f2[j + 1], f2[j + 1], r23);
        }
        {
          float a0, a1, a2, a3;
          unpack_f32x2(add_f32x2(p01, p23), a0, a1);
          unpack_f32x2(add_f32x2(r01, r23), a2, a3);
          sg[g] += a0 + a1;
          ssg[g] += a2 + a3;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float lo, hi;
          unpack_f32x2(f2[j], lo, hi);
          pk[st][j] = pack_act2(lo, hi);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed_remote(tempty_bar(acc), 0);  // orders TMEM reads only, not the stores
      SDD_PROF_LAP(1);
      // (The statistics come BEFORE the stores: behind them, their first instruction had to wait for the store unit to read
      // the 256-bit stores' data registers it reuses -- a write-after-read stall worth 9 % of the epilogue's samples.)
      // warp reduction of (sg0, ssg0, sg1, ssg1) in 6 shuffles: halve the value count while halving the lanes.
      // lane bit 4 selects the group it keeps, bit 3 the statistic; bits 2..0 are summed out.
      if constexpr (!kSplit) {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        const float keep_s = up16 ? sg[1] : sg[0], keep_ss = up16 ? ssg[1] : ssg[0];
        const float send_s = up16 ? sg[0] : sg[1], send_ss = up16 ? ssg[0] : ssg[1];
        const float s2 = keep_s + __shfl_xor_sync(0xffffffffu, send_s, 16);
        const float ss2 = keep_ss + __shfl_xor_sync(0xffffffffu, send_ss, 16);
        float val = (up8 ? ss2 : s2) + __shfl_xor_sync(0xffffffffu, up8 ? s2 : ss2, 8);
        val += __shfl_xor_sync(0xffffffffu, val, 4);
        val += __shfl_xor_sync(0xffffffffu, val, 2);
        val += __shfl_xor_sync(0xffffffffu, val, 1);
        if ((lane & 7) == 0 && valid && a.out_sums)
          gn_red_add(a.out_sums + ((size_t)n * 4 + hcol * 2 + (lane >> 4)) * 2 + ((lane >> 3) & 1), val);
      } else {
        // four groups x (sum, sum of squares) in 9 shuffles: lane bit 4 selects the group PAIR it keeps, bit 3 the group of
        // the pair, bit 2 the statistic; bits 1..0 are summed out
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
        float k_s[2], k_ss[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float keep_s = up16 ? sg[2 + j] : sg[j], send_s = up16 ? sg[j] : sg[2 + j];
          const float keep_ss = up16 ? ssg[2 + j] : ssg[j], send_ss = up16 ? ssg[j] : ssg[2 + j];
          k_s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, 16);
          k_ss[j] = keep_ss + __shfl_xor_sync(0xffffffffu, send_ss, 16);
        }
        const float s2 = (up8 ? k_s[1] : k_s[0]) + __shfl_xor_sync(0xffffffffu, up8 ? k_s[0] : k_s[1], 8);
        const float ss2 = (up8 ? k_ss[1] : k_ss[0]) + __shfl_xor_sync(0xffffffffu, up8 ? k_ss[0] : k_ss[1], 8);
        float val = (up4 ? ss2 : s2) + __shfl_xor_sync(0xffffffffu, up4 ? s2 : ss2, 4);
        val += __shfl_xor_sync(0xffffffffu, val, 2);
        val += __shfl_xor_sync(0xffffffffu, val, 1);
        if ((lane & 3) == 0 && valid && a.out_sums)
          gn_red_add(a.out_sums + ((size_t)n * 4 + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)) * 2 + ((lane >> 2) & 1), val);
      }
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
        act_t* obase = orow - (size_t)(lane & (G - 1)) * COUT + (lane & (G - 1)) * 16;
#pragma unroll
        for (int i = 0; i < G; ++i) st_global_v8(obase + (size_t)i * COUT, pk[i]);
      }
      SDD_PROF_LAP(2);
      if (it + it_step < (int)SDD_CFG(6)) {  // advance to this warp's next (valid) tile
        tw += e_dtw; th += e_dth; n += e_dn;
        if (tw >= a.tiles_w) { tw -= a.tiles_w; ++th; }
        if (th >= tiles_h_e) { th -= tiles_h_e; ++n; }
      }
    }
    SDD_PROF_FLUSH(blockIdx.x == 0 && warp == 4 && lane == 0, 8, 4);
  } else {
    // ===================== loaders: global -> registers -> GroupNorm+SiLU -> swizzled shared memory ==========
    // TWO groups of four warps work on ALTERNATING items (group g: items g, g+2, ...).  Measured on the single-group
    // version: ~2500 cycles of serial control code per item plus up to ~4000 cycles of exposed load latency, against
    // 2304 cycles of MMA -- so one group's chain is allowed to take two item times while the other group feeds the
    // tensor core.  A thread owns twelve 16-byte vectors of its item (halo rows col + 16 i) in ONE register buffer:
    // the loads of the group's next item are issued right after its arrive, i.e. after the MEMBAR inside
    // fence.proxy.async (which would otherwise wait for them), and have the other group's whole item to land.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    constexpr int kVecs = 12;
    const int tt = threadIdx.x - 384;        // 0..255
    const int grp = tt >> 7, tg = tt & 127;  // group, thread in group
    const int piece = tg & 7, col = tg >> 3; // col 0..15
    const bool fuse = a.in_ab != nullptr;
    const uint8_t* in_bytes = reinterpret_cast<const uint8_t*>(a.in);
    const int tiles_h = a.H / kTileH;
    // vector i <-> halo row r = col + 16 i (pixel (r / 10, r % 10) of the 18 x 10 box); r & 7 == col & 7 for every i
    const int nvec = (col < kHaloRowsV2 - 16 * (kVecs - 1)) ? kVecs : kVecs - 1;  // rows >= 180 do not exist
    uint32_t goff[kVecs];  // byte offset of the vector from the box origin (chunk 0)
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int r = col + 16 * i, hr = r / kHaloW, wr = r - hr * kHaloW;
      goff[i] = (uint32_t)((hr * a.W + wr) * CIN * 2 + piece * 16);
      asm volatile("" : "+r"(goff[i]));  // opaque: ptxas otherwise rebuilds each offset from (hr, W) in front of its load
    }
    const uint32_t soff = (uint32_t)col * 128u + (uint32_t)((piece ^ (col & 7)) << 4);  // + i * 2048

    // item j of this CTA = (pair iteration j / kchunks, chunk j % kchunks); this group's items are grp, grp + 2, ...
    // so with two chunks a group always has the same chunk, with one chunk it takes every other tile
    const int my_items = ((pair_end - pair0 + pair_stride - 1) / pair_stride) * kchunks;
    const int kc = (kchunks == 2) ? grp : 0;
    const int gstep_pairs = (2 / kchunks) * pair_stride;  // pairs between two items of a group
    struct Cursor { int n, th, tw; };
    const int step = 2 * gstep_pairs;                     // tiles
    const int d_n = step / a.tiles_per_sample, d_r = step - d_n * a.tiles_per_sample;
    const int d_th = d_r / a.tiles_w, d_tw = d_r - d_th * a.tiles_w;
    auto cursor_of = [&](int pair) {
      bool valid;
      const int tile = tile_of(pair, valid);
      Cursor c;
      c.n = tile / a.tiles_per_sample;
      const int tr = tile - c.n * a.tiles_per_sample;
      c.th = tr / a.tiles_w; c.tw = tr - c.th * a.tiles_w;
      return c;
    };
    auto cursor_next = [&](Cursor c, int next_pair) {
      if (2 * next_pair + (int)rank >= a.num_tiles) {   // dummy tile of an odd count, or past the end: any valid tile
        return c;
      }
      c.tw += d_tw; c.th += d_th; c.n += d_n;
      if (c.tw >= a.tiles_w) { c.tw -= a.tiles_w; ++c.th; }
      if (c.th >= tiles_h) { c.th -= tiles_h; ++c.n; }
      return c;
    };
    // this thread's vectors in the first / last halo row and column: the ones a tile on the image border loses
    uint32_t m_top = 0, m_bot = 0, m_left = 0, m_right = 0;
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int r = col + 16 * i, hr = r / kHaloW, wr = r - hr * kHaloW;
      if (hr == 0) m_top |= 1u << i;
      if (hr == kTileH + 1) m_bot |= 1u << i;
      if (wr == 0) m_left |= 1u << i;
      if (wr == kHaloW - 1) m_right |= 1u << i;
    }
    // pointer to the box origin of a tile (chunk kc) and the mask of this thread's in-image vectors
    auto tile_src = [&](const Cursor& c, const uint8_t*& base, uint32_t& okmask) {
      const int h0 = c.th * kTileH - 1, w0 = c.tw * kTileW - 1;
      base = in_bytes + (((long long)c.n * a.H + h0) * a.W + w0) * (long long)(CIN * 2) + kc * 128;
      uint32_t lost = 0;  // (the per-vector compare loop this replaces cost ~170 instructions on every border item)
      if (c.th == 0) lost |= m_top;
      if (c.th == tiles_h - 1) lost |= m_bot;
      if (c.tw == 0) lost |= m_left;
      if (c.tw == a.tiles_w - 1) lost |= m_right;
      okmask = ((1u << nvec) - 1u) & ~lost;
    };
    uint4 r[kVecs];
    auto issue_loads = [&](const uint8_t* base, uint32_t okmask) {
      if (okmask == (1u << nvec) - 1u) {  // every vector this thread owns is inside the image
#pragma unroll
        for (int i = 0; i < kVecs - 1; ++i) r[i] = ldg_nc_v4(base + goff[i]);
        r[kVecs - 1] = (nvec == kVecs) ? ldg_nc_v4(base + goff[kVecs - 1]) : make_uint4(0u, 0u, 0u, 0u);
      } else {
#pragma unroll
        for (int i = 0; i < kVecs; ++i)
          r[i] = ((okmask >> i) & 1u) ? ldg_nc_v4(base + goff[i]) : make_uint4(0u, 0u, 0u, 0u);
      }
    };

    int item = grp;                                      // this group's current item
    int pair = pair0 + (item / kchunks) * pair_stride;
    Cursor c = cursor_of(pair < pair_end ? pair : (pair0 < a.num_pairs ? pair0 : 0));
    const uint8_t* base = nullptr; uint32_t ok_c = 0;
    if (item < my_items) { tile_src(c, base, ok_c); if constexpr (!kRaw) issue_loads(base, ok_c); }
    int stage = grp % a.stages; uint32_t phase = (uint32_t)((grp / a.stages) & 1);
    int cur_n = -1;
    uint64_t ga2[4] = {0ull, 0ull, 0ull, 0ull}, gb2[4] = {0ull, 0ull, 0ull, 0ull};
    mbar_wait(w_bar, 0);  // this CTA's weights have landed (the MMA warp relies on the loaders for this)

    // silu(gn(v)) of one fp16 pair, silu(v) = h + h tanh(h), h = v / 2 (scale / shift arrive pre-halved):
    //   fp32 math: 2 cvt + FFMA2 + 2 MUFU.TANH + FFMA2 + 1 cvt.f16x2;  fp16 math: HFMA2 + 2 MUFU.TANH.F16 + PRMT + HFMA2
    uint32_t gah[4] = {0u, 0u, 0u, 0u}, gbh[4] = {0u, 0u, 0u, 0u};  // kHalfMath: packed fp16 scale / shift
    auto xform_pair = [&](uint32_t u, int j) -> uint32_t {
      if constexpr (kHalfMath) {
        uint32_t h, t, r;
        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(h) : "r"(u), "r"(gah[j]), "r"(gbh[j]));
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
        asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(h), "r"(t));
        return r;
      } else {
        float vl, vh;
        unpack_act2(u, vl, vh);
        const uint64_t h = fma_f32x2(pack_f32x2(vl, vh), ga2[j], gb2[j]);
        float hl, hh;
        unpack_f32x2(h, hl, hh);
        const uint64_t q2 = fma_f32x2(h, pack_f32x2(tanh_approx(hl), tanh_approx(hh)), h);
        float rl, rh;
        unpack_f32x2(q2, rl, rh);
        return pack_act2(rl, rh);
      }
    };

    SDD_PROF_DECL(8);
    while (item < my_items) {
      SDD_PROF_LAP(7);
      // ---- GroupNorm scale / shift: this group's chunk is fixed, so the registers only change with the sample
      if (fuse && c.n != cur_n) {
        // precomputed by gn_scale_shift_kernel: this thread's eight channels, four 16-byte loads, no barrier
        const float* pa = a.in_ab + ((size_t)c.n * CIN + kc * 64 + piece * 8);
        const float* pb = pa + (size_t)a.B * CIN;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(pa + 2 * j));
          const float4 y = __ldg(reinterpret_cast<const float4*>(pb + 2 * j));
          if constexpr (kHalfMath) {
            gah[j] = pack_act2(x.x, x.y); gah[j + 1] = pack_act2(x.z, x.w);
            gbh[j] = pack_act2(y.x, y.y); gbh[j + 1] = pack_act2(y.z, y.w);
          } else {
            ga2[j] = pack_f32x2(x.x, x.y); ga2[j + 1] = pack_f32x2(x.z, x.w);
            gb2[j] = pack_f32x2(y.x, y.y); gb2[j + 1] = pack_f32x2(y.z, y.w);
          }
        }
        cur_n = c.n;
      }
      if constexpr (kRaw) {
        // the raw box of this item (TMA, swizzled exactly like the operand stage): same offsets as the stores below;
        // out-of-image pixels arrive as zeros (TMA fill) and stay zeros (ok_c)
        const int slot = item % a.raw_slots;
        mbar_wait(raw_full_bar(slot), (uint32_t)((item / a.raw_slots) & 1));
        const uint32_t src = raw_base + (uint32_t)slot * kHaloBytes + soff;
#pragma unroll
        for (int i = 0; i < kVecs; ++i)
          if (i < nvec) r[i] = lds_v4(src + (uint32_t)i * 2048u);
        SDD_PROF_LAP(0);
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      SDD_PROF_LAP(1);
#ifdef SDD_CONV_PROF
      {  // consume every load: the lap below is the exposed load latency
        uint32_t x_ = 0;
#pragma unroll
        for (int i = 0; i < kVecs; ++i) x_ ^= r[i].x ^ r[i].w;
        SDD_PROF_SINK(x_);
      }
      SDD_PROF_LAP(2);
#endif
      const uint32_t dst = a_base + (uint32_t)stage * kHaloBytes + soff;
      if constexpr (!kRaw) {
        if (!fuse) {
          // un-fused launches: padding pixels hold the zeros they were "loaded" as
#pragma unroll
          for (int i = 0; i < kVecs; ++i)
            if (i < nvec) sts_v4(dst + (uint32_t)i * 2048u, r[i]);
        } else if (ok_c == (1u << nvec) - 1u) {
          // interior tile: straight-line code, the vectors' chains interleave freely.  (Three of a group's four warps own
          // eleven vectors, not twelve -- rows >= 180 do not exist -- and used to fall into the masked path below on every
          // item: 78 % of all loader executions, ~1.7x the instructions.)
#pragma unroll
          for (int i = 0; i < kVecs - 1; ++i)
            sts_v4(dst + (uint32_t)i * 2048u,
                   make_uint4(xform_pair(r[i].x, 0), xform_pair(r[i].y, 1), xform_pair(r[i].z, 2), xform_pair(r[i].w, 3)));
          if (nvec == kVecs)  // warp-uniform: col = tg >> 3, so only the first warp of a group has the twelfth vector
            sts_v4(dst + (uint32_t)(kVecs - 1) * 2048u,
                   make_uint4(xform_pair(r[kVecs - 1].x, 0), xform_pair(r[kVecs - 1].y, 1), xform_pair(r[kVecs - 1].z, 2),
                              xform_pair(r[kVecs - 1].w, 3)));
        } else {
          // border tile (18 % of the tiles at 256^2): the SAME straight-line transform, then out-of-image vectors are zeroed
          // with a mask -- padding must be zero AFTER the activation.  No branch inside the unrolled loop: with the `fuse`
          // test per vector ptxas kept the twelve chains serial (MUFU latency exposed: a border item cost 2.1x an interior
          // one in the round-2 128->128 capture, profiles/r2_stalls_conv.md).
          auto xform_masked = [&](int i) {
            const uint32_t m = 0u - ((ok_c >> i) & 1u);
            sts_v4(dst + (uint32_t)i * 2048u, make_uint4(xform_pair(r[i].x, 0) & m, xform_pair(r[i].y, 1) & m,
                                                        xform_pair(r[i].z, 2) & m, xform_pair(r[i].w, 3) & m));
          };
#pragma unroll
          for (int i = 0; i < kVecs - 1; ++i) xform_masked(i);
          if (nvec == kVecs) xform_masked(kVecs - 1);
        }
      } else {
        // raw-ring layers (the vectors come from shared memory): per-vector guards, which ptxas schedules next to each
        // vector's LDS -- same-box A/B, the straight-line border form above is 1.3 % slower here
        if (fuse && ok_c == (1u << nvec) - 1u) {
#pragma unroll
          for (int i = 0; i < kVecs - 1; ++i)
            sts_v4(dst + (uint32_t)i * 2048u,
                   make_uint4(xform_pair(r[i].x, 0), xform_pair(r[i].y, 1), xform_pair(r[i].z, 2), xform_pair(r[i].w, 3)));
          if (nvec == kVecs)
            sts_v4(dst + (uint32_t)(kVecs - 1) * 2048u,
                   make_uint4(xform_pair(r[kVecs - 1].x, 0), xform_pair(r[kVecs - 1].y, 1), xform_pair(r[kVecs - 1].z, 2),
                              xform_pair(r[kVecs - 1].w, 3)));
        } else {
#pragma unroll
          for (int i = 0; i < kVecs; ++i) {
            if (i < nvec) {
              uint4 v = r[i];
              if (fuse) {
                const uint32_t m = ((ok_c >> i) & 1u) ? 0xffffffffu : 0u;
                v = make_uint4(xform_pair(v.x, 0) & m, xform_pair(v.y, 1) & m, xform_pair(v.z, 2) & m, xform_pair(v.w, 3) & m);
              }
              sts_v4(dst + (uint32_t)i * 2048u, v);
            }
          }
        }
      }
      SDD_PROF_LAP(3);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      SDD_PROF_LAP(4);
      // relaxed: the proxy fence has completed this thread's shared-memory writes and made them visible to the async
      // proxy of THIS CTA's tensor core, the only reader
      if (lane == 0) mbar_arrive_relaxed_remote(ready_bar(stage), 0);
      if constexpr (kRaw) {  // every lane's LDS results were consumed by the stores above: the raw slot is free
        if (lane == 0) mbar_arrive(raw_empty_bar(item % a.raw_slots));
      }
      // ---- this group's next item: coordinates, then its loads (nothing of this thread is in flight at a MEMBAR)
      item += 2;
      stage += 2; if (stage >= a.stages) { stage -= a.stages; phase ^= 1u; }
      if (a.stages == 1) phase ^= 1u;  // two wraps per step when there is a single stage
      pair += gstep_pairs;
      if (item < my_items) {
        c = cursor_next(c, pair);
        tile_src(c, base, ok_c);
        if constexpr (!kRaw) issue_loads(base, ok_c);
        // pull the box of this group's item AFTER that one into L2 (TMA prefetch: no smem, no barrier).  With a single
        // register buffer the loads above are on the group's chain; this turns their DRAM latency into an L2 hit.
        if (!kRaw && a.prefetch > 0 && tg == 0 && item + 2 < my_items) {
          const Cursor cp = cursor_next(c, pair + gstep_pairs);
          tma_prefetch_l2_4d(&tmA, kc * 64, cp.tw * kTileW - 1, cp.th * kTileH - 1, cp.n);
        }
      }
      SDD_PROF_LAP(5);
    }
    SDD_PROF_FLUSH(blockIdx.x == 0 && (tg >> 5) == 0 && lane == 0, 16 + 8 * grp, 8);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading our smem / arriving on our barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(SDD_TMEM_BASE()), "n"(kTmemCols) : "memory");
#undef SDD_TMEM_BASE
#undef SDD_CFG
  }
}

}  // namespace sdd
