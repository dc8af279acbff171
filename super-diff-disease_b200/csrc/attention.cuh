// Fused flash-style self-attention core on tcgen05 / TMEM / TMA  (SURVEY.md 8(a) A8 / 8(f) N2 -- a north-star-only
// extension: the reference UNet has no attention, src/models/unet.py:37-65; oracle = ours, oracle.attention_core).
//
//   out[bh, q, :] = softmax_k(scale * <Q[bh,q,:], K[bh,k,:]>) @ V[bh,k,:]        head_dim = 64, S % 128 == 0
//
// sized for the feature maps the north star names (32^2 = 1024 and 16^2 = 256 tokens).  The S x S score matrix never
// leaves the SM: per CTA one (batch*head, 128-query) tile; per 128-key block
//   MMA1  S[128 x 128]  = Q K^T        tcgen05.mma kind::f16, A = Q (smem, loaded once), B = K block (smem), D in TMEM
//   softmax warps: tcgen05.ld S -> running row max / sum (online softmax, exp2 with the scale folded in) -> P as fp16
//         straight into the K-major SWIZZLE_128B layout the next MMA reads as its A operand (fence.proxy.async)
//   MMA2  PV[128 x 64]  = P V          A = P (smem), B = V^T block (smem, keys contiguous), D in TMEM
//   softmax warps: O = O * exp2(m_old - m_new) + PV   (O lives in registers: 64 fp32 per row)
// MMA1 of block j+1 is issued right behind MMA2 of block j, so the tensor core computes the next scores while the
// softmax warps fold PV_j into O.  K / V^T blocks arrive through a 2-stage TMA ring filled by a dedicated producer lane.
// V is taken TRANSPOSED (vt[bh, d, k]) so that both operands of MMA2 are plain K-major tiles; the projection GEMM that
// produces V writes it that way.
// Warp roles (192 threads): warp 0 = MMA issuer (one lane), warp 1 = TMEM allocator + TMA producer (one lane),
// warps 2..5 = softmax / epilogue (TMEM lane quadrant = warp % 4, thread = query row).
#pragma once
#include "common.cuh"
#include "conv_common.cuh"

namespace sdd {

constexpr int kAttnD = 64;          // head dim: one 128-byte swizzle row
constexpr int kAttnBM = 128;        // queries per CTA
constexpr int kAttnBN = 128;        // keys per block
constexpr int kAttnThreads = 192;
constexpr int kAttnQBytes = kAttnBM * 128;          // 16 KB
constexpr int kAttnKBytes = kAttnBN * 128;          // 16 KB
constexpr int kAttnVBytes = 2 * kAttnD * 128;       // two 64-key halves of V^T: 16 KB
constexpr int kAttnPBytes = 2 * kAttnBM * 128;      // two 64-key halves of P: 32 KB
constexpr int kAttnStages = 2;
// 112 KB of tiles + 128 B of barriers and NO alignment slack: two CTAs per SM need 2 x (dynamic + 1 KB reserved) <= 228 KB,
// and with the round-1 "+ 1024 + 256" the kernel missed that by 512 bytes and ran ONE CTA per SM (ncu: 9 % warps active).
// The kernel has no static shared memory, so its dynamic window starts at CTA-relative address 0 (1024-aligned, as the
// SWIZZLE_128B tiles need); it verifies that and traps otherwise.
constexpr int kAttnSmem = kAttnQBytes + kAttnStages * (kAttnKBytes + kAttnVBytes) + kAttnPBytes + 128;
constexpr int kAttnTmemCols = 256;  // S: columns [0,128), PV: [128,192)

struct AttnArgs {
  act_t* out;  // [BH][S][64]
  int S, BH;
  float scale_log2e;   // softmax scale * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmVt, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t attn_smem_raw[];
  const uint32_t smem_base = smem_u32(attn_smem_raw);
  if ((smem_base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("sdd: attention_fwd_kernel needs a 1024-byte aligned dynamic shared-memory window\n");
    __trap();
  }
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int s) { return smem_base + kAttnQBytes + (uint32_t)s * (kAttnKBytes + kAttnVBytes); };
  auto v_smem = [&](int s) { return k_smem(s) + kAttnKBytes; };
  const uint32_t p_smem = smem_base + kAttnQBytes + kAttnStages * (kAttnKBytes + kAttnVBytes);
  const uint32_t bar_base = p_smem + kAttnPBytes;
  const uint32_t bar_q = bar_base;
  auto bar_kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto bar_kv_free = [&](int s) { return bar_base + 8u * (3 + s); };
  const uint32_t bar_s = bar_base + 8u * 5, bar_p = bar_base + 8u * 6, bar_pv = bar_base + 8u * 7;
  const uint32_t tmem_slot = bar_base + 8u * 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(attn_smem_raw + (tmem_slot - smem_u32(attn_smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x, bh = blockIdx.y;
  const int nblk = a.S / kAttnBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    mbar_init(bar_q, 1);
    for (int s = 0; s < kAttnStages; ++s) { mbar_init(bar_kv_full(s), 1); mbar_init(bar_kv_free(s), 1); }
    mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kAttnTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t s_tmem = tmem_base, pv_tmem = tmem_base + 128u;

  if (warp == 1) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, kAttnQBytes);
      tma_load_3d(q_smem, &tmQ, bar_q, 0, qtile * kAttnBM, bh);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kAttnStages;
        if (j >= kAttnStages) mbar_wait(bar_kv_free(s), (uint32_t)(((j / kAttnStages) - 1) & 1));
        mbar_arrive_expect_tx(bar_kv_full(s), kAttnKBytes + kAttnVBytes);
        tma_load_3d(k_smem(s), &tmK, bar_kv_full(s), 0, j * kAttnBN, bh);
        tma_load_3d(v_smem(s), &tmVt, bar_kv_full(s), j * kAttnBN, 0, bh);
        tma_load_3d(v_smem(s) + kAttnD * 128, &tmVt, bar_kv_full(s), j * kAttnBN + 64, 0, bh);
      }
    }
  } else if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = SDD_ACT_IDESC(kAttnBM, kAttnBN);
      constexpr uint32_t idesc_o = SDD_ACT_IDESC(kAttnBM, kAttnD);
      auto issue_s = [&](int j) {
        const int s = j % kAttnStages;
        mbar_wait(bar_kv_full(s), (uint32_t)((j / kAttnStages) & 1));
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(q_smem), bdesc = umma_desc_sw128(k_smem(s));
#pragma unroll
        for (int k = 0; k < kAttnD / 16; ++k)
          umma_f16(s_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        umma_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kAttnStages;
        mbar_wait(bar_p, (uint32_t)(j & 1));  // P_j is in shared memory, S_j and PV_{j-1} have been read
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAttnBN / 16; ++k) {
          const uint64_t adesc = umma_desc_sw128(p_smem + (uint32_t)(k >> 2) * (kAttnBM * 128)) + (uint64_t)((k & 3) * 2);
          const uint64_t bdesc = umma_desc_sw128(v_smem(s) + (uint32_t)(k >> 2) * (kAttnD * 128)) + (uint64_t)((k & 3) * 2);
          umma_f16(pv_tmem, adesc, bdesc, idesc_o, k ? 1u : 0u);
        }
        umma_commit(bar_pv);
        umma_commit(bar_kv_free(s));
        if (j + 1 < nblk) issue_s(j + 1);  // the next scores are computed while the softmax warps fold PV_j into O
      }
    }
  } else {
    // ===================== softmax / epilogue: thread = query row =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.0f;
    float o[kAttnD];
#pragma unroll
    for (int i = 0; i < kAttnD; ++i) o[i] = 0.0f;
    // byte offset of this row inside a [128 x 128 B] swizzled region, and its swizzle key
    const uint32_t p_row = p_smem + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t key = (uint32_t)(row & 7);

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(bar_s, (uint32_t)(j & 1));
      tc_fence_after();
      // pass 1: row max of the 128 scores
      float mx = -INFINITY;
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(s_tmem + lane_off + (uint32_t)(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * a.scale_log2e);  // exp2(-inf) = 0 on the first block
      const float mb = m_new * a.scale_log2e;
      // pass 2: p = exp2(s * c - m_new * c), row sum, fp16 P into the swizzled A-operand layout
      float lsum = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32(s_tmem + lane_off + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), a.scale_log2e, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), a.scale_log2e, -mb));
          pk[i] = pack_act2(p0, p1);
          // the row sum uses the rounded values the tensor core will multiply (keeps rows normalised)
          float r0, r1;
          unpack_act2(pk[i], r0, r1);
          lsum += r0 + r1;
        }
        // keys c*32 .. c*32+31 = 64 bytes = pieces (c&1)*4 .. +3 of region c>>1
        const uint32_t reg_base = p_row + (uint32_t)(c >> 1) * (kAttnBM * 128);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t piece = (uint32_t)((c & 1) * 4 + pc);
          sts_v4(reg_base + ((piece ^ key) << 4), make_uint4(pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]));
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      tc_fence_before();          // the TMEM reads of S_j are complete (wait::ld above) before MMA1 of block j+1
      fence_proxy_async_smem();   // P_j visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // fold PV_j into O
      mbar_wait(bar_pv, (uint32_t)(j & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(pv_tmem + lane_off + (uint32_t)(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    // epilogue: out = O / l, fp16, this thread's 128-byte row
    const float inv = 1.0f / l_run;
    act_t* orow = a.out + ((size_t)bh * a.S + (size_t)qtile * kAttnBM + row) * kAttnD;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_act2(o[c * 16 + 2 * i] * inv, o[c * 16 + 2 * i + 1] * inv);
      st_global_v8(orow + c * 16, pk);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

}  // namespace sdd
