// Self-attention BLOCK around attention_fwd_kernel (SURVEY.md 8(a) A8 / 8(f) N2 -- extension, no reference code; oracle =
// oracle.attention_block):   out = x + proj(attention(q, k, v)),  [q | k | v] = GroupNorm(4, C)(x) W_qkv^T + b_qkv
// for C = 128 channels = 2 heads x 64 at the 32^2 / 16^2 feature maps the north star names.  The projections are
// M x 128 x {384, 128} GEMMs (a few GFLOP: 1-2 % of one 3x3 conv launch), done here on bf16 mma.sync.m16n8k16 with the
// surrounding data movement fused in:
//   MODE_QKV : A = x (raw bf16 NHWC rows) with the GroupNorm affine applied in registers on load; the epilogue adds the
//              bias and scatters straight into the layouts attention_fwd_kernel consumes -- q, k as [B*heads][S][64],
//              v TRANSPOSED as [B*heads][64][S] -- so no separate split / transpose pass exists;
//   MODE_PROJ: A = attention output [B*heads][S][64] read head-merged; the epilogue adds bias and the residual x.
#pragma once
#include "common.cuh"
#include "unet_kernels.cuh"

namespace sdd {

constexpr int kAbC = 128;      // channels
constexpr int kAbHeads = 2;    // heads of 64
constexpr int kAbRows = 64;    // rows (tokens) per CTA: 4 warps x 16

struct AttnBlockGemmArgs {
  const __nv_bfloat16* a;        // MODE_QKV: x [B*S][128];  MODE_PROJ: attention out [B*heads][S][64]
  const float* w;                // [N][128] fp32 (rounded to bf16 on load)
  const float* bias;             // [N]
  const float* meanrstd;         // MODE_QKV: [B][4][2]
  const float* gamma;            // MODE_QKV: [128]
  const float* beta;             // MODE_QKV: [128]
  const __nv_bfloat16* resid;    // MODE_PROJ: x [B*S][128]
  __nv_bfloat16* q;              // MODE_QKV outputs
  __nv_bfloat16* k;
  __nv_bfloat16* vt;
  __nv_bfloat16* out;            // MODE_PROJ output [B*S][128]
  int B, S;
};

// grid: (B*S / 64, N / 64); 128 threads.  Warp w owns rows 16w..16w+15 of the CTA's 64 and all 64 columns of its slab.
template <int MODE>
__global__ void __launch_bounds__(128) attn_block_gemm_kernel(const AttnBlockGemmArgs g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & 3, j = lane >> 2;
  const int row0 = blockIdx.x * kAbRows + warp * 16;  // first token row of this warp (global over B*S)
  const int col0 = blockIdx.y * 64;                   // first output column of this CTA
  const int b = row0 / g.S;                           // S % 64 == 0: a CTA never straddles two samples
  __shared__ float s_sc[kAbC], s_sh[kAbC];
  if (MODE == 0) {
    for (int c = threadIdx.x; c < kAbC; c += 128) {
      const int grp = c / (kAbC / 4);
      const float mean = g.meanrstd[(b * 4 + grp) * 2], rstd = g.meanrstd[(b * 4 + grp) * 2 + 1];
      const float sc = rstd * g.gamma[c];
      s_sc[c] = sc;
      s_sh[c] = g.beta[c] - mean * sc;
    }
  }
  __syncthreads();

  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  const int r_lo = row0 + j, r_hi = row0 + j + 8;
#pragma unroll
  for (int ks = 0; ks < kAbC / 16; ++ks) {
    const int k0 = ks * 16 + 2 * t;  // this lane's k pairs: k0, k0+1 and k0+8, k0+9
    uint32_t af[4];
    auto load_a = [&](int r, int kk) -> uint32_t {
      if (MODE == 0) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(g.a + (size_t)r * kAbC + kk);
        const float lo = __uint_as_float(u << 16), hi = __uint_as_float(u & 0xffff0000u);
        return pack_bf16x2(fmaf(lo, s_sc[kk], s_sh[kk]), fmaf(hi, s_sc[kk + 1], s_sh[kk + 1]));
      } else {
        const int s = r - b * g.S, h = kk >> 6, d = kk & 63;  // head-merged read of [B*heads][S][64]
        return *reinterpret_cast<const uint32_t*>(g.a + (((size_t)(b * kAbHeads + h) * g.S + s) << 6) + d);
      }
    };
    af[0] = load_a(r_lo, k0); af[1] = load_a(r_hi, k0); af[2] = load_a(r_lo, k0 + 8); af[3] = load_a(r_hi, k0 + 8);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float* wr = g.w + (size_t)(col0 + n * 8 + j) * kAbC;  // B operand "col": W[n][k]
      const float2 w0 = *reinterpret_cast<const float2*>(wr + k0), w1 = *reinterpret_cast<const float2*>(wr + k0 + 8);
      mma_bf16_16816(acc[n], af, pack_bf16x2(w0.x, w0.y), pack_bf16x2(w1.x, w1.y));
    }
  }
  // epilogue: lane holds columns col0 + 8n + 2t, +1 of rows r_lo (acc[n][0..1]) and r_hi (acc[n][2..3])
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = col0 + n * 8 + 2 * t;
    const float b0 = g.bias[c], b1 = g.bias[c + 1];
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int r = rh ? r_hi : r_lo;
      const float v0 = acc[n][2 * rh] + b0, v1 = acc[n][2 * rh + 1] + b1;
      if (MODE == 0) {
        const int which = c >> 7, h = (c & 127) >> 6, d = c & 63, s = r - b * g.S;
        const size_t bh = (size_t)(b * kAbHeads + h);
        if (which == 2) {  // V transposed: [bh][d][s]
          g.vt[(bh * 64 + d) * g.S + s] = __float2bfloat16(v0);
          g.vt[(bh * 64 + d + 1) * g.S + s] = __float2bfloat16(v1);
        } else {
          __nv_bfloat16* dst = (which == 0 ? g.q : g.k) + ((bh * g.S + s) << 6) + d;
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v0, v1);
        }
      } else {
        const uint32_t xr = *reinterpret_cast<const uint32_t*>(g.resid + (size_t)r * kAbC + c);
        const float x0 = __uint_as_float(xr << 16), x1 = __uint_as_float(xr & 0xffff0000u);
        *reinterpret_cast<uint32_t*>(g.out + (size_t)r * kAbC + c) = pack_bf16x2(v0 + x0, v1 + x1);
      }
    }
  }
}

}  // namespace sdd
