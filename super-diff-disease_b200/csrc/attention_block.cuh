// Self-attention BLOCK around attention_fwd_kernel (SURVEY.md 8(a) A8 / 8(f) N2 -- extension, no reference code; oracle =
// oracle.attention_block):   out = x + proj(attention(q, k, v)),  [q | k | v] = GroupNorm(4, C)(x) W_qkv^T + b_qkv
// for C = 128 channels = 2 heads x 64 at the 32^2 / 16^2 feature maps the north star names.  The projections are
// M x 128 x {384, 128} GEMMs (a few GFLOP: 1-2 % of one 3x3 conv launch), done here on fp16 mma.sync.m16n8k16 with the
// surrounding data movement fused in:
//   MODE_QKV : A = x (raw fp16 NHWC rows) with the GroupNorm affine applied in registers on load (statistics either as
//              (mean, rstd) floats or as the producer's fixed-point sums, common.cuh); the epilogue adds the bias and
//              scatters straight into the layouts attention_fwd_kernel consumes -- q, k as [B*heads][S][64], v TRANSPOSED
//              as [B*heads][64][S] -- so no separate split / transpose pass exists;
//   MODE_PROJ: A = attention output [B*heads][S][64] read head-merged; the epilogue adds bias and the residual x and, when
//              asked, accumulates the OUTPUT's GroupNorm(4,128) sums for the next layer (fixed-point REDs).
// No allocation, no synchronisation: all scratch comes from the caller, so the block can sit inside a captured graph.
#pragma once
#include "common.cuh"
#include "unet_kernels.cuh"

namespace sdd {

constexpr int kAbC = 128;      // channels
constexpr int kAbHeads = 2;    // heads of 64
constexpr int kAbRows = 128;   // rows (tokens) per CTA: two passes of 4 warps x 16
constexpr int kAbWStride = kAbC + 8;  // halfs per staged weight row: 272 B, so a warp's B-fragment loads hit 32 distinct banks

struct AttnBlockGemmArgs {
  const act_t* a;                // MODE_QKV: x [B*S][128];  MODE_PROJ: attention out [B*heads][S][64]
  const float* w;                // [N][128] fp32 (rounded to fp16 on load)
  const float* bias;             // [N]
  const float* meanrstd;         // MODE_QKV: [B][4][2], or
  const long long* in_sums;      //           [B][4][2] fixed-point (sum, sumsq) of x per group (count = S * 32)
  const float* gamma;            // MODE_QKV: [128]
  const float* beta;             // MODE_QKV: [128]
  const act_t* resid;            // MODE_PROJ: x [B*S][128]
  act_t* q;                      // MODE_QKV outputs
  act_t* k;
  act_t* vt;
  act_t* out;                    // MODE_PROJ output [B*S][128]
  long long* out_sums;           // MODE_PROJ: [B][4][2] accumulators of the output's GroupNorm(4,128) sums, or nullptr
  int B, S;
};

__device__ __forceinline__ void mma_act_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifndef SDD_ACT_BF16
  mma_f16_16816(d, a, b0, b1);
#else
  mma_bf16_16816(d, a, b0, b1);
#endif
}

// grid: (B*S / 128, N / 64); 128 threads.  The CTA's 64 x 128 weight slab is staged ONCE in shared memory as fp16 (the
// first version re-read fp32 weights from global memory for every MMA: 114 us per qkv launch at 64 x 32^2 tokens, 9 % of
// the variant's step); per pass warp w owns rows 16w..16w+15 of 64 and all 64 columns of the slab.
template <int MODE>
__global__ void __launch_bounds__(128) attn_block_gemm_kernel(const AttnBlockGemmArgs g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & 3, j = lane >> 2;
  const int col0 = blockIdx.y * 64;                   // first output column of this CTA
  const int b = (blockIdx.x * kAbRows) / g.S;         // S % 128 == 0: a CTA never straddles two samples
  __shared__ float s_sc[kAbC], s_sh[kAbC];
  __shared__ float s_red[4][2][2];
  __shared__ __align__(16) act_t s_w[64 * kAbWStride];
  for (int i = threadIdx.x; i < 64 * (kAbC / 2); i += 128) {
    const int n = i / (kAbC / 2), kp = i - n * (kAbC / 2);
    const float2 wv = *reinterpret_cast<const float2*>(g.w + (size_t)(col0 + n) * kAbC + 2 * kp);
    *reinterpret_cast<uint32_t*>(&s_w[n * kAbWStride + 2 * kp]) = pack_act2(wv.x, wv.y);
  }
  if (MODE == 0) {
    for (int c = threadIdx.x; c < kAbC; c += 128) {
      const int grp = c / (kAbC / 4);
      float mean, rstd;
      if (g.in_sums) gn_mean_rstd_from_sums(g.in_sums + ((size_t)b * 4 + grp) * 2, (double)g.S * (double)(kAbC / 4), kGnEps, mean, rstd);
      else { mean = g.meanrstd[(b * 4 + grp) * 2]; rstd = g.meanrstd[(b * 4 + grp) * 2 + 1]; }
      const float sc = rstd * g.gamma[c];
      s_sc[c] = sc;
      s_sh[c] = g.beta[c] - mean * sc;
    }
  }
  __syncthreads();

  float gs[2] = {0.f, 0.f}, gss[2] = {0.f, 0.f};  // MODE_PROJ: this CTA's 64 columns = GroupNorm groups col0/32, col0/32 + 1
#pragma unroll 1
  for (int pass = 0; pass < kAbRows / 64; ++pass) {
  const int row0 = blockIdx.x * kAbRows + pass * 64 + warp * 16;  // first token row of this warp (global over B*S)
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
        float lo, hi;
        unpack_act2(u, lo, hi);
        return pack_act2(fmaf(lo, s_sc[kk], s_sh[kk]), fmaf(hi, s_sc[kk + 1], s_sh[kk + 1]));
      } else {
        const int s = r - b * g.S, h = kk >> 6, d = kk & 63;  // head-merged read of [B*heads][S][64]
        return *reinterpret_cast<const uint32_t*>(g.a + (((size_t)(b * kAbHeads + h) * g.S + s) << 6) + d);
      }
    };
    af[0] = load_a(r_lo, k0); af[1] = load_a(r_hi, k0); af[2] = load_a(r_lo, k0 + 8); af[3] = load_a(r_hi, k0 + 8);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const act_t* wr = &s_w[(n * 8 + j) * kAbWStride + k0];  // B operand "col": W[n][k]
      mma_act_16816(acc[n], af, *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
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
          g.vt[(bh * 64 + d) * g.S + s] = float_to_act(v0);
          g.vt[(bh * 64 + d + 1) * g.S + s] = float_to_act(v1);
        } else {
          act_t* dst = (which == 0 ? g.q : g.k) + ((bh * g.S + s) << 6) + d;
          *reinterpret_cast<uint32_t*>(dst) = pack_act2(v0, v1);
        }
      } else {
        const uint32_t xr = *reinterpret_cast<const uint32_t*>(g.resid + (size_t)r * kAbC + c);
        float x0, x1;
        unpack_act2(xr, x0, x1);
        const float o0 = v0 + x0, o1 = v1 + x1;
        *reinterpret_cast<uint32_t*>(g.out + (size_t)r * kAbC + c) = pack_act2(o0, o1);
        gs[n >> 2] += o0 + o1;
        gss[n >> 2] = fmaf(o0, o0, fmaf(o1, o1, gss[n >> 2]));
      }
    }
  }
  }  // pass
  if (MODE == 1 && g.out_sums) {
    // statistics of the fp32 values before rounding (as the conv epilogue does): warp shuffle, fixed-order sum over the
    // four warps, one fixed-point RED per (group, statistic)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs[i] += __shfl_xor_sync(0xffffffffu, gs[i], o);
        gss[i] += __shfl_xor_sync(0xffffffffu, gss[i], o);
      }
    }
    if (lane == 0) { s_red[warp][0][0] = gs[0]; s_red[warp][0][1] = gss[0]; s_red[warp][1][0] = gs[1]; s_red[warp][1][1] = gss[1]; }
    __syncthreads();
    if (threadIdx.x < 4) {
      const int gi = threadIdx.x >> 1, st = threadIdx.x & 1;
      const float v = (s_red[0][gi][st] + s_red[1][gi][st]) + (s_red[2][gi][st] + s_red[3][gi][st]);
      gn_red_add(g.out_sums + ((size_t)b * 4 + (col0 >> 5) + gi) * 2 + st, v);
    }
  }
}

// ------------------------------------------------------------------ resampling between the levels of the multi-resolution
// UNet variant (extension N2; oracle.unet_attn): fp16 NHWC, fp32 arithmetic, one 16-byte vector (8 channels) per thread
// and output pixel, GroupNorm(4,C) sums of the OUTPUT accumulated for the next layer.  HBM-bound elementwise passes.
//   pool2 : out[n,h,w,:] = mean of the 2x2 block in[n,2h..2h+1,2w..2w+1,:]                      (F.avg_pool2d(., 2))
//   up2add: out[n,h,w,:] = in[n,h/2,w/2,:] + skip[n,h,w,:]      (nearest-neighbour upsampling + additive skip connection)
// grid = (blocks, B): a block stays inside one sample; per-thread partial sums are a fixed function of the shape, the
// cross-lane / cross-warp sums run in a fixed order, and blocks meet in integer (fixed-point) REDs, so the statistics are
// bit-reproducible.  (A first version had every thread add into shared-memory 64-bit atomics: the contended CAS loops
// made the full-resolution up+skip pass cost 2.5 ms per 64 samples, 30 % of the variant's step.)
template <bool kUp>
__global__ void __launch_bounds__(256) resample_kernel(const act_t* __restrict__ in, const act_t* __restrict__ skip,
                                                       act_t* __restrict__ out, long long* out_sums, int Ho, int Wo,
                                                       int C) {
  const int b = blockIdx.y;
  const int cv = C >> 3;                     // 16-byte vectors per pixel
  const int nvec = Ho * Wo * cv;
  const int Hi = kUp ? Ho / 2 : Ho * 2, Wi = kUp ? Wo / 2 : Wo * 2;
  const uint4* inb = reinterpret_cast<const uint4*>(in + (size_t)b * Hi * Wi * C);
  const uint4* skb = kUp ? reinterpret_cast<const uint4*>(skip + (size_t)b * Ho * Wo * C) : nullptr;
  uint4* outb = reinterpret_cast<uint4*>(out + (size_t)b * Ho * Wo * C);
  __shared__ float s_red[8][4][2];
  // a thread's vector index v = i % cv = threadIdx.x % cv is the same for every i it visits (256 % cv == 0, cv = 8 or 16)
  const int stride = gridDim.x * 256;
  float s = 0.f, ss = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
    const int v = i % cv, p = i / cv, w = p % Wo, h = p / Wo;
    float f[8];
    if (kUp) {
      const uint4 a = __ldg(inb + ((size_t)(h >> 1) * Wi + (w >> 1)) * cv + v);
      const uint4 k = __ldg(skb + (size_t)p * cv + v);
      const uint32_t au[4] = {a.x, a.y, a.z, a.w}, ku[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a0, a1, k0, k1;
        unpack_act2(au[j], a0, a1);
        unpack_act2(ku[j], k0, k1);
        f[2 * j] = a0 + k0; f[2 * j + 1] = a1 + k1;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const uint4 a = __ldg(inb + ((size_t)(2 * h + dy) * Wi + (2 * w + dx)) * cv + v);
          const uint32_t au[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a0, a1;
            unpack_act2(au[j], a0, a1);
            f[2 * j] += a0; f[2 * j + 1] += a1;
          }
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= 0.25f;
    }
    uint4 o;
    o.x = pack_act2(f[0], f[1]); o.y = pack_act2(f[2], f[3]); o.z = pack_act2(f[4], f[5]); o.w = pack_act2(f[6], f[7]);
    outb[i] = o;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
  }
  if (out_sums) {
    // lanes of one GroupNorm group inside a warp: lane % cv in [g * cv/4, (g+1) * cv/4)  ->  sum over the lane bits that
    // do not select the group (cv = 8: bits 0, 3, 4; cv = 16: bits 0, 1, 4), then over the 8 warps in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s += __shfl_xor_sync(0xffffffffu, s, 1);  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 16);
    const int o3 = cv == 8 ? 8 : 2;
    s += __shfl_xor_sync(0xffffffffu, s, o3); ss += __shfl_xor_sync(0xffffffffu, ss, o3);
    const int gshift = cv == 8 ? 1 : 2;       // first lane of group g: g << gshift
    if (lane < 4 * (1 << gshift) && (lane & ((1 << gshift) - 1)) == 0) {
      s_red[warp][lane >> gshift][0] = s; s_red[warp][lane >> gshift][1] = ss;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float a = 0.f;
      for (int wq = 0; wq < 8; ++wq) a += s_red[wq][threadIdx.x >> 1][threadIdx.x & 1];
      gn_red_add(out_sums + (size_t)b * 8 + threadIdx.x, a);
    }
  }
}

}  // namespace sdd
