// Training-side reuse of the sampling kernels: forward noising q_sample (ddpm.py:13-17) and the
// eps-MSE of p_losses (ddpm.py:20-24), forward only -- validation-loss evaluation on the sampler's UNet kernels.
#pragma once
#include "common.cuh"

namespace sdd {

// out[b,:] = ca[b] * x0[b,:] + cb[b] * noise[b,:] with the reference's expression tree (two products, one sum, no FMA
// contraction): bit-identical to torch.sqrt(alpha_bar) * x_start + torch.sqrt(1 - alpha_bar) * noise, ddpm.py:17.
// ca / cb are evaluated by the caller with the same torch ops on the same fp32 table (ddpm.py:16-17).
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const float* __restrict__ ca, const float* __restrict__ cb,
                                                       float* __restrict__ out, int D) {
  const int b = blockIdx.y;
  const float a = ca[b], c = cb[b];
  const int nq = D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x0 + (size_t)b * D);
  const float4* n4 = reinterpret_cast<const float4*>(noise + (size_t)b * D);
  float4* o4 = reinterpret_cast<float4*>(out + (size_t)b * D);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const float4 x = __ldcs(x4 + q), n = __ldcs(n4 + q);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(c, n.x));
    o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(c, n.y));
    o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(c, n.z));
    o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(c, n.w));
    o4[q] = o;
  }
}

// mean((pred - target)^2) over n elements (F.mse_loss, reduction="mean", ddpm.py:24): per-CTA partial sums in double
// written to partials[gridDim.x], reduced in a fixed order by the last launch (no atomics -> deterministic).
constexpr int kMseBlocks = 1024;
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          size_t n4, double* __restrict__ partials) {
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float acc = 0.0f;
  double dacc = 0.0;
  int cnt = 0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
    const float4 p = __ldcs(p4 + q), t = __ldcs(t4 + q);
    const float d0 = p.x - t.x, d1 = p.y - t.y, d2 = p.z - t.z, d3 = p.w - t.w;
    acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    if (++cnt == 64) { dacc += (double)acc; acc = 0.0f; cnt = 0; }  // bounded fp32 run length
  }
  dacc += (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w];
    partials[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(256) mse_final_kernel(const double* __restrict__ partials, int nparts, double count,
                                                        float* __restrict__ out) {
  __shared__ double red[256];
  double v = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) v += partials[i];
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(red[0] / count);
}

}  // namespace sdd
