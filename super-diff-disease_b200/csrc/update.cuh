// Fused superposition update: one coalesced, vectorised HBM pass per step.
//   reads  x[B,D], eps[M,B,D], (noise[B,D] | Philox)      writes x'[B,D]
//   per-sample reductions <eps_m,dx>, <x,eps_m>, |eps_m|^2, sum x', sum x'^2 via warp shuffles,
//   fixed-order cross-CTA reduce by the last CTA of each sample -> logq', kappa, GN(1,1) stats of x'.
// Algorithmic bytes / element: 4 + 4M + (4 if noise tensor) + 4.
#pragma once
#include "common.cuh"

namespace sdd {

constexpr int kMaxModels = 4;
constexpr int kUpdThreads = 256;

// Per-timestep scalars, either passed by value (operator API) or read from a device table
// indexed by the device-side step counter (captured step graph).
struct StepScalars {
  float alpha, alpha_bar, beta;
  int draw_index;  // Philox draw index / noise-stack slice for this step; < 0 => z = 0 (t == 0)
  // coefficients of the x update, evaluated ONCE on the host in fp32 with IEEE sqrt / divide -- the same correctly
  // rounded operations torch performs for ddpm.py:42-44, so every thread no longer repeats ~50 instructions of them:
  float c1;  // 1 / sqrt(alpha)
  float c2;  // (1 - alpha) / sqrt(1 - alpha_bar)
  float c3;  // sqrt(beta)
  float pad;
};
inline void step_scalars_fill(StepScalars& s) {
  s.c1 = 1.0f / sqrtf(s.alpha);
  s.c2 = (1.0f - s.alpha) / sqrtf(1.0f - s.alpha_bar);
  s.c3 = sqrtf(s.beta);
  s.pad = 0.f;
}

struct UpdateArgs {
  const float* x_in;
  float* x_out;
  const float* eps;          // [M,B,D]
  const float* noise;        // [B,D] slice base or nullptr
  int64_t noise_step_stride; // elements between consecutive draw indices in a noise stack (0 = single slice)
  const float* logq;         // [B,M]
  float* logq_out;           // [B,M]
  float* kappa_out;          // [B,M] or nullptr
  float* xstats_out;         // [B,2] (mean, rstd) or nullptr
  float* kappa_traj;         // [T,B,M] or nullptr (row = step)
  float* logq_traj;          // [T+1,B,M] or nullptr (row = step+1)
  const StepScalars* table;  // device table or nullptr
  const int* step_ptr;       // device step counter or nullptr
  StepScalars sc;            // used when table == nullptr
  float temperature;
  const float* bias;         // [M] or nullptr
  uint64_t seed;
  int64_t sample_offset;
  float* partials;           // [B][nblk][kPartialsPerBlock]
  // "AND" mode: kappa solved per sample from a first reduction pass instead of the softmax
  float* and_partials;       // [B][nblk][kAndPartialsPerBlock] or nullptr
  float* kappa_in;           // [B][M]: written by superpose_and_solve_kernel, read by the update / finalize kernels
  int mode;                  // 0 = OR (softmax of log q), 1 = AND (equal log-density increments)
  int B, D, M, nblk;
};

constexpr int kPartialsPerBlock = 3 * kMaxModels + 2;
constexpr int kAndPartialsPerBlock = kMaxModels * (kMaxModels + 1) / 2 + 2 * kMaxModels;  // G_ij (i <= j), <eps,x>, <eps,z>

// ------------------------------------------------------------------------------------- Philox
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)r * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
// 4 standard normals for elements 4q..4q+3 of (global sample, draw); definition in oracle.philox_normal.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t q, uint32_t gsample, uint32_t draw) {
  uint4 r = philox4x32_10(make_uint4(q, gsample, draw, 0x5D1FFu), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float4 o;
  // accurate logf (the radius is ill-conditioned near u = 1); fast sqrt and sin/cos (angle in (-pi, pi], where the
  // MUFU approximations are good to ~4e-7 absolute).  Normals agree with the float64 oracle to < 5e-6.
  float rad0, rad1;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(rad0) : "f"(-2.0f * logf(u01(r.x))));
  asm("sqrt.approx.f32 %0, %1;" : "=f"(rad1) : "f"(-2.0f * logf(u01(r.z))));
  float s0, c0, s1, c1;
  __sincosf(3.14159265358979f * (2.0f * u01(r.y) - 1.0f), &s0, &c0);
  __sincosf(3.14159265358979f * (2.0f * u01(r.w) - 1.0f), &s1, &c1);
  o.x = rad0 * s0; o.y = rad0 * c0; o.z = rad1 * s1; o.w = rad1 * c1;
  return o;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x' for one element, with the reference's expression tree and no FMA contraction (ddpm.py:42-44):
//   (1/sqrt(alpha)) * (x - ((1-alpha)/sqrt(1-alpha_bar)) * eps_bar) + sqrt(beta) * z
__device__ __forceinline__ float ddpm_x_update(float x, float eb, float z, float c1, float c2, float c3) {
  return __fadd_rn(__fmul_rn(c1, __fsub_rn(x, __fmul_rn(c2, eb))), __fmul_rn(c3, z));
}

// One CTA = one segment of one sample = STEPS sub-steps of kSegVecPerThread float4 per thread and array (STEPS = 2:
// 4096 elements).  Per sub-step all loads are issued before any use; the per-thread sums are carried across the
// sub-steps, so the ~320 instructions of per-thread prologue + reduction are paid once per 32 elements (they were 43 %
// of all issued instructions at 8 elements per thread).  The first sub-step's x / eps loads are issued BEFORE the
// dependent chain step counter -> schedule row -> log q -> kappa is walked (two L2 round trips).  Segment boundaries
// and every reduction order are fixed functions of D alone, so logq / statistics are bit-identical for any batch
// sharding.  No atomics, no fences: per-segment partial sums go to a small buffer and superpose_finalize_kernel (one
// CTA per sample) reduces them in a fixed order.
// Measured (B200, rotating buffers, D = 65536, M = 2; us per launch Philox / noise tensor): STEPS 1: B=64 18.8 / 16.5,
// B=256 59.5 / 51.9; STEPS 2: 18.4 / 17.4, 53.0 / 53.3; STEPS 4: 18.5 / 17.5, 52.4 / 53.6 (B=16: 8.2, 8.2, 10.3).
// Tried and rejected: persistent CTAs (2-3 per SM) fed by a cp.async.bulk + mbarrier shared-memory ring -- 22.3 / 19.1
// at B=64 and 65.4 / 58.0 at B=256: the Philox variant is issue-bound and wants the 32 warps per SM this version has.
constexpr int kSegVecPerThread = 2;
constexpr int kSegSteps = 2;
constexpr int kSubVec = kUpdThreads * kSegVecPerThread;  // float4 per sub-step
constexpr int kSegVec = kSubVec * kSegSteps;             // float4 per segment

template <int M>
__device__ __forceinline__ void softmax_kappa(const UpdateArgs& a, int b, float (&kap)[M]) {
  if (a.mode == 1) {  // AND: solved by superpose_and_solve_kernel earlier on the stream
#pragma unroll
    for (int m = 0; m < M; ++m) kap[m] = a.kappa_in[b * M + m];
    return;
  }
  float lg[M], mx = -INFINITY;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    lg[m] = a.temperature * a.logq[b * M + m] + (a.bias ? a.bias[m] : 0.0f);
    mx = fmaxf(mx, lg[m]);
  }
  // fast exp / divide (relative error ~2^-21, far inside the 2e-6 kappa tolerance; exp(0) = 1 and 1/2 stay exact, so
  // self-superposition still yields kappa == 0.5 bit-for-bit).  The update and the finalize kernel share this function.
  float den = 0.0f;
#pragma unroll
  for (int m = 0; m < M; ++m) { kap[m] = __expf(lg[m] - mx); den += kap[m]; }
#pragma unroll
  for (int m = 0; m < M; ++m) kap[m] = __fdividef(kap[m], den);
}

template <int M, int STEPS>
__global__ void __launch_bounds__(kUpdThreads, STEPS == 1 ? 5 : 4) superpose_update_kernel(const UpdateArgs a) {
  const int b = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nq = a.D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(a.x_in + (size_t)b * a.D);
  const float4* e4 = reinterpret_cast<const float4*>(a.eps + (size_t)b * a.D);
  const size_t e_stride = (size_t)a.B * (size_t)nq;  // float4 between models
  float4* xo4 = reinterpret_cast<float4*>(a.x_out + (size_t)b * a.D);
  const int q0 = seg * (kSubVec * STEPS) + tid;

  float4 xv[kSegVecPerThread], ev[M][kSegVecPerThread], zv[kSegVecPerThread];
  auto load_xe = [&](int s) {
#pragma unroll
    for (int i = 0; i < kSegVecPerThread; ++i) {
      const int q = q0 + s * kSubVec + i * kUpdThreads;
      const bool ok = q < nq;
      xv[i] = ok ? __ldcs(x4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int m = 0; m < M; ++m) ev[m][i] = ok ? __ldcs(e4 + m * e_stride + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_xe(0);  // in flight while the scalar chain below is walked

  const int step = a.step_ptr ? *a.step_ptr : 0;
  const StepScalars sc = a.table ? a.table[step] : a.sc;
  const bool have_noise = sc.draw_index >= 0;
  const bool noise_tensor = a.noise != nullptr && have_noise;
  const float4* n4 = noise_tensor ? reinterpret_cast<const float4*>(a.noise + (size_t)sc.draw_index * a.noise_step_stride + (size_t)b * a.D) : nullptr;
  float kap[M];
  softmax_kappa<M>(a, b, kap);
  const float c1 = sc.c1, c2 = sc.c2, c3 = sc.c3;
  const uint32_t gsample = (uint32_t)(a.sample_offset + b);

  float accA[M], accB[M], accC[M], sx = 0.0f, sxx = 0.0f;
#pragma unroll
  for (int m = 0; m < M; ++m) accA[m] = accB[m] = accC[m] = 0.0f;

#pragma unroll 1
  for (int s = 0; s < STEPS; ++s) {
    if (s > 0) load_xe(s);
#pragma unroll
    for (int i = 0; i < kSegVecPerThread; ++i) {
      const int q = q0 + s * kSubVec + i * kUpdThreads;
      zv[i] = (noise_tensor && q < nq) ? __ldcs(n4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kSegVecPerThread; ++i) {
      const int q = q0 + s * kSubVec + i * kUpdThreads;
      if (q >= nq) continue;
      float4 z = zv[i];
      if (have_noise && !noise_tensor) z = philox_normal4(a.seed, (uint32_t)q, gsample, (uint32_t)sc.draw_index);
      const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      const float zs[4] = {z.x, z.y, z.z, z.w};
      float xn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float es[M];
#pragma unroll
        for (int m = 0; m < M; ++m)
          es[m] = (j == 0 ? ev[m][i].x : j == 1 ? ev[m][i].y : j == 2 ? ev[m][i].z : ev[m][i].w);
        float eb = __fmul_rn(kap[0], es[0]);
#pragma unroll
        for (int m = 1; m < M; ++m) eb = __fadd_rn(eb, __fmul_rn(kap[m], es[m]));
        xn[j] = ddpm_x_update(xs[j], eb, zs[j], c1, c2, c3);
        const float dx = xn[j] - xs[j];
#pragma unroll
        for (int m = 0; m < M; ++m) {
          accA[m] = fmaf(es[m], dx, accA[m]);
          accB[m] = fmaf(xs[j], es[m], accB[m]);
          accC[m] = fmaf(es[m], es[m], accC[m]);
        }
        sx += xn[j];
        sxx = fmaf(xn[j], xn[j], sxx);
      }
      __stcs(xo4 + q, make_float4(xn[0], xn[1], xn[2], xn[3]));
    }
  }

  // segment reduce: shuffle within warps, fixed-order sum across the 8 warps
  __shared__ float red[kUpdThreads / 32][3 * M + 2];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    accA[m] = warp_sum(accA[m]); accB[m] = warp_sum(accB[m]); accC[m] = warp_sum(accC[m]);
  }
  sx = warp_sum(sx); sxx = warp_sum(sxx);
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < M; ++m) { red[warp][3 * m] = accA[m]; red[warp][3 * m + 1] = accB[m]; red[warp][3 * m + 2] = accC[m]; }
    red[warp][3 * M] = sx; red[warp][3 * M + 1] = sxx;
  }
  __syncthreads();
  if (tid < 3 * M + 2) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kUpdThreads / 32; ++w) v += red[w][tid];
    a.partials[((size_t)b * a.nblk + seg) * kPartialsPerBlock + tid] = v;
  }
}

// ------------------------------------------------------------------------------------------------ AND mode
// First pass of the "AND" step (oracle.and_kappa): per-sample G_ij = <eps_i, eps_j>, a_i = <eps_i, x>, b_i = <eps_i, z>
// with the update kernel's segmentation and reduction tree (shard-invariant, no atomics).  Reads x, eps, z once more
// than the OR step: 4 + 4M (+ 4 with a noise tensor) extra bytes per element, L2 hits at the sampler's sizes.
template <int M>
__global__ void __launch_bounds__(kUpdThreads, 4) superpose_and_gram_kernel(const UpdateArgs a) {
  constexpr int NG = M * (M + 1) / 2, NVAL = NG + 2 * M;
  const int b = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nq = a.D >> 2;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const StepScalars sc = a.table ? a.table[step] : a.sc;
  const bool have_noise = sc.draw_index >= 0;
  const bool noise_tensor = a.noise != nullptr && have_noise;
  const float4* x4 = reinterpret_cast<const float4*>(a.x_in + (size_t)b * a.D);
  const float4* e4 = reinterpret_cast<const float4*>(a.eps + (size_t)b * a.D);
  const size_t e_stride = (size_t)a.B * (size_t)nq;
  const float4* n4 = noise_tensor ? reinterpret_cast<const float4*>(a.noise + (size_t)sc.draw_index * a.noise_step_stride + (size_t)b * a.D) : nullptr;
  const uint32_t gsample = (uint32_t)(a.sample_offset + b);
  const int per_seg = nq / a.nblk + ((nq % a.nblk) ? 1 : 0);  // == kSubVec * steps for full segments
  float acc[NVAL];
#pragma unroll
  for (int j = 0; j < NVAL; ++j) acc[j] = 0.0f;
  const int q_end = min(nq, (seg + 1) * per_seg);
#pragma unroll 1
  for (int q = seg * per_seg + tid; q < q_end; q += kUpdThreads) {
    const float4 xv = __ldg(x4 + q);
    float4 ev[M];
#pragma unroll
    for (int m = 0; m < M; ++m) ev[m] = __ldg(e4 + m * e_stride + q);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (noise_tensor) z = __ldg(n4 + q);
    else if (have_noise) z = philox_normal4(a.seed, (uint32_t)q, gsample, (uint32_t)sc.draw_index);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const float zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float es[M];
#pragma unroll
      for (int m = 0; m < M; ++m) es[m] = (j == 0 ? ev[m].x : j == 1 ? ev[m].y : j == 2 ? ev[m].z : ev[m].w);
      int g = 0;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int k = i; k < M; ++k) { acc[g] = fmaf(es[i], es[k], acc[g]); ++g; }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        acc[NG + m] = fmaf(es[m], xs[j], acc[NG + m]);
        acc[NG + M + m] = fmaf(es[m], zs[j], acc[NG + M + m]);
      }
    }
  }
  __shared__ float red[kUpdThreads / 32][NVAL];
#pragma unroll
  for (int j = 0; j < NVAL; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NVAL; ++j) red[warp][j] = acc[j];
  }
  __syncthreads();
  if (tid < NVAL) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kUpdThreads / 32; ++w) v += red[w][tid];
    a.and_partials[((size_t)b * a.nblk + seg) * kAndPartialsPerBlock + tid] = v;
  }
}

// One CTA per sample: fixed-order reduce of the Gram partials, then the M x M solve in double (Gaussian elimination
// with partial pivoting); a singular system (identical models) yields the uniform weights, as in the oracle.
template <int M>
__global__ void __launch_bounds__(256) superpose_and_solve_kernel(const UpdateArgs a) {
  constexpr int NG = M * (M + 1) / 2, NVAL = NG + 2 * M;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const StepScalars sc = a.table ? a.table[step] : a.sc;
  __shared__ float fin[8][32];
  __shared__ double tot[NVAL];
  {
    const int j = tid & 31, g = tid >> 5;
    float v = 0.0f;
    if (j < NVAL) {
      const float* src = a.and_partials + (size_t)b * a.nblk * kAndPartialsPerBlock + j;
      for (int p = g; p < a.nblk; p += 8) v += src[(size_t)p * kAndPartialsPerBlock];
    }
    fin[g][j] = v;
  }
  __syncthreads();
  if (tid < NVAL) {
    double v = 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += (double)fin[g][tid];
    tot[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double G[M][M], av[M], bv[M];
    int g = 0;
    for (int i = 0; i < M; ++i)
      for (int k = i; k < M; ++k) { G[i][k] = G[k][i] = tot[g]; ++g; }
    for (int m = 0; m < M; ++m) { av[m] = tot[NG + m]; bv[m] = tot[NG + M + m]; }
    const double c1 = (double)sc.c1, c2 = (double)sc.c2, c3 = (double)sc.c3, beta = (double)sc.beta;
    const double sig = sqrt(1.0 - (double)sc.alpha_bar);
    double r[M];
    for (int i = 0; i < M; ++i)
      r[i] = -((c1 - 1.0) * av[i] + c3 * bv[i]) / sig + beta / (2.0 * sig) * av[i] - beta / (2.0 * sig * sig) * G[i][i];
    double A[M][M + 1];
    for (int j = 0; j < M; ++j) A[0][j] = 1.0;
    A[0][M] = 1.0;
    const double coef = c1 * c2 / sig;
    double scale = 1.0;
    for (int i = 1; i < M; ++i) {
      for (int j = 0; j < M; ++j) { A[i][j] = coef * (G[i][j] - G[0][j]); scale = fmax(scale, fabs(A[i][j])); }
      A[i][M] = r[0] - r[i];
    }
    bool singular = false;
    for (int c = 0; c < M && !singular; ++c) {
      int piv = c;
      for (int i = c + 1; i < M; ++i) if (fabs(A[i][c]) > fabs(A[piv][c])) piv = i;
      if (fabs(A[piv][c]) <= 1e-12 * scale) { singular = true; break; }
      if (piv != c) for (int j = 0; j <= M; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
      for (int i = c + 1; i < M; ++i) {
        const double f = A[i][c] / A[c][c];
        for (int j = c; j <= M; ++j) A[i][j] -= f * A[c][j];
      }
    }
    double kap[M];
    if (!singular) {
      for (int i = M - 1; i >= 0; --i) {
        double v = A[i][M];
        for (int j = i + 1; j < M; ++j) v -= A[i][j] * kap[j];
        kap[i] = v / A[i][i];
      }
    } else {
      for (int m = 0; m < M; ++m) kap[m] = 1.0 / M;
    }
    for (int m = 0; m < M; ++m) a.kappa_in[b * M + m] = (float)kap[m];
  }
}

// One CTA (256 threads) per sample: fixed-order reduce of the segment partials (16 strided lanes per value, then the
// 16 lanes in order), Ito log-density increment in double, kappa, GroupNorm(1,1) statistics of x'.
template <int M>
__global__ void __launch_bounds__(256) superpose_finalize_kernel(const UpdateArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int step = a.step_ptr ? *a.step_ptr : 0;
  const StepScalars sc = a.table ? a.table[step] : a.sc;
  __shared__ float fin[16][16];
  __shared__ double tot[3 * M + 2];
  {
    const int j = tid & 15, g = tid >> 4;
    float v = 0.0f;
    if (j < 3 * M + 2) {
      const float* src = a.partials + (size_t)b * a.nblk * kPartialsPerBlock + j;
      for (int p = g; p < a.nblk; p += 16) v += src[(size_t)p * kPartialsPerBlock];
    }
    fin[g][j] = v;
  }
  __syncthreads();
  if (tid < 3 * M + 2) {
    double v = 0.0;
#pragma unroll
    for (int g = 0; g < 16; ++g) v += (double)fin[g][tid];
    tot[tid] = v;
  }
  __syncthreads();
  if (tid < M) {
    float kap[M];
    softmax_kappa<M>(a, b, kap);
    const double beta = (double)sc.beta;
    const double inv_sig = 1.0 / sqrt(1.0 - (double)sc.alpha_bar);
    const double A = tot[3 * tid], Bx = tot[3 * tid + 1], C = tot[3 * tid + 2];
    // s = -eps * inv_sig:  <s,dx> - beta D/2 - beta/2 <x,s> - beta/2 |s|^2
    const double inc = -inv_sig * A - 0.5 * beta * (double)a.D + 0.5 * beta * inv_sig * Bx - 0.5 * beta * inv_sig * inv_sig * C;
    const float lq_new = (float)((double)a.logq[b * M + tid] + inc);
    float kv = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) if (m == tid) kv = kap[m];
    __syncwarp((1u << M) - 1u);  // every lane has read logq[b,:] (for kappa) before any lane overwrites it in place
    a.logq_out[b * M + tid] = lq_new;
    if (a.kappa_out) a.kappa_out[b * M + tid] = kv;
    if (a.kappa_traj) a.kappa_traj[((size_t)step * a.B + b) * M + tid] = kv;
    if (a.logq_traj) a.logq_traj[((size_t)(step + 1) * a.B + b) * M + tid] = lq_new;
  }
  if (tid == 32 && a.xstats_out) {
    const double mean = tot[3 * M] / (double)a.D;
    double var = tot[3 * M + 1] / (double)a.D - mean * mean;
    if (var < 0.0) var = 0.0;
    a.xstats_out[b * 2 + 0] = (float)mean;
    a.xstats_out[b * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
  }
}

// Segments per sample depend on D only (steps * 2048 elements each).
inline int update_blocks_per_sample(int D, int steps = kSegSteps) {
  int nq = D / 4;
  int nb = (nq + kSubVec * steps - 1) / (kSubVec * steps);
  return nb < 1 ? 1 : nb;
}

// workspace = [update partials | AND partials | AND kappa[B][kMaxModels]], each 256-byte aligned
inline size_t update_ws_part_bytes(int B, int D) {
  return ((size_t)B * update_blocks_per_sample(D, 1) * kPartialsPerBlock * sizeof(float) + 255) & ~(size_t)255;  // any steps
}
inline size_t update_ws_and_bytes(int B, int D) {
  return ((size_t)B * update_blocks_per_sample(D, 1) * kAndPartialsPerBlock * sizeof(float) + 255) & ~(size_t)255;
}
inline size_t update_workspace_bytes(int B, int D, int /*M*/) {
  return update_ws_part_bytes(B, D) + update_ws_and_bytes(B, D) + (((size_t)B * kMaxModels * sizeof(float) + 255) & ~(size_t)255);
}

int launch_superpose_update(UpdateArgs a, void* workspace, cudaStream_t stream, cudaEvent_t after_update, bool finalize);

__global__ void philox_normal_kernel(float* out, int B, int D, uint64_t seed, int64_t sample_offset, int draw);
__global__ void copy_f32_kernel(float* dst, const float* src, size_t n);

}  // namespace sdd
