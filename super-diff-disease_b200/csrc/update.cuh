// Fused superposition update: ONE launch per step, one coalesced, vectorised HBM pass.
//   reads  x[B,D], eps[M,B,D], (noise[B,D] | Philox)      writes x'[B,D]
//   per-sample reductions <eps_m,dx>, <x,eps_m>, |eps_m|^2, sum x', sum x'^2 via warp shuffles -> per-segment partials,
//   reduced in ONE canonical order (common.cuh: reduce_partials) -> logq', kappa, GN(1,1) statistics of x':
//     operator API: by the last-arriving CTA of each sample, inside the launch (results are there when it returns);
//     sampler step graph (defer): by the prologue of the NEXT step's launch and by the next forward's first layer, so
//     that the launch has no serial tail; the last CTA to have read the device-side step counter advances it.
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

// Per-CALL values.  The sampler keeps one copy in device memory and rewrites it at the start of every run, so the
// captured step graph (which only holds the pointer) survives a new seed / shard offset / trajectory buffer; the
// operator API passes the struct by value.
struct RunParams {
  const float* noise;        // [B,D] slice (operator API), noise stack [T,B,D] (sampler; draw_index selects) or nullptr
  int64_t noise_step_stride; // elements between consecutive draw indices of a noise stack (0 = single slice)
  uint64_t seed;             // Philox key when noise == nullptr
  int64_t sample_offset;     // global index of local sample 0
  float temperature;
  int noise_ring;            // > 0: `noise` is a RING of this many slices (slot = draw index % noise_ring): the sampler's
                             // streamed host-noise path refills it in step-range chunks while earlier steps compute
  const float* bias;         // [M] or nullptr
  float* kappa_traj;         // [T,B,M] or nullptr (row = step)
  float* logq_traj;          // [T+1,B,M] or nullptr (row = step+1)
  float* x_traj;             // [T+1,B,D] or nullptr (row = step+1 receives x'; row 0 = x_T is written by the sampler)
};

__device__ __forceinline__ int noise_slot(int draw_index, int ring) { return ring > 0 ? draw_index % ring : draw_index; }

struct UpdateArgs {
  const float* x_in;
  float* x_out;
  const float* eps;          // [M,B,D]
  const float* logq;         // [B,M]
  float* logq_out;           // [B,M]
  float* kappa_out;          // [B,M] or nullptr
  float* xstats_out;         // [B,2] (mean, rstd) or nullptr
  const StepScalars* table;  // device table or nullptr
  int* step_ptr;             // device step counter or nullptr; advanced by the kernel when `advance_step`
  StepScalars sc;            // used when table == nullptr
  const RunParams* rp;       // device copy (sampler) or nullptr
  RunParams rv;              // used when rp == nullptr
  float* partials;           // [B][nblk][kPartialsPerBlock]  (defer: two such buffers, parity_stride floats apart)
  int* counters;             // [B + 2]: per-sample arrivals, finalised samples, CTAs that have read the step counter
  // defer (the sampler's step graph): a sample's finalisation -- reduce the partials, Ito increment, log q -- is NOT done
  // by the launch that produced the partials but in the prologue of the NEXT step's launch, where it overlaps the x / eps
  // loads already in flight; the launch then ends with its partial stores (no fence, no arrival, no serial tail).
  // logq is double-buffered by step parity ([2][B][M]); sdd_sampler_run closes a run with finish_run_kernel.
  int defer;
  size_t parity_stride;      // floats between the two partial buffers
  // "AND" mode: kappa solved per sample from a first reduction pass instead of the softmax
  float* and_partials;       // [B][nblk][kAndPartialsPerBlock] or nullptr
  float* kappa_in;           // [B][M]: written by superpose_and_solve_kernel, read by the update kernel
  int mode;                  // 0 = OR (softmax of log q), 1 = AND (equal log-density increments)
  int advance_step;
  int B, D, M, nblk;
};

constexpr int kPartialsPerBlock = 3 * kMaxModels + 2;
constexpr int kAndPartialsPerBlock = kMaxModels * (kMaxModels + 1) / 2 + 2 * kMaxModels;  // G_ij (i <= j), <eps,x>, <eps,z>

__device__ __forceinline__ RunParams load_run_params(const UpdateArgs& a) {
  if (a.rp) return *a.rp;
  return a.rv;
}

// ------------------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al., SC'11).  mul.wide gives hi:lo of a round's product in one IMAD.WIDE.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c.z;
    c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)r * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
// -2 ln(u) for u in (0, 1].  MUFU.LG2 (absolute error 2^-22 on [0.5, 2), relative elsewhere) is accurate enough except
// next to u = 1, where ln u -> 0 and the radius sqrt(-2 ln u) is ill-conditioned: there (u > 0.998, 0.2 % of the
// draws, warp-divergent but rare) a three-term series in t = 1 - u (exact in fp32) is used instead.  The radius then
// agrees with the float64 oracle to < 3e-6 absolute everywhere (the accurate logf this replaces cost ~40 of the ~180
// instructions of a four-normal draw).
__device__ __forceinline__ float neg2_log(float u) {
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
  float v = -1.3862943611198906f * l2;
  if (u > 0.998f) {
    const float t = 1.0f - u;
    v = 2.0f * t * fmaf(t, fmaf(t, 0.3333333333f, 0.5f), 1.0f);
  }
  return v;
}
// 4 standard normals for elements 4q..4q+3 of (global sample, draw); definition in oracle.philox_normal.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t q, uint32_t gsample, uint32_t draw) {
  uint4 r = philox4x32_10(make_uint4(q, gsample, draw, 0x5D1FFu), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float4 o;
  // fast sqrt and sin / cos (angle in (-pi, pi], where the MUFU approximations are good to ~5e-7 absolute)
  float rad0, rad1;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad0) : "f"(neg2_log(u01(r.x))));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad1) : "f"(neg2_log(u01(r.z))));
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

// packed fp32 pairs without contraction: the reference's expression tree (ddpm.py:42-44) is mul, mul, sub, mul, add
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// One CTA = one segment of one sample = STEPS sub-steps of kSegVecPerThread float4 per thread and array (STEPS = 2:
// 4096 elements).  Per sub-step all loads are issued before any use; the per-thread sums are carried across the
// sub-steps.  The first sub-step's x / eps loads are issued BEFORE the dependent chain step counter -> schedule row ->
// log q -> kappa is walked (two L2 round trips).  Segment boundaries and every reduction order are fixed functions of
// D alone, so logq / statistics are bit-identical for any batch sharding.  No floating-point atomics: per-segment
// partial sums go to a small buffer, an integer arrival counter per sample elects the last CTA, and that CTA reduces
// the partials in a fixed order.  The arithmetic of an element runs on packed fp32 pairs (FMUL2 / FADD2 / FFMA2): two
// elements per instruction with the scalar code's per-element rounding (mul, mul, sub, mul, add -- no contraction).
constexpr int kSegVecPerThread = 2;
constexpr int kSegSteps = 2;
constexpr int kSubVec = kUpdThreads * kSegVecPerThread;  // float4 per sub-step
constexpr int kSegVec = kSubVec * kSegSteps;             // float4 per segment

template <int M>
__device__ __forceinline__ void softmax_kappa(const UpdateArgs& a, const RunParams& rp, int b, float (&kap)[M]) {
  if (a.mode == 1) {  // AND: solved by superpose_and_solve_kernel earlier on the stream
#pragma unroll
    for (int m = 0; m < M; ++m) kap[m] = a.kappa_in[b * M + m];
    return;
  }
  float lg[M], mx = -INFINITY;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    lg[m] = rp.temperature * a.logq[b * M + m] + (rp.bias ? rp.bias[m] : 0.0f);
    mx = fmaxf(mx, lg[m]);
  }
  // fast exp / divide (relative error ~2^-21, far inside the 2e-6 kappa tolerance; exp(0) = 1 and 1/2 stay exact, so
  // self-superposition still yields kappa == 0.5 bit-for-bit).
  float den = 0.0f;
#pragma unroll
  for (int m = 0; m < M; ++m) { kap[m] = __expf(lg[m] - mx); den += kap[m]; }
#pragma unroll
  for (int m = 0; m < M; ++m) kap[m] = __fdividef(kap[m], den);
}

// Ito log-density increment of one model and step from the three per-sample sums A = <eps, dx>, Bx = <x, eps>,
// C = |eps|^2, with s = -eps / sqrt(1 - alpha_bar):  <s,dx> - beta D/2 - beta/2 <x,s> - beta/2 |s|^2, in double.
__device__ __forceinline__ double ito_increment(const StepScalars& sc, int D, double A, double Bx, double C) {
  const double beta = (double)sc.beta;
  const double inv_sig = 1.0 / sqrt(1.0 - (double)sc.alpha_bar);
  return -inv_sig * A - 0.5 * beta * (double)D + 0.5 * beta * inv_sig * Bx - 0.5 * beta * inv_sig * inv_sig * C;
}

// Immediate finalisation (operator API): the canonical reduce of one sample's segment partials (common.cuh) by its
// last-arriving CTA, Ito increment, kappa / log q trajectory rows, GroupNorm(1,1) statistics of x'.  `kap` is the kappa
// this step used (every CTA of the sample computed the same values).
template <int M>
__device__ __forceinline__ void finalize_sample(const UpdateArgs& a, const RunParams& rp, const StepScalars& sc, int step,
                                                int b, const float (&kap)[M]) {
  const int tid = threadIdx.x;
  __shared__ double tot[3 * M + 2];
  if (tid < 3 * M + 2)
    tot[tid] = reduce_partials(a.partials + (size_t)b * a.nblk * kPartialsPerBlock, a.nblk, kPartialsPerBlock, tid);
  __syncthreads();
  if (tid < M) {
    const float lq_new = (float)((double)a.logq[b * M + tid] + ito_increment(sc, a.D, tot[3 * tid], tot[3 * tid + 1], tot[3 * tid + 2]));
    float kv = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) if (m == tid) kv = kap[m];
    a.logq_out[b * M + tid] = lq_new;  // in place is fine: every CTA of this sample read logq before it arrived
    if (a.kappa_out) a.kappa_out[b * M + tid] = kv;
    if (rp.kappa_traj) rp.kappa_traj[((size_t)step * a.B + b) * M + tid] = kv;
    if (rp.logq_traj) rp.logq_traj[((size_t)(step + 1) * a.B + b) * M + tid] = lq_new;
  }
  if (tid == 32 && a.xstats_out) {
    const double mean = tot[3 * M] / (double)a.D;
    double var = tot[3 * M + 1] / (double)a.D - mean * mean;
    if (var < 0.0) var = 0.0;
    a.xstats_out[b * 2 + 0] = (float)mean;
    a.xstats_out[b * 2 + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
  }
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
  const RunParams rp = load_run_params(a);
  const bool have_noise = sc.draw_index >= 0;
  const bool noise_tensor = rp.noise != nullptr && have_noise;
  const float4* n4 = noise_tensor ? reinterpret_cast<const float4*>(rp.noise + (size_t)noise_slot(sc.draw_index, rp.noise_ring) * rp.noise_step_stride + (size_t)b * a.D) : nullptr;
  float kap[M];
  if (a.defer) {
    // ---- the previous step's finalisation, here: log q_k = log q_{k-1} + increment(partials of step k-1).  Every CTA of
    // the sample computes the same values (canonical reduction order); segment 0 publishes them.  Then the step counter:
    // the CTA that is the last of the launch to have READ it advances it (nobody can still observe the old value).
    __shared__ float s_lq[M];
    __shared__ float s_stage[kPartialStage][3 * M];
    __shared__ double s_tot[3 * M];
    const float* lq_prev = a.logq + (size_t)((step + 1) & 1) * a.B * M;   // parity (step - 1) & 1
    float* lq_cur = a.logq_out + (size_t)(step & 1) * a.B * M;
    if (step > 0) {
      // one independent load per thread (a single L2 round trip), then the canonical sequential sums from shared memory
      const float* pb = a.partials + (size_t)((step + 1) & 1) * a.parity_stride + (size_t)b * a.nblk * kPartialsPerBlock;
      if (tid < 3 * M) s_tot[tid] = 0.0;
      for (int p0 = 0; p0 < a.nblk; p0 += kPartialStage) {
        const int n = min(kPartialStage, a.nblk - p0);
        __syncthreads();
        for (int i = tid; i < n * 3 * M; i += kUpdThreads) {
          const int p = i / (3 * M), j = i - p * (3 * M);
          s_stage[p][j] = __ldcg(pb + (size_t)(p0 + p) * kPartialsPerBlock + j);
        }
        __syncthreads();
        if (tid < 3 * M) {
          double v = s_tot[tid];
          for (int p = 0; p < n; ++p) v += (double)s_stage[p][tid];
          s_tot[tid] = v;
        }
      }
      __syncthreads();
    }
    if (tid < M) {
      float lq = 0.0f;
      if (step > 0)
        lq = (float)((double)lq_prev[b * M + tid] +
                     ito_increment(a.table[step - 1], a.D, s_tot[3 * tid], s_tot[3 * tid + 1], s_tot[3 * tid + 2]));
      s_lq[tid] = lq;
      if (seg == 0) {
        lq_cur[b * M + tid] = lq;
        if (rp.logq_traj && step > 0) rp.logq_traj[((size_t)step * a.B + b) * M + tid] = lq;
      }
    }
    __syncthreads();
    // (every thread of this CTA has issued its schedule-row load by now, whose address needs the step counter's value: the
    // CTA's reads of the counter are complete before it reports in)
    if (tid == 32 && a.advance_step) {
      const int total = (int)(gridDim.x * gridDim.y);
      if (atomicAdd(&a.counters[a.B + 1], 1) == total - 1) {
        a.counters[a.B + 1] = 0;
        *a.step_ptr = step + 1;
      }
    }
    if (a.mode == 1) {
#pragma unroll
      for (int m = 0; m < M; ++m) kap[m] = a.kappa_in[b * M + m];
    } else {
      float lg[M], mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < M; ++m) { lg[m] = rp.temperature * s_lq[m] + (rp.bias ? rp.bias[m] : 0.0f); mx = fmaxf(mx, lg[m]); }
      float den = 0.0f;
#pragma unroll
      for (int m = 0; m < M; ++m) { kap[m] = __expf(lg[m] - mx); den += kap[m]; }
#pragma unroll
      for (int m = 0; m < M; ++m) kap[m] = __fdividef(kap[m], den);
    }
    if (seg == 0 && tid < M && rp.kappa_traj) {
      float kv = 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) if (m == tid) kv = kap[m];
      rp.kappa_traj[((size_t)step * a.B + b) * M + tid] = kv;
    }
  } else {
    softmax_kappa<M>(a, rp, b, kap);
  }
  const uint64_t c1 = pack_f32x2(sc.c1, sc.c1), c2 = pack_f32x2(sc.c2, sc.c2), c3 = pack_f32x2(sc.c3, sc.c3);
  uint64_t kap2[M];
#pragma unroll
  for (int m = 0; m < M; ++m) kap2[m] = pack_f32x2(kap[m], kap[m]);
  const uint32_t gsample = (uint32_t)(rp.sample_offset + b);
  float4* xt4 = rp.x_traj ? reinterpret_cast<float4*>(rp.x_traj + ((size_t)(step + 1) * a.B + b) * (size_t)a.D) : nullptr;

  // packed (even element, odd element) accumulators
  uint64_t accA[M], accB[M], accC[M], sx = 0ull, sxx = 0ull;
#pragma unroll
  for (int m = 0; m < M; ++m) accA[m] = accB[m] = accC[m] = 0ull;

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
      if (have_noise && !noise_tensor) z = philox_normal4(rp.seed, (uint32_t)q, gsample, (uint32_t)sc.draw_index);
      uint64_t xn2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // elements (0,1) then (2,3) of the float4
        const uint64_t x2 = h ? pack_f32x2(xv[i].z, xv[i].w) : pack_f32x2(xv[i].x, xv[i].y);
        const uint64_t z2 = h ? pack_f32x2(z.z, z.w) : pack_f32x2(z.x, z.y);
        uint64_t e2[M];
#pragma unroll
        for (int m = 0; m < M; ++m) e2[m] = h ? pack_f32x2(ev[m][i].z, ev[m][i].w) : pack_f32x2(ev[m][i].x, ev[m][i].y);
        uint64_t eb = mul_f32x2(kap2[0], e2[0]);
#pragma unroll
        for (int m = 1; m < M; ++m) eb = add_f32x2(eb, mul_f32x2(kap2[m], e2[m]));
        // (1/sqrt(alpha)) * (x - ((1-alpha)/sqrt(1-alpha_bar)) * eps_bar) + sqrt(beta) * z, ddpm.py:42-44
        const uint64_t xn = add_f32x2(mul_f32x2(c1, sub_f32x2(x2, mul_f32x2(c2, eb))), mul_f32x2(c3, z2));
        const uint64_t dx = sub_f32x2(xn, x2);
#pragma unroll
        for (int m = 0; m < M; ++m) {
          accA[m] = fma_f32x2(e2[m], dx, accA[m]);
          accB[m] = fma_f32x2(x2, e2[m], accB[m]);
          accC[m] = fma_f32x2(e2[m], e2[m], accC[m]);
        }
        sx = add_f32x2(sx, xn);
        sxx = fma_f32x2(xn, xn, sxx);
        xn2[h] = xn;
      }
      float4 o;
      unpack_f32x2(xn2[0], o.x, o.y);
      unpack_f32x2(xn2[1], o.z, o.w);
      __stcs(xo4 + q, o);
      if (xt4) __stcs(xt4 + q, o);  // optional trajectory copy (visualisation strips, parity tests)
    }
  }

  // segment reduce: (even + odd), shuffle within warps, fixed-order sum across the 8 warps
  __shared__ float red[kUpdThreads / 32][3 * M + 2];
  auto fold = [](uint64_t v) { float lo, hi; unpack_f32x2(v, lo, hi); return lo + hi; };
  float rA[M], rB[M], rC[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    rA[m] = warp_sum(fold(accA[m])); rB[m] = warp_sum(fold(accB[m])); rC[m] = warp_sum(fold(accC[m]));
  }
  const float rsx = warp_sum(fold(sx)), rsxx = warp_sum(fold(sxx));
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m < M; ++m) { red[warp][3 * m] = rA[m]; red[warp][3 * m + 1] = rB[m]; red[warp][3 * m + 2] = rC[m]; }
    red[warp][3 * M] = rsx; red[warp][3 * M + 1] = rsxx;
  }
  __syncthreads();
  if (tid < 3 * M + 2) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kUpdThreads / 32; ++w) v += red[w][tid];
    float* pdst = a.partials + (a.defer ? (size_t)(step & 1) * a.parity_stride : 0);
    __stcg(pdst + ((size_t)b * a.nblk + seg) * kPartialsPerBlock + tid, v);
    if (a.defer) return;  // consumed by the next launch (stream order): nothing to wait for, no tail
    __threadfence();      // this thread's partial is visible device-wide before the arrival below
  }
  if (a.defer) return;
  // ---- arrival: the last CTA of this sample finalises it (fixed reduction order: the result does not depend on which
  // CTA is last); the CTA that finalises the last sample advances the step counter (every CTA of the launch read it
  // before arriving, so nobody can still observe the old value)
  __shared__ int s_last;
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&a.counters[b], 1) == a.nblk - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  finalize_sample<M>(a, rp, sc, step, b, kap);
  if (tid == 0) {
    a.counters[b] = 0;  // ready for the next launch
    __threadfence();
    if (atomicAdd(&a.counters[a.B], 1) == a.B - 1) {
      a.counters[a.B] = 0;
      if (a.advance_step && a.step_ptr) *a.step_ptr = step + 1;
    }
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
  const RunParams rp = load_run_params(a);
  const bool have_noise = sc.draw_index >= 0;
  const bool noise_tensor = rp.noise != nullptr && have_noise;
  const float4* x4 = reinterpret_cast<const float4*>(a.x_in + (size_t)b * a.D);
  const float4* e4 = reinterpret_cast<const float4*>(a.eps + (size_t)b * a.D);
  const size_t e_stride = (size_t)a.B * (size_t)nq;
  const float4* n4 = noise_tensor ? reinterpret_cast<const float4*>(rp.noise + (size_t)noise_slot(sc.draw_index, rp.noise_ring) * rp.noise_step_stride + (size_t)b * a.D) : nullptr;
  const uint32_t gsample = (uint32_t)(rp.sample_offset + b);
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
    else if (have_noise) z = philox_normal4(rp.seed, (uint32_t)q, gsample, (uint32_t)sc.draw_index);
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

// Closes a deferred run (one launch per sampling call, not per step): log q_T from the last step's partials.
template <int M>
__global__ void finish_run_kernel(const UpdateArgs a, int T) {
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid >= M) return;
  const RunParams rp = load_run_params(a);
  const float* pb = a.partials + (size_t)((T - 1) & 1) * a.parity_stride + (size_t)b * a.nblk * kPartialsPerBlock;
  const float* lq_prev = a.logq + (size_t)((T - 1) & 1) * a.B * M;
  const float lq = (float)((double)lq_prev[b * M + tid] +
                           ito_increment(a.table[T - 1], a.D, reduce_partials(pb, a.nblk, kPartialsPerBlock, 3 * tid),
                                         reduce_partials(pb, a.nblk, kPartialsPerBlock, 3 * tid + 1),
                                         reduce_partials(pb, a.nblk, kPartialsPerBlock, 3 * tid + 2)));
  a.logq_out[(size_t)(T & 1) * a.B * M + b * M + tid] = lq;
  if (rp.logq_traj) rp.logq_traj[((size_t)T * a.B + b) * M + tid] = lq;
}

// Segments per sample depend on D only (steps * 2048 elements each).
inline int update_blocks_per_sample(int D, int steps = kSegSteps) {
  int nq = D / 4;
  int nb = (nq + kSubVec * steps - 1) / (kSubVec * steps);
  return nb < 1 ? 1 : nb;
}

// workspace = [update partials x 2 (step parity) | AND partials | AND kappa[B][kMaxModels] | counters[B + 2]], 256-byte aligned
inline size_t update_ws_part_bytes(int B, int D) {  // ONE parity buffer
  return ((size_t)B * update_blocks_per_sample(D, 1) * kPartialsPerBlock * sizeof(float) + 255) & ~(size_t)255;  // any steps
}
inline size_t update_ws_and_bytes(int B, int D) {
  return ((size_t)B * update_blocks_per_sample(D, 1) * kAndPartialsPerBlock * sizeof(float) + 255) & ~(size_t)255;
}
inline size_t update_ws_kappa_bytes(int B) { return ((size_t)B * kMaxModels * sizeof(float) + 255) & ~(size_t)255; }
inline size_t update_workspace_bytes(int B, int D, int /*M*/) {
  return 2 * update_ws_part_bytes(B, D) + update_ws_and_bytes(B, D) + update_ws_kappa_bytes(B) +
         ((((size_t)B + 2) * sizeof(int) + 255) & ~(size_t)255);
}

__global__ void philox_normal_kernel(float* out, int B, int D, uint64_t seed, int64_t sample_offset, int draw);
__global__ void copy_f32_kernel(float* dst, const float* src, size_t n);

}  // namespace sdd
