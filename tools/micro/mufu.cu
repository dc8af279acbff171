// MUFU throughput microbenchmark: cycles per warp-instruction with 1 warp per SMSP (4 warps / CTA, 1 CTA).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
  uint32_t b[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = __float_as_uint(a[i]); }
  long long t0 = clock64();
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 6) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(b[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(b[i]);
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
  const char* names[] = {"tanh.f32", "ex2.f32", "rcp.f32", "tanh.bf16x2", "ex2.bf16x2", "fma.f32", "fma.bf16x2"};
  for (int warps = 4; warps <= 8; warps += 4) {
    for (int m = 0; m < 7; ++m) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) {
          case 0: k<0><<<1, warps * 32>>>(o, c, 0.5f); break; case 1: k<1><<<1, warps * 32>>>(o, c, 0.5f); break;
          case 2: k<2><<<1, warps * 32>>>(o, c, 0.5f); break; case 3: k<3><<<1, warps * 32>>>(o, c, 0.5f); break;
          case 4: k<4><<<1, warps * 32>>>(o, c, 0.5f); break; case 5: k<5><<<1, warps * 32>>>(o, c, 0.5f); break;
          case 6: k<6><<<1, warps * 32>>>(o, c, 0.5f); break;
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      }
      printf("%d warps/CTA %-12s: %.2f cycles per warp-instruction (per SMSP: %d warp%s)\n", warps, names[m], h / (256.0 * 8), warps / 4, warps > 4 ? "s" : "");
    }
  }
  return 0;
}
