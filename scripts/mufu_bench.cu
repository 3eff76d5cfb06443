// Micro-benchmark: MUFU.EX2 / FFMA / f16x2 ex2 throughput per SM on this GPU (run under gpurun).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  unsigned h[8];
  for (int i = 0; i < 8; ++i) h[i] = 0x3c003c00u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 5 && (i & 1) == 0) { float2 t = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(a[i], a[i + 1]), make_float2(a[i + 1], a[i])); a[i] = t.x; a[i + 1] = t.y; }
      if (MODE == 6) { if (i < 6) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_op) {
  float* out; cudaMalloc(&out, 148 * 1024 * 8 * sizeof(float));
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm : {1, 2}) {
    k<MODE><<<148 * blocks_per_sm, 1024>>>(out, 16);
    cudaEventRecord(e0);
    k<MODE><<<148 * blocks_per_sm, 1024>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ops = (double)148 * blocks_per_sm * 1024 * 8.0 * iters * per_op;
    printf("%-28s blocks/SM=%d: %.3f ms  %.2f results/clk/SM (at %d MHz nominal)\n", name, blocks_per_sm, ms,
           ops / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
  }
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<4>("ex2.approx.f32", 1);
  run<2>("ex2.approx.ftz.f16x2", 2);
  run<3>("lg2.approx.ftz.f32", 1);
  run<1>("fma.rn.f32", 1);
  run<5>("fma.rn.f32x2 (FFMA2)", 1);
  run<6>("6 FFMA + 2 EX2 interleaved", 1);
  return 0;
}
