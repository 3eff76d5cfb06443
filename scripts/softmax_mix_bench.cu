// Micro-benchmark: exponentials per clock and SM for the instruction mix of the attention kernel's softmax inner loop
// (csrc/attn_tc.cu), in isolation: no TMEM, no MMA, no barriers.  One 512-thread block per SM (16 warps = the softmax group of the
// kernel), 64 scores per thread in registers, cycles from %clock64.  Variants strip the mix down to find which companion
// instruction costs MUFU throughput.   nvcc -arch=sm_100a -O3 -o softmax_mix_bench softmax_mix_bench.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// MODE bits: 1 = FFMA2 scale/shift, 2 = FADD2 row sum, 4 = F2FP pack, 8 = PRMT pack instead, 16 = FMNMX3 max pass, 32 = unpacked FFMA/FADD,
//            64 = no MUFU (x passed through)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* in, float* out, long long* cyc, int iters) {
  float v[64];
  for (int i = 0; i < 64; ++i) v[i] = in[(threadIdx.x * 64 + i) & 4095];
  float2 sc2 = make_float2(in[1], in[1]), nm2 = make_float2(in[2], in[2]);
  float2 rs2 = make_float2(0.f, 0.f);
  float mx = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE & 16) {
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 64; i += 2) m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], v[i], v[i + 1]);
      mx += fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      nm2.x -= mx * 1e-9f;
    }
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float2 x = make_float2(v[i], v[i + 1]);
      if (MODE & 1) {
        if (MODE & 32) { x.x = fmaf(x.x, sc2.x, nm2.x); x.y = fmaf(x.y, sc2.y, nm2.y); }
        else x = __ffma2_rn(x, sc2, nm2);
      } else { x.x += nm2.x; }
      float2 ab = (MODE & 64) ? x : make_float2(ex2(x.x), ex2(x.y));
      if (MODE & 2) {
        if (MODE & 32) { rs2.x += ab.x; rs2.y += ab.y; }
        else rs2 = __fadd2_rn(rs2, ab);
      }
      unsigned pk;
      if (MODE & 4) { __nv_bfloat162 p2 = __floats2bfloat162_rn(ab.x, ab.y); pk = *reinterpret_cast<unsigned*>(&p2); }
      else if (MODE & 8) pk = __byte_perm(__float_as_uint(ab.x), __float_as_uint(ab.y), 0x7632);
      else pk = __float_as_uint(ab.x) ^ __float_as_uint(ab.y);
      asm volatile("" ::"r"(pk));
    }
    nm2.x += 1e-7f; nm2.y += 1e-7f;                                  // loop-carried: nothing can be hoisted
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = rs2.x + rs2.y + mx;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads = 512) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -0.001f * (i % 977); h[1] = 0.18f; h[2] = -0.5f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(in, out, cyc, 10);
  k<MODE><<<148, threads>>>(in, out, cyc, iters);
  long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  long long mxc = 0; for (int i = 0; i < 148; ++i) mxc = hc[i] > mxc ? hc[i] : mxc;
  printf("%-70s %6.2f elements / clk / SM   (%.1f clk per 128x128 block pair-of-tiles equivalent = 32768 elements)\n", name,
         (double)threads * 64 * iters / (double)mxc, (double)mxc / iters * 512.0 / threads);
}
int main() {
  run<0>("MUFU.EX2 only");
  run<1>("FFMA2 + MUFU");
  run<1 | 2>("FFMA2 + MUFU + FADD2");
  run<1 | 2 | 4>("FFMA2 + MUFU + FADD2 + F2FP           (the kernel's mix)");
  run<1 | 2 | 8>("FFMA2 + MUFU + FADD2 + PRMT");
  run<1 | 2 | 4 | 32>("FFMA + MUFU + FADD + F2FP (unpacked)");
  run<1 | 2 | 4 | 16>("FMNMX3 pass + FFMA2 + MUFU + FADD2 + F2FP");
  run<1 | 2 | 4 | 64>("FFMA2 + FADD2 + F2FP, no MUFU");
  run<4>("MUFU + F2FP");
  run<2>("MUFU + FADD2");
  run<1 | 2 | 4>("the kernel's mix, 8 warps per SM (2 per scheduler)", 256);
  run<1 | 2 | 4>("the kernel's mix, 4 warps per SM (1 per scheduler)", 128);
  run<1 | 2 | 4 | 16>("max pass + the kernel's mix, 8 warps per SM", 256);
  run<1 | 2 | 4 | 16>("max pass + the kernel's mix, 4 warps per SM", 128);
  return 0;
}
