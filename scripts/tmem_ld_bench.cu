// Micro-benchmark: tcgen05.ld throughput (TMEM -> registers) per SM.  Build + run under gpurun:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem scripts/tmem_ld_bench.cu && /tmp/tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NWARPS, int X>
__global__ void __launch_bounds__(NWARPS * 32) k(long long* out, int iters) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32];
    if (X == 32) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(t + (uint32_t)((it & 3) * 32)) : "memory");
    } else {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t + (uint32_t)((it & 7) * 16)) : "memory");
      for (int i = 16; i < 32; ++i) r[i] = 0;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < X; ++i) acc ^= r[i];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256) : "memory");
}
template <int NWARPS, int X> void run(const char* name, int ctas_per_sm) {
  long long* out; cudaMalloc(&out, 2048 * sizeof(long long));
  const int iters = 4096;
  k<NWARPS, X><<<148 * ctas_per_sm, NWARPS * 32>>>(out, 16);
  k<NWARPS, X><<<148 * ctas_per_sm, NWARPS * 32>>>(out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double bytes = (double)NWARPS * 32 * X * 4 * iters * ctas_per_sm;
  printf("%-36s ctas/SM=%d: %lld clk -> %.1f B/clk/SM (%s)\n", name, ctas_per_sm, h[0], bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  run<4, 32>("4 warps, 32x32b.x32 + wait each", 1);
  run<4, 32>("4 warps, 32x32b.x32 + wait each", 2);
  run<8, 32>("8 warps, 32x32b.x32 + wait each", 1);
  run<8, 32>("8 warps, 32x32b.x32 + wait each", 2);
  run<8, 16>("8 warps, 32x32b.x16 + wait each", 2);
  return 0;
}
