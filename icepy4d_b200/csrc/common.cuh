// Shared helpers for the icepy4d_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define I4D_OK 0
#define I4D_ERR_INVALID (-1)
#define I4D_ERR_CUDA (-2)
#define I4D_ERR_WORKSPACE (-3)
#define I4D_ERR_UNSUPPORTED (-4)
#define I4D_ERR_OVERFLOW (-5)

void i4d_set_error(const char* fmt, ...);

#define I4D_CHECK_ARG(cond, msg)                                        \
  do {                                                                  \
    if (!(cond)) {                                                      \
      i4d_set_error("%s: invalid argument: %s", __func__, msg);         \
      return I4D_ERR_INVALID;                                           \
    }                                                                   \
  } while (0)

#define I4D_CUDA_LAUNCH_CHECK()                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      i4d_set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e__));    \
      return I4D_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define I4D_CUDA_CALL(x)                                                                 \
  do {                                                                                   \
    cudaError_t e__ = (x);                                                               \
    if (e__ != cudaSuccess) {                                                            \
      i4d_set_error("%s: %s failed: %s", __func__, #x, cudaGetErrorString(e__));         \
      return I4D_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// cudaFuncSetAttribute is per device: a process that drives several GPUs must set it on each one (the benign race between two
// host threads of one process only repeats an idempotent call).  Returns true the first time it is called on the current device
// for the given call-site flag array.
static inline bool i4d_first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

static inline int i4d_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// streaming 128-bit load that does not pollute L1 (HBM-bound single-use data)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

int i4d_num_sms();
