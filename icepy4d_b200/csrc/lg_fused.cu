// LightGlue glue kernels of the tensor-core path (sm_100a): the element-wise work between the tcgen05 GEMMs / attention, fused so
// that every activation makes ONE trip through HBM between two tensor-core kernels.
//
//   i4d_lg_rotary_cast_bf16     f32 [n, 768] fused QKV projection -> rotary embedding on q and k (lightglue.py:49-57, 155-159:
//                               pairs (x[2p], x[2p+1]) of every head rotated by the keypoint's angle p) -> bf16 [n, 768] in the
//                               layout attn_tc reads.  Replaces two rotary passes + one cast pass.
//   i4d_layernorm_gelu_bf16     f32 [n, C] -> LayerNorm(C) -> GELU(erf) (lightglue.py:144-149, ffn[1], ffn[2]) -> bf16 [n, C], the A
//                               operand of the second FFN GEMM.  Replaces a LayerNorm+GELU pass + one cast pass.
// Both are HBM-bound streaming kernels: 16-byte loads, 8- / 16-byte stores, one warp per row.
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// one warp per row: 768 columns = 6 float4 per lane.  Column c of q / k (c < 512) belongs to pair p = (c % 64) / 2.
__global__ void __launch_bounds__(256) lg_rotary_cast_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ cs, int n,
                                                             __nv_bfloat16* __restrict__ Y, int ldy) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = X + (size_t)row * ldx;
  const float* c = cs + (size_t)row * 64;
  __nv_bfloat16* y = Y + (size_t)row * ldy;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int col = (k * 32 + lane) * 4;                       // 4 consecutive columns = 2 rotary pairs
    float4 v = ldg_stream(reinterpret_cast<const float4*>(x + col));
    if (col < 512) {
      const int p = (col & 63) >> 1;
      const float2 cc = __ldg(reinterpret_cast<const float2*>(c + p)), ss = __ldg(reinterpret_cast<const float2*>(c + 32 + p));
      v = make_float4(v.x * cc.x - v.y * ss.x, v.y * cc.x + v.x * ss.x, v.z * cc.y - v.w * ss.y, v.w * cc.y + v.z * ss.y);
    }
    *reinterpret_cast<uint2*>(y + col) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_lg_rotary_cast_bf16(const float* qkv, int ldx, const float* cs, int n, void* out,
                                                                             int ldy, void* stream) {
  I4D_CHECK_ARG(qkv && cs && out && n >= 0, "null pointer");
  I4D_CHECK_ARG((ldx & 3) == 0 && (ldy & 3) == 0 && ldx >= 768 && ldy >= 768, "pitches must be multiples of 4 and >= 768");
  if (n == 0) return I4D_OK;
  lg_rotary_cast_kernel<<<i4d_cdiv((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(qkv, ldx, cs, n, reinterpret_cast<__nv_bfloat16*>(out), ldy);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// one warp per row, C = 512: the row stays in registers (4 float4 per lane) between the statistics and the output pass
__global__ void __launch_bounds__(256) layernorm_gelu_bf16_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int n, float eps,
                                                                  __nv_bfloat16* __restrict__ Y, int ldy) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* x = X + (size_t)row * ldx;
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = ldg_stream(reinterpret_cast<const float4*>(x + (k * 32 + lane) * 4));
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = warp_sum(s) * (1.f / 512.f);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 512.f) + eps);
  __nv_bfloat16* y = Y + (size_t)row * ldy;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int col = (k * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col)), b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float t[4] = {(v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y, (v[k].z - mean) * rstd * g.z + b.z,
                  (v[k].w - mean) * rstd * g.w + b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = 0.5f * t[e] * (1.f + erff(t[e] * 0.70710678118654752440f));
    *reinterpret_cast<uint2*>(y + col) = make_uint2(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]));
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_layernorm_gelu_bf16(const float* X, int ldx, const float* gamma, const float* beta,
                                                                             int n, int C, float eps, void* out, int ldy, void* stream) {
  I4D_CHECK_ARG(X && gamma && beta && out && n >= 0, "null pointer");
  I4D_CHECK_ARG(C == 512 && (ldx & 3) == 0 && (ldy & 3) == 0, "C must be 512 (LightGlue FFN width), pitches multiples of 4");
  if (n == 0) return I4D_OK;
  layernorm_gelu_bf16_kernel<<<i4d_cdiv((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, gamma, beta, n, eps,
                                                                                                 reinterpret_cast<__nv_bfloat16*>(out), ldy);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
