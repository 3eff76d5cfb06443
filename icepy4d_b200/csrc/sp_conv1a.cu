// SuperPoint conv1a (1 -> 64 channels, 3x3, pad 1) + bias + ReLU, hand-written: the layer is purely bandwidth bound
// (9 MACs per output, 1 GB of f32 output per 1999x1999 tile) and the library path costs a 1 GB NCHW->NHWC transpose on
// top of a CUDA-core convolution.  This kernel reads the grey tile once (through L1) and writes the activation directly
// in channels-last layout, f32 or f16/bf16, with 16-byte stores.
//
// Reference behaviour replaced: `self.relu(self.conv1a(data["image"]))`, thirdparty/SuperGlue/models/superpoint.py:154
// (LightGlue copy: thirdparty/LightGlue/lightglue/superpoint.py:155).
#include "common.cuh"
#include <cuda_fp16.h>
#include "../../include/icepy4d_b200.h"

template <typename OutT> struct Pack4;
template <> struct Pack4<float> {
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
};
template <> struct Pack4<__half> {
  static __device__ __forceinline__ void store(__half* p, float a, float b, float c, float d) {
    __half2 x = __floats2half2_rn(a, b), y = __floats2half2_rn(c, d);
    uint2 v; v.x = *reinterpret_cast<uint32_t*>(&x); v.y = *reinterpret_cast<uint32_t*>(&y);
    *reinterpret_cast<uint2*>(p) = v;
  }
};
template <> struct Pack4<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
    uint2 v; v.x = *reinterpret_cast<uint32_t*>(&x); v.y = *reinterpret_cast<uint32_t*>(&y);
    *reinterpret_cast<uint2*>(p) = v;
  }
};

// thread = (pixel, channel quad): 16 consecutive threads write the 64 channels of one pixel (256 B f32 / 128 B half)
template <typename OutT>
__global__ void __launch_bounds__(256) sp_conv1a_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, OutT* __restrict__ out) {
  __shared__ float w_s[64 * 9], b_s[64];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) w_s[i] = wgt[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) b_s[i] = bias[i];
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gid >> 4;
  const int cq = (int)(gid & 15);
  if (pix >= (long long)H * W) return;
  const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
  float v[9];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      v[(dy + 1) * 3 + dx + 1] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (size_t)yy * W + xx) : 0.f;
    }
  float o[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float* w = w_s + (cq * 4 + c) * 9;
    float acc = b_s[cq * 4 + c];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc = fmaf(v[t], w[t], acc);
    o[c] = fmaxf(acc, 0.f);
  }
  Pack4<OutT>::store(out + (size_t)pix * 64 + cq * 4, o[0], o[1], o[2], o[3]);
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_conv1a_relu(const float* image, int H, int W, const float* weight,
                                                                        const float* bias, void* out_nhwc, int out_dtype,
                                                                        void* stream) {
  I4D_CHECK_ARG(image && weight && bias && out_nhwc && H > 0 && W > 0, "null pointer or empty image");
  I4D_CHECK_ARG(out_dtype >= 0 && out_dtype <= 2, "out_dtype: 0 = f32, 1 = f16, 2 = bf16");
  const long long threads = (long long)H * W * 16;
  const int grid = i4d_cdiv(threads, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == 0) sp_conv1a_kernel<float><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<float*>(out_nhwc));
  else if (out_dtype == 1) sp_conv1a_kernel<__half><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__half*>(out_nhwc));
  else sp_conv1a_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__nv_bfloat16*>(out_nhwc));
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
