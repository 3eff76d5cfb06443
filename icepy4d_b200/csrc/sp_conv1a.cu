// SuperPoint conv1a (1 -> 64 channels, 3x3, pad 1) + bias + ReLU, hand-written: the layer is purely bandwidth bound
// (9 MACs per output, 1 GB of f32 output per 1999x1999 tile) and the library path costs a 1 GB NCHW->NHWC transpose on
// top of a CUDA-core convolution.  This kernel reads the grey tile once (through L1) and writes the activation directly
// in channels-last layout, f32 or f16/bf16, with 16-byte stores.
//
// Reference behaviour replaced: `self.relu(self.conv1a(data["image"]))`, thirdparty/SuperGlue/models/superpoint.py:154
// (LightGlue copy: thirdparty/LightGlue/lightglue/superpoint.py:155).
#include "common.cuh"
#include <cuda_fp16.h>
#include <type_traits>
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

// split bf16 planes (hi at p, lo at p + plane): the operand format of the tcgen05 convolutions in conv_tc.cu
struct SplitBf16 { __nv_bfloat16 v; };
template <> struct Pack4<SplitBf16> {
  static __device__ __forceinline__ void store(SplitBf16* p, size_t plane, float a, float b, float c, float d) {
    __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
    const float2 xf = __bfloat1622float2(x), yf = __bfloat1622float2(y);
    __nv_bfloat162 xl = __floats2bfloat162_rn(a - xf.x, b - xf.y), yl = __floats2bfloat162_rn(c - yf.x, d - yf.y);
    uint2 v; v.x = *reinterpret_cast<uint32_t*>(&x); v.y = *reinterpret_cast<uint32_t*>(&y);
    uint2 l; l.x = *reinterpret_cast<uint32_t*>(&xl); l.y = *reinterpret_cast<uint32_t*>(&yl);
    *reinterpret_cast<uint2*>(p) = v;
    *reinterpret_cast<uint2*>(p + plane) = l;
  }
};

// thread = (pixel, 16-channel group): 4 consecutive threads write the 64 channels of one pixel; the 3x3 neighbourhood is
// loaded once per thread for 16 output channels.
template <typename OutT>
__global__ void __launch_bounds__(256) sp_conv1a_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, OutT* __restrict__ out) {
  __shared__ float w_s[64 * 9], b_s[64];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) w_s[i] = wgt[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) b_s[i] = bias[i];
  __syncthreads();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gid >> 2;
  const int cg = (int)(gid & 3);
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
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = cg * 16 + q * 4 + c;
      const float* w = w_s + ch * 9;
      float acc = b_s[ch];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(v[t], w[t], acc);
      o[c] = fmaxf(acc, 0.f);
    }
    if constexpr (std::is_same<OutT, SplitBf16>::value)
      Pack4<OutT>::store(out + (size_t)pix * 64 + cg * 16 + q * 4, (size_t)H * W * 64, o[0], o[1], o[2], o[3]);
    else
      Pack4<OutT>::store(out + (size_t)pix * 64 + cg * 16 + q * 4, o[0], o[1], o[2], o[3]);
  }
}

// 2x2 / stride-2 max pooling on a channels-last tensor [H,W,C] -> [H/2,W/2,C] (floor), 16 bytes per thread.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) maxpool2_nhwc_kernel(const T* __restrict__ in, int H, int W, int C, T* __restrict__ out) {
  const int Ho = H >> 1, Wo = W >> 1, cv = C / VEC;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)Ho * Wo * cv) return;
  const int c = (int)(gid % cv);
  const long long p = gid / cv;
  const int xo = (int)(p % Wo), yo = (int)(p / Wo);
  const T* base = in + ((size_t)(2 * yo) * W + 2 * xo) * C + c * VEC;
  uint4 a = *reinterpret_cast<const uint4*>(base), b = *reinterpret_cast<const uint4*>(base + C);
  uint4 d = *reinterpret_cast<const uint4*>(base + (size_t)W * C), e = *reinterpret_cast<const uint4*>(base + (size_t)W * C + C);
  uint4 r;
  if (sizeof(T) == 4) {
    const float* fa = reinterpret_cast<const float*>(&a); const float* fb = reinterpret_cast<const float*>(&b);
    const float* fd = reinterpret_cast<const float*>(&d); const float* fe = reinterpret_cast<const float*>(&e);
    float* fr = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) fr[i] = fmaxf(fmaxf(fa[i], fb[i]), fmaxf(fd[i], fe[i]));
  } else if (sizeof(T) == 2 && VEC == 8) {
    // post-ReLU activations are >= 0: for non-negative IEEE half / bfloat16 values the bit patterns order like the values
    const unsigned short* ha = reinterpret_cast<const unsigned short*>(&a); const unsigned short* hb = reinterpret_cast<const unsigned short*>(&b);
    const unsigned short* hd = reinterpret_cast<const unsigned short*>(&d); const unsigned short* he = reinterpret_cast<const unsigned short*>(&e);
    unsigned short* hr = reinterpret_cast<unsigned short*>(&r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned short m1 = ha[i] > hb[i] ? ha[i] : hb[i], m2 = hd[i] > he[i] ? hd[i] : he[i];
      hr[i] = m1 > m2 ? m1 : m2;
    }
  }
  *reinterpret_cast<uint4*>(out + ((size_t)yo * Wo + xo) * C + c * VEC) = r;
}

extern "C" __attribute__((visibility("default"))) int i4d_maxpool2x2_nhwc(const void* in, int H, int W, int C, void* out,
                                                                         int elem_bytes, int nonneg, void* stream) {
  I4D_CHECK_ARG(in && out && H >= 2 && W >= 2 && C > 0, "bad arguments");
  I4D_CHECK_ARG(elem_bytes == 4 || elem_bytes == 2, "element size must be 4 (f32) or 2 (f16/bf16)");
  I4D_CHECK_ARG(elem_bytes == 4 || nonneg, "16-bit pooling compares bit patterns: inputs must be non-negative (post-ReLU)");
  const int vec = 16 / elem_bytes;
  I4D_CHECK_ARG(C % vec == 0, "C must be a multiple of 16 bytes worth of elements");
  const long long n = (long long)(H >> 1) * (W >> 1) * (C / vec);
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 4) maxpool2_nhwc_kernel<float, 4><<<i4d_cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<const float*>(in), H, W, C, reinterpret_cast<float*>(out));
  else maxpool2_nhwc_kernel<unsigned short, 8><<<i4d_cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<const unsigned short*>(in), H, W, C, reinterpret_cast<unsigned short*>(out));
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_conv1a_relu(const float* image, int H, int W, const float* weight,
                                                                        const float* bias, void* out_nhwc, int out_dtype,
                                                                        void* stream) {
  I4D_CHECK_ARG(image && weight && bias && out_nhwc && H > 0 && W > 0, "null pointer or empty image");
  I4D_CHECK_ARG(out_dtype >= 0 && out_dtype <= 3, "out_dtype: 0 = f32, 1 = f16, 2 = bf16, 3 = split bf16 planes");
  const long long threads = (long long)H * W * 4;
  const int grid = i4d_cdiv(threads, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == 0) sp_conv1a_kernel<float><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<float*>(out_nhwc));
  else if (out_dtype == 1) sp_conv1a_kernel<__half><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__half*>(out_nhwc));
  else if (out_dtype == 2) sp_conv1a_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__nv_bfloat16*>(out_nhwc));
  else sp_conv1a_kernel<SplitBf16><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<SplitBf16*>(out_nhwc));
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
