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
    Pack4<OutT>::store(out + (size_t)pix * 64 + cg * 16 + q * 4, o[0], o[1], o[2], o[3]);
  }
}

// Split-plane variant, register-blocked: thread = (4 consecutive pixels, 8 channels).  The 72 weights of the 8 channels are
// fetched once (18 LDS.128 from a [tap][channel] table) and reused for the 4 pixels, so the kernel is bound by its 16-byte
// stores (8 lanes = the 128 contiguous bytes of one pixel per plane) instead of by shared-memory weight reads.
__global__ void __launch_bounds__(256) sp_conv1a_split_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ wgt,
                                                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int fmt) {
  __shared__ __align__(16) float w_s[9 * 64];
  __shared__ __align__(16) float b_s[64];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) w_s[(i % 9) * 64 + i / 9] = wgt[i];      // [ch][tap] -> [tap][ch]
  for (int i = threadIdx.x; i < 64; i += blockDim.x) b_s[i] = bias[i];
  __syncthreads();
  const int wq = (W + 3) >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(gid & 7);
  const long long grp = gid >> 3;
  if (grp >= (long long)H * wq) return;
  const int y = (int)(grp / wq), x0 = (int)(grp - (long long)y * wq) * 4;
  float v[3][6];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 6; ++dx) {
      const int yy = y + dy - 1, xx = x0 + dx - 1;
      v[dy][dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (size_t)yy * W + xx) : 0.f;
    }
  // packed f32x2 accumulators (two channels per FFMA2: same rounding as two FFMAs, half the issue slots)
  float2 acc[4][4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(b_s + cg * 8), b1 = *reinterpret_cast<const float4*>(b_s + cg * 8 + 4);
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      acc[px][0] = make_float2(b0.x, b0.y); acc[px][1] = make_float2(b0.z, b0.w);
      acc[px][2] = make_float2(b1.x, b1.y); acc[px][3] = make_float2(b1.z, b1.w);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = *reinterpret_cast<const float4*>(w_s + t * 64 + cg * 8), w1 = *reinterpret_cast<const float4*>(w_s + t * 64 + cg * 8 + 4);
    const float2 w[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const float a = v[t / 3][px + t % 3];
      const float2 a2 = make_float2(a, a);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[px][c] = __ffma2_rn(a2, w[c], acc[px][c]);
    }
  }
  const size_t plane = (size_t)H * W * 64;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    if (x0 + px >= W) break;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const float a = fmaxf(acc[px][c >> 1].x, 0.f), b = fmaxf(acc[px][c >> 1].y, 0.f);
      if (fmt) {                                                     // bf16 (hi, lo)
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        const float2 hf = __bfloat1622float2(h);
        const float2 rem = __fadd2_rn(make_float2(a, b), make_float2(-hf.x, -hf.y));
        const __nv_bfloat162 l = __floats2bfloat162_rn(rem.x, rem.y);
        hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&l);
      } else {                                                       // IEEE half (hi, lo); values beyond the half range saturate
        const float as = fminf(a, 65504.f), bs = fminf(b, 65504.f);
        const __half2 h = __floats2half2_rn(as, bs);
        const float2 hf = __half22float2(h);
        const float2 rem = __fadd2_rn(make_float2(as, bs), make_float2(-hf.x, -hf.y));
        const __half2 l = __floats2half2_rn(rem.x, rem.y);
        hi[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        lo[c >> 1] = *reinterpret_cast<const uint32_t*>(&l);
      }
    }
    __nv_bfloat16* o = out + ((size_t)y * W + x0 + px) * 64 + cg * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
  I4D_CHECK_ARG(out_dtype >= 0 && out_dtype <= 4, "out_dtype: 0 = f32, 1 = f16, 2 = bf16, 3 = split bf16 planes, 4 = split f16 planes");
  const long long threads = (long long)H * W * 4;
  const int grid = i4d_cdiv(threads, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == 0) sp_conv1a_kernel<float><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<float*>(out_nhwc));
  else if (out_dtype == 1) sp_conv1a_kernel<__half><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__half*>(out_nhwc));
  else if (out_dtype == 2) sp_conv1a_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__nv_bfloat16*>(out_nhwc));
  else {
    const long long t8 = (long long)H * ((W + 3) >> 2) * 8;
    sp_conv1a_split_kernel<<<i4d_cdiv(t8, 256), 256, 0, st>>>(image, H, W, weight, bias, reinterpret_cast<__nv_bfloat16*>(out_nhwc), out_dtype == 3 ? 1 : 0);
  }
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
