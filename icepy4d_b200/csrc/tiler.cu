// Tiler front end on sm_100a: patch extraction fused with grey conversion and u8 -> f32 /255 conversion, so a tile
// goes from the resident u8 image straight to the network input tensor in one HBM pass (reads th*tw*C bytes,
// writes th*tw*4 bytes).
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d/matching):
//   tiling.py:123-135 extract_patch;  matchers.py:911-917 cv2.cvtColor(RGB2GRAY) + :263-274 `image / 255.` (f64 divide,
//   cast to f32);  matchers.py:1212-1220 + thirdparty/LightGlue/lightglue/utils.py:35-36 (per-channel /255 then
//   0.299/0.587/0.114 weighted sum in f32, summed r, g, b in that order).
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

__device__ __forceinline__ float u8_to_unit(unsigned int v) { return (float)((double)v / 255.0); }

__global__ void __launch_bounds__(256) tile_gray_kernel(const unsigned char* __restrict__ img, int W, int C, int x0, int y0,
                                                        int tw, int th, int mode, float* __restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= tw || y >= th) return;
  const unsigned char* p = img + ((size_t)(y0 + y) * W + (x0 + x)) * C;
  float v;
  if (C == 1) {
    v = u8_to_unit(p[0]);
  } else if (mode == 0) {
    // OpenCV 8-bit RGB2GRAY: (R*9798 + G*19235 + B*3735 + 2^14) >> 15
    unsigned int g = (p[0] * 9798u + p[1] * 19235u + p[2] * 3735u + 16384u) >> 15;
    v = u8_to_unit(g);
  } else {
    float r = __fmul_rn(u8_to_unit(p[0]), 0.299f), g = __fmul_rn(u8_to_unit(p[1]), 0.587f), b = __fmul_rn(u8_to_unit(p[2]), 0.114f);
    v = __fadd_rn(__fadd_rn(r, g), b);
  }
  out[(size_t)y * tw + x] = v;
}

extern "C" __attribute__((visibility("default"))) int i4d_tile_to_gray_f32(const unsigned char* image, int H, int W, int C,
                                                                          int x0, int y0, int tw, int th, int mode,
                                                                          float* out, void* stream) {
  I4D_CHECK_ARG(image && out, "null pointer");
  I4D_CHECK_ARG(C == 1 || C == 3, "image must have 1 or 3 channels");
  I4D_CHECK_ARG(x0 >= 0 && y0 >= 0 && tw > 0 && th > 0 && x0 + tw <= W && y0 + th <= H, "tile outside the image");
  I4D_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (cv2 RGB2GRAY) or 1 (float weights)");
  dim3 grid(i4d_cdiv(tw, 256), th);
  tile_gray_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, W, C, x0, y0, tw, th, mode, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Gaussian pyramid on u8 images, bit-exact with OpenCV's 8-bit cv2.pyrDown / cv2.pyrUp (used by the reference for
// Quality.{LOW,MEDIUM,HIGHEST} and for the PRESELECTION low-resolution pass: matching/matchers.py:583-610, 516-531).
//   pyrDown: 5x5 kernel [1 4 6 4 1] (x) [1 4 6 4 1], integer, (sum + 128) >> 8, BORDER_REFLECT_101, out = ((W+1)/2, (H+1)/2)
//   pyrUp  : even samples p[i-1] + 6 p[i] + p[i+1], odd samples 4 p[i] + 4 p[i+1] per axis, (sum + 32) >> 6;
//            reflect-101 at the low border, replicate at the high border (OpenCV's rule), out = (2W, 2H)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int refl101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : i;
}

__global__ void __launch_bounds__(256) pyr_down_kernel(const unsigned char* __restrict__ in, int H, int W, int C,
                                                       unsigned char* __restrict__ out, int Ho, int Wo) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)Ho * Wo * C) return;
  const int c = (int)(gid % C);
  const long long p = gid / C;
  const int xo = (int)(p % Wo), yo = (int)(p / Wo);
  const int wgt[5] = {1, 4, 6, 4, 1};
  int acc = 0;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    const int yy = refl101(2 * yo + dy - 2, H);
    int row = 0;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) row += wgt[dx] * (int)__ldg(in + ((size_t)yy * W + refl101(2 * xo + dx - 2, W)) * C + c);
    acc += wgt[dy] * row;
  }
  out[gid] = (unsigned char)((acc + 128) >> 8);
}

__global__ void __launch_bounds__(256) pyr_up_kernel(const unsigned char* __restrict__ in, int H, int W, int C,
                                                     unsigned char* __restrict__ out) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)Ho * Wo * C) return;
  const int c = (int)(gid % C);
  const long long p = gid / C;
  const int xo = (int)(p % Wo), yo = (int)(p / Wo);
  const int i = yo >> 1, j = xo >> 1;
  const int il = i > 0 ? i - 1 : (H > 1 ? 1 : 0), ir = i + 1 < H ? i + 1 : i;      // reflect-101 low, replicate high
  const int jl = j > 0 ? j - 1 : (W > 1 ? 1 : 0), jr = j + 1 < W ? j + 1 : j;
  auto px = [&](int y, int x) { return (int)__ldg(in + ((size_t)y * W + x) * C + c); };
  auto hrow = [&](int y) { return (xo & 1) ? 4 * px(y, j) + 4 * px(y, jr) : px(y, jl) + 6 * px(y, j) + px(y, jr); };
  const int acc = (yo & 1) ? 4 * hrow(i) + 4 * hrow(ir) : hrow(il) + 6 * hrow(i) + hrow(ir);
  out[gid] = (unsigned char)((acc + 32) >> 6);
}

extern "C" __attribute__((visibility("default"))) int i4d_pyr_down_u8(const unsigned char* in, int H, int W, int C,
                                                                     unsigned char* out, void* stream) {
  I4D_CHECK_ARG(in && out && H > 0 && W > 0 && (C == 1 || C == 3), "bad arguments");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long n = (long long)Ho * Wo * C;
  pyr_down_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, H, W, C, out, Ho, Wo);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_pyr_up_u8(const unsigned char* in, int H, int W, int C,
                                                                   unsigned char* out, void* stream) {
  I4D_CHECK_ARG(in && out && H > 0 && W > 0 && (C == 1 || C == 3), "bad arguments");
  const long long n = (long long)4 * H * W * C;
  pyr_up_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, H, W, C, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- PRESELECTION decision (matchers.py:495-499, 541-556) ---------------------------------------------------------------
// counts[t0 * T1 + t1] = #{ matches i : valid[i], scale * kp0[i] strictly inside lims0[t0] and scale * kp1[i] strictly inside
// lims1[t1] }.  One thread per match builds the two membership bit sets (<= 64 tiles per image) and adds 1 to every pair in
// their product (integer atomics: order-independent, bit-reproducible).
__global__ void __launch_bounds__(256) tile_pair_count_kernel(const float* __restrict__ kp0, const float* __restrict__ kp1,
                                                              const unsigned char* __restrict__ valid, int n, float scale,
                                                              const float* __restrict__ lims0, int T0,
                                                              const float* __restrict__ lims1, int T1, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (valid && !valid[i])) return;
  const float x0 = kp0[2 * i] * scale, y0 = kp0[2 * i + 1] * scale, x1 = kp1[2 * i] * scale, y1 = kp1[2 * i + 1] * scale;
  unsigned long long m1 = 0ull;
  for (int t = 0; t < T1; ++t) {
    const float* r = lims1 + 4 * t;
    if (x1 > r[0] && y1 > r[1] && x1 < r[2] && y1 < r[3]) m1 |= 1ull << t;
  }
  if (!m1) return;
  for (int t = 0; t < T0; ++t) {
    const float* r = lims0 + 4 * t;
    if (!(x0 > r[0] && y0 > r[1] && x0 < r[2] && y0 < r[3])) continue;
    for (unsigned long long b = m1; b; b &= b - 1) atomicAdd(counts + t * T1 + (__ffsll((long long)b) - 1), 1);
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_tile_pair_counts(const float* kp0, const float* kp1,
                                                                          const unsigned char* valid, int n, float scale,
                                                                          const float* lims0, int T0, const float* lims1, int T1,
                                                                          int* counts, void* stream) {
  I4D_CHECK_ARG(counts && lims0 && lims1 && (n == 0 || (kp0 && kp1)), "null pointer");
  I4D_CHECK_ARG(T0 >= 1 && T0 <= 64 && T1 >= 1 && T1 <= 64, "1..64 tiles per image");
  I4D_CUDA_CALL(cudaMemsetAsync(counts, 0, (size_t)T0 * T1 * sizeof(int), (cudaStream_t)stream));
  if (n > 0) {
    tile_pair_count_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(kp0, kp1, valid, n, scale, lims0, T0, lims1, T1, counts);
    I4D_CUDA_LAUNCH_CHECK();
  }
  return I4D_OK;
}
