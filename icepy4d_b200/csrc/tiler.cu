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
