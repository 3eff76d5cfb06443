// SuperPoint post-processing on sm_100a: softmax(65)+pixel-shuffle, exact 3-pass NMS with smem halo tiles,
// threshold + border + block-aggregated compaction, radix-select top-k + bitonic sort, and the bilinear
// descriptor gather with both L2 normalisations folded in.
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d/thirdparty):
//   SuperGlue/models/superpoint.py:169-172 (softmax, shuffle)  :48-64 (simple_nms)  :176-203 (extract, borders, top-k)
//   SuperGlue/models/superpoint.py:82-97,208 (sample_descriptors)   LightGlue/lightglue/superpoint.py:176-200 (LG variant)
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

// ---------------------------------------------------------------------------------------------------
// 1. softmax over 65 channels, drop dustbin, 8x8 pixel shuffle.  One thread per coarse cell.
//    HBM-bound: reads 65*h*w*4 B (coalesced per channel plane), writes 64*h*w*4 B as float4 pairs.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sp_score_map_kernel(const float* __restrict__ logits, int h, int w,
                                                           float* __restrict__ scores) {
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= h * w) return;
  int cy = cell / w, cx = cell - cy * w;
  const size_t plane = (size_t)h * w;
  float v[65];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 65; ++c) {
    v[c] = __ldg(logits + c * plane + cell);
    mx = fmaxf(mx, v[c]);
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < 65; ++c) {
    v[c] = expf(v[c] - mx);
    sum += v[c];
  }
  const int W8 = w * 8;
#pragma unroll
  for (int dy = 0; dy < 8; ++dy) {
    float4 a, b;
    a.x = v[dy * 8 + 0] / sum; a.y = v[dy * 8 + 1] / sum; a.z = v[dy * 8 + 2] / sum; a.w = v[dy * 8 + 3] / sum;
    b.x = v[dy * 8 + 4] / sum; b.y = v[dy * 8 + 5] / sum; b.z = v[dy * 8 + 6] / sum; b.w = v[dy * 8 + 7] / sum;
    float4* dst = reinterpret_cast<float4*>(scores + (size_t)(cy * 8 + dy) * W8 + cx * 8);
    dst[0] = a;
    dst[1] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// 2. NMS (exact simple_nms semantics) + threshold + border + compaction.
//    Tile of 64x64 outputs with a 5r halo staged in shared memory: the final mask at a pixel depends on
//    scores within 5r (5 chained (2r+1)^2 max-pools), so one HBM read of the tile(+halo) suffices.
// ---------------------------------------------------------------------------------------------------
#define NMS_T 64
#define NMS_THREADS 256

struct NmsSmem {
  float* S0; float* T1; float* X; unsigned char* M; unsigned char* P;
};

template <typename F>
__device__ __forceinline__ void nms_rowmax(const F& src, float* __restrict__ dst, int D, int r) {
  for (int i = threadIdx.x; i < D * D; i += NMS_THREADS) {
    int y = i / D, x = i - y * D;
    int x0 = max(x - r, 0), x1 = min(x + r, D - 1);
    float m = -INFINITY;
    for (int xx = x0; xx <= x1; ++xx) m = fmaxf(m, src(y * D + xx));
    dst[i] = m;
  }
}
__device__ __forceinline__ float nms_colmax(const float* __restrict__ t, int D, int r, int y, int x) {
  int y0 = max(y - r, 0), y1 = min(y + r, D - 1);
  float m = -INFINITY;
  for (int yy = y0; yy <= y1; ++yy) m = fmaxf(m, t[yy * D + x]);
  return m;
}

__global__ void __launch_bounds__(NMS_THREADS) sp_nms_kernel(const float* __restrict__ scores, int H, int W, int r,
                                                             float thr, int border,
                                                             unsigned long long* __restrict__ cand, int cand_cap,
                                                             int* __restrict__ cand_count, float* __restrict__ nms_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int halo = 5 * r;
  const int D = NMS_T + 2 * halo;
  float* S0 = reinterpret_cast<float*>(smem_raw);
  float* T1 = S0 + D * D;
  float* X = T1 + D * D;
  unsigned char* M = reinterpret_cast<unsigned char*>(X + D * D);
  unsigned char* P = M + D * D;
  __shared__ int s_n, s_base;

  const int ty0 = blockIdx.y * NMS_T - halo, tx0 = blockIdx.x * NMS_T - halo;
  auto inimg = [&](int i) {
    int y = i / D, x = i - y * D;
    int gy = ty0 + y, gx = tx0 + x;
    return gy >= 0 && gy < H && gx >= 0 && gx < W;
  };
  for (int i = threadIdx.x; i < D * D; i += NMS_THREADS) {
    int y = i / D, x = i - y * D;
    int gy = ty0 + y, gx = tx0 + x;
    S0[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(scores + (size_t)gy * W + gx) : -INFINITY;
  }
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  nms_rowmax([&](int j) { return S0[j]; }, T1, D, r);
  __syncthreads();
  for (int i = threadIdx.x; i < D * D; i += NMS_THREADS) {
    int y = i / D, x = i - y * D;
    M[i] = (inimg(i) && S0[i] == nms_colmax(T1, D, r, y, x)) ? 1 : 0;
  }
  __syncthreads();
  for (int it = 0; it < 2; ++it) {
    nms_rowmax([&](int j) { return M[j] ? 1.f : 0.f; }, T1, D, r);
    __syncthreads();
    for (int i = threadIdx.x; i < D * D; i += NMS_THREADS) {
      int y = i / D, x = i - y * D;
      bool supp = nms_colmax(T1, D, r, y, x) > 0.f;
      P[i] = supp ? 1 : 0;
      X[i] = inimg(i) ? (supp ? 0.f : S0[i]) : -INFINITY;
    }
    __syncthreads();
    nms_rowmax([&](int j) { return X[j]; }, T1, D, r);
    __syncthreads();
    for (int i = threadIdx.x; i < D * D; i += NMS_THREADS) {
      int y = i / D, x = i - y * D;
      bool nm = inimg(i) && (X[i] == nms_colmax(T1, D, r, y, x));
      if (nm && !P[i]) M[i] = 1;
    }
    __syncthreads();
  }
  // threshold + border + block-aggregated compaction (X is free now: reuse it as the per-CTA key list)
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(X);
  for (int i = threadIdx.x; i < NMS_T * NMS_T; i += NMS_THREADS) {
    int y = i / NMS_T, x = i - y * NMS_T;
    int gy = blockIdx.y * NMS_T + y, gx = blockIdx.x * NMS_T + x;
    if (gy >= H || gx >= W) continue;
    int si = (y + halo) * D + (x + halo);
    float s = M[si] ? S0[si] : 0.f;
    if (nms_out) nms_out[(size_t)gy * W + gx] = s;
    if (s > thr && gy >= border && gy < H - border && gx >= border && gx < W - border) {
      int slot = atomicAdd(&s_n, 1);
      unsigned int idx = (unsigned int)(gy * W + gx);
      keys[slot] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s_base = s_n ? atomicAdd(cand_count, s_n) : 0;
  __syncthreads();
  for (int i = threadIdx.x; i < s_n; i += NMS_THREADS) {
    int g = s_base + i;
    if (g < cand_cap) cand[g] = keys[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// 3. top-k (radix select on 64-bit keys = score bits | inverted index => deterministic tie-break:
//    score descending, then linear index ascending) + bitonic sort + emit.
//    Output order mirrors the reference: score-descending when top-k applies (torch.topk sorted),
//    row-major (y, x) when all candidates are kept (torch.nonzero order).
// ---------------------------------------------------------------------------------------------------
#define SEL_THREADS 1024
#define SEL_MAX_SMEM_KEYS 16384

__device__ void bitonic_sort_desc_smem(unsigned long long* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = a[i], y = a[ixj];
          bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// mode 0: keys are (score|~idx), mode 1: keys are (~idx|score)
__device__ __forceinline__ void emit_keypoint(unsigned long long key, int mode, int W, float* kp, float* sc) {
  unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)(key & 0xFFFFFFFFull);
  unsigned int idx = 0xFFFFFFFFu - (mode ? hi : lo);
  float s = __uint_as_float(mode ? lo : hi);
  int y = idx / W, x = idx - y * W;
  kp[0] = (float)x;
  kp[1] = (float)y;
  *sc = s;
}

__global__ void __launch_bounds__(SEL_THREADS) sp_topk_kernel(const unsigned long long* __restrict__ cand,
                                                              const int* __restrict__ cand_count, int cand_cap, int k,
                                                              int W, float* __restrict__ kpts, float* __restrict__ sc,
                                                              int out_cap, int* __restrict__ n_out,
                                                              unsigned long long* __restrict__ spill) {
  extern __shared__ __align__(16) unsigned long long skeys[];  // SEL_MAX_SMEM_KEYS
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_cnt;
  const int n = min(*cand_count, cand_cap);
  const bool keep_all = (k < 0) || (n <= k);
  int m = keep_all ? n : k;
  if (m > out_cap) m = out_cap;  // caller sized the outputs; never write past them
  const int mode = keep_all ? 1 : 0;

  unsigned long long kth = 0ull;
  if (!keep_all) {
    if (threadIdx.x == 0) { s_prefix = 0ull; s_remaining = k; }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const int shift = pass * 8;
      const unsigned long long pmask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
      const unsigned long long prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned long long key = cand[i];
        if ((key & pmask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int rem = s_remaining;
        int b = 255;
        for (; b > 0; --b) {
          if ((int)hist[b] >= rem) break;
          rem -= (int)hist[b];
        }
        s_remaining = rem;
        s_prefix = prefix | ((unsigned long long)b << shift);
      }
      __syncthreads();
    }
    kth = s_prefix;
  }
  if (threadIdx.x == 0) { s_cnt = 0; *n_out = m; }
  __syncthreads();

  if (m <= SEL_MAX_SMEM_KEYS) {
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) skeys[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned long long key = cand[i];
      if (keep_all) {
        if (i < m) skeys[i] = (key << 32) | (key >> 32);
      } else if (key >= kth) {
        int slot = atomicAdd(&s_cnt, 1);
        if (slot < m) skeys[slot] = key;
      }
    }
    __syncthreads();
    bitonic_sort_desc_smem(skeys, p2);
    for (int i = threadIdx.x; i < m; i += blockDim.x) emit_keypoint(skeys[i], mode, W, kpts + 2 * i, sc + i);
  } else {
    // large result (only reachable with max_keypoints < 0 or > 16384): spill to global, sorted by the
    // follow-up global bitonic kernels (host side checks n_out and launches them).
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned long long key = cand[i];
      if (keep_all) {
        if (i < m) spill[i] = (key << 32) | (key >> 32);
      } else if (key >= kth) {
        int slot = atomicAdd(&s_cnt, 1);
        if (slot < m) spill[slot] = key;
      }
    }
  }
}

__global__ void bitonic_global_step(unsigned long long* a, int n_pow2, int k, int j) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pow2) return;
  int ixj = i ^ j;
  if (ixj > i) {
    unsigned long long x = a[i], y = a[ixj];
    bool desc = (i & k) == 0;
    if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
  }
}
__global__ void fill_u64(unsigned long long* a, int from, int to) {
  int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) a[i] = 0ull;
}
__global__ void emit_global(const unsigned long long* a, int m, int mode, int W, float* kpts, float* sc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) emit_keypoint(a[i], mode, W, kpts + 2 * i, sc + i);
}

// ---------------------------------------------------------------------------------------------------
// 4. descriptor gather: one warp per keypoint, 256 channels = 8 floats per lane (two float4), HWC map so a
//    cell's 256 channels are 1 KB contiguous.  Dense L2 normalisation of the 4 taps, bilinear blend
//    (align_corners=True, zero padding), final L2 normalisation.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sp_sample_desc_kernel(const float* __restrict__ desc_hwc, int h, int w,
                                                             const float* __restrict__ kpts,
                                                             const int* __restrict__ n_ptr, int n_max,
                                                             float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = n_ptr ? min(*n_ptr, n_max) : n_max;
  if (warp >= n) return;
  const float kx = kpts[2 * warp], ky = kpts[2 * warp + 1];
  // reference: g = (kp - 3.5) / (dim*8 - 4.5) * 2 - 1 ; grid_sample(align_corners=True): ix = (g + 1)/2 * (dim - 1)
  float gx = (kx - 3.5f) / ((float)(w * 8) - 4.5f) * 2.f - 1.f;
  float gy = (ky - 3.5f) / ((float)(h * 8) - 4.5f) * 2.f - 1.f;
  float ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  float wnw = ((fx + 1.f) - ix) * ((fy + 1.f) - iy), wne = (ix - fx) * ((fy + 1.f) - iy);
  float wsw = ((fx + 1.f) - ix) * (iy - fy), wse = (ix - fx) * (iy - fy);
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  auto tap = [&](int yy, int xx, float wt) {
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) return;  // zero padding (warp-uniform branch)
    const float4* p = reinterpret_cast<const float4*>(desc_hwc + ((size_t)yy * w + xx) * 256);
    float4 a = __ldg(p + lane), b = __ldg(p + 32 + lane);
    float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    ss = warp_sum(ss);
    float inv = wt / fmaxf(sqrtf(ss), 1e-12f);
    acc[0] += a.x * inv; acc[1] += a.y * inv; acc[2] += a.z * inv; acc[3] += a.w * inv;
    acc[4] += b.x * inv; acc[5] += b.y * inv; acc[6] += b.z * inv; acc[7] += b.w * inv;
  };
  tap(y0, x0, wnw);
  tap(y0, x1, wne);
  tap(y1, x0, wsw);
  tap(y1, x1, wse);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) ss += acc[c] * acc[c];
  ss = warp_sum(ss);
  float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  float4* o = reinterpret_cast<float4*>(out + (size_t)warp * 256);
  o[lane] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  o[32 + lane] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int i4d_sp_score_map(const float* logits, int h, int w, float* scores, void* stream) {
  I4D_CHECK_ARG(logits && scores && h > 0 && w > 0, "null pointer or empty map");
  int cells = h * w;
  sp_score_map_kernel<<<i4d_cdiv(cells, 128), 128, 0, (cudaStream_t)stream>>>(logits, h, w, scores);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_nms_candidates(const float* scores, int H, int W, int nms_radius, float thr, int border,
                                     unsigned long long* cand_keys, int cand_cap, int* cand_count, float* nms_out,
                                     void* stream) {
  I4D_CHECK_ARG(scores && cand_keys && cand_count && H > 0 && W > 0, "null pointer or empty map");
  I4D_CHECK_ARG(nms_radius >= 0 && nms_radius <= 4, "nms_radius must be in [0, 4]");
  I4D_CHECK_ARG((long long)H * W < 0x7fffffffLL, "score map too large for 32-bit indices");
  cudaStream_t st = (cudaStream_t)stream;
  I4D_CUDA_CALL(cudaMemsetAsync(cand_count, 0, sizeof(int), st));
  int D = NMS_T + 10 * nms_radius;
  size_t smem = (size_t)D * D * (3 * sizeof(float) + 2);
  static bool attr_set = false;
  if (!attr_set) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(i4d_cdiv(W, NMS_T), i4d_cdiv(H, NMS_T));
  sp_nms_kernel<<<grid, NMS_THREADS, smem, st>>>(scores, H, W, nms_radius, thr, border, cand_keys, cand_cap,
                                                 cand_count, nms_out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_select_topk(const unsigned long long* cand_keys, const int* cand_count, int cand_cap, int k,
                                  int W, float* kpts, float* scores, int out_cap, int* n_out,
                                  unsigned long long* spill, void* stream) {
  I4D_CHECK_ARG(cand_keys && cand_count && kpts && scores && n_out && spill, "null pointer");
  I4D_CHECK_ARG(out_cap > 0 && cand_cap > 0 && W > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SEL_MAX_SMEM_KEYS * 8));
    attr_set = true;
  }
  sp_topk_kernel<<<1, SEL_THREADS, SEL_MAX_SMEM_KEYS * 8, st>>>(cand_keys, cand_count, cand_cap, k, W, kpts, scores,
                                                                out_cap, n_out, spill);
  I4D_CUDA_LAUNCH_CHECK();
  // The in-kernel path covers every result of up to 16384 keypoints.  Larger results are only possible when
  // the caller asked for them (k < 0 or k > 16384): then the result size must be read back to sort globally.
  if (k >= 0 && k <= SEL_MAX_SMEM_KEYS) return I4D_OK;
  int m = 0, cnt = 0;
  I4D_CUDA_CALL(cudaMemcpyAsync(&m, n_out, sizeof(int), cudaMemcpyDeviceToHost, st));
  I4D_CUDA_CALL(cudaMemcpyAsync(&cnt, cand_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  I4D_CUDA_CALL(cudaStreamSynchronize(st));
  if (m <= SEL_MAX_SMEM_KEYS) return I4D_OK;
  if (cnt > cand_cap) cnt = cand_cap;
  const int mode = (k < 0 || cnt <= k) ? 1 : 0;
  int p2 = 1;
  while (p2 < m) p2 <<= 1;
  I4D_CHECK_ARG(p2 <= cand_cap, "spill buffer (cand_cap entries) too small for the padded sort");
  if (p2 > m) fill_u64<<<i4d_cdiv(p2 - m, 256), 256, 0, st>>>(spill, m, p2);
  for (int kk = 2; kk <= p2; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) bitonic_global_step<<<i4d_cdiv(p2, 256), 256, 0, st>>>(spill, p2, kk, j);
  emit_global<<<i4d_cdiv(m, 256), 256, 0, st>>>(spill, m, mode, W, kpts, scores);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_sample_descriptors(const float* desc_hwc, int h, int w, const float* kpts, const int* n_dev,
                                         int n_max, float* out, void* stream) {
  I4D_CHECK_ARG(desc_hwc && kpts && out && h > 0 && w > 0, "null pointer or empty map");
  if (n_max <= 0) return I4D_OK;
  sp_sample_desc_kernel<<<i4d_cdiv((long long)n_max * 32, 256), 256, 0, (cudaStream_t)stream>>>(desc_hwc, h, w, kpts,
                                                                                                 n_dev, n_max, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
