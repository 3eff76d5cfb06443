// SuperPoint backbone convolutions (3x3 pad 1, and the 1x1 heads) as implicit GEMMs on tcgen05 with SPLIT bf16
// operands ("bf16x3"): every f32 value v is carried as hi = bf16(v), lo = bf16(v - hi), and
//     x * w  ~=  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi        (relative error ~2^-16, f32 accumulation in TMEM)
// A plain TF32/bf16 convolution flips enough near-tie keypoints to drop the end-to-end match IoU against the f32
// reference to 0.94-0.98; the split form measures 0.996-0.999 (scripts/split_precision_iou.py) at bf16 tensor-core rate.
//
// Layout.  Activations are channels-last bf16 planes [H][W][C] (one hi plane, one lo plane).  A CTA works on a region of
// (8*T) x 16 output pixels.  For each 64-channel chunk one TMA box load per plane brings the region plus its 1-pixel halo
// into shared memory as (8T+2)*18 rows of 128 bytes (128B swizzle); out-of-image coordinates are zero-filled by TMA,
// which IS the convolution's zero padding.  The nine taps are then nine *views* of that one buffer: the UMMA descriptor's
// start address is shifted by (dy*(8T+2) + dx) rows and its stride-byte-offset is one halo row, so operand row m maps to
// pixel (m%8, m/8) of the tile (the swizzle XOR is a function of the absolute shared-memory address, so row-shifted
// starts are legal: scripts/umma_shift_probe.cu).  Weights are packed per (tap, chunk) as [w_hi rows | w_lo rows] x 64
// so that  x_hi * [w_hi | w_lo]  is ONE N = 2*N_T MMA; x_lo * w_hi is a second N = N_T MMA into the first half of the
// same accumulator; the epilogue adds the two halves.
//
// Roles (384 threads, one persistent CTA per SM): warp 0 (one elected lane) = activation TMA producer, warp 1 = weight TMA
// producer, warp 2 = MMA issuer, warps 4-11 = epilogue (tcgen05.ld -> +bias, ReLU, optional fused 2x2 max-pool via
// warp shuffles -> split back to hi/lo bf16 or f32).  Accumulators are double-buffered in TMEM (all 512 columns), so the
// epilogue of one region overlaps the MMAs of the next.
//
// FUSE variant (conv1b only): the activation planes of conv1a never exist in HBM.  Four extra warps compute relu(conv1a) of the
// region's 18 x 18 halo box straight from the grey image (1 -> 64 channels, 9 f32 FMAs per value in the order of sp_conv1a.cu,
// so the values are bit-identical to the stand-alone kernel's), split them into (hi, lo) and write them into the A slot in the
// very layout the TMA box load produces (rows of 128 bytes, 128B swizzle: 16-byte chunk c of row r at chunk c ^ (r & 7));
// pixels outside the image are zeros (conv1b's padding).  That removes conv1a's 1 GB of stores per 2000 x 2000 tile and
// conv1b's 1.3 GB of loads; the producer's ~2 k instructions per region hide behind the region's ~7 k clk of MMA work.
//
// Reference behaviour replaced: conv1b..convDb of thirdparty/SuperGlue/models/superpoint.py:154-168,193-196 (and the
// LightGlue copy, lightglue/superpoint.py:155-170,189-192), which run as f32 cuDNN/MKL convolutions there.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/icepy4d_b200.h"

#define CV_THREADS 384
#define CV_TILE_H 16
#define CV_A_SLOTS 2
#define CV_B_STAGES 3

struct ConvTcParams {
  int H, W, KC, taps, NT, tiles_x, tiles_y, halo, relu, pool;
  const float* bias;
  __nv_bfloat16* y_hi; __nv_bfloat16* y_lo; int ld16;
  float* y32; int ld32; int planar;
  int cout;
  int fmt;          // 16-bit operand format of the split planes and weights: 0 = IEEE half (f16x3), 1 = bfloat16 (bf16x3)
  const float* img; const float* w1a; const float* b1a;     // FUSE: grey image [H][W], conv1a weights [64][9] and bias [64]
};

// v -> (hi, lo) pair of 16-bit values, packed two by two: hi = round16(v), lo = round16(v - hi).  bf16x3 carries 16 mantissa
// bits, f16x3 22 (exact f32 products of the three partial MMAs either way); halves are clamped to the finite f16 range.
__device__ __forceinline__ void cv_split2(float a, float b, int fmt, uint32_t& hi, uint32_t& lo) {
  if (fmt) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    a = fminf(fmaxf(a, -65504.f), 65504.f); b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

template <int N_T, int T> struct ConvCfg {
  static constexpr int ACC_COLS = 2 * N_T;                       // [x_hi*w_hi + x_lo*w_hi | x_hi*w_lo]
  static constexpr int TMEM_COLS = 2 * T * ACC_COLS;             // double-buffered
  static constexpr int BW_MAX = 8 * T + 2, BH_MAX = CV_TILE_H + 2;
  static constexpr int A_PLANE = ((BW_MAX * BH_MAX * 128) + 1023) & ~1023;
  static constexpr int A_SLOT = 2 * A_PLANE;
  static constexpr int B_STAGE = 2 * N_T * 128;
  static constexpr int SMEM = CV_A_SLOTS * A_SLOT + CV_B_STAGES * B_STAGE + 1024;
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256, "TMEM allocation must be a power of two");
  static_assert(SMEM <= 232448 - 512, "shared memory budget");
};

#define CV_FUSE_THREADS 128      // conv1a producer warps of the FUSE variant (warps 12-15)

template <int N_T, int T, bool FUSE>
__global__ void __launch_bounds__(CV_THREADS + (FUSE ? CV_FUSE_THREADS : 0), 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmHi,
                                                                 const __grid_constant__ CUtensorMap tmLo,
                                                                 const __grid_constant__ CUtensorMap tmW, ConvTcParams p) {
  using C = ConvCfg<N_T, T>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + CV_A_SLOTS * C::A_SLOT;
  __shared__ __align__(8) uint64_t a_full[CV_A_SLOTS], a_empty[CV_A_SLOTS], b_full[CV_B_STAGES], b_empty[CV_B_STAGES],
      t_full[2], t_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[512];
  __shared__ float patch_s[FUSE ? 2 : 1][FUSE ? (8 * T + 4) * (CV_TILE_H + 4) : 1];      // FUSE: grey patch under the halo box, double-buffered

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bw = 8 * T + 2 * p.halo, bh = CV_TILE_H + 2 * p.halo;
  const int tiles = p.tiles_x * p.tiles_y;
  const int units = tiles * p.NT;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmHi); tc::prefetch_tmap(&tmLo); tc::prefetch_tmap(&tmW);
    for (int s = 0; s < CV_A_SLOTS; ++s) { tc::mbar_init(&a_full[s], FUSE ? CV_FUSE_THREADS : 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < CV_B_STAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 256); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.NT * N_T; i += blockDim.x) bias_s[i] = p.bias[i];
  if (warp == 3) tc::tmem_alloc(&tmem_base_s, C::TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  bool is_producer = false;
  if constexpr (FUSE) if (warp >= CV_THREADS / 32) {
    is_producer = true;
    // ---- conv1a producer (FUSE): thread = (8-channel group cg, every 16th pixel of the 18 x 18 halo box); the 72 weights of
    // the group stay in registers ----
    const int pt = threadIdx.x - CV_THREADS, cg = pt & 7;
    float w[9][8], b[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      b[c] = __ldg(p.b1a + cg * 8 + c);
#pragma unroll
      for (int t = 0; t < 9; ++t) w[t][c] = __ldg(p.w1a + (cg * 8 + c) * 9 + t);
    }
    constexpr int BW = 8 * T + 2, BH = CV_TILE_H + 2, PW = BW + 2, PH = BH + 2;     // halo box of conv1b; grey patch under it
    static_assert(BW % 3 == 0 && (BH * (BW / 3) * 8) <= 7 * CV_FUSE_THREADS, "item decomposition");
    int ai = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ai) {
      const int rem = u % tiles;
      const int x0 = (rem % p.tiles_x) * 8 * T - 1, y0 = (rem / p.tiles_x) * CV_TILE_H - 1;     // origin of the halo box
      const int s = ai % CV_A_SLOTS;
      float* patch = patch_s[ai & 1];
      // grey patch (PH x PW, zero outside the image = conv1a's padding): one coalesced pass, then everything reads shared memory
      for (int i = pt; i < PW * PH; i += CV_FUSE_THREADS) {
        const int yy = y0 - 1 + i / PW, xx = x0 - 1 + i % PW;
        patch[i] = (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) ? __ldg(p.img + (size_t)yy * p.W + xx) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CV_FUSE_THREADS) : "memory");   // patch complete; everybody is done with the patch of 2 tiles ago
      tc::mbar_wait(&a_empty[s], ((ai / CV_A_SLOTS) & 1) ^ 1);
      uint8_t* dst = a_base + s * C::A_SLOT;
      // item = (halo row, 3-pixel segment, channel group): BH * BW/3 * 8 = 864 items, 7 rounds of 128 threads
#pragma unroll 1
      for (int it = pt; it < BH * (BW / 3) * 8; it += CV_FUSE_THREADS) {
        const int q = it >> 3, row = q / (BW / 3), px0 = (q - row * (BW / 3)) * 3;
        float v[3][5];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) v[dy][dx] = patch[(row + dy) * PW + px0 + dx];
        const int yy = y0 + row;
#pragma unroll
        for (int px = 0; px < 3; ++px) {
          const int xx = x0 + px0 + px;
          uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {           // outside the image the ACTIVATION is zero (conv1b's padding)
            float2 acc[4] = {make_float2(b[0], b[1]), make_float2(b[2], b[3]), make_float2(b[4], b[5]), make_float2(b[6], b[7])};
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float a = v[t / 3][px + t % 3];
              const float2 a2 = make_float2(a, a);
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[c] = __ffma2_rn(a2, make_float2(w[t][2 * c], w[t][2 * c + 1]), acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) cv_split2(fmaxf(acc[c].x, 0.f), fmaxf(acc[c].y, 0.f), p.fmt, hi[c], lo[c]);
          }
          const int r = row * BW + px0 + px;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((cg ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + C::A_PLANE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      tc::fence_proxy_async_smem();          // my generic-proxy stores before the tensor core's (async proxy) reads
      tc::mbar_arrive(&a_full[s]);
    }
  }
  if (is_producer) {
    // (done above)
  } else if (warp == 0 && !FUSE) {
    // ---- activation producer: one (hi, lo) halo box per 64-channel chunk ----
    if (tc::elect_one()) {
      const uint32_t bytes = 2u * (uint32_t)(bw * bh * 128);
      int ai = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int rem = u % tiles;
        const int x0 = (rem % p.tiles_x) * 8 * T - p.halo, y0 = (rem / p.tiles_x) * CV_TILE_H - p.halo;
        for (int kc = 0; kc < p.KC; ++kc, ++ai) {
          const int s = ai % CV_A_SLOTS;
          tc::mbar_wait(&a_empty[s], ((ai / CV_A_SLOTS) & 1) ^ 1);
          uint8_t* dst = a_base + s * C::A_SLOT;
          tc::mbar_arrive_expect_tx(&a_full[s], bytes);
          tc::tma_load_3d(dst, &tmHi, &a_full[s], kc * 64, x0, y0);
          tc::tma_load_3d(dst + C::A_PLANE, &tmLo, &a_full[s], kc * 64, x0, y0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- weight producer: one [w_hi | w_lo] x 64 tile per (tap, chunk) ----
    if (tc::elect_one()) {
      int bi = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = u / tiles;
        for (int kc = 0; kc < p.KC; ++kc)
          for (int tap = 0; tap < p.taps; ++tap, ++bi) {
            const int s = bi % CV_B_STAGES;
            tc::mbar_wait(&b_empty[s], ((bi / CV_B_STAGES) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&b_full[s], C::B_STAGE);
            tc::tma_load_2d(b_base + s * C::B_STAGE, &tmW, &b_full[s], 0, ((nt * p.taps + tap) * p.KC + kc) * 2 * N_T);
          }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---- MMA issuer: one elected lane; descriptors advance by adding (bytes >> 4) to a precomputed low word ----
    if (tc::elect_one()) {
      const uint32_t idesc_cat = tc::make_idesc(128, 2 * N_T, 0, 0, p.fmt);
      const uint32_t idesc_hi = tc::make_idesc(128, N_T, 0, 0, p.fmt);
      constexpr uint32_t b_hiword = tc::desc_hi_sw128(1024);
      const uint32_t a_hiword = tc::desc_hi_sw128((uint32_t)bw * 128);
      const uint32_t a_lo0 = tc::desc_lo_sw128(tc::smem_u32(a_base)), b_lo0 = tc::desc_lo_sw128(tc::smem_u32(b_base));
      const uint32_t row_step = (uint32_t)bw * 8;                      // one halo row, in 16-byte units
      int ai = 0, bi = 0, ui = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int ab = ui & 1;
        tc::mbar_wait(&t_empty[ab], ((ui >> 1) & 1) ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(ab * T * C::ACC_COLS);
        for (int kc = 0; kc < p.KC; ++kc, ++ai) {
          const int as = ai % CV_A_SLOTS;
          tc::mbar_wait(&a_full[as], (ai / CV_A_SLOTS) & 1);
          tc::tcgen05_fence_after();
          uint32_t a_tap = a_lo0 + (uint32_t)(as * (C::A_SLOT >> 4));   // tap (0,0); taps advance by 8 (dx) / row_step (dy)
          int dx = 0;
          for (int tap = 0; tap < p.taps; ++tap, ++bi) {
            const int bs = bi % CV_B_STAGES;
            tc::mbar_wait(&b_full[bs], (bi / CV_B_STAGES) & 1);
            tc::tcgen05_fence_after();
            const uint32_t b_lo = b_lo0 + (uint32_t)(bs * (C::B_STAGE >> 4));
#pragma unroll
            for (int t = 0; t < T; ++t) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tc::umma_f16_parts(d0 + t * C::ACC_COLS, a_tap + t * 64 + k * 2, a_hiword, b_lo + k * 2, b_hiword, idesc_cat,
                                   (kc | tap | k) ? 1u : 0u);
                tc::umma_f16_parts(d0 + t * C::ACC_COLS, a_tap + (C::A_PLANE >> 4) + t * 64 + k * 2, a_hiword, b_lo + k * 2, b_hiword,
                                   idesc_hi, 1u);
              }
            }
            tc::umma_commit(&b_empty[bs]);
            if (++dx == 3) { dx = 0; a_tap += row_step - 16; } else a_tap += 8;
          }
          tc::umma_commit(&a_empty[as]);
        }
        tc::umma_commit(&t_full[ab]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---- epilogue: 8 warps.  TMEM lane quarter q = warp % 4 <-> tile rows 4q..4q+3 (8 pixels each); the two warps of a
    // quarter split the unit's T * N_T/32 (tile, 32-channel chunk) items, two each, and keep the second item's TMEM load
    // in flight while the first is written out. ----
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int px = m & 7, py = m >> 3;
    const int Ho = p.pool ? (p.H >> 1) : p.H, Wo = p.pool ? (p.W >> 1) : p.W;
    constexpr int CH = N_T / 32;                       // chunks per tile
    static_assert(T * CH == 4, "two items per epilogue warp");
    const bool odd = lane & 1, up = (lane >> 3) & 1;
    int ui = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
      const int ab = ui & 1;
      const int nt = u / tiles, rem = u % tiles;
      const int x0 = (rem % p.tiles_x) * 8 * T, y0 = (rem / p.tiles_x) * CV_TILE_H;
      tc::mbar_wait(&t_full[ab], (ui >> 1) & 1);
      tc::tcgen05_fence_after();
      const uint32_t d_unit = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * T * C::ACC_COLS);
      uint32_t v1[32], v2[32];
      float f[32];
      auto issue = [&](int item) {
        const uint32_t d = d_unit + (uint32_t)((item / CH) * C::ACC_COLS + (item % CH) * 32);
        tc::tmem_ld32(d, v1);
        tc::tmem_ld32(d + N_T, v2);
      };
      auto combine = [&](int item) {
        const int ch0 = nt * N_T + (item % CH) * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v1[j]) + __uint_as_float(v2[j]) + bias_s[ch0 + j];
          f[j] = p.relu ? fmaxf(x, 0.f) : x;
        }
      };
      auto write = [&](int item) {
        const int t = item / CH, ch0 = nt * N_T + (item % CH) * 32;
        int gx = x0 + 8 * t + px, gy = y0 + py;
        if (p.pool) {
          // 2x2 max over lanes (m, m^1, m^8, m^9): each exchange hands over half of the channels, so 24 shuffles instead of
          // 64 and every lane ends up owning 8 pooled channels of the output pixel (16-byte stores from all lanes)
          float k[16], z[8];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float r = __shfl_xor_sync(0xffffffffu, odd ? f[j] : f[j + 16], 1);
            k[j] = fmaxf(odd ? f[j + 16] : f[j], r);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float r = __shfl_xor_sync(0xffffffffu, up ? k[j] : k[j + 8], 8);
            z[j] = fmaxf(up ? k[j + 8] : k[j], r);
          }
          gx >>= 1; gy >>= 1;
          if (gx < Wo && gy < Ho) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) cv_split2(z[j], z[j + 1], p.fmt, hi[j >> 1], lo[j >> 1]);
            const size_t o = ((size_t)gy * Wo + gx) * p.ld16 + ch0 + (odd ? 16 : 0) + (up ? 8 : 0);
            *reinterpret_cast<uint4*>(p.y_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(p.y_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          return;
        }
        if (gx >= p.W || gy >= p.H) return;
        const size_t pix = (size_t)gy * p.W + gx;
        if (p.y_hi) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) cv_split2(f[j], f[j + 1], p.fmt, hi[j >> 1], lo[j >> 1]);
          uint4* oh = reinterpret_cast<uint4*>(p.y_hi + pix * p.ld16 + ch0);
          uint4* ol = reinterpret_cast<uint4*>(p.y_lo + pix * p.ld16 + ch0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
        if (p.y32) {
          if (p.planar) {
            const size_t plane = (size_t)p.H * p.W;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ch0 + j < p.cout) p.y32[(size_t)(ch0 + j) * plane + pix] = f[j];
          } else if (ch0 + 32 <= p.cout && (p.ld32 & 3) == 0) {
            float4* o = reinterpret_cast<float4*>(p.y32 + pix * p.ld32 + ch0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ch0 + j < p.cout) p.y32[pix * p.ld32 + ch0 + j] = f[j];
          }
        }
      };
      issue(2 * g);
      tc::tmem_ld_wait();
      combine(2 * g);
      issue(2 * g + 1);
      write(2 * g);
      __syncwarp();
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      combine(2 * g + 1);
      tc::mbar_arrive(&t_empty[ab]);          // accumulator values are in registers: the MMA warp may overwrite them
      write(2 * g + 1);
      __syncwarp();
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 3) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int N_T, int T, bool FUSE = false>
static int launch_conv(const CUtensorMap& tmHi, const CUtensorMap& tmLo, const CUtensorMap& tmW, ConvTcParams p, cudaStream_t st) {
  using C = ConvCfg<N_T, T>;
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(conv_tc_kernel<N_T, T, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  }
  p.tiles_x = i4d_cdiv(p.W, 8 * T);
  p.tiles_y = i4d_cdiv(p.H, CV_TILE_H);
  const long long units = (long long)p.tiles_x * p.tiles_y * p.NT;
  const int grid = (int)(units < i4d_num_sms() ? units : i4d_num_sms());
  conv_tc_kernel<N_T, T, FUSE><<<grid, CV_THREADS + (FUSE ? CV_FUSE_THREADS : 0), C::SMEM, st>>>(tmHi, tmLo, tmW, p);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_conv_tile_cout(int cout_pad) { return cout_pad == 64 ? 64 : 128; }

extern "C" __attribute__((visibility("default"))) int i4d_conv_bf16x3_tc(
    const void* x_hi, const void* x_lo, int H, int W, int Cin, const void* w_packed, const float* bias, int cout_pad, int cout,
    int ksize, int relu, int pool, void* y_hi, void* y_lo, float* y32, int ld32, int y32_planar, int operand_format, void* stream) {
  I4D_CHECK_ARG(x_hi && x_lo && w_packed && bias, "null pointer");
  I4D_CHECK_ARG(H > 0 && W > 0 && Cin > 0 && Cin % 64 == 0, "Cin must be a positive multiple of 64");
  I4D_CHECK_ARG(ksize == 1 || ksize == 3, "kernel size must be 1 or 3");
  I4D_CHECK_ARG(operand_format == 0 || operand_format == 1, "operand_format: 0 = f16 split planes, 1 = bf16 split planes");
  I4D_CHECK_ARG(cout_pad > 0 && cout_pad % 64 == 0 && cout_pad <= 512 && cout > 0 && cout <= cout_pad, "bad output channel counts");
  I4D_CHECK_ARG((y_hi != nullptr) == (y_lo != nullptr) && (y_hi || y32), "need split planes and/or an f32 output");
  I4D_CHECK_ARG(!pool || (!y32 && H >= 2 && W >= 2), "the fused 2x2 max-pool writes split planes only");
  I4D_CHECK_ARG(!y_hi || cout == cout_pad, "split-plane output needs cout == cout_pad");
  I4D_CHECK_ARG(!y32 || y32_planar || ld32 >= cout, "ld32 too small");
  const int n_t = i4d_conv_tile_cout(cout_pad);
  I4D_CHECK_ARG(cout_pad % n_t == 0, "cout_pad must be 64 or a multiple of 128");
  const int taps = ksize * ksize, halo = ksize / 2, T = n_t == 64 ? 2 : 1;
  ConvTcParams p{};
  p.H = H; p.W = W; p.KC = Cin / 64; p.taps = taps; p.NT = cout_pad / n_t; p.halo = halo; p.relu = relu; p.pool = pool;
  p.bias = bias; p.y_hi = reinterpret_cast<__nv_bfloat16*>(y_hi); p.y_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
  p.ld16 = cout_pad; p.y32 = y32; p.ld32 = ld32; p.planar = y32_planar; p.cout = cout; p.fmt = operand_format;
  CUtensorMap tmHi, tmLo, tmW;
  const uint32_t bw = 8 * T + 2 * halo, bh = CV_TILE_H + 2 * halo;
  if (int rc = i4d_make_tmap_hwc_bf16(&tmHi, x_hi, (uint64_t)H, (uint64_t)W, (uint64_t)Cin, bh, bw)) return rc;
  if (int rc = i4d_make_tmap_hwc_bf16(&tmLo, x_lo, (uint64_t)H, (uint64_t)W, (uint64_t)Cin, bh, bw)) return rc;
  const uint64_t wrows = (uint64_t)p.NT * taps * p.KC * 2 * n_t;
  if (int rc = i4d_make_tmap_2d_bf16(&tmW, w_packed, wrows, 64, 64, 2 * n_t, 64)) return rc;
  if (n_t == 64) return launch_conv<64, 2>(tmHi, tmLo, tmW, p, (cudaStream_t)stream);
  return launch_conv<128, 1>(tmHi, tmLo, tmW, p, (cudaStream_t)stream);
}

// conv1a + conv1b in one kernel (FUSE variant above): image [H][W] f32 -> relu(conv1b(relu(conv1a(image)))) as split planes,
// optionally 2x2 max-pooled.  conv1a's weights are f32 [64][9] + bias [64]; conv1b's are packed as for i4d_conv_bf16x3_tc.
extern "C" __attribute__((visibility("default"))) int i4d_sp_conv1ab_tc(
    const float* image, int H, int W, const float* w1a, const float* b1a, const void* w1b_packed, const float* b1b, int pool,
    void* y_hi, void* y_lo, int operand_format, void* stream) {
  I4D_CHECK_ARG(image && w1a && b1a && w1b_packed && b1b && y_hi && y_lo, "null pointer");
  I4D_CHECK_ARG(H > 0 && W > 0 && (!pool || (H >= 2 && W >= 2)), "bad image size");
  I4D_CHECK_ARG(operand_format == 0 || operand_format == 1, "operand_format: 0 = f16 split planes, 1 = bf16 split planes");
  ConvTcParams p{};
  p.H = H; p.W = W; p.KC = 1; p.taps = 9; p.NT = 1; p.halo = 1; p.relu = 1; p.pool = pool;
  p.bias = b1b; p.y_hi = reinterpret_cast<__nv_bfloat16*>(y_hi); p.y_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
  p.ld16 = 64; p.y32 = nullptr; p.ld32 = 0; p.planar = 0; p.cout = 64; p.fmt = operand_format;
  p.img = image; p.w1a = w1a; p.b1a = b1a;
  CUtensorMap tmW;
  if (int rc = i4d_make_tmap_2d_bf16(&tmW, w1b_packed, (uint64_t)9 * 2 * 64, 64, 64, 2 * 64, 64)) return rc;
  return launch_conv<64, 2, true>(tmW, tmW, tmW, p, (cudaStream_t)stream);      // the activation tensor maps are unused
}
