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
#define NMS_THREADS 1024
#define NMS_WARPS (NMS_THREADS / 32)

// 2-D sweep of the D x D staging tile: warp w takes rows w, w + 32, ...; lanes stride the row (no integer division)
#define NMS_FOR_TILE(y, x, D) for (int y = (int)(threadIdx.x >> 5); y < (D); y += NMS_WARPS) for (int x = (int)(threadIdx.x & 31); x < (D); x += 32)

// One (2R+1)^2 max-pool of the staging tile, separable, with register sliding windows: a thread produces 8 consecutive outputs of
// a row (then of a column) from 8 + 2R inputs — 1.75 shared-memory reads per output at R = 3 instead of 7.  Row pass: lane ->
// row (the row pitch Dp is odd, so the 32 lanes hit 32 different banks); column pass: lane -> column.  `src(i)` reads element i
// of the [D][Dp] tile, `dst(y, x, m)` consumes the pooled value; positions outside the tile count as -inf.
// K = position of this pool in the chain of five (1..5): the final mask is needed in the central NMS_T x NMS_T box only, pool k's
// output therefore in [kR, D - kR)^2 (each later pool reaches R further).  The values inside that box are the ones a full-tile
// pool would compute (its inputs lie in the previous pool's box), everything outside would never be read.  At R = 3 this is 11
// rounds of 1024 tasks per tile instead of 20.
template <int R, int K, typename Src, typename Dst>
__device__ __forceinline__ void nms_pool(const Src& src, float* __restrict__ T1, const Dst& dst, int D, int Dp) {
  const int lo = K * R, hi = D - K * R, w = hi - lo;                      // output box [lo, hi)^2
  const int ry0 = lo - R, nrows = w + 2 * R;                              // rows of the row pass: [lo - R, hi + R), inside [0, D)
  const int NS = (w + 7) >> 3;
  for (int task = threadIdx.x; task < nrows * NS; task += NMS_THREADS) {
    const int st = task / nrows, y = ry0 + (task - st * nrows), x0 = lo + st * 8;
    float v[8 + 2 * R];
#pragma unroll
    for (int k = 0; k < 8 + 2 * R; ++k) {
      const int xx = x0 - R + k;
      v[k] = (xx >= 0 && xx < D) ? src(y * Dp + xx) : -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float m = v[i];
#pragma unroll
      for (int k = 1; k <= 2 * R; ++k) m = fmaxf(m, v[i + k]);
      if (x0 + i < hi) T1[y * Dp + x0 + i] = m;
    }
  }
  __syncthreads();
  for (int task = threadIdx.x; task < w * NS; task += NMS_THREADS) {
    const int st = task / w, x = lo + (task - st * w), y0 = lo + st * 8;
    float v[8 + 2 * R];
#pragma unroll
    for (int k = 0; k < 8 + 2 * R; ++k) {
      const int yy = y0 - R + k;
      v[k] = (yy >= 0 && yy < D) ? T1[yy * Dp + x] : -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float m = v[i];
#pragma unroll
      for (int k = 1; k <= 2 * R; ++k) m = fmaxf(m, v[i + k]);
      if (y0 + i < hi) dst(y0 + i, x, m);
    }
  }
  __syncthreads();
}

template <int V> struct nms_const { static constexpr int value = V; };

template <int R>
__global__ void __launch_bounds__(NMS_THREADS) sp_nms_kernel(const float* __restrict__ scores, int H, int W,
                                                             float thr, int border,
                                                             unsigned long long* __restrict__ cand, int cand_cap,
                                                             int* __restrict__ cand_count, float* __restrict__ nms_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int r = R;
  constexpr int halo = 5 * R;
  constexpr int D = NMS_T + 2 * halo;
  constexpr int Dp = D | 1;                                      // odd row pitch: conflict-free lane -> row accesses
  float* S0 = reinterpret_cast<float*>(smem_raw);
  float* T1 = S0 + D * Dp;
  float* X = T1 + D * Dp;
  unsigned char* M = reinterpret_cast<unsigned char*>(X + D * Dp);
  unsigned char* P = M + D * Dp;
  __shared__ int s_n, s_base;
  (void)r;

  const int ty0 = blockIdx.y * NMS_T - halo, tx0 = blockIdx.x * NMS_T - halo;
  auto inimg = [&](int y, int x) {
    int gy = ty0 + y, gx = tx0 + x;
    return gy >= 0 && gy < H && gx >= 0 && gx < W;
  };
  NMS_FOR_TILE(y, x, D) S0[y * Dp + x] = inimg(y, x) ? __ldg(scores + (size_t)(ty0 + y) * W + (tx0 + x)) : -INFINITY;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  nms_pool<R, 1>([&](int j) { return S0[j]; }, T1,
                 [&](int y, int x, float m) { M[y * Dp + x] = (inimg(y, x) && S0[y * Dp + x] == m) ? 1 : 0; }, D, Dp);
  auto suppress_and_add = [&](auto k_supp, auto k_max) {                  // one iteration of simple_nms (superpoint.py:56-62)
    nms_pool<R, decltype(k_supp)::value>([&](int j) { return M[j] ? 1.f : 0.f; }, T1,
                [&](int y, int x, float m) {
                  const bool supp = m > 0.f;
                  P[y * Dp + x] = supp ? 1 : 0;
                  X[y * Dp + x] = inimg(y, x) ? (supp ? 0.f : S0[y * Dp + x]) : -INFINITY;
                }, D, Dp);
    nms_pool<R, decltype(k_max)::value>([&](int j) { return X[j]; }, T1,
                [&](int y, int x, float m) {
                  const bool nm = inimg(y, x) && (X[y * Dp + x] == m);
                  if (nm && !P[y * Dp + x]) M[y * Dp + x] = 1;
                }, D, Dp);
  };
  suppress_and_add(nms_const<2>{}, nms_const<3>{});
  suppress_and_add(nms_const<4>{}, nms_const<5>{});
  // threshold + border + block-aggregated compaction (X is free now: reuse it as the per-CTA key list)
  // with R >= 1 the surviving maxima are at least R + 1 apart (<= 1024 per tile: 8 KB fits X); R = 0 keeps up to 4096 keys
  // in a dedicated area behind the masks
  unsigned long long* keys = (R == 0)
      ? reinterpret_cast<unsigned long long*>(smem_raw + (((size_t)D * Dp * (3 * sizeof(float) + 2) + 15) & ~(size_t)15))
      : reinterpret_cast<unsigned long long*>(X);
  NMS_FOR_TILE(y, x, NMS_T) {
    int gy = blockIdx.y * NMS_T + y, gx = blockIdx.x * NMS_T + x;
    if (gy >= H || gx >= W) continue;
    int si = (y + halo) * Dp + (x + halo);
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

// ---------------------------------------------------------------------------------------------------
// 3b. multi-CTA top-k for 1 <= k <= 16384 (the production path: k = 2048..16384 of ~10^5 candidates).
//     Radix select on the 32 score bits in three grid-wide histogram passes (11 + 11 + 10 bits), one compaction pass
//     (score > T straight to the output list, score == T to a tie list), and a single-CTA finish (ties by lowest
//     index, bitonic sort in shared memory, emit).  Same result and order as sp_topk_kernel.
// ---------------------------------------------------------------------------------------------------
#define TK_BINS 2048
struct TopkState {
  unsigned int hist[3][TK_BINS];
  unsigned long long prefix;      // selected high bits of the 64-bit key so far
  int remaining;                  // how many keys are still needed inside the current prefix bucket
  int n, m, keep_all;
  int n_gt, n_tie;
};
__device__ __forceinline__ int tk_shift(int pass) { return pass == 0 ? 53 : (pass == 1 ? 42 : 32); }
__device__ __forceinline__ unsigned int tk_digit(unsigned long long key, int pass) {
  return pass == 2 ? (unsigned int)(key >> 32) & 1023u : (unsigned int)(key >> tk_shift(pass)) & 2047u;
}

__global__ void topk_init_kernel(const int* __restrict__ cand_count, int cand_cap, int k, int out_cap, TopkState* st,
                                 int* __restrict__ n_out) {
  for (int i = threadIdx.x; i < 3 * TK_BINS; i += blockDim.x) (&st->hist[0][0])[i] = 0u;
  if (threadIdx.x == 0) {
    int n = min(*cand_count, cand_cap);
    int keep_all = n <= k;
    int m = keep_all ? n : k;
    if (m > out_cap) m = out_cap;
    st->n = n; st->m = m; st->keep_all = keep_all; st->prefix = 0ull; st->remaining = k; st->n_gt = 0; st->n_tie = 0;
    *n_out = m;
  }
}

__global__ void __launch_bounds__(256) topk_hist_kernel(const unsigned long long* __restrict__ cand, TopkState* st, int pass) {
  __shared__ unsigned int h[TK_BINS];
  if (st->keep_all) return;
  for (int i = threadIdx.x; i < TK_BINS; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const int n = st->n;
  const unsigned long long prefix = st->prefix;
  const int hs = pass == 0 ? 64 : tk_shift(pass - 1);          // bits above this are fixed by the prefix
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long key = cand[i];
    bool match = (hs >= 64) || ((key >> hs) == (prefix >> hs));
    if (match) atomicAdd(&h[tk_digit(key, pass)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TK_BINS; i += blockDim.x)
    if (h[i]) atomicAdd(&st->hist[pass][i], h[i]);
}

__global__ void __launch_bounds__(1024) topk_scan_kernel(TopkState* st, int pass) {
  // find the highest bin b with (count of keys in bins > b) < remaining <= (count in bins >= b)
  __shared__ unsigned int part[1024];
  if (st->keep_all) return;
  const int t = threadIdx.x;
  // two bins per thread, processed from the top: thread t owns bins 2047-2t and 2046-2t
  unsigned int c0 = st->hist[pass][TK_BINS - 1 - 2 * t], c1 = st->hist[pass][TK_BINS - 2 - 2 * t];
  part[t] = c0 + c1;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {                  // inclusive scan from the top
    unsigned int v = (t >= off) ? part[t - off] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const unsigned int rem = (unsigned int)st->remaining;
  const unsigned int above = part[t] - (c0 + c1);              // keys in bins above my pair
  if (above < rem && rem <= part[t]) {
    int b; unsigned int gt;
    if (above + c0 >= rem) { b = TK_BINS - 1 - 2 * t; gt = above; }
    else { b = TK_BINS - 2 - 2 * t; gt = above + c0; }
    st->prefix |= (unsigned long long)b << tk_shift(pass);
    st->remaining = (int)(rem - gt);
  }
}

__global__ void __launch_bounds__(256) topk_compact_kernel(const unsigned long long* __restrict__ cand, TopkState* st,
                                                           unsigned long long* __restrict__ sel, unsigned long long* __restrict__ ties) {
  if (st->keep_all) return;
  const int n = st->n;
  const unsigned int T = (unsigned int)(st->prefix >> 32);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long key = cand[i];
    unsigned int sc = (unsigned int)(key >> 32);
    if (sc > T) sel[atomicAdd(&st->n_gt, 1)] = key;
    else if (sc == T) ties[atomicAdd(&st->n_tie, 1)] = key;
  }
}

// block-wide: k-th largest 64-bit key of a[0..n) (1 <= k <= n), 8 passes of 8 bits
__device__ unsigned long long block_radix_kth(const unsigned long long* __restrict__ a, int n, int k, unsigned int* hist,
                                              unsigned long long* s_prefix, int* s_remaining) {
  if (threadIdx.x == 0) { *s_prefix = 0ull; *s_remaining = k; }
  __syncthreads();
  for (int pass = 7; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int shift = pass * 8;
    const unsigned long long pmask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
    const unsigned long long prefix = *s_prefix;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned long long key = a[i];
      if ((key & pmask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int rem = *s_remaining;
      int b = 255;
      for (; b > 0; --b) {
        if ((int)hist[b] >= rem) break;
        rem -= (int)hist[b];
      }
      *s_remaining = rem;
      *s_prefix = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  return *s_prefix;
}

// Gathers the result set (unordered) into `sel[0, m)`: the keys above the k-th value are there already (compaction order), the
// `need` ties with the lowest linear indices are appended; in keep-all mode every candidate goes in with its halves swapped
// (index-major keys: the reference's row-major nonzero() order).
__global__ void __launch_bounds__(SEL_THREADS) topk_finish_kernel(const unsigned long long* __restrict__ cand, TopkState* st,
                                                                  unsigned long long* __restrict__ sel,
                                                                  const unsigned long long* __restrict__ ties) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_cnt;
  const int m = st->m, keep_all = st->keep_all;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (keep_all) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) { unsigned long long key = cand[i]; sel[i] = (key << 32) | (key >> 32); }
    return;
  }
  const int n_gt = min(st->n_gt, m), n_tie = st->n_tie, need = min(st->remaining, m - n_gt);
  if (need > 0 && n_tie > 0) {
    // the `need` ties with the largest keys = the lowest linear indices
    unsigned long long kth = (need >= n_tie) ? 0ull : block_radix_kth(ties, n_tie, need, hist, &s_prefix, &s_remaining);
    __syncthreads();
    for (int i = threadIdx.x; i < n_tie; i += blockDim.x) {
      unsigned long long key = ties[i];
      if (key >= kth) {
        int slot = atomicAdd(&s_cnt, 1);
        if (slot < need) sel[n_gt + slot] = key;
      }
    }
  }
}

// Sort + emit without a sort: the keys are distinct (the linear index is part of every key), so the output position of a key is
// the number of keys greater than it.  CTA = 32 keys x 8 threads each (a thread compares its key with every 8th key of the set in
// shared memory: consecutive lanes read consecutive words, no bank conflicts), 3 shuffles, one write.  m <= 16384 keys: 256 / 512
// independent CTAs of ~1 k iterations instead of ONE CTA running a 91-step bitonic network (105 us for 8192 keys).
#define RANK_THREADS 256
__global__ void __launch_bounds__(RANK_THREADS) topk_rank_emit_kernel(const unsigned long long* __restrict__ keys, const TopkState* st,
                                                                      int W, float* __restrict__ kpts, float* __restrict__ sc) {
  extern __shared__ __align__(16) unsigned long long rk[];
  const int m = st->m, mode = st->keep_all ? 1 : 0;
  if ((int)blockIdx.x * (RANK_THREADS / 8) >= m) return;
  for (int i = threadIdx.x; i < m; i += RANK_THREADS) rk[i] = keys[i];
  __syncthreads();
  const int part = threadIdx.x & 7, idx = blockIdx.x * (RANK_THREADS / 8) + (threadIdx.x >> 3);
  const unsigned long long mine = idx < m ? rk[idx] : 0ull;
  int cnt = 0;
#pragma unroll 8
  for (int j = part; j < m; j += 8) cnt += rk[j] > mine ? 1 : 0;
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
  if (part == 0 && idx < m) emit_keypoint(mine, mode, W, kpts + 2 * cnt, sc + cnt);
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
  const int D = NMS_T + 10 * nms_radius, Dp = D | 1;
  const size_t smem = (size_t)D * Dp * (3 * sizeof(float) + 2) + (nms_radius == 0 ? 16 + NMS_T * NMS_T * 8 : 0);
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_nms_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid(i4d_cdiv(W, NMS_T), i4d_cdiv(H, NMS_T));
  switch (nms_radius) {
    case 0: sp_nms_kernel<0><<<grid, NMS_THREADS, smem, st>>>(scores, H, W, thr, border, cand_keys, cand_cap, cand_count, nms_out); break;
    case 1: sp_nms_kernel<1><<<grid, NMS_THREADS, smem, st>>>(scores, H, W, thr, border, cand_keys, cand_cap, cand_count, nms_out); break;
    case 2: sp_nms_kernel<2><<<grid, NMS_THREADS, smem, st>>>(scores, H, W, thr, border, cand_keys, cand_cap, cand_count, nms_out); break;
    case 3: sp_nms_kernel<3><<<grid, NMS_THREADS, smem, st>>>(scores, H, W, thr, border, cand_keys, cand_cap, cand_count, nms_out); break;
    default: sp_nms_kernel<4><<<grid, NMS_THREADS, smem, st>>>(scores, H, W, thr, border, cand_keys, cand_cap, cand_count, nms_out); break;
  }
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_sp_select_topk(const unsigned long long* cand_keys, const int* cand_count, int cand_cap, int k,
                                  int W, float* kpts, float* scores, int out_cap, int* n_out,
                                  unsigned long long* spill, void* stream) {
  I4D_CHECK_ARG(cand_keys && cand_count && kpts && scores && n_out && spill, "null pointer");
  I4D_CHECK_ARG(out_cap > 0 && cand_cap > 0 && W > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    I4D_CUDA_CALL(cudaFuncSetAttribute(sp_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SEL_MAX_SMEM_KEYS * 8));
  }
  if (k >= 1 && k <= SEL_MAX_SMEM_KEYS) {
    // multi-CTA radix select; scratch: `spill` holds [TopkState | selected keys (k) | ties (rest)]
    static bool attr2_seen[64] = {};
    if (i4d_first_use_on_device(attr2_seen)) {
      I4D_CUDA_CALL(cudaFuncSetAttribute(topk_rank_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_MAX_SMEM_KEYS * 8));
    }
    const size_t st_words = (sizeof(TopkState) + 7) / 8;
    // spill layout (2 * cand_cap + 4096 keys, see the header): [TopkState (< 4096 words) | selected keys (k <= 16384 <= cand_cap) | ties (cand_cap)]
    static_assert(sizeof(TopkState) <= 4096 * 8, "TopkState must fit the reserved head of the spill buffer");
    I4D_CHECK_ARG(cand_cap >= SEL_MAX_SMEM_KEYS, "cand_cap must be at least 16384");
    TopkState* tks = reinterpret_cast<TopkState*>(spill);
    unsigned long long* sel = spill + 4096;
    unsigned long long* ties = sel + cand_cap;
    (void)st_words;
    const int G = 2 * i4d_num_sms();
    topk_init_kernel<<<1, 1024, 0, st>>>(cand_count, cand_cap, k, out_cap, tks, n_out);
    for (int pass = 0; pass < 3; ++pass) {
      topk_hist_kernel<<<G, 256, 0, st>>>(cand_keys, tks, pass);
      topk_scan_kernel<<<1, 1024, 0, st>>>(tks, pass);
    }
    topk_compact_kernel<<<G, 256, 0, st>>>(cand_keys, tks, sel, ties);
    topk_finish_kernel<<<1, SEL_THREADS, 0, st>>>(cand_keys, tks, sel, ties);
    const int m_max = k < out_cap ? k : out_cap;                     // the result holds at most min(k, out_cap) keypoints
    topk_rank_emit_kernel<<<i4d_cdiv(m_max, RANK_THREADS / 8), RANK_THREADS, (size_t)m_max * 8, st>>>(sel, tks, W, kpts, scores);
    I4D_CUDA_LAUNCH_CHECK();
    return I4D_OK;
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
