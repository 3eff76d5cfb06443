// Assignment kernels on sm_100a: log-domain Sinkhorn (SuperGlue), dual-softmax (LightGlue) and mutual-NN.
// All are HBM-bound streaming reductions over an M x N f32 score matrix that does not fit L2 at the target
// sizes (8192^2 * 4 B = 268 MB, 16384^2 * 4 B = 1.07 GB).
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d/thirdparty):
//   SuperGlue/models/superglue.py:152-186  log_sinkhorn_iterations / log_optimal_transport
//   SuperGlue/models/superglue.py:288-298  mutual-NN + threshold           LightGlue/lightglue/lightglue.py:290-306
//   LightGlue/lightglue/lightglue.py:253-266 sigmoid_log_double_softmax
//
// The (M+1) x (N+1) coupling matrix is never materialised: its dustbin row/column are the constant
// `bin_score`, so their contribution to every log-sum-exp is added analytically.
//
// Building blocks (S is M x N row-major with row pitch ld >= N floats; 128-bit loads whenever ld % 4 == 0):
//   row_reduce : out_i = LSE_j / max_j ( scale * S_ij + coloff_j )      one warp per row, float4 streaming loads
//   col_reduce : out_j = LSE_i / max_i ( scale * S_ij + rowoff_i )      column strips, partials + combine
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

#define LOG2E 1.4426950408889634f
#define LN2 0.6931471805599453f

struct LseAcc {  // running (max, sum of exp(x - max))
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add4(float a, float b, float c, float d) {
    float cm = fmaxf(fmaxf(a, b), fmaxf(c, d));
    if (cm > m) { s *= exp2f((m - cm) * LOG2E); m = cm; }
    s += exp2f((a - m) * LOG2E) + exp2f((b - m) * LOG2E) + exp2f((c - m) * LOG2E) + exp2f((d - m) * LOG2E);
  }
  __device__ __forceinline__ void add1(float a) {
    if (a > m) { s *= exp2f((m - a) * LOG2E); m = a; }
    s += exp2f((a - m) * LOG2E);
  }
  __device__ __forceinline__ void merge(float om, float os) {
    float nm = fmaxf(m, om);
    if (nm == -INFINITY) return;
    s = s * exp2f((m - nm) * LOG2E) + os * exp2f((om - nm) * LOG2E);
    m = nm;
  }
  __device__ __forceinline__ float value() const { return m + logf(s); }
};

struct ArgAcc {  // running (max, first index of max)
  float m; int i;
  __device__ __forceinline__ void init() { m = -INFINITY; i = 0x7fffffff; }
  __device__ __forceinline__ void add(float a, int idx) { if (a > m || (a == m && idx < i)) { m = a; i = idx; } }
};

// ------------------------------------------------------------------------------------------------
// row pass.  ROWS_PER_CTA warps per CTA, one warp per row.  mode 0: LSE -> out_val[i] = base_i - LSE
//            (Sinkhorn: u_i = log_mu_i - LSE_j(S_ij + v_j), with the dustbin column folded in)
//            mode 1: plain LSE -> out_val[i] ; mode 2: argmax -> out_val[i] = max, out_idx[i]
// ------------------------------------------------------------------------------------------------
#define ROW_WARPS 8

template <int MODE>
__global__ void __launch_bounds__(ROW_WARPS * 32) row_reduce_kernel(const float* __restrict__ S, int M, int N, int ld,
                                                                    float scale, const float* __restrict__ coloff,
                                                                    const float* __restrict__ extra_ptr, float extra_add, float base,
                                                                    float* __restrict__ out_val,
                                                                    int* __restrict__ out_idx, const int* __restrict__ run_if = nullptr) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M || (run_if && *run_if == 0)) return;                  // run_if: a device flag that switches the whole pass on
  const float* r = S + (size_t)row * ld;
  const int Nv = (ld & 3) ? 0 : (N & ~3);        // columns covered by 128-bit loads; the rest (tail, or everything) is scalar
  if (MODE == 2) {
    ArgAcc a; a.init();
    for (int j = lane * 4; j < Nv; j += 128) {
      float4 x = ldg_stream(reinterpret_cast<const float4*>(r + j));
      float4 o = coloff ? __ldg(reinterpret_cast<const float4*>(coloff + j)) : make_float4(0, 0, 0, 0);
      a.add(fmaf(scale, x.x, o.x), j); a.add(fmaf(scale, x.y, o.y), j + 1);
      a.add(fmaf(scale, x.z, o.z), j + 2); a.add(fmaf(scale, x.w, o.w), j + 3);
    }
    for (int j = Nv + lane; j < N; j += 32) a.add(fmaf(scale, r[j], coloff ? coloff[j] : 0.f), j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, a.m, o);
      int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
      a.add(om, oi);
    }
    if (lane == 0) { out_val[row] = a.m; out_idx[row] = a.i; }
  } else {
    LseAcc a; a.init();
    for (int j = lane * 4; j < Nv; j += 128) {
      float4 x = ldg_stream(reinterpret_cast<const float4*>(r + j));
      float4 o = coloff ? __ldg(reinterpret_cast<const float4*>(coloff + j)) : make_float4(0, 0, 0, 0);
      a.add4(fmaf(scale, x.x, o.x), fmaf(scale, x.y, o.y), fmaf(scale, x.z, o.z), fmaf(scale, x.w, o.w));
    }
    for (int j = Nv + lane; j < N; j += 32) a.add1(fmaf(scale, r[j], coloff ? coloff[j] : 0.f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, a.m, o);
      float os = __shfl_xor_sync(0xffffffffu, a.s, o);
      a.merge(om, os);
    }
    if (lane == 0) {
      if (extra_ptr) a.add1(__ldg(extra_ptr) + extra_add);
      float l = a.value();
      out_val[row] = (MODE == 0) ? base - l : l;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column pass.  CTA = 128 threads x 4 columns each (512-column strip, float4 loads), a slab of rows per
// blockIdx.y.  Partials (max,sum) or (max,idx) go to a workspace [splits][N]; a combine kernel finishes.
// ------------------------------------------------------------------------------------------------
#define COL_THREADS 128

template <int MODE>  // 0/1: LSE partials, 2: argmax partials
__global__ void __launch_bounds__(COL_THREADS) col_partial_kernel(const float* __restrict__ S, int M, int N, int ld,
                                                                  float scale, const float* __restrict__ rowoff,
                                                                  int rows_per_split, float* __restrict__ pm,
                                                                  float* __restrict__ ps, int* __restrict__ pi,
                                                                  const int* __restrict__ run_if = nullptr) {
  if (run_if && *run_if == 0) return;
  const int j0 = (blockIdx.x * COL_THREADS + threadIdx.x) * 4;
  const int i0 = blockIdx.y * rows_per_split, i1 = min(M, i0 + rows_per_split);
  if (j0 >= N) return;
  const bool vec = ((ld & 3) == 0);           // 128-bit loads stay inside the row pitch; components >= nv are ignored
  const int nv = min(4, N - j0);
  if (MODE == 2) {
    ArgAcc a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c].init();
    for (int i = i0; i < i1; ++i) {
      float o = rowoff ? __ldg(rowoff + i) : 0.f;
      const float* p = S + (size_t)i * ld + j0;
      if (vec) {
        float4 x = ldg_stream(reinterpret_cast<const float4*>(p));
        a[0].add(fmaf(scale, x.x, o), i); a[1].add(fmaf(scale, x.y, o), i);
        a[2].add(fmaf(scale, x.z, o), i); a[3].add(fmaf(scale, x.w, o), i);      // columns >= nv: computed, never stored
      } else {
        for (int c = 0; c < nv; ++c) a[c].add(fmaf(scale, p[c], o), i);
      }
    }
    for (int c = 0; c < nv; ++c) {
      pm[(size_t)blockIdx.y * N + j0 + c] = a[c].m;
      pi[(size_t)blockIdx.y * N + j0 + c] = a[c].i;
    }
  } else {
    LseAcc a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c].init();
    int i = i0;
    if (vec) {
      // 4 rows at a time: 4 independent 128-bit loads in flight per thread, one rescale per column per 4 rows
      for (; i + 4 <= i1; i += 4) {
        float4 x[4]; float o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          x[r] = ldg_stream(reinterpret_cast<const float4*>(S + (size_t)(i + r) * ld + j0));
          o[r] = rowoff ? __ldg(rowoff + i + r) : 0.f;
        }
        a[0].add4(fmaf(scale, x[0].x, o[0]), fmaf(scale, x[1].x, o[1]), fmaf(scale, x[2].x, o[2]), fmaf(scale, x[3].x, o[3]));
        a[1].add4(fmaf(scale, x[0].y, o[0]), fmaf(scale, x[1].y, o[1]), fmaf(scale, x[2].y, o[2]), fmaf(scale, x[3].y, o[3]));
        a[2].add4(fmaf(scale, x[0].z, o[0]), fmaf(scale, x[1].z, o[1]), fmaf(scale, x[2].z, o[2]), fmaf(scale, x[3].z, o[3]));
        a[3].add4(fmaf(scale, x[0].w, o[0]), fmaf(scale, x[1].w, o[1]), fmaf(scale, x[2].w, o[2]), fmaf(scale, x[3].w, o[3]));
      }
    }
    for (; i < i1; ++i) {
      float o = rowoff ? __ldg(rowoff + i) : 0.f;
      const float* p = S + (size_t)i * ld + j0;
      for (int c = 0; c < nv; ++c) a[c].add1(fmaf(scale, p[c], o));
    }
    for (int c = 0; c < nv; ++c) {
      pm[(size_t)blockIdx.y * N + j0 + c] = a[c].m;
      ps[(size_t)blockIdx.y * N + j0 + c] = a[c].s;
    }
  }
}

template <int MODE>
__global__ void col_combine_kernel(const float* __restrict__ pm, const float* __restrict__ ps,
                                   const int* __restrict__ pi, int splits, int N, const float* __restrict__ extra_ptr,
                                   float extra_add, float base, float* __restrict__ out_val, int* __restrict__ out_idx,
                                   const int* __restrict__ run_if = nullptr) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N || (run_if && *run_if == 0)) return;
  if (MODE == 2) {
    ArgAcc a; a.init();
    for (int s = 0; s < splits; ++s) a.add(pm[(size_t)s * N + j], pi[(size_t)s * N + j]);
    out_val[j] = a.m; out_idx[j] = a.i;
  } else {
    LseAcc a; a.init();
    for (int s = 0; s < splits; ++s) a.merge(pm[(size_t)s * N + j], ps[(size_t)s * N + j]);
    if (extra_ptr) a.add1(__ldg(extra_ptr) + extra_add);
    float l = a.value();
    out_val[j] = (MODE == 0) ? base - l : l;
  }
}

// ------------------------------------------------------------------------------------------------
// Row AND column arg-max in ONE read of S (the mutual-nearest-neighbour step of both matchers needs both; two separate passes
// stream the matrix twice: 2 x 1.07 GB at 16384^2).  Same tiling as the column pass (128 threads x 4 columns, a slab of rows per
// blockIdx.y, 128-bit loads); column partials as there.  The row side costs two warp reductions per row: every warp owns a
// 128-column strip, its lanes' four candidates are reduced with redux.sync (max over an order-preserving integer key, then min
// over the indices that hold the maximum: first index on ties, like torch.max), and lane 0 writes the strip's (max, index) to
// a [strips][M] workspace that row_strip_combine_kernel folds.  Arithmetic identical to row_reduce_kernel<2> /
// col_partial_kernel<2> (one fmaf per element and side), so the results are bit-identical to the two-pass path.
// ------------------------------------------------------------------------------------------------
#ifndef RC_ROWS
#define RC_ROWS 4                 // rows (= independent 128-bit loads) in flight per thread in the one-read passes
#endif
__device__ __forceinline__ unsigned f32_order_key(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_order_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(COL_THREADS) rowcol_argmax_kernel(const float* __restrict__ S, int M, int N, int ld, float scale,
                                                                    const float* __restrict__ coloff,   // added on the ROW side
                                                                    const float* __restrict__ rowoff,   // added on the COLUMN side
                                                                    int rows_per_split, float* __restrict__ pm, int* __restrict__ pi,
                                                                    float* __restrict__ rm, int* __restrict__ ri) {
  const int j0 = (blockIdx.x * COL_THREADS + threadIdx.x) * 4;
  const int strip = blockIdx.x * (COL_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (strip * 128 >= N) return;                                      // whole warp beyond the matrix
  const int i0 = blockIdx.y * rows_per_split, i1 = min(M, i0 + rows_per_split);
  const int nv = max(0, min(4, N - j0));                             // my valid columns (0 for the lanes past a ragged edge)
  const int jl = min(j0, ld - 4);                                    // a load address inside the row pitch for those lanes
  float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
  if (coloff && nv == 4) co = __ldg(reinterpret_cast<const float4*>(coloff + j0));
  else if (coloff) { if (nv > 0) co.x = coloff[j0]; if (nv > 1) co.y = coloff[j0 + 1]; if (nv > 2) co.z = coloff[j0 + 2]; }
  float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int ci[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
  auto one_row = [&](int i, const float4 x, const float ro) {
    // column side: rows arrive in increasing order, so a strict comparison keeps the first maximum
    const float c0 = fmaf(scale, x.x, ro), c1 = fmaf(scale, x.y, ro), c2 = fmaf(scale, x.z, ro), c3 = fmaf(scale, x.w, ro);
    if (c0 > cm[0]) { cm[0] = c0; ci[0] = i; }
    if (c1 > cm[1]) { cm[1] = c1; ci[1] = i; }
    if (c2 > cm[2]) { cm[2] = c2; ci[2] = i; }
    if (c3 > cm[3]) { cm[3] = c3; ci[3] = i; }
    // row side: my best of four (first index on ties), then the strip's
    float b = nv > 0 ? fmaf(scale, x.x, co.x) : -INFINITY;
    int bj = j0;
    const float r1 = nv > 1 ? fmaf(scale, x.y, co.y) : -INFINITY, r2 = nv > 2 ? fmaf(scale, x.z, co.z) : -INFINITY,
                r3 = nv > 3 ? fmaf(scale, x.w, co.w) : -INFINITY;
    if (r1 > b) { b = r1; bj = j0 + 1; }
    if (r2 > b) { b = r2; bj = j0 + 2; }
    if (r3 > b) { b = r3; bj = j0 + 3; }
    const unsigned key = f32_order_key(b);
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    const unsigned jmin = __reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)bj : 0x7fffffffu);
    if (lane == 0) { rm[(size_t)strip * M + i] = f32_from_order_key(kmax); ri[(size_t)strip * M + i] = (int)jmin; }
  };
  int i = i0;
  for (; i + RC_ROWS <= i1; i += RC_ROWS) {                          // RC_ROWS independent 128-bit loads in flight per thread
    float4 x[RC_ROWS]; float ro[RC_ROWS];
#pragma unroll
    for (int r = 0; r < RC_ROWS; ++r) {
      x[r] = ldg_stream(reinterpret_cast<const float4*>(S + (size_t)(i + r) * ld + jl));
      ro[r] = rowoff ? __ldg(rowoff + i + r) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < RC_ROWS; ++r) one_row(i + r, x[r], ro[r]);
  }
  for (; i < i1; ++i)
    one_row(i, ldg_stream(reinterpret_cast<const float4*>(S + (size_t)i * ld + jl)), rowoff ? __ldg(rowoff + i) : 0.f);
  for (int c = 0; c < nv; ++c) {
    pm[(size_t)blockIdx.y * N + j0 + c] = cm[c];
    pi[(size_t)blockIdx.y * N + j0 + c] = ci[c];
  }
}

__global__ void row_strip_combine_kernel(const float* __restrict__ rm, const int* __restrict__ ri, int strips, int M,
                                         float* __restrict__ out_val, int* __restrict__ out_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  ArgAcc a; a.init();
  for (int s = 0; s < strips; ++s) a.add(rm[(size_t)s * M + i], ri[(size_t)s * M + i]);
  out_val[i] = a.m; out_idx[i] = a.i;
}

// ------------------------------------------------------------------------------------------------
// Row AND column log-sum-exp of S in ONE read (first half of LightGlue's double softmax).  Tiling as above.  Per row, a warp
// takes the maximum w_i of its 128-column strip (redux.sync) and e_ij = exp(x_ij - w_i): ONE exponential per element serves
// both sides.  Row side: the strip's (w_i, sum_j e_ij) goes to the [strips][M] workspace.  Column side: the thread's four columns
// accumulate sum_i e_ij * exp(w_i - C) against a warp-uniform reference C = the largest w_i seen so far (one more exponential
// per row and lane; a growing C rescales the four sums), and the slab's (C, sum) pairs are merged like the exact partials.
// e * exp(w_i - C) <= 1, so nothing overflows; a column whose entries all lie > 69 nats below its strip's maximum would
// underflow — then `flag` is raised and the exact two-pass kernels (launched behind it with run_if = flag) redo both sides.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COL_THREADS) rowcol_lse_kernel(const float* __restrict__ S, int M, int N, int ld,
                                                                 int rows_per_split, float* __restrict__ pm, float* __restrict__ ps,
                                                                 float* __restrict__ rm, float* __restrict__ rs, int* __restrict__ flag) {
  const int j0 = (blockIdx.x * COL_THREADS + threadIdx.x) * 4;
  const int strip = blockIdx.x * (COL_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (strip * 128 >= N) return;
  const int i0 = blockIdx.y * rows_per_split, i1 = min(M, i0 + rows_per_split);
  const int nv = max(0, min(4, N - j0));
  const int jl = min(j0, ld - 4);
  float ca[4] = {0.f, 0.f, 0.f, 0.f};
  float C = -INFINITY;                                               // warp-uniform
  auto one_row = [&](int i, float4 x) {
    if (nv < 4) { if (nv < 1) x.x = -INFINITY; if (nv < 2) x.y = -INFINITY; if (nv < 3) x.z = -INFINITY; x.w = -INFINITY; }
    const float b = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
    const float w = f32_from_order_key(__reduce_max_sync(0xffffffffu, f32_order_key(b)));
    const float nw = (w == -INFINITY) ? 0.f : -w * LOG2E;             // a strip of -inf only: every term is exp2(-inf) = 0
    const float e0 = exp2f(fmaf(x.x, LOG2E, nw)), e1 = exp2f(fmaf(x.y, LOG2E, nw)), e2 = exp2f(fmaf(x.z, LOG2E, nw)),
                e3 = exp2f(fmaf(x.w, LOG2E, nw));
    float sum = (e0 + e1) + (e2 + e3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) { rm[(size_t)strip * M + i] = w; rs[(size_t)strip * M + i] = sum; }
    float f;
    if (w > C) {                                                      // uniform over the warp
      const float g = exp2f((C - w) * LOG2E);                         // 0 the first time (C = -inf)
      ca[0] *= g; ca[1] *= g; ca[2] *= g; ca[3] *= g;
      C = w; f = 1.f;
    } else {
      f = (w == -INFINITY) ? 0.f : exp2f((w - C) * LOG2E);
    }
    ca[0] = fmaf(e0, f, ca[0]); ca[1] = fmaf(e1, f, ca[1]); ca[2] = fmaf(e2, f, ca[2]); ca[3] = fmaf(e3, f, ca[3]);
  };
  int i = i0;
  for (; i + RC_ROWS <= i1; i += RC_ROWS) {
    float4 x[RC_ROWS];
#pragma unroll
    for (int r = 0; r < RC_ROWS; ++r) x[r] = ldg_stream(reinterpret_cast<const float4*>(S + (size_t)(i + r) * ld + jl));
#pragma unroll
    for (int r = 0; r < RC_ROWS; ++r) one_row(i + r, x[r]);
  }
  for (; i < i1; ++i) one_row(i, ldg_stream(reinterpret_cast<const float4*>(S + (size_t)i * ld + jl)));
  bool bad = false;
  for (int c = 0; c < nv; ++c) {
    pm[(size_t)blockIdx.y * N + j0 + c] = C;
    ps[(size_t)blockIdx.y * N + j0 + c] = ca[c];
    bad |= (i1 > i0) && C > -INFINITY && !(ca[c] > 1e-30f);          // underflow (or NaN): this slab's share of column j is lost
  }
  if (bad) atomicOr(flag, 1);
}

__global__ void row_strip_lse_combine_kernel(const float* __restrict__ rm, const float* __restrict__ rs, int strips, int M,
                                             float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  LseAcc a; a.init();
  for (int s = 0; s < strips; ++s) a.merge(rm[(size_t)s * M + i], rs[(size_t)s * M + i]);
  out[i] = a.value();
}

// out = base - LSE_i(add + vec[i]), i in [0, n)   (single CTA; dustbin row/column of the coupling matrix)
__global__ void __launch_bounds__(1024) vec_lse_kernel(const float* __restrict__ vec, int n, float add, float base,
                                                       float* __restrict__ out) {
  __shared__ float sm[32], ss[32];
  LseAcc a; a.init();
  for (int i = threadIdx.x; i < n; i += blockDim.x) a.add1(vec[i] + add);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, a.m, o);
    float os = __shfl_xor_sync(0xffffffffu, a.s, o);
    a.merge(om, os);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = a.m; ss[threadIdx.x >> 5] = a.s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    LseAcc t; t.init();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t.merge(sm[w], ss[w]);
    *out = base - t.value();
  }
}

static int col_splits_for(int M, int N) {
  int strips = i4d_cdiv(N, COL_THREADS * 4);
  static int per_sm = 0;
  if (!per_sm) { const char* e = getenv("I4D_COL_CTAS_PER_SM"); per_sm = e ? atoi(e) : 6; if (per_sm < 1) per_sm = 6; }
  int target = per_sm * i4d_num_sms();  // several 128-thread CTAs resident per SM keep enough 128-bit loads in flight
  int splits = i4d_cdiv(target, strips);
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  if (splits > M) splits = M;
  return splits;
}

// workspace layout (floats): pm[256*N] ps[256*N] pi[256*N] | val0[M] idx0[M] idx1[N] val1[N] rowoff[M+1] coloff[N+1] rl[M] cl[N]
//                           | rm[strips*M] ri[strips*M] flag   (strips = ceil(N / 128): row partials of the one-read passes)
extern "C" __attribute__((visibility("default"))) size_t i4d_assignment_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return ((size_t)256 * N * 3 + 4 * (size_t)(M + 8) + 4 * (size_t)(N + 8) + 64 + 2 * (size_t)i4d_cdiv(N, 128) * M + 16) * sizeof(float);
}

struct AssignWs {
  float *pm, *ps; int* pi;
  float* val0; int* idx0; int* idx1; float* val1; float* rowoff; float* coloff; float* rl; float* cl;
  float* rm; int* ri; int* flag;
  AssignWs(void* w, int M, int N) {
    float* f = reinterpret_cast<float*>(w);
    pm = f; ps = f + (size_t)256 * N; pi = reinterpret_cast<int*>(f + (size_t)512 * N);
    float* q = f + (size_t)768 * N;
    const size_t ms = ((size_t)M + 1 + 3) & ~(size_t)3, ns = ((size_t)N + 1 + 3) & ~(size_t)3;  // keep 16 B alignment (float4 loads)
    val0 = q; q += ms; idx0 = reinterpret_cast<int*>(q); q += ms; rowoff = q; q += ms; rl = q; q += ms;
    idx1 = reinterpret_cast<int*>(q); q += ns; val1 = q; q += ns; coloff = q; q += ns; cl = q; q += ns;
    rm = q; ri = reinterpret_cast<int*>(q + (size_t)i4d_cdiv(N, 128) * M);
    flag = ri + (size_t)i4d_cdiv(N, 128) * M;
  }
};

template <int MODE>
static void launch_row(const float* S, int M, int N, int ld, float scale, const float* coloff, const float* extra_ptr,
                       float extra_add, float base, float* out_val, int* out_idx, cudaStream_t st, const int* run_if = nullptr) {
  row_reduce_kernel<MODE><<<i4d_cdiv(M, ROW_WARPS), ROW_WARPS * 32, 0, st>>>(S, M, N, ld, scale, coloff, extra_ptr,
                                                                             extra_add, base, out_val, out_idx, run_if);
}
template <int MODE>
static void launch_col(const float* S, int M, int N, int ld, float scale, const float* rowoff, const float* extra_ptr,
                       float extra_add, float base, float* out_val, int* out_idx, AssignWs& w, cudaStream_t st,
                       const int* run_if = nullptr) {
  int splits = col_splits_for(M, N);
  int rps = i4d_cdiv(M, splits);
  rps = (rps + 3) & ~3;
  splits = i4d_cdiv(M, rps);
  dim3 grid(i4d_cdiv(N, COL_THREADS * 4), splits);
  col_partial_kernel<MODE><<<grid, COL_THREADS, 0, st>>>(S, M, N, ld, scale, rowoff, rps, w.pm, w.ps, w.pi, run_if);
  col_combine_kernel<MODE><<<i4d_cdiv(N, 256), 256, 0, st>>>(w.pm, w.ps, w.pi, splits, N, extra_ptr, extra_add, base,
                                                             out_val, out_idx, run_if);
}

// row arg-max of (scale S + coloff_j) -> (val0, idx0) and column arg-max of (scale S + rowoff_i) -> (val1, idx1), one read of S
static void launch_rowcol_argmax(const float* S, int M, int N, int ld, float scale, const float* coloff, const float* rowoff,
                                 float* val0, int* idx0, float* val1, int* idx1, AssignWs& w, cudaStream_t st) {
  if ((ld & 3) || ld < 4 || (reinterpret_cast<uintptr_t>(S) & 15)) {  // no 128-bit loads: the two separate passes
    launch_row<2>(S, M, N, ld, scale, coloff, nullptr, 0.f, 0.f, val0, idx0, st);
    launch_col<2>(S, M, N, ld, scale, rowoff, nullptr, 0.f, 0.f, val1, idx1, w, st);
    return;
  }
  int splits = col_splits_for(M, N);
  int rps = i4d_cdiv(M, splits);
  rps = (rps + 3) & ~3;
  splits = i4d_cdiv(M, rps);
  dim3 grid(i4d_cdiv(N, COL_THREADS * 4), splits);
  rowcol_argmax_kernel<<<grid, COL_THREADS, 0, st>>>(S, M, N, ld, scale, coloff, rowoff, rps, w.pm, w.pi, w.rm, w.ri);
  col_combine_kernel<2><<<i4d_cdiv(N, 256), 256, 0, st>>>(w.pm, w.ps, w.pi, splits, N, nullptr, 0.f, 0.f, val1, idx1);
  row_strip_combine_kernel<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.rm, w.ri, i4d_cdiv(N, 128), M, val0, idx0);
}

// rl_i = LSE_j S_ij and cl_j = LSE_i S_ij, one read of S (exact two-pass kernels behind a device flag, see rowcol_lse_kernel)
static void launch_rowcol_lse(const float* S, int M, int N, int ld, float* rl, float* cl, AssignWs& w, cudaStream_t st) {
  if ((ld & 3) || ld < 4 || (reinterpret_cast<uintptr_t>(S) & 15)) {
    launch_row<1>(S, M, N, ld, 1.f, nullptr, nullptr, 0.f, 0.f, rl, nullptr, st);
    launch_col<1>(S, M, N, ld, 1.f, nullptr, nullptr, 0.f, 0.f, cl, nullptr, w, st);
    return;
  }
  int splits = col_splits_for(M, N);
  int rps = i4d_cdiv(M, splits);
  rps = (rps + 3) & ~3;
  splits = i4d_cdiv(M, rps);
  dim3 grid(i4d_cdiv(N, COL_THREADS * 4), splits);
  cudaMemsetAsync(w.flag, 0, sizeof(int), st);
  rowcol_lse_kernel<<<grid, COL_THREADS, 0, st>>>(S, M, N, ld, rps, w.pm, w.ps, w.rm, reinterpret_cast<float*>(w.ri), w.flag);
  col_combine_kernel<1><<<i4d_cdiv(N, 256), 256, 0, st>>>(w.pm, w.ps, w.pi, splits, N, nullptr, 0.f, 0.f, cl, nullptr);
  row_strip_lse_combine_kernel<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.rm, reinterpret_cast<float*>(w.ri), i4d_cdiv(N, 128), M, rl);
  launch_row<1>(S, M, N, ld, 1.f, nullptr, nullptr, 0.f, 0.f, rl, nullptr, st, w.flag);
  launch_col<1>(S, M, N, ld, 1.f, nullptr, nullptr, 0.f, 0.f, cl, nullptr, w, st, w.flag);
}

static int check_ws(const char* fn, int M, int N, size_t have) {
  size_t need = i4d_assignment_workspace_bytes(M, N);
  if (have < need) {
    i4d_set_error("%s: workspace too small (%zu < %zu bytes)", fn, have, need);
    return I4D_ERR_WORKSPACE;
  }
  return 0;
}

// ---- primitives exposed for tests / roofline measurements -------------------------------------------
extern "C" __attribute__((visibility("default"))) int i4d_row_lse(const float* S, int M, int N, float scale, const float* coloff, float* out, void* stream) {
  I4D_CHECK_ARG(S && out && M > 0 && N > 0, "null pointer or empty matrix");
  launch_row<1>(S, M, N, N, scale, coloff, nullptr, 0.f, 0.f, out, nullptr, (cudaStream_t)stream);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
extern "C" __attribute__((visibility("default"))) int i4d_col_lse(const float* S, int M, int N, float scale, const float* rowoff, float* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(S && out && workspace && M > 0 && N > 0, "null pointer or empty matrix");
  if (int rc = check_ws("i4d_col_lse", M, N, workspace_bytes)) return rc;
  AssignWs w(workspace, M, N);
  launch_col<1>(S, M, N, N, scale, rowoff, nullptr, 0.f, 0.f, out, nullptr, w, (cudaStream_t)stream);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- Sinkhorn potentials ------------------------------------------------------------------------------
static int sinkhorn_fused_launch(float* S, int M, int N, int ld, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st);
static int g_sinkhorn_mode = 0;   // 0 = fused when the shape allows, 1 = always the two-pass kernels (tests / comparison)

static void sinkhorn_iterations(float* S, int M, int N, int ld, float alpha, int iters, float* u, float* v, AssignWs& w,
                                cudaStream_t st) {
  if (g_sinkhorn_mode == 0 && sinkhorn_fused_launch(S, M, N, ld, alpha, iters, u, v, w, st) == 0) return;
  const float norm = -logf((float)M + (float)N);
  const float log_mu_last = logf((float)N) + norm, log_nu_last = logf((float)M) + norm;
  cudaMemsetAsync(u, 0, (size_t)(M + 1) * sizeof(float), st);
  cudaMemsetAsync(v, 0, (size_t)(N + 1) * sizeof(float), st);
  for (int it = 0; it < iters; ++it) {
    // u_i = log_mu_i - LSE_j(Z_ij + v_j): rows i < M stream S; the dustbin row is a vector LSE
    launch_row<0>(S, M, N, ld, 1.f, v, v + N, alpha, norm, u, nullptr, st);
    vec_lse_kernel<<<1, 1024, 0, st>>>(v, N + 1, alpha, log_mu_last, u + M);
    // v_j = log_nu_j - LSE_i(Z_ij + u_i)
    launch_col<0>(S, M, N, ld, 1.f, u, u + M, alpha, norm, v, nullptr, w, st);
    vec_lse_kernel<<<1, 1024, 0, st>>>(u, M + 1, alpha, log_nu_last, v + N);
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_sinkhorn(float* scores, int M, int N, int ld, float bin_score, int iters, float* u, float* v,
                            void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(scores && u && v && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && ld >= N && iters >= 0, "bad sizes");
  if (int rc = check_ws("i4d_sinkhorn", M, N, workspace_bytes)) return rc;
  AssignWs w(workspace, M, N);
  sinkhorn_iterations(scores, M, N, ld, bin_score, iters, u, v, w, (cudaStream_t)stream);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- mutual nearest neighbour + threshold (superglue.py:288-298 / lightglue.py:290-306) -----------------
__global__ void mutual_kernel0(const float* __restrict__ val0, const int* __restrict__ idx0,
                               const float* __restrict__ rowadd, float add_const, const int* __restrict__ idx1, int M,
                               float thr, int* __restrict__ m0, float* __restrict__ ms0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  int j = idx0[i];
  bool mutual = idx1[j] == i;
  float sc = mutual ? expf(val0[i] + rowadd[i] + add_const) : 0.f;
  ms0[i] = sc;
  m0[i] = (mutual && sc > thr) ? j : -1;
}
__global__ void mutual_kernel1(const int* __restrict__ idx0, const int* __restrict__ idx1,
                               const int* __restrict__ m0, const float* __restrict__ ms0, int N,
                               int* __restrict__ m1, float* __restrict__ ms1) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  int i = idx1[j];
  bool mutual = idx0[i] == j;
  ms1[j] = mutual ? ms0[i] : 0.f;
  m1[j] = (mutual && m0[i] >= 0) ? i : -1;
}

extern "C" __attribute__((visibility("default"))) int i4d_sg_assign(float* scores, int M, int N, int ld, float bin_score, int iters, float match_threshold,
                             int* matches0, int* matches1, float* mscores0, float* mscores1, float* u, float* v,
                             void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(scores && matches0 && matches1 && mscores0 && mscores1 && u && v && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && ld >= N && iters >= 0, "bad sizes");
  if (int rc = check_ws("i4d_sg_assign", M, N, workspace_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  AssignWs w(workspace, M, N);
  sinkhorn_iterations(scores, M, N, ld, bin_score, iters, u, v, w, st);
  const float norm = -logf((float)M + (float)N);
  // P_ij = S_ij + u_i + v_j - norm over the core block: row argmax ignores u_i, column argmax ignores v_j
  launch_rowcol_argmax(scores, M, N, ld, 1.f, v, u, w.val0, w.idx0, w.val1, w.idx1, w, st);
  mutual_kernel0<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.val0, w.idx0, u, -norm, w.idx1, M, match_threshold, matches0,
                                                   mscores0);
  mutual_kernel1<<<i4d_cdiv(N, 256), 256, 0, st>>>(w.idx0, w.idx1, matches0, mscores0, N, matches1, mscores1);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- LightGlue dual-softmax assignment ---------------------------------------------------------------
// scores_ij = log_softmax_j(sim)_ij + log_softmax_i(sim)_ij + logsigmoid(z0_i) + logsigmoid(z1_j)
// pass 1: row LSE rl_i and column LSE cl_j in ONE read of sim (rowcol_lse_kernel).  pass 2: row argmax of
// (2 sim_ij + logsig(z1_j) - cl_j) and column argmax of (2 sim_ij + logsig(z0_i) - rl_i) in one more read (rowcol_argmax_kernel);
// the per-row constant (logsig(z0_i) - rl_i) is added afterwards.  The (M+1)x(N+1) matrix is never written.
__device__ __forceinline__ float logsigmoidf_(float z) { return fminf(z, 0.f) - log1pf(expf(-fabsf(z))); }
__global__ void lg_offsets_kernel(const float* __restrict__ z, const float* __restrict__ lse, int n,
                                  float* __restrict__ off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] = logsigmoidf_(z[i]) - lse[i];
}

extern "C" __attribute__((visibility("default"))) int i4d_lg_assign(const float* sim, int M, int N, int ld, const float* z0, const float* z1, float filter_threshold,
                             int* matches0, int* matches1, float* mscores0, float* mscores1, void* workspace,
                             size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(sim && z0 && z1 && matches0 && matches1 && mscores0 && mscores1 && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && ld >= N, "bad sizes");
  if (int rc = check_ws("i4d_lg_assign", M, N, workspace_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  AssignWs w(workspace, M, N);
  launch_rowcol_lse(sim, M, N, ld, w.rl, w.cl, w, st);
  lg_offsets_kernel<<<i4d_cdiv(M, 256), 256, 0, st>>>(z0, w.rl, M, w.rowoff);
  lg_offsets_kernel<<<i4d_cdiv(N, 256), 256, 0, st>>>(z1, w.cl, N, w.coloff);
  launch_rowcol_argmax(sim, M, N, ld, 2.f, w.coloff, w.rowoff, w.val0, w.idx0, w.val1, w.idx1, w, st);
  mutual_kernel0<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.val0, w.idx0, w.rowoff, 0.f, w.idx1, M, filter_threshold,
                                                   matches0, mscores0);
  mutual_kernel1<<<i4d_cdiv(N, 256), 256, 0, st>>>(w.idx0, w.idx1, matches0, mscores0, N, matches1, mscores1);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
// ====================================================================================================================
// Fused persistent Sinkhorn (N <= 8192, N % 4 == 0): ONE read of the score matrix per iteration instead of two, and part
// of that read served by L2.
//
// One CTA per SM (cooperative launch), each owning a contiguous band of rows.  Rows are streamed HBM -> shared memory
// with cp.async.bulk (1-D TMA) through a 3-stage mbarrier ring of 2 rows (<= 64 KB) per stage.  The band is walked
// top-down on even iterations and bottom-up on odd ones (boustrophedon): the rows touched last in iteration t are the
// first of iteration t+1, so the tail of every pass is still resident in the 126 MB L2 when it is needed again.
// 512 threads; each thread owns the same 16 columns for the whole solve:
//   phase A  (stage s)    e_ij = 2^(x_ij + v_j - m_i) from shared memory into registers, per-warp partial row sums ->
//                         shared memory, one mbarrier arrive per warp
//   phase B  (stage s-1)  after its own phase A of stage s a warp waits for the partials of stage s-1 (complete long before,
//                         unless a warp lags by a whole stage), finishes the row sums itself -> row factor a_i = kfac / sum,
//                         and updates its column accumulators += e_ij * a_i from the registers kept since stage s-1
//   duty                  one warp per stage (round-robin) additionally writes u_i, checks the sums and refills the
//                         shared-memory buffer the stage has released; nobody waits for it
// No block-wide barrier in the steady state: warps run up to a stage apart.
//
// Two arithmetic modes, same result up to f32 rounding:
//   exact : running (max, sum) per accumulator — used for the first two iterations (and always if the fast mode trips);
//   fast  : a-priori stabilisers instead of running maxima.  After a column update sum_i exp(S_ij + u_i + v_j) = nu_j, so
//           S_ij + v_j <= log nu_j - u_i: m_i = log(nu_max) - u_i(previous) bounds every term of row i from above; likewise
//           m_j = log(mu_max) - v_j(previous) for the columns.  The column pass needs no second exponential:
//           2^(x_ij + u_i - m_j) = e_ij * 2^(u_i + m_i - c_mu) — a rank-1 rescaling of the same kernel matrix.  Packed
//           f32x2 arithmetic (FFMA2/FADD2): 3.25 issue slots per matrix element.  A sum that underflows to 0 (potential
//           jump > ~80 nats between two iterations) or is not finite raises a flag: the whole solve restarts in exact
//           mode, on the device, without host involvement.
// Everything is kept in the log2 domain.  HBM traffic per iteration <= M*N*4 bytes (algorithmic count of SURVEY.md §8d: 2*M*N*4).
// ====================================================================================================================
#include <stdlib.h>
static int g_sinkhorn_fast = 1;   // 1 = a-priori stabilisers after the first two iterations, 0 = running maxima throughout
// Two instantiations of the fused kernel (sinkhorn_fused.inl): 256 threads x 32 columns for the straight-line path of wide matrices
// (N >= 7168: 41.4 us per iteration at 8192^2 against 43.3 with 512 threads), 512 threads x 16 columns below (37.2 against 41.6 us
// at 6000^2: the narrower rows leave the wider threads idle).
namespace sk512 {
#define SK_THREADS 512
#include "sinkhorn_fused.inl"
}  // namespace sk512
namespace sk256 {
#define SK_THREADS 256
#include "sinkhorn_fused.inl"
}  // namespace sk256
// returns 0 when the fused kernel ran, 1 when the shape is unsupported (caller falls back to the two-pass kernels)
static int sinkhorn_fused_launch(float* S, int M, int N, int ld, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st) {
  static int force = -1;                                            // I4D_SK_THREADS=256 / 512 overrides (experiments)
  if (force < 0) { const char* e = getenv("I4D_SK_THREADS"); force = e ? atoi(e) : 0; }
  const bool wide = force ? force == 256 : ((N + 3) & ~3) * 8 >= 8192 * 7;
  return wide ? sk256::sinkhorn_fused_launch(S, M, N, ld, alpha, iters, u, v, w, st)
              : sk512::sinkhorn_fused_launch(S, M, N, ld, alpha, iters, u, v, w, st);
}

// test / comparison hook: 0 = fused Sinkhorn when possible (default), 1 = always the two-pass kernels
extern "C" __attribute__((visibility("default"))) int i4d_set_sinkhorn_mode(int mode) {
  I4D_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0 (fused, fast stabilisers), 1 (two-pass kernels) or 2 (fused, running maxima)");
  g_sinkhorn_mode = mode == 1 ? 1 : 0;
  g_sinkhorn_fast = mode == 0 ? 1 : 0;
  return I4D_OK;
}
