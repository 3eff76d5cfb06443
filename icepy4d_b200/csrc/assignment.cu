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

#define SK_THREADS 512
#define SK_WARPS (SK_THREADS / 32)
#define SK_ROWS 2                 // rows per stage
#define SK_STAGES 3
#define SK_MAXN 8192
#define SK_GROUPS (SK_MAXN / 4 / SK_THREADS)   // float4 column groups per thread = 4
#define SK_EXACT_ITERS 1
#define SK_MAX_BAND 1024            // rows per CTA the previous-u staging buffer can hold
#define SK_KEEP_PCT_DEFAULT 0
// row pitch of the shared-memory stages for NP (= N rounded up to 4) columns
__host__ __device__ inline int sk_smem_pitch(int NP) { return (NP * 8 >= SK_MAXN * 7) ? SK_MAXN : NP; }

__device__ __forceinline__ float sk_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct L2Acc {   // running (max, sum) in the log2 domain, one exp per update
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void add(float x) {
    float d = x - m;
    float e = sk_ex2(-fabsf(d));
    s = (d > 0.f) ? fmaf(s, e, 1.f) : (s + e);
    m = fmaxf(m, x);
  }
  __device__ __forceinline__ void merge(float om, float os) {
    float nm = fmaxf(m, om);
    if (nm == -INFINITY) return;
    s = s * exp2f(m - nm) + os * exp2f(om - nm);
    m = nm;
  }
  __device__ __forceinline__ float lse_log2() const { return m + log2f(s); }
  __device__ __forceinline__ float lse_ln() const { return (m + log2f(s)) * LN2; }
};

__device__ __forceinline__ void sk_mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void sk_mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
  uint32_t ok = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void sk_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "l"(policy)
               : "memory");
}

struct SkCtx {
  const float* S; float* stage_buf; uint64_t* full; uint64_t* bpart;
  float (*part_m)[SK_ROWS][SK_WARPS]; float (*part_s)[SK_ROWS][SK_WARPS];
  float* u; const float* v; float* pm; float* ps; int* flag;
  float* unew_s;         // this iteration's u of the band, log2 domain (shared memory; the next iteration's uold_s)
  int M, N, NP, NS, ld, n4, row0, row1, nst, cta, keep_rows, pf, tid, warp, lane; uint32_t total;   // NP = 4 * n4: N rounded up to whole float4 groups; NS = row pitch of the smem stages
  float norm, c_mu, c_nu, extra_row, kfac;
  const float* uold_s;   // previous-iteration u of this CTA's band, log2 domain (shared memory)
  uint32_t row_bytes;
};
// ring / barrier phase bookkeeping carried across stages, bands and iterations (one copy per thread, all identical)
// Timeline instrumentation (-DSK_TRACE, scripts/sinkhorn_trace.py): CTA 0 stamps %clock64 at the synchronisation points of its
// fast-mode stages (warps 0 and 9) and of the iteration boundary (thread 0).
#ifdef SK_TRACE
#define SK_TRACE_MAX 512
__device__ unsigned long long sk_trace_buf[2][SK_TRACE_MAX][8];
#define SK_STAMP(fq_, k) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 288) && (fq_) < 440) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); sk_trace_buf[threadIdx.x ? 1 : 0][(fq_)][k] = c_; } } while (0)
#define SK_BSTAMP(it_, k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (it_) < 64) { unsigned long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); sk_trace_buf[1][448 + (it_)][k] = c_; } } while (0)
extern "C" __attribute__((visibility("default"))) int i4d_sinkhorn_trace_dump(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, sk_trace_buf, sizeof(sk_trace_buf));
}
#else
#define SK_STAMP(fq_, k) do { } while (0)
#define SK_BSTAMP(it_, k) do { } while (0)
#endif
struct SkRing {
  uint32_t seq;        // stages consumed so far (all iterations)
  uint32_t buf;        // seq % SK_STAGES
  uint32_t full_par;   // bit b: parity the next wait on full[b] expects
  uint32_t fq;         // fast-mode stages so far: partial-sum slot fq & 1, duty warp fq % SK_WARPS
  uint32_t bp_par;     // bit p: parity the next wait on bpart[p] expects
  uint32_t total;      // stages to run in total
  __device__ __forceinline__ void advance() { ++seq; buf = (buf == SK_STAGES - 1) ? 0u : buf + 1u; }
};

// Stages are numbered consecutively across iterations (S never changes, so the ring keeps streaming through the grid
// barriers).  Stage `seq` lives in buffer seq % SK_STAGES and covers row block sk_idx(seq) of the band: forward on even
// passes, backward on odd ones.
__device__ __forceinline__ int sk_idx(uint32_t seq, int nst) {
  const uint32_t pass = seq / (uint32_t)nst, r = seq - pass * (uint32_t)nst;
  return (pass & 1u) ? (nst - 1 - (int)r) : (int)r;
}
// one thread: start the bulk loads of stage `seq`.  Invariant kept by the callers: stage seq + SK_STAGES is issued as soon as
// every thread has taken stage seq out of shared memory.
__device__ __forceinline__ void sk_issue(const SkCtx& c, uint32_t seq) {
  const int buf = seq % SK_STAGES;
  const int r = c.row0 + sk_idx(seq, c.nst) * SK_ROWS;
  const int nr = min(SK_ROWS, c.row1 - r);
  // the first `keep_rows` rows of the band are asked to stay in L2 (evict_last), the rest to stream through (evict_first)
  uint64_t policy;
  if (r - c.row0 < c.keep_rows) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  sk_mbar_expect(&c.full[buf], nr * c.row_bytes);
  for (int k = 0; k < nr; ++k)
    sk_bulk_load(c.stage_buf + ((size_t)buf * SK_ROWS + k) * c.NS, c.S + (size_t)(r + k) * c.ld, c.row_bytes, &c.full[buf], policy);
  if (c.pf > 0 && seq + (uint32_t)c.pf < c.total) {      // pull a later stage of the band from HBM into L2 ahead of its smem fill
    const int rp = c.row0 + sk_idx(seq + (uint32_t)c.pf, c.nst) * SK_ROWS;
    const int np = min(SK_ROWS, c.row1 - rp);
    for (int k = 0; k < np; ++k)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c.S + (size_t)(rp + k) * c.ld), "r"(c.row_bytes) : "memory");
  }
}
__device__ __forceinline__ void sk_wait_full(const SkCtx& c, SkRing& rg) {
  sk_mbar_wait(&c.full[rg.buf], (rg.full_par >> rg.buf) & 1u);
  rg.full_par ^= 1u << rg.buf;
}

// ---- exact mode: one stage (SK_ROWS rows): elements of my columns -> registers, row-pass partials, ONE block barrier, then
// every warp finishes the row reduction itself and runs the column pass from registers.
// FULL = all SK_ROWS rows and all SK_GROUPS column groups are present (straight-line code without predicates).
template <bool FULL>
__device__ __forceinline__ void sk_stage_exact(const SkCtx& c, int idx, SkRing& rg, const float (&vl)[SK_GROUPS][4],
                                               L2Acc (&col)[SK_GROUPS][4]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = c.NS, n4 = c.n4;
  const int pp = rg.seq & 1;
  const int r_base = c.row0 + idx * SK_ROWS;
  const int nr = FULL ? SK_ROWS : min(SK_ROWS, c.row1 - r_base);
  const float* sb = c.stage_buf + (size_t)rg.buf * SK_ROWS * N;
  sk_wait_full(c, rg);
  float x[SK_ROWS][SK_GROUPS][4];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k)
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      float4 t = (FULL || (k < nr && gi < n4)) ? reinterpret_cast<const float4*>(sb + (size_t)k * N)[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[k][g][0] = t.x * LOG2E; x[k][g][1] = t.y * LOG2E; x[k][g][2] = t.z * LOG2E; x[k][g][3] = t.w * LOG2E;
    }
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      L2Acc a, a1; a.init(); a1.init();
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        if (FULL || g * SK_THREADS + tid < n4) {
          a.add(x[k][g][0] + vl[g][0]); a1.add(x[k][g][1] + vl[g][1]); a.add(x[k][g][2] + vl[g][2]); a1.add(x[k][g][3] + vl[g][3]);
        }
      }
      a.merge(a1.m, a1.s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      if (lane == 0) { c.part_m[pp][k][warp] = a.m; c.part_s[pp][k][warp] = a.s; }
    }
  }
  __syncthreads();              // partials published; every thread has its elements in registers -> the buffer is free
  if (tid == 0 && rg.seq + SK_STAGES < rg.total) sk_issue(c, rg.seq + SK_STAGES);
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      L2Acc a; a.init();
      if (lane < SK_WARPS) { a.m = c.part_m[pp][k][lane]; a.s = c.part_s[pp][k][lane]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      a.add(c.extra_row);
      const float ui = c.norm - a.lse_ln();
      if (tid == 0) c.u[r_base + k] = ui;
      const float ul = ui * LOG2E;
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        if (FULL || g * SK_THREADS + tid < n4) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) col[g][cc].add(x[k][g][cc] + ul);
        }
      }
    }
  }
  rg.advance();
}

__device__ __forceinline__ void sk_band_exact(const SkCtx& c, SkRing& rg, const float4 (&vraw)[SK_GROUPS]) {
  const int tid = threadIdx.x;
  const int N = c.NP, n4 = c.n4;
  float vl[SK_GROUPS][4];                                          // old v of my columns, log2 domain
  L2Acc col[SK_GROUPS][4];
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const float4 t = vraw[g];
    vl[g][0] = t.x * LOG2E; vl[g][1] = t.y * LOG2E; vl[g][2] = t.z * LOG2E; vl[g][3] = t.w * LOG2E;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) col[g][cc].init();
  }
  const bool full_cols = c.NS == SK_MAXN;
  const bool rev = ((rg.seq / (uint32_t)c.nst) & 1u) != 0;
#pragma unroll 1
  for (int st = 0; st < c.nst; ++st) {
    const int idx = rev ? c.nst - 1 - st : st;
    if (full_cols && c.row0 + (idx + 1) * SK_ROWS <= c.row1) sk_stage_exact<true>(c, idx, rg, vl, col);
    else sk_stage_exact<false>(c, idx, rg, vl, col);
  }
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const int gi = g * SK_THREADS + tid;
    if (gi < n4) {
      reinterpret_cast<float4*>(c.pm + (size_t)c.cta * N)[gi] = make_float4(col[g][0].m, col[g][1].m, col[g][2].m, col[g][3].m);
      reinterpret_cast<float4*>(c.ps + (size_t)c.cta * N)[gi] = make_float4(col[g][0].s, col[g][1].s, col[g][2].s, col[g][3].s);
    }
  }
}

// ---- fast mode --------------------------------------------------------------------------------------------------------
// Column pass of a stage (row block `idx`, stage number seq, fast-stage number fq) from the registers kept since its
// phase A: cs_j += e_ij * a_i.  Every warp finishes the row sums itself as soon as all partials are published (mbarrier
// bpart): a_i = 2^(u_i + m_i - c_mu) = kfac / rowsum_i, one reciprocal.  The stage's duty warp (round-robin) also writes u_i,
// checks the sums and refills the shared-memory buffer the stage has released; nobody waits for it.
template <bool FULL>
__device__ __forceinline__ void sk_col_accum(const SkCtx& c, int idx, uint32_t seq, uint32_t fq, SkRing& rg,
                                             const float (&e)[SK_ROWS][SK_GROUPS][4], float2 (&cs)[SK_GROUPS][2], float& csN,
                                             bool refill = true) {
  const int warp = c.warp, lane = c.lane;
  const uint32_t pp = fq & 1u, slot = fq & 3u;
  const int row = lane >> 4;
  const bool valid = FULL || (c.row0 + idx * SK_ROWS + row < c.row1);
  const float mr = valid ? (c.c_nu - c.uold_s[idx * SK_ROWS + row]) : 0.f;
  const float ex = sk_ex2(c.extra_row - mr);                        // dustbin column term
  sk_mbar_wait(&c.bpart[pp], (rg.bp_par >> pp) & 1u);               // all warps: partials published, buffer emptied
  SK_STAMP(fq + 1u, 4);
  rg.bp_par ^= 1u << pp;
  const float2 pr = *reinterpret_cast<const float2*>(&c.part_s[slot][row][(lane & 7) * 2]);
  float t = pr.x + pr.y;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  const float sm = t + ex;
  const float a = valid ? __fdividef(c.kfac, sm) : 0.f;
  csN = fmaf(ex, a, csN);                                           // dustbin column (every lane; lanes 0 / 16 of warp 0 are read)
  const float a0 = __shfl_sync(0xffffffffu, a, 0), a1 = __shfl_sync(0xffffffffu, a, 16);
  const float2 a0v = make_float2(a0, a0), a1v = make_float2(a1, a1);
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    cs[g][0] = __ffma2_rn(make_float2(e[0][g][0], e[0][g][1]), a0v, cs[g][0]);
    cs[g][1] = __ffma2_rn(make_float2(e[0][g][2], e[0][g][3]), a0v, cs[g][1]);
  }
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    cs[g][0] = __ffma2_rn(make_float2(e[1][g][0], e[1][g][1]), a1v, cs[g][0]);
    cs[g][1] = __ffma2_rn(make_float2(e[1][g][2], e[1][g][3]), a1v, cs[g][1]);
  }
  if (warp == (int)(fq & (SK_WARPS - 1))) {                                               // duty: off everybody's critical path
    if (refill && lane == 0 && seq + SK_STAGES < rg.total) sk_issue(c, seq + SK_STAGES);
    if ((lane & 15) == 0 && valid) {
      if (!(sm > 0.f && sm < INFINITY)) atomicExch(c.flag, 1);
      const float ul = c.norm * LOG2E - (mr + __log2f(sm));
      c.unew_s[idx * SK_ROWS + row] = ul;
      c.u[c.row0 + idx * SK_ROWS + row] = ul * LN2;
    }
  }
}

// One stage: phase A of row block `idx` (results in e_cur), then the column pass of the previous stage (from e_prev): a
// warp only waits for the slowest warp of the PREVIOUS stage after finishing its own share of this one.
// FULL = both rows exist and every thread owns SK_GROUPS complete float4 column groups (N == SK_MAXN).
template <bool FULL, bool HAVE_PREV, bool PREV_FULL>
__device__ __forceinline__ void sk_stage_fast(const SkCtx& c, int idx, int idx_prev, SkRing& rg,
                                              const float2 (&vl)[SK_GROUPS][2], float2 (&cs)[SK_GROUPS][2], float& csN,
                                              float (&e_cur)[SK_ROWS][SK_GROUPS][4], const float (&e_prev)[SK_ROWS][SK_GROUPS][4]) {
  const int tid = c.tid, warp = c.warp, lane = c.lane;
  const int N = c.NS, n4 = c.n4;
  // partial sums go to slot fq % 4: a warp may run a stage ahead of another one that has arrived for stage s+1 but not yet read
  // the partials of stage s, so a slot is only rewritten four stages later (behind the wait on stage s+2's barrier)
  const uint32_t pp = rg.fq & 1u, slot = rg.fq & 3u;
  const int nr = FULL ? SK_ROWS : min(SK_ROWS, c.row1 - (c.row0 + idx * SK_ROWS));
  const float* sb = c.stage_buf + (size_t)rg.buf * SK_ROWS * N;
  float mrow[SK_ROWS], rsum[SK_ROWS];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) mrow[k] = (FULL || k < nr) ? (c.c_nu - c.uold_s[idx * SK_ROWS + k]) : 0.f;   // previous u
  SK_STAMP(rg.fq, 0);
  sk_wait_full(c, rg);
  SK_STAMP(rg.fq, 1);
  const float2 l2e = make_float2(LOG2E, LOG2E);
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    float2 s2 = make_float2(0.f, 0.f);
    const float2 nm = make_float2(-mrow[k], -mrow[k]);
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      if (FULL || (k < nr && gi < n4)) {
        const float4 t = reinterpret_cast<const float4*>(sb + (size_t)k * N)[gi];
        const float2 a01 = __ffma2_rn(make_float2(t.x, t.y), l2e, __fadd2_rn(vl[g][0], nm));
        const float2 a23 = __ffma2_rn(make_float2(t.z, t.w), l2e, __fadd2_rn(vl[g][1], nm));
        e_cur[k][g][0] = sk_ex2(a01.x); e_cur[k][g][1] = sk_ex2(a01.y);
        e_cur[k][g][2] = sk_ex2(a23.x); e_cur[k][g][3] = sk_ex2(a23.y);
        s2 = __fadd2_rn(s2, make_float2(e_cur[k][g][0], e_cur[k][g][1]));
        s2 = __fadd2_rn(s2, make_float2(e_cur[k][g][2], e_cur[k][g][3]));
      } else {
        e_cur[k][g][0] = e_cur[k][g][1] = e_cur[k][g][2] = e_cur[k][g][3] = 0.f;
      }
    }
    rsum[k] = s2.x + s2.y;
  }
  SK_STAMP(rg.fq, 2);
  // transposed warp reduction of the two row sums: lanes 0-15 end up with row 0, lanes 16-31 with row 1
  const bool hi = (lane & 16) != 0;
  float keep = hi ? rsum[1] : rsum[0];
  const float send = hi ? rsum[0] : rsum[1];
  keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
  if ((lane & 15) == 0) c.part_s[slot][lane >> 4][warp] = keep;
  __syncwarp();
  if (lane == 0) sk_mbar_arrive(&c.bpart[pp]);                       // release: this warp's partials, and its reads of the buffer
  SK_STAMP(rg.fq, 3);
  if (HAVE_PREV) sk_col_accum<PREV_FULL>(c, idx_prev, rg.seq - 1, rg.fq - 1, rg, e_prev, cs, csN);
  SK_STAMP(rg.fq, 5);
  rg.advance();
  ++rg.fq;
}

// One pass over the band with the potentials vl (log2 domain) of this thread's columns.  Results stay in registers: cs = column
// sums relative to the stabiliser m_j = c_mu - v_j, csN = the dustbin column's (valid in lanes 0 / 16 of warp 0: rows 0 / 1 of
// every stage).  The shared-memory buffer of the band's LAST stage is not refilled here: the caller stages the reduction of the
// column sums through it and refills it afterwards (stage rg.seq - 1 + SK_STAGES).
__device__ __forceinline__ void sk_band_fast(const SkCtx& c, SkRing& rg, const float2 (&vl)[SK_GROUPS][2], float2 (&cs)[SK_GROUPS][2],
                                             float& csN) {
  const int nst = c.nst;
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) cs[g][0] = cs[g][1] = make_float2(0.f, 0.f);
  csN = 0.f;
  const bool rev = ((rg.seq / (uint32_t)nst) & 1u) != 0;
  // everything but (possibly) the band's last row block is complete: the ragged block — the first stage of a backward
  // pass, the last of a forward one — takes the predicated instantiation
  const bool all_full = (c.NS == SK_MAXN) && (c.row0 + nst * SK_ROWS == c.row1);
  const int ragged = all_full ? -1 : ((c.NS == SK_MAXN) ? nst - 1 : -2);                // -2: every stage is predicated
  float eA[SK_ROWS][SK_GROUPS][4], eB[SK_ROWS][SK_GROUPS][4];      // registers rotate between the two (loop unrolled by 2)
#define SK_IDX(st) (rev ? nst - 1 - (st) : (st))
#define SK_ISFULL(idx) (ragged == -1 || (ragged >= 0 && (idx) != ragged))
#define SK_STAGE(st, cur, prev)                                                                              \
  {                                                                                                          \
    const int i_ = SK_IDX(st), ip_ = SK_IDX((st) - 1);                                                       \
    const bool f_ = SK_ISFULL(i_), fp_ = SK_ISFULL(ip_);                                                     \
    if (f_ && fp_) sk_stage_fast<true, true, true>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                 \
    else if (f_) sk_stage_fast<true, true, false>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                  \
    else sk_stage_fast<false, true, false>(c, i_, ip_, rg, vl, cs, csN, cur, prev);                         \
  }
  int st = 1;
  if (SK_ISFULL(SK_IDX(0))) sk_stage_fast<true, false, false>(c, SK_IDX(0), 0, rg, vl, cs, csN, eA, eB);
  else sk_stage_fast<false, false, false>(c, SK_IDX(0), 0, rg, vl, cs, csN, eA, eB);
#pragma unroll 1
  for (; st + 1 < nst; st += 2) {
    SK_STAGE(st, eB, eA);
    SK_STAGE(st + 1, eA, eB);
  }
  if (st < nst) {
    SK_STAGE(st, eB, eA);
    sk_col_accum<false>(c, SK_IDX(nst - 1), rg.seq - 1, rg.fq - 1, rg, eB, cs, csN, false);
  } else {
    sk_col_accum<false>(c, SK_IDX(nst - 1), rg.seq - 1, rg.fq - 1, rg, eA, cs, csN, false);
  }
#undef SK_STAGE
#undef SK_ISFULL
#undef SK_IDX
}

__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_fused_kernel(const float* __restrict__ S, int M, int N, int ld, float alpha,
                                                                       int iters, float* u, float* v, float* pm, float* ps,
                                                                       int* flag, unsigned long long* acc_base, int fx_shift, int rows_per_cta, int allow_fast,
                                                                       int keep_pct, int pf_stages) {
  extern __shared__ __align__(128) unsigned char sk_smem[];
  float* stage_buf = reinterpret_cast<float*>(sk_smem);                         // [SK_STAGES][SK_ROWS][N]
  __shared__ __align__(8) uint64_t full[SK_STAGES], bpart[2];
  __shared__ __align__(8) float part_m[2][SK_ROWS][SK_WARPS], part_s[4][SK_ROWS][SK_WARPS];
  __shared__ float red_m[SK_WARPS], red_s[SK_WARPS];
  __shared__ __align__(8) float uold_s[2][SK_MAX_BAND];           // u of the band, log2 domain: previous / this iteration
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int row0 = min(M, cta * rows_per_cta), row1 = min(M, row0 + rows_per_cta);
  const int nst = (row1 - row0 + SK_ROWS - 1) / SK_ROWS;                       // stages per iteration for this CTA (>= 1)
  const float norm = -logf((float)M + (float)N);
  const float log_mu_last = logf((float)N) + norm, log_nu_last = logf((float)M) + norm;
  const float c_mu = fmaxf(norm, log_mu_last) * LOG2E, c_nu = fmaxf(norm, log_nu_last) * LOG2E;   // log2 of the largest marginals
  const int n4 = (N + 3) >> 2, NP = n4 * 4;      // columns [N, NP) hold -1e30 (filled by the launcher): they add 0 to every sum
  // Wide matrices (N >= 7/8 of the maximum) use the maximal smem row pitch: the bulk copies fill the first NP floats of a row,
  // the tail [NP, SK_MAXN) is set to -1e30 once, and every stage takes the straight-line (unpredicated) instantiation.
  const int NS = sk_smem_pitch(NP);

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) sk_mbar_init(&full[s], 1);
    for (int s = 0; s < 2; ++s) sk_mbar_init(&bpart[s], SK_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (NS != NP)
    for (int i = tid; i < SK_STAGES * SK_ROWS * (NS - NP); i += SK_THREADS)
      stage_buf[(size_t)(i / (NS - NP)) * NS + NP + i % (NS - NP)] = -1e30f;
  __syncthreads();

  SkCtx ctx;
  ctx.S = S; ctx.stage_buf = stage_buf; ctx.full = full; ctx.bpart = bpart;
  ctx.part_m = part_m; ctx.part_s = part_s; ctx.u = u; ctx.v = v;
  ctx.pm = pm; ctx.ps = ps; ctx.flag = flag; ctx.M = M; ctx.N = N; ctx.NP = NP; ctx.NS = NS; ctx.ld = ld; ctx.n4 = n4; ctx.row0 = row0; ctx.row1 = row1;
  ctx.uold_s = uold_s[0]; ctx.unew_s = uold_s[1];
  ctx.nst = nst; ctx.cta = cta; ctx.norm = norm; ctx.c_mu = c_mu; ctx.c_nu = c_nu; ctx.row_bytes = (uint32_t)NP * 4u;
  ctx.keep_rows = (row1 - row0) * keep_pct / 100;
  ctx.tid = tid; ctx.warp = warp; ctx.lane = lane;
  ctx.pf = pf_stages;
  ctx.total = (uint32_t)iters * (uint32_t)nst;
  ctx.kfac = exp2f(norm * LOG2E - c_mu);                             // a_i = 2^(u_i + m_i - c_mu) = kfac / rowsum_i

  // the band is re-streamed every iteration: stages are counted across iterations
  SkRing rg;
  rg.seq = 0; rg.buf = 0; rg.full_par = 0; rg.fq = 0; rg.bp_par = 0;
  rg.total = (uint32_t)iters * (uint32_t)nst;
  if (tid == 0)
    for (uint32_t s = 0; s < rg.total && s < SK_STAGES; ++s) sk_issue(ctx, s);
  bool fast_ok = allow_fast != 0;

  unsigned int* gbar = reinterpret_cast<unsigned int*>(flag) + 1;      // monotonic arrival counter of the grid barrier (zeroed by the host)
  unsigned int gtarget = 0;
  // Grid barrier: one release-add + relaxed polling per CTA.  Cumulativity through the two block barriers makes every
  // thread's earlier writes visible to every thread of the grid afterwards (readers use ld.global.cg).
  auto grid_barrier = [&]() {
    __syncthreads();
    gtarget += (unsigned int)G;
    if (tid == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
      unsigned int seen;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gbar) : "memory");
      } while (seen < gtarget);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
  };

  // ---- state of the fast iterations (registers / shared memory, never through HBM) ----
  float2 vl[SK_GROUPS][2];               // v of my 16 columns, log2 domain
  float vN_l = 0.f, uM_l = 0.f;          // dustbin column / dustbin row potentials, log2 domain
  bool v_regs = false;                   // vl / vN_l / uM_l / uold are current (else: reload from u, v in global memory)
  uint32_t fit = 0;                      // fast iterations so far: sums of iteration f go through acc[f % 3]
  float* uold = uold_s[0];
  float* unew = uold_s[1];
  const float alpha_l = alpha * LOG2E, norm_l = norm * LOG2E;
  const float fx_up = __int_as_float((127 + fx_shift) << 23), fx_dn = __int_as_float((127 - fx_shift) << 23);   // 2^shift, 2^-shift
  const size_t acc_stride = (size_t)NP + 2;                           // u64 per accumulation buffer: NP columns, [NP] = dustbin column

  for (int it = 0; it < iters; ++it) {
    const bool fast = fast_ok && it >= SK_EXACT_ITERS;
    if (fast) {
      // ================= fast iteration: ONE grid barrier =================
      // Column sums leave the CTA as 64-bit fixed-point numbers and are added up by the L2 (one bulk reduction per CTA): integer
      // addition is associative, so the result does not depend on the order in which the CTAs arrive (bit-reproducible), and
      // after the barrier EVERY CTA derives the new v of its threads' columns itself — no combine phase, no second barrier.
      float uM_prev;
      if (!v_regs) {
        // first fast iteration (or after exact ones): potentials come from global memory
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int gi = g * SK_THREADS + tid;
          const float4 t = (gi < n4) ? __ldcg(reinterpret_cast<const float4*>(v) + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
          vl[g][0] = make_float2(t.x * LOG2E, t.y * LOG2E); vl[g][1] = make_float2(t.z * LOG2E, t.w * LOG2E);
        }
        vN_l = __ldcg(v + N) * LOG2E;
        uM_prev = __ldcg(u + M) * LOG2E;
        for (int i = tid; i < row1 - row0; i += SK_THREADS) uold[i] = __ldcg(u + row0 + i) * LOG2E;
      } else {
        uM_prev = uM_l;
      }
      // dustbin row: u_M = log_mu_last - LSE_{j <= N}(alpha + v_j).  alpha + v_j <= m_M = c_nu - u_M(previous) (same a-priori
      // bound as for the other rows), so plain sums of 2^(alpha + v_j - m_M) are safe; every CTA computes it (same order: same bits)
      const float m_M = c_nu - uM_prev;
      {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int j = (g * SK_THREADS + tid) * 4;
          if (j + 0 < N) t += sk_ex2(alpha_l + vl[g][0].x - m_M);
          if (j + 1 < N) t += sk_ex2(alpha_l + vl[g][0].y - m_M);
          if (j + 2 < N) t += sk_ex2(alpha_l + vl[g][1].x - m_M);
          if (j + 3 < N) t += sk_ex2(alpha_l + vl[g][1].y - m_M);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red_s[warp] = t;
      }
      ctx.extra_row = alpha_l + vN_l;                                 // dustbin column term of every row sum
      ctx.uold_s = uold; ctx.unew_s = unew;
      __syncthreads();
      bool bad = false;
      {
        float t = sk_ex2(alpha_l + vN_l - m_M);
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) t += red_s[w];
        bad = !(t > 0.f && t < INFINITY);
        uM_l = log_mu_last * LOG2E - (m_M + __log2f(t));
      }
      float2 cs[SK_GROUPS][2];
      float csN;
      SK_BSTAMP(fit, 0);
      sk_band_fast(ctx, rg, vl, cs, csN);
      SK_BSTAMP(fit, 1);
      // ---- column sums -> fixed point -> the shared-memory buffer of the band's last stage (free: every warp has passed the wait
      //      for that stage's partials, i.e. every warp has taken its elements out of it) -> one bulk reduction into acc ----
      unsigned long long* acc = acc_base + (size_t)(fit % 3u) * acc_stride;
      {
        const uint32_t last_buf = (rg.buf == 0 ? SK_STAGES : rg.buf) - 1u;
        // The staging area is the DATA part of the buffer's two rows (16 n4 bytes each = half of the 32 n4 bytes of sums): the
        // -1e30 tails [NP, NS) of the rows must survive.  16-byte unit w (two columns) goes to row w / n4, offset (w % n4) * 16.
        unsigned char* stg0 = reinterpret_cast<unsigned char*>(stage_buf + (size_t)last_buf * SK_ROWS * NS);
        unsigned char* stg1 = stg0 + (size_t)NS * 4;
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          const int gi = g * SK_THREADS + tid;
          if (gi < n4) {
            const int w0 = 2 * gi, w1 = 2 * gi + 1;
            *reinterpret_cast<ulonglong2*>(w0 < n4 ? stg0 + (size_t)w0 * 16 : stg1 + (size_t)(w0 - n4) * 16) =
                make_ulonglong2(__float2ull_rn(cs[g][0].x * fx_up), __float2ull_rn(cs[g][0].y * fx_up));
            *reinterpret_cast<ulonglong2*>(w1 < n4 ? stg0 + (size_t)w1 * 16 : stg1 + (size_t)(w1 - n4) * 16) =
                make_ulonglong2(__float2ull_rn(cs[g][1].x * fx_up), __float2ull_rn(cs[g][1].y * fx_up));
          }
        }
        asm volatile("fence.proxy.async;" ::: "memory");             // my generic-proxy writes (shared and global) before the bulk engine's
        const float csN1 = __shfl_sync(0xffffffffu, csN, 16);
        __syncthreads();
        SK_BSTAMP(fit, 2);
        if (tid == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                       ::"l"(acc), "r"((uint32_t)__cvta_generic_to_shared(stg0)), "r"((uint32_t)n4 * 16u) : "memory");
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                       ::"l"(acc + 2 * (size_t)n4), "r"((uint32_t)__cvta_generic_to_shared(stg1)), "r"((uint32_t)n4 * 16u) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          atomicAdd(acc + NP, __float2ull_rn((csN + csN1) * fx_up));  // dustbin column: one more 64-bit integer add
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer read: hand it back to the ring
          if (rg.seq - 1u + SK_STAGES < rg.total) sk_issue(ctx, rg.seq - 1u + SK_STAGES);
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the reduction is performed before this CTA arrives
        }
      }
      SK_BSTAMP(fit, 3);
      grid_barrier();
      SK_BSTAMP(fit, 4);
      // ---- every CTA: new v of my columns from the reduced sums; recycle the buffer of two iterations ago ----
      {
        unsigned long long* old = acc_base + (size_t)((fit + 2u) % 3u) * acc_stride;
        for (size_t i = (size_t)cta * SK_THREADS + tid; i < acc_stride; i += (size_t)G * SK_THREADS) old[i] = 0ull;
      }
      const float extra_col = alpha_l + uM_l;
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        const int gi = g * SK_THREADS + tid;
        if (gi < n4) {
          const ulonglong2 s01 = __ldcg(reinterpret_cast<const ulonglong2*>(acc) + 2 * gi);
          const ulonglong2 s23 = __ldcg(reinterpret_cast<const ulonglong2*>(acc) + 2 * gi + 1);
          const float fs[4] = {__ull2float_rn(s01.x) * fx_dn, __ull2float_rn(s01.y) * fx_dn, __ull2float_rn(s23.x) * fx_dn,
                               __ull2float_rn(s23.y) * fx_dn};
          float vo[4] = {vl[g][0].x, vl[g][0].y, vl[g][1].x, vl[g][1].y};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float mj = c_mu - vo[cc];                            // the stabiliser used in the band (old v)
            const float sj = fs[cc] + sk_ex2(extra_col - mj);          // + dustbin row
            if (gi * 4 + cc < N && !(sj > 0.f && sj < INFINITY)) bad = true;
            vo[cc] = norm_l - (mj + __log2f(sj));
          }
          vl[g][0] = make_float2(vo[0], vo[1]); vl[g][1] = make_float2(vo[2], vo[3]);
        }
      }
      {
        const float mN = c_mu - vN_l;
        const float sN = __ull2float_rn(__ldcg(acc + NP)) * fx_dn + sk_ex2(extra_col - mN);
        if (!(sN > 0.f && sN < INFINITY)) bad = true;
        vN_l = log_nu_last * LOG2E - (mN + __log2f(sN));
      }
      { float* t = uold; uold = unew; unew = t; }
      v_regs = true;
      SK_BSTAMP(fit, 5);
      ++fit;
      if (__syncthreads_or((bad || __ldcg(flag) != 0) ? 1 : 0)) {
        // the fast mode tripped (potential jump beyond the f32 exponent range): restart the whole solve in exact mode.  Every
        // CTA sees the same sums and the same flag, so all of them take this branch together.
        fast_ok = false; v_regs = false;
        for (int j = cta * SK_THREADS + tid; j <= N; j += G * SK_THREADS) v[j] = 0.f;
        for (int i = cta * SK_THREADS + tid; i <= M; i += G * SK_THREADS) u[i] = 0.f;
        const uint32_t old_total = rg.total;
        rg.total = rg.seq + (uint32_t)iters * (uint32_t)nst;
        ctx.total = rg.total;
        if (tid == 0)      // stages below min(old_total, seq + SK_STAGES) are already in flight
          for (uint32_t sq = min(old_total, rg.seq + SK_STAGES); sq < rg.total && sq < rg.seq + SK_STAGES; ++sq) sk_issue(ctx, sq);
        it = -1;
        grid_barrier();
      }
      continue;
    }
    // ================= exact iteration: running maxima, column partials in global memory, two grid barriers =================
    v_regs = false;
    float4 vraw[SK_GROUPS];
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      vraw[g] = (gi < n4) ? __ldcg(reinterpret_cast<const float4*>(v) + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ctx.extra_row = (alpha + __ldcg(v + N)) * LOG2E;                  // dustbin column term of every row LSE (old v)
    // dustbin row: u[M] = log_mu_last - LSE_j(alpha + v_j), j in [0, N]   (last CTA: its band is the short one)
    if (cta == G - 1) {
      L2Acc a; a.init();
      for (int j0 = tid; j0 <= N; j0 += 16 * SK_THREADS) {
        float vv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int j = j0 + k * SK_THREADS; vv[k] = (j <= N) ? __ldcg(v + j) : -INFINITY; }
#pragma unroll
        for (int k = 0; k < 16; ++k) if (vv[k] != -INFINITY) a.add((alpha + vv[k]) * LOG2E);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
      __syncthreads();
      if (tid == 0) {
        L2Acc t; t.init();
        for (int w = 0; w < SK_WARPS; ++w) t.merge(red_m[w], red_s[w]);
        u[M] = log_mu_last - t.lse_ln();
      }
    }
    __syncthreads();
    sk_band_exact(ctx, rg, vraw);
    grid_barrier();
    // ---- combine: v_j = log_nu_j - LSE_i(S_ij + u_i) incl. the dustbin row; v[N] from all u ----
    {
      const float extra_col = (alpha + __ldcg(u + M)) * LOG2E;
      // one warp per 4-column tile, lane = 4 * sub + column: 8 lanes share a column
      for (int tile = cta * SK_WARPS + warp; tile < n4; tile += G * SK_WARPS) {
        const int sub = lane >> 2, j = tile * 4 + (lane & 3);
        L2Acc a; a.init();
#pragma unroll 2
        for (int c = sub; c < G; c += 8) a.merge(__ldcg(pm + (size_t)c * NP + j), __ldcg(ps + (size_t)c * NP + j));
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        if (sub == 0 && j < N) {
          a.add(extra_col);
          v[j] = norm - a.lse_ln();
        }
      }
      if (cta == G - 1) {                                             // v[N] = log_nu_last - LSE_{i <= M}(alpha + u_i)
        L2Acc a; a.init();
        for (int i0 = tid; i0 <= M; i0 += 16 * SK_THREADS) {
          float uu[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) { const int i = i0 + k * SK_THREADS; uu[k] = (i <= M) ? __ldcg(u + i) : -INFINITY; }
#pragma unroll
          for (int k = 0; k < 16; ++k) if (uu[k] != -INFINITY) a.add((alpha + uu[k]) * LOG2E);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
        __syncthreads();
        if (warp == 0) {
          L2Acc t; t.init();
          if (lane < SK_WARPS) { t.m = red_m[lane]; t.s = red_s[lane]; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t.merge(__shfl_xor_sync(0xffffffffu, t.m, o), __shfl_xor_sync(0xffffffffu, t.s, o));
          if (lane == 0) v[N] = log_nu_last - t.lse_ln();
        }
      }
    }
    grid_barrier();
  }
  // the potentials of the last fast iteration are still in registers: CTA 0 writes them out (u_i, i < M, went out row by row)
  if (v_regs && cta == 0) {
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int j = (g * SK_THREADS + tid) * 4;
      if (j + 0 < N) v[j + 0] = vl[g][0].x * LN2;
      if (j + 1 < N) v[j + 1] = vl[g][0].y * LN2;
      if (j + 2 < N) v[j + 2] = vl[g][1].x * LN2;
      if (j + 3 < N) v[j + 3] = vl[g][1].y * LN2;
    }
    if (tid == 0) { v[N] = vN_l * LN2; u[M] = uM_l * LN2; }
  }
}

// the row pitch must hold whole float4 groups (bulk copies are 16-byte granular); N itself may be anything in [64, SK_MAXN]
static bool sinkhorn_fused_ok(const float* S, int M, int N, int ld) {
  return (ld & 3) == 0 && ld >= ((N + 3) & ~3) && (reinterpret_cast<uintptr_t>(S) & 15) == 0 && N <= SK_MAXN && N >= 64 && M >= 1;
}
// columns [N, round_up(N, 4)) of every row <- -1e30: the fused kernel then treats them as ordinary columns of weight 0
__global__ void sk_pad_fill_kernel(float* S, int M, int N, int ld) {
  const int npad = ((N + 3) & ~3) - N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * npad) return;
  S[(size_t)(i / npad) * ld + N + i % npad] = -1e30f;
}
static int g_sinkhorn_fast = 1;   // 1 = a-priori stabilisers after the first two iterations, 0 = running maxima throughout

// returns 0 when the fused kernel ran, 1 when the shape is unsupported (caller falls back to the two-pass kernels)
static int sinkhorn_fused_launch(float* S, int M, int N, int ld, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st) {
  if (!sinkhorn_fused_ok(S, M, N, ld) || iters <= 0) return 1;
  static int coop = -1, sms = 0;   // B200 boxes are homogeneous: queried once
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!coop || sms <= 0 || sms > 256) return 1;
  const int NP = (N + 3) & ~3;
  const size_t smem = (size_t)SK_STAGES * SK_ROWS * sk_smem_pitch(NP) * sizeof(float);
  static bool attr_seen[64] = {};
  if (i4d_first_use_on_device(attr_seen)) {
    if (cudaFuncSetAttribute(sinkhorn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SK_STAGES * SK_ROWS * SK_MAXN * (int)sizeof(float)) != cudaSuccess) { cudaGetLastError(); return 1; }
  }
  int G = sms;
  int rpc = (M + G - 1) / G;
  rpc = (rpc + SK_ROWS - 1) / SK_ROWS * SK_ROWS;
  if (rpc > SK_MAX_BAND) return 1;
  G = (M + rpc - 1) / rpc;                      // CTAs that actually own rows (<= sms <= 256 partial slots)
  cudaMemsetAsync(u, 0, (size_t)(M + 1) * sizeof(float), st);
  cudaMemsetAsync(v, 0, (size_t)(NP + 1) * sizeof(float), st);       // v[N+1 .. NP] are read (never written) as the pad columns' potentials
  if (NP != N) sk_pad_fill_kernel<<<i4d_cdiv(M * (NP - N), 256), 256, 0, st>>>(S, M, N, ld);
  float* pm = w.pm; float* ps = w.ps;
  int* flag = w.pi;
  cudaMemsetAsync(flag, 0, 2 * sizeof(int), st);   // [0] fast-mode trip flag, [1] grid-barrier arrival counter
  // three rotating accumulation buffers of (NP + 2) 64-bit fixed-point column sums (the argmax-index partials of the workspace
  // are free during the solve).  Column sums relative to their stabilisers are bounded by M / N (total row mass over the
  // largest column marginal): scale 2^shift with (M / N) * 2^shift < 2^62.
  unsigned long long* acc_base = reinterpret_cast<unsigned long long*>(w.pi + 4);   // (the first 16 bytes hold flag / barrier counter)
  cudaMemsetAsync(acc_base, 0, 3 * ((size_t)NP + 2) * sizeof(unsigned long long), st);
  int fx_shift = 61;
  for (long long r = 1; r * (long long)N < (long long)M; r *= 2) --fx_shift;
  int allow_fast = g_sinkhorn_fast;
  static int keep_pct = -1;      // share of every band pinned in L2 with evict_last (I4D_SK_KEEP_PCT overrides, for experiments)
  if (keep_pct < 0) {
    const char* e = getenv("I4D_SK_KEEP_PCT");
    keep_pct = e ? atoi(e) : SK_KEEP_PCT_DEFAULT;
    if (keep_pct < 0 || keep_pct > 100) keep_pct = SK_KEEP_PCT_DEFAULT;
  }
  // L2 prefetch distance in stages (cp.async.bulk.prefetch.L2 of a later stage with every shared-memory refill): 1 measured
  // best at 8192^2 (48.8 -> 46.3 us per iteration; 2: 47.3, 3: 48.4).  I4D_SK_PF overrides, for experiments.
  static int pf_stages = -1;
  if (pf_stages < 0) { const char* e = getenv("I4D_SK_PF"); pf_stages = e ? atoi(e) : 1; if (pf_stages < 0 || pf_stages > 16) pf_stages = 1; }
  void* args[] = {(void*)&S, (void*)&M, (void*)&N, (void*)&ld, (void*)&alpha, (void*)&iters, (void*)&u, (void*)&v, (void*)&pm, (void*)&ps,
                  (void*)&flag, (void*)&acc_base, (void*)&fx_shift, (void*)&rpc, (void*)&allow_fast, (void*)&keep_pct, (void*)&pf_stages};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)sinkhorn_fused_kernel, dim3(G), dim3(SK_THREADS), args, smem, st);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }
  return 0;
}

// test / comparison hook: 0 = fused Sinkhorn when possible (default), 1 = always the two-pass kernels
extern "C" __attribute__((visibility("default"))) int i4d_set_sinkhorn_mode(int mode) {
  I4D_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0 (fused, fast stabilisers), 1 (two-pass kernels) or 2 (fused, running maxima)");
  g_sinkhorn_mode = mode == 1 ? 1 : 0;
  g_sinkhorn_fast = mode == 0 ? 1 : 0;
  return I4D_OK;
}
