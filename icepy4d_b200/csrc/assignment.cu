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
// Building blocks (S is M x N row-major, ld = N):
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
__global__ void __launch_bounds__(ROW_WARPS * 32) row_reduce_kernel(const float* __restrict__ S, int M, int N,
                                                                    float scale, const float* __restrict__ coloff,
                                                                    const float* __restrict__ extra_ptr, float extra_add, float base,
                                                                    float* __restrict__ out_val,
                                                                    int* __restrict__ out_idx) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* r = S + (size_t)row * N;
  if (MODE == 2) {
    ArgAcc a; a.init();
    if ((N & 3) == 0) {
      for (int j = lane * 4; j < N; j += 128) {
        float4 x = ldg_stream(reinterpret_cast<const float4*>(r + j));
        float4 o = coloff ? __ldg(reinterpret_cast<const float4*>(coloff + j)) : make_float4(0, 0, 0, 0);
        a.add(fmaf(scale, x.x, o.x), j); a.add(fmaf(scale, x.y, o.y), j + 1);
        a.add(fmaf(scale, x.z, o.z), j + 2); a.add(fmaf(scale, x.w, o.w), j + 3);
      }
    } else {
      for (int j = lane; j < N; j += 32) a.add(fmaf(scale, r[j], coloff ? coloff[j] : 0.f), j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, a.m, o);
      int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
      a.add(om, oi);
    }
    if (lane == 0) { out_val[row] = a.m; out_idx[row] = a.i; }
  } else {
    LseAcc a; a.init();
    if ((N & 3) == 0) {
      for (int j = lane * 4; j < N; j += 128) {
        float4 x = ldg_stream(reinterpret_cast<const float4*>(r + j));
        float4 o = coloff ? __ldg(reinterpret_cast<const float4*>(coloff + j)) : make_float4(0, 0, 0, 0);
        a.add4(fmaf(scale, x.x, o.x), fmaf(scale, x.y, o.y), fmaf(scale, x.z, o.z), fmaf(scale, x.w, o.w));
      }
    } else {
      for (int j = lane; j < N; j += 32) a.add1(fmaf(scale, r[j], coloff ? coloff[j] : 0.f));
    }
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
__global__ void __launch_bounds__(COL_THREADS) col_partial_kernel(const float* __restrict__ S, int M, int N,
                                                                  float scale, const float* __restrict__ rowoff,
                                                                  int rows_per_split, float* __restrict__ pm,
                                                                  float* __restrict__ ps, int* __restrict__ pi) {
  const int j0 = (blockIdx.x * COL_THREADS + threadIdx.x) * 4;
  const int i0 = blockIdx.y * rows_per_split, i1 = min(M, i0 + rows_per_split);
  if (j0 >= N) return;
  const bool vec = ((N & 3) == 0);
  const int nv = vec ? 4 : min(4, N - j0);
  if (MODE == 2) {
    ArgAcc a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c].init();
    for (int i = i0; i < i1; ++i) {
      float o = rowoff ? __ldg(rowoff + i) : 0.f;
      const float* p = S + (size_t)i * N + j0;
      if (vec) {
        float4 x = ldg_stream(reinterpret_cast<const float4*>(p));
        a[0].add(fmaf(scale, x.x, o), i); a[1].add(fmaf(scale, x.y, o), i);
        a[2].add(fmaf(scale, x.z, o), i); a[3].add(fmaf(scale, x.w, o), i);
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
          x[r] = ldg_stream(reinterpret_cast<const float4*>(S + (size_t)(i + r) * N + j0));
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
      const float* p = S + (size_t)i * N + j0;
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
                                   float extra_add, float base, float* __restrict__ out_val, int* __restrict__ out_idx) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
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
  int target = 6 * i4d_num_sms();  // several 128-thread CTAs resident per SM keep enough 128-bit loads in flight
  int splits = i4d_cdiv(target, strips);
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  if (splits > M) splits = M;
  return splits;
}

// workspace layout (floats): pm[256*N] ps[256*N] pi[256*N] | val0[M] idx0[M] idx1[N] val1[N] rowoff[M+1] coloff[N+1] rl[M] cl[N]
extern "C" __attribute__((visibility("default"))) size_t i4d_assignment_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return ((size_t)256 * N * 3 + 4 * (size_t)(M + 8) + 4 * (size_t)(N + 8) + 64) * sizeof(float);
}

struct AssignWs {
  float *pm, *ps; int* pi;
  float* val0; int* idx0; int* idx1; float* val1; float* rowoff; float* coloff; float* rl; float* cl;
  AssignWs(void* w, int M, int N) {
    float* f = reinterpret_cast<float*>(w);
    pm = f; ps = f + (size_t)256 * N; pi = reinterpret_cast<int*>(f + (size_t)512 * N);
    float* q = f + (size_t)768 * N;
    const size_t ms = ((size_t)M + 1 + 3) & ~(size_t)3, ns = ((size_t)N + 1 + 3) & ~(size_t)3;  // keep 16 B alignment (float4 loads)
    val0 = q; q += ms; idx0 = reinterpret_cast<int*>(q); q += ms; rowoff = q; q += ms; rl = q; q += ms;
    idx1 = reinterpret_cast<int*>(q); q += ns; val1 = q; q += ns; coloff = q; q += ns; cl = q;
  }
};

template <int MODE>
static void launch_row(const float* S, int M, int N, float scale, const float* coloff, const float* extra_ptr,
                       float extra_add, float base, float* out_val, int* out_idx, cudaStream_t st) {
  row_reduce_kernel<MODE><<<i4d_cdiv(M, ROW_WARPS), ROW_WARPS * 32, 0, st>>>(S, M, N, scale, coloff, extra_ptr,
                                                                             extra_add, base, out_val, out_idx);
}
template <int MODE>
static void launch_col(const float* S, int M, int N, float scale, const float* rowoff, const float* extra_ptr,
                       float extra_add, float base, float* out_val, int* out_idx, AssignWs& w, cudaStream_t st) {
  int splits = col_splits_for(M, N);
  int rps = i4d_cdiv(M, splits);
  rps = (rps + 3) & ~3;
  splits = i4d_cdiv(M, rps);
  dim3 grid(i4d_cdiv(N, COL_THREADS * 4), splits);
  col_partial_kernel<MODE><<<grid, COL_THREADS, 0, st>>>(S, M, N, scale, rowoff, rps, w.pm, w.ps, w.pi);
  col_combine_kernel<MODE><<<i4d_cdiv(N, 256), 256, 0, st>>>(w.pm, w.ps, w.pi, splits, N, extra_ptr, extra_add, base,
                                                             out_val, out_idx);
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
  launch_row<1>(S, M, N, scale, coloff, nullptr, 0.f, 0.f, out, nullptr, (cudaStream_t)stream);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
extern "C" __attribute__((visibility("default"))) int i4d_col_lse(const float* S, int M, int N, float scale, const float* rowoff, float* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(S && out && workspace && M > 0 && N > 0, "null pointer or empty matrix");
  if (int rc = check_ws("i4d_col_lse", M, N, workspace_bytes)) return rc;
  AssignWs w(workspace, M, N);
  launch_col<1>(S, M, N, scale, rowoff, nullptr, 0.f, 0.f, out, nullptr, w, (cudaStream_t)stream);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- Sinkhorn potentials ------------------------------------------------------------------------------
static void sinkhorn_iterations(const float* S, int M, int N, float alpha, int iters, float* u, float* v, AssignWs& w,
                                cudaStream_t st) {
  const float norm = -logf((float)M + (float)N);
  const float log_mu_last = logf((float)N) + norm, log_nu_last = logf((float)M) + norm;
  cudaMemsetAsync(u, 0, (size_t)(M + 1) * sizeof(float), st);
  cudaMemsetAsync(v, 0, (size_t)(N + 1) * sizeof(float), st);
  for (int it = 0; it < iters; ++it) {
    // u_i = log_mu_i - LSE_j(Z_ij + v_j): rows i < M stream S; the dustbin row is a vector LSE
    launch_row<0>(S, M, N, 1.f, v, v + N, alpha, norm, u, nullptr, st);
    vec_lse_kernel<<<1, 1024, 0, st>>>(v, N + 1, alpha, log_mu_last, u + M);
    // v_j = log_nu_j - LSE_i(Z_ij + u_i)
    launch_col<0>(S, M, N, 1.f, u, u + M, alpha, norm, v, nullptr, w, st);
    vec_lse_kernel<<<1, 1024, 0, st>>>(u, M + 1, alpha, log_nu_last, v + N);
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_sinkhorn(const float* scores, int M, int N, float bin_score, int iters, float* u, float* v,
                            void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(scores && u && v && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && iters >= 0, "bad sizes");
  if (int rc = check_ws("i4d_sinkhorn", M, N, workspace_bytes)) return rc;
  AssignWs w(workspace, M, N);
  sinkhorn_iterations(scores, M, N, bin_score, iters, u, v, w, (cudaStream_t)stream);
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

extern "C" __attribute__((visibility("default"))) int i4d_sg_assign(const float* scores, int M, int N, float bin_score, int iters, float match_threshold,
                             int* matches0, int* matches1, float* mscores0, float* mscores1, float* u, float* v,
                             void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(scores && matches0 && matches1 && mscores0 && mscores1 && u && v && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0 && iters >= 0, "bad sizes");
  if (int rc = check_ws("i4d_sg_assign", M, N, workspace_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  AssignWs w(workspace, M, N);
  sinkhorn_iterations(scores, M, N, bin_score, iters, u, v, w, st);
  const float norm = -logf((float)M + (float)N);
  // P_ij = S_ij + u_i + v_j - norm over the core block: row argmax ignores u_i, column argmax ignores v_j
  launch_row<2>(scores, M, N, 1.f, v, nullptr, 0.f, 0.f, w.val0, w.idx0, st);
  launch_col<2>(scores, M, N, 1.f, u, nullptr, 0.f, 0.f, w.val1, w.idx1, w, st);
  mutual_kernel0<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.val0, w.idx0, u, -norm, w.idx1, M, match_threshold, matches0,
                                                   mscores0);
  mutual_kernel1<<<i4d_cdiv(N, 256), 256, 0, st>>>(w.idx0, w.idx1, matches0, mscores0, N, matches1, mscores1);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- LightGlue dual-softmax assignment ---------------------------------------------------------------
// scores_ij = log_softmax_j(sim)_ij + log_softmax_i(sim)_ij + logsigmoid(z0_i) + logsigmoid(z1_j)
// pass 1: row LSE rl_i and column LSE cl_j (two streams over sim).  pass 2: row argmax of
// (2 sim_ij + logsig(z1_j) - cl_j) and column argmax of (2 sim_ij + logsig(z0_i) - rl_i) (two more streams);
// the per-row constant (logsig(z0_i) - rl_i) is added afterwards.  The (M+1)x(N+1) matrix is never written.
__device__ __forceinline__ float logsigmoidf_(float z) { return fminf(z, 0.f) - log1pf(expf(-fabsf(z))); }
__global__ void lg_offsets_kernel(const float* __restrict__ z, const float* __restrict__ lse, int n,
                                  float* __restrict__ off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] = logsigmoidf_(z[i]) - lse[i];
}

extern "C" __attribute__((visibility("default"))) int i4d_lg_assign(const float* sim, int M, int N, const float* z0, const float* z1, float filter_threshold,
                             int* matches0, int* matches1, float* mscores0, float* mscores1, void* workspace,
                             size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(sim && z0 && z1 && matches0 && matches1 && mscores0 && mscores1 && workspace, "null pointer");
  I4D_CHECK_ARG(M > 0 && N > 0, "bad sizes");
  if (int rc = check_ws("i4d_lg_assign", M, N, workspace_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  AssignWs w(workspace, M, N);
  launch_row<1>(sim, M, N, 1.f, nullptr, nullptr, 0.f, 0.f, w.rl, nullptr, st);
  launch_col<1>(sim, M, N, 1.f, nullptr, nullptr, 0.f, 0.f, w.cl, nullptr, w, st);
  lg_offsets_kernel<<<i4d_cdiv(M, 256), 256, 0, st>>>(z0, w.rl, M, w.rowoff);
  lg_offsets_kernel<<<i4d_cdiv(N, 256), 256, 0, st>>>(z1, w.cl, N, w.coloff);
  launch_row<2>(sim, M, N, 2.f, w.coloff, nullptr, 0.f, 0.f, w.val0, w.idx0, st);
  launch_col<2>(sim, M, N, 2.f, w.rowoff, nullptr, 0.f, 0.f, w.val1, w.idx1, w, st);
  mutual_kernel0<<<i4d_cdiv(M, 256), 256, 0, st>>>(w.val0, w.idx0, w.rowoff, 0.f, w.idx1, M, filter_threshold,
                                                   matches0, mscores0);
  mutual_kernel1<<<i4d_cdiv(N, 256), 256, 0, st>>>(w.idx0, w.idx1, matches0, mscores0, N, matches1, mscores1);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
