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
static int sinkhorn_fused_launch(const float* S, int M, int N, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st);
static int g_sinkhorn_mode = 0;   // 0 = fused when the shape allows, 1 = always the two-pass kernels (tests / comparison)

static void sinkhorn_iterations(const float* S, int M, int N, float alpha, int iters, float* u, float* v, AssignWs& w,
                                cudaStream_t st) {
  if (g_sinkhorn_mode == 0 && sinkhorn_fused_launch(S, M, N, alpha, iters, u, v, w, st) == 0) return;
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

// ====================================================================================================================
// Fused persistent Sinkhorn (N <= 8192, N % 4 == 0): ONE read of the score matrix per iteration instead of two.
//
// One CTA per SM (cooperative launch), each owning a contiguous band of rows.  Rows are streamed HBM -> shared memory
// with cp.async.bulk (1-D TMA) through a 3-stage mbarrier ring of 2 rows (<= 64 KB) per stage.  For every stage:
//   phase A  row log-sum-exp of (S_ij + v_j) from shared memory (8 warps per row, warp-shuffle reduction)  -> u_i
//   phase B  the SAME shared-memory rows, now with the fresh u_i added, update the per-thread accumulators of the
//            16 columns this thread owns (registers, for the whole band)
// Phase A of stage st+1 shares a barrier interval with phase B of stage st (software pipeline).  After the band: column
// partials -> workspace, grid barrier, all CTAs combine the partials into v_j (+ the dustbin terms), grid barrier.
//
// Two arithmetic modes, same result up to f32 rounding:
//   exact : running (max, sum) per accumulator — used for the first two iterations (and always if the fast mode trips);
//   fast  : a-priori stabilisers instead of running maxima.  After a column update sum_i exp(S_ij + u_i + v_j) = nu_j, so
//           S_ij + v_j <= log nu_j - u_i: m_i = log(nu_max) - u_i(previous) bounds every term of row i from above; likewise
//           m_j = log(mu_max) - v_j(previous) for the columns.  One FFMA + FADD + MUFU.EX2 + FADD per element and pass, no
//           dependent max chain.  A sum that underflows to 0 (potential jump > ~80 nats between two iterations) or is not
//           finite raises a flag: the whole solve restarts in exact mode, on the device, without host involvement.
// Everything is kept in the log2 domain.  HBM traffic per iteration = M*N*4 bytes (algorithmic count of SURVEY.md §8d: 2*M*N*4).
// ====================================================================================================================
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define SK_THREADS 1024
#define SK_ROWS 3                 // rows per stage
#define SK_STAGES 2
#define SK_MAXN 8192
#define SK_GROUPS (SK_MAXN / 4 / SK_THREADS)   // float4 column groups per thread = 4
#define SK_EXACT_ITERS 2
#define SK_MAX_BAND 1024            // rows per CTA the previous-u staging buffer can hold

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
__device__ __forceinline__ void sk_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}


struct SkCtx {
  const float* S; float* stage_buf; uint64_t* full; float (*part_m)[SK_ROWS][SK_THREADS / 32]; float (*part_s)[SK_ROWS][SK_THREADS / 32];
  float* u; const float* v; float* pm; float* ps; int* flag;
  int M, N, n4, row0, row1, nstage_total, cta;
  float norm, c_mu, c_nu, extra_row;
  const float* uold_s;   // previous-iteration u of this CTA's band, pre-scaled by log2(e) (shared memory)
  uint32_t row_bytes;
};

// One stage (SK_ROWS rows) of the band: elements of my columns -> registers, row-pass partials, ONE block barrier, then
// every warp finishes the row reduction itself and runs the column pass from registers.  FAST = a-priori stabilisers,
// FULL = all SK_ROWS rows and all SK_GROUPS column groups are present (straight-line code without predicates).
template <bool FAST, bool FULL>
__device__ __forceinline__ void sk_stage(const SkCtx& c, int st, uint32_t& consumed, uint32_t& issued, uint32_t total_seq,
                                         const float (&vl)[SK_GROUPS][4], L2Acc (&col)[SK_GROUPS][4]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = c.N, n4 = c.n4;
  const uint32_t seq = consumed;
  const int buf = seq % SK_STAGES;
  const int r_base = c.row0 + st * SK_ROWS;
  const int nr = FULL ? SK_ROWS : min(SK_ROWS, c.row1 - r_base);
  const float* sb = c.stage_buf + (size_t)buf * SK_ROWS * N;
  float mrow[SK_ROWS];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) mrow[k] = (FAST && (FULL || k < nr)) ? (c.c_nu - c.uold_s[st * SK_ROWS + k]) : 0.f;   // previous u
  sk_mbar_wait(&c.full[buf], (seq / SK_STAGES) & 1);
  float x[SK_ROWS][SK_GROUPS][4];
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k)
#pragma unroll
    for (int g = 0; g < SK_GROUPS; ++g) {
      const int gi = g * SK_THREADS + tid;
      float4 t = (FULL || (k < nr && gi < n4)) ? reinterpret_cast<const float4*>(sb + (size_t)k * N)[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[k][g][0] = t.x * LOG2E; x[k][g][1] = t.y * LOG2E; x[k][g][2] = t.z * LOG2E; x[k][g][3] = t.w * LOG2E;
    }
  // ---- row pass partials (old v) ----
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      if (FAST) {
        // e_ij = 2^(x_ij + v_j - m_i) is kept in registers: the column pass needs no second exponential, because
        // 2^(x_ij + u_i - m_j) = e_ij * 2^(u_i + m_i - c_mu)   (m_j = c_mu - v_j): a rank-1 rescaling of the same kernel matrix.
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int g = 0; g < SK_GROUPS; ++g) {
          if (FULL || g * SK_THREADS + tid < n4) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) x[k][g][cc] = sk_ex2(x[k][g][cc] + (vl[g][cc] - mrow[k]));
            s0 += x[k][g][0] + x[k][g][1];
            s1 += x[k][g][2] + x[k][g][3];
          }
        }
        float sm = warp_sum(s0 + s1);
        if (lane == 0) c.part_s[st & 1][k][warp] = sm;
      } else {
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
        if (lane == 0) { c.part_m[st & 1][k][warp] = a.m; c.part_s[st & 1][k][warp] = a.s; }
      }
    }
  }
  __syncthreads();              // partials published; every thread has its elements in registers -> the buffer is free
  ++consumed;
  if (tid == 0 && issued < total_seq) {   // refill the buffer just released (wraps into the next iteration: S never changes)
    const int st_idx = issued % c.nstage_total, nbuf = issued % SK_STAGES;
    const int r = c.row0 + st_idx * SK_ROWS;
    const int nrr = min(SK_ROWS, c.row1 - r);
    sk_mbar_expect(&c.full[nbuf], nrr * c.row_bytes);
    for (int k = 0; k < nrr; ++k)
      sk_bulk_load(c.stage_buf + ((size_t)nbuf * SK_ROWS + k) * N, c.S + (size_t)(r + k) * N, c.row_bytes, &c.full[nbuf]);
    ++issued;
  }
  // ---- every warp finishes the row reduction itself (no second barrier), then the column pass from registers ----
#pragma unroll
  for (int k = 0; k < SK_ROWS; ++k) {
    if (FULL || k < nr) {
      float ui;
      if (FAST) {
        float sm = warp_sum(c.part_s[st & 1][k][lane]) + sk_ex2(c.extra_row - mrow[k]);
        if (!(sm > 0.f && sm < INFINITY) && tid == 0) atomicExch(c.flag, 1);
        ui = c.norm - (mrow[k] + __log2f(sm)) * LN2;
      } else {
        L2Acc a;
        a.m = c.part_m[st & 1][k][lane]; a.s = c.part_s[st & 1][k][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        a.add(c.extra_row);
        ui = c.norm - a.lse_ln();
      }
      if (tid == 0) c.u[r_base + k] = ui;
      const float ul = ui * LOG2E;
      const float ak = FAST ? sk_ex2(ul + mrow[k] - c.c_mu) : 0.f;     // row factor of the rank-1 rescaling (fast mode)
#pragma unroll
      for (int g = 0; g < SK_GROUPS; ++g) {
        if (FULL || g * SK_THREADS + tid < n4) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (FAST) col[g][cc].s = fmaf(x[k][g][cc], ak, col[g][cc].s);
            else col[g][cc].add(x[k][g][cc] + ul);
          }
        }
      }
    }
  }
}

// One iteration's pass over this CTA's band (row pass + column pass), templated on the arithmetic mode so that no
// per-element branch survives in the inner loops.
template <bool FAST>
__device__ __forceinline__ void sk_band(const SkCtx& c, uint32_t& consumed, uint32_t& issued, uint32_t total_seq) {
  const int tid = threadIdx.x;
  const int N = c.N, n4 = c.n4;
  float vl[SK_GROUPS][4];                                          // old v of my columns, log2 domain
  L2Acc col[SK_GROUPS][4];                                         // exact: (max, sum); fast: (stabiliser m_j, sum)
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const int gi = g * SK_THREADS + tid;
    float4 t = (gi < n4) ? __ldcg(reinterpret_cast<const float4*>(c.v) + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    vl[g][0] = t.x * LOG2E; vl[g][1] = t.y * LOG2E; vl[g][2] = t.z * LOG2E; vl[g][3] = t.w * LOG2E;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) { col[g][cc].init(); if (FAST) col[g][cc].m = c.c_mu - vl[g][cc]; }
  }
  const bool full_cols = n4 == SK_GROUPS * SK_THREADS;
  for (int st = 0; st < c.nstage_total; ++st) {
    if (full_cols && c.row0 + (st + 1) * SK_ROWS <= c.row1) sk_stage<FAST, true>(c, st, consumed, issued, total_seq, vl, col);
    else sk_stage<FAST, false>(c, st, consumed, issued, total_seq, vl, col);
  }
  // ---- column partials of this CTA ----
#pragma unroll
  for (int g = 0; g < SK_GROUPS; ++g) {
    const int gi = g * SK_THREADS + tid;
    if (gi < n4) {
      if (!FAST) reinterpret_cast<float4*>(c.pm + (size_t)c.cta * N)[gi] = make_float4(col[g][0].m, col[g][1].m, col[g][2].m, col[g][3].m);
      reinterpret_cast<float4*>(c.ps + (size_t)c.cta * N)[gi] = make_float4(col[g][0].s, col[g][1].s, col[g][2].s, col[g][3].s);
    }
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_fused_kernel(const float* __restrict__ S, int M, int N, float alpha,
                                                                       int iters, float* u, float* v, float* pm, float* ps,
                                                                       int* flag, int rows_per_cta, int allow_fast) {
  extern __shared__ __align__(128) unsigned char sk_smem[];
  float* stage_buf = reinterpret_cast<float*>(sk_smem);                         // [SK_STAGES][SK_ROWS][N]
  __shared__ __align__(8) uint64_t full[SK_STAGES];
  __shared__ float part_m[2][SK_ROWS][SK_THREADS / 32], part_s[2][SK_ROWS][SK_THREADS / 32];
  __shared__ float red_m[SK_THREADS / 32], red_s[SK_THREADS / 32];
  __shared__ float uold_s[SK_MAX_BAND];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int row0 = min(M, cta * rows_per_cta), row1 = min(M, row0 + rows_per_cta);
  const int nstage_total = (row1 - row0 + SK_ROWS - 1) / SK_ROWS;              // stages per iteration for this CTA
  const float norm = -logf((float)M + (float)N);
  const float log_mu_last = logf((float)N) + norm, log_nu_last = logf((float)M) + norm;
  const float c_mu = fmaxf(norm, log_mu_last) * LOG2E, c_nu = fmaxf(norm, log_nu_last) * LOG2E;   // log2 of the largest marginals
  const uint32_t row_bytes = (uint32_t)N * 4u;
  const int n4 = N >> 2;

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) sk_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int st_idx, int buf) {       // thread 0 only: load stage `st_idx` of this CTA's band into buffer `buf`
    const int r = row0 + st_idx * SK_ROWS;
    const int nr = min(SK_ROWS, row1 - r);
    sk_mbar_expect(&full[buf], nr * row_bytes);
    for (int k = 0; k < nr; ++k)
      sk_bulk_load(stage_buf + ((size_t)buf * SK_ROWS + k) * N, S + (size_t)(r + k) * N, row_bytes, &full[buf]);
  };
  // the band is re-streamed every iteration (S never changes): stages are numbered consecutively across iterations
  uint32_t total_seq = (uint32_t)iters * (uint32_t)nstage_total;
  uint32_t issued = 0, consumed = 0;
  if (tid == 0)
    while (issued < total_seq && issued < SK_STAGES) { issue(issued % nstage_total, issued % SK_STAGES); ++issued; }
  bool fast_ok = allow_fast != 0;

  SkCtx ctx;
  ctx.S = S; ctx.stage_buf = stage_buf; ctx.full = full; ctx.part_m = part_m; ctx.part_s = part_s; ctx.u = u; ctx.v = v;
  ctx.pm = pm; ctx.ps = ps; ctx.flag = flag; ctx.M = M; ctx.N = N; ctx.n4 = n4; ctx.row0 = row0; ctx.row1 = row1;
  ctx.uold_s = uold_s;
  ctx.nstage_total = nstage_total; ctx.cta = cta; ctx.norm = norm; ctx.c_mu = c_mu; ctx.c_nu = c_nu; ctx.row_bytes = row_bytes;

  for (int it = 0; it < iters; ++it) {
    const bool fast = fast_ok && it >= SK_EXACT_ITERS;
    ctx.extra_row = (alpha + __ldcg(v + N)) * LOG2E;                  // dustbin column term of every row LSE (old v)
    // dustbin row: u[M] = log_mu_last - LSE_j(alpha + v_j), j in [0, N]   (last CTA; small)
    if (cta == G - 1) {
      L2Acc a; a.init();
      for (int j = tid; j <= N; j += SK_THREADS) a.add((alpha + __ldcg(v + j)) * LOG2E);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
      if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
      __syncthreads();
      if (tid == 0) {
        L2Acc t; t.init();
        for (int w = 0; w < SK_THREADS / 32; ++w) t.merge(red_m[w], red_s[w]);
        u[M] = log_mu_last - t.lse_ln();
      }
      __syncthreads();
    }
    // Each thread owns the same SK_GROUPS float4 column groups in BOTH passes: a stage's elements are read from shared memory
    // once into registers, used for the row pass (old v, in registers for the whole iteration) and — after the block-wide row
    // reduction gives u — again for the column pass.
    for (int i = tid; i < row1 - row0; i += SK_THREADS) uold_s[i] = __ldcg(u + row0 + i) * LOG2E;
    __syncthreads();
    if (fast) sk_band<true>(ctx, consumed, issued, total_seq);
    else sk_band<false>(ctx, consumed, issued, total_seq);
    grid.sync();
    // ---- combine: v_j = log_nu_j - LSE_i(S_ij + u_i) incl. the dustbin row; v[N] from all u ----
    {
      const float extra_col = (alpha + __ldcg(u + M)) * LOG2E;
      // one warp per 2-column tile: lane = 2 * sub + col; each lane reduces the partials of CTAs c = sub, sub + 16, ...,
      // then a 4-step shuffle over `sub`.  No shared memory, no block barrier, ~all warps of the grid busy.
      const int sub = lane >> 1, cl = lane & 1;
      for (int tile = cta * (SK_THREADS / 32) + warp; tile * 2 < N; tile += G * (SK_THREADS / 32)) {
        const int j = tile * 2 + cl;
        if (fast) {
          float sm = 0.f;
          if (j < N) {
#pragma unroll 4
            for (int c = sub; c < G; c += 16) sm += __ldcg(ps + (size_t)c * N + j);
          }
#pragma unroll
          for (int o = 2; o < 32; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
          if (sub == 0 && j < N) {
            const float mj = c_mu - __ldcg(v + j) * LOG2E;            // the stabiliser used above (old v)
            sm += sk_ex2(extra_col - mj);
            if (!(sm > 0.f && sm < INFINITY)) atomicExch(flag, 1);
            v[j] = norm - (mj + __log2f(sm)) * LN2;
          }
        } else {
          L2Acc a; a.init();
          if (j < N) {
#pragma unroll 2
            for (int c = sub; c < G; c += 16) a.merge(__ldcg(pm + (size_t)c * N + j), __ldcg(ps + (size_t)c * N + j));
          }
#pragma unroll
          for (int o = 2; o < 32; o <<= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
          if (sub == 0 && j < N) {
            a.add(extra_col);
            v[j] = norm - a.lse_ln();
          }
        }
      }
      if (cta == G - 1) {
        L2Acc a; a.init();
        for (int i = tid; i <= M; i += SK_THREADS) a.add((alpha + __ldcg(u + i)) * LOG2E);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a.merge(__shfl_xor_sync(0xffffffffu, a.m, o), __shfl_xor_sync(0xffffffffu, a.s, o));
        if (lane == 0) { red_m[warp] = a.m; red_s[warp] = a.s; }
        __syncthreads();
        if (tid == 0) {
          L2Acc t; t.init();
          for (int w = 0; w < SK_THREADS / 32; ++w) t.merge(red_m[w], red_s[w]);
          v[N] = log_nu_last - t.lse_ln();
        }
      }
    }
    grid.sync();
    if (fast_ok && __ldcg(flag) != 0) {
      // the fast mode tripped (potential jump beyond the f32 exponent range): restart the whole solve in exact mode
      fast_ok = false;
      for (int j = cta * SK_THREADS + tid; j <= N; j += G * SK_THREADS) v[j] = 0.f;
      for (int i = cta * SK_THREADS + tid; i <= M; i += G * SK_THREADS) u[i] = 0.f;
      total_seq = consumed + (uint32_t)iters * (uint32_t)nstage_total;
      if (tid == 0)
        while (issued < total_seq && issued - consumed < SK_STAGES) { issue(issued % nstage_total, issued % SK_STAGES); ++issued; }
      it = -1;
      grid.sync();
    }
  }
}

static bool sinkhorn_fused_ok(int M, int N) { return N % 4 == 0 && N <= SK_MAXN && N >= 64 && M >= 1; }
static int g_sinkhorn_fast = 1;   // 1 = a-priori stabilisers after the first two iterations, 0 = running maxima throughout

// returns 0 when the fused kernel ran, 1 when the shape is unsupported (caller falls back to the two-pass kernels)
static int sinkhorn_fused_launch(const float* S, int M, int N, float alpha, int iters, float* u, float* v, AssignWs& w,
                                 cudaStream_t st) {
  if (!sinkhorn_fused_ok(M, N) || iters <= 0) return 1;
  static int coop = -1, sms = 0;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!coop || sms <= 0 || sms > 256) return 1;
  const size_t smem = (size_t)SK_STAGES * SK_ROWS * N * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(sinkhorn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SK_STAGES * SK_ROWS * SK_MAXN * (int)sizeof(float)) != cudaSuccess) { cudaGetLastError(); return 1; }
    attr_set = true;
  }
  int G = sms;
  int rpc = (M + G - 1) / G;
  rpc = (rpc + SK_ROWS - 1) / SK_ROWS * SK_ROWS;
  if (rpc > SK_MAX_BAND) return 1;
  G = (M + rpc - 1) / rpc;                      // CTAs that actually own rows (<= sms <= 256 partial slots)
  cudaMemsetAsync(u, 0, (size_t)(M + 1) * sizeof(float), st);
  cudaMemsetAsync(v, 0, (size_t)(N + 1) * sizeof(float), st);
  float* pm = w.pm; float* ps = w.ps;
  int* flag = w.pi;
  cudaMemsetAsync(flag, 0, sizeof(int), st);
  int allow_fast = g_sinkhorn_fast;
  void* args[] = {(void*)&S, (void*)&M, (void*)&N, (void*)&alpha, (void*)&iters, (void*)&u, (void*)&v, (void*)&pm, (void*)&ps,
                  (void*)&flag, (void*)&rpc, (void*)&allow_fast};
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
