// Five-point essential-matrix RANSAC on sm_100a: seeded batches of minimal samples solved one per thread (five_point.cuh,
// Nister's solver — the one behind cv2.findEssentialMat), every real solution scored by one warp counting Sampson inliers
// over all correspondences.
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d):
//   sfm/geometry.py:63-65   cv2.findEssentialMat(kpts0, kpts1, np.eye(3), threshold=norm_thresh, prob=conf, method=cv2.RANSAC)
// OpenCV's RANSAC keeps the minimal-sample model with the most inliers (error = squared Sampson distance in normalised
// coordinates, inlier if < threshold^2) and stops once  log(1 - prob) / log(1 - w^5)  samples have been drawn; so does this file.
// Ties are broken towards the smaller summed inlier error (OpenCV: first found).
//
//   e5_hypotheses_kernel  one thread per sample: 5 distinct seeded draws -> up to 10 essential matrices (f64)
//   e5_score_kernel       one warp per (sample, solution): inlier count and summed error over all correspondences
//   e5_update_kernel      one CTA: best model of the round vs the incumbent, adaptive stopping rule
#include "common.cuh"
#include "five_point.cuh"
#include "../../include/icepy4d_b200.h"

#define E5_BATCH 512          // samples per round
#define E5_MAX_ROUNDS 64

struct E5State {
  double bestE[9];
  double best_err;
  int best_inl;
  int tested;                 // samples drawn so far
  int done;
};

__device__ __forceinline__ unsigned int e5_hash(unsigned int a, unsigned int b, unsigned int c) {
  unsigned int x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

__global__ void e5_init_kernel(E5State* st) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) st->bestE[i] = 0.0;
    st->best_err = 1e300; st->best_inl = -1; st->tested = 0; st->done = 0;
  }
}

__global__ void __launch_bounds__(64) e5_hypotheses_kernel(const float* __restrict__ xn0, const float* __restrict__ xn1, int n,
                                                           unsigned int seed, int round, const E5State* st,
                                                           double* __restrict__ hypE, int* __restrict__ hyp_n) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= E5_BATCH) return;
  if (st->done) { hyp_n[h] = 0; return; }
  int idx[5];
  unsigned int ctr = 0;
  for (int j = 0; j < 5; ++j) {
    for (;;) {
      const int c = (int)(e5_hash(seed ^ 0x5bd1e995u, (unsigned)(round * E5_BATCH + h), ctr++) % (unsigned)n);
      bool dup = false;
      for (int t = 0; t < j; ++t) dup |= (idx[t] == c);
      if (!dup) { idx[j] = c; break; }
    }
  }
  double a[5][2], b[5][2];
  for (int j = 0; j < 5; ++j) {
    a[j][0] = xn0[2 * idx[j]]; a[j][1] = xn0[2 * idx[j] + 1];
    b[j][0] = xn1[2 * idx[j]]; b[j][1] = xn1[2 * idx[j] + 1];
  }
  double E[10][9];
  const int ns = fivept::solve(a, b, E);
  for (int k = 0; k < ns; ++k)
    for (int i = 0; i < 9; ++i) hypE[((size_t)h * 10 + k) * 9 + i] = E[k][i];
  hyp_n[h] = ns;
}

__global__ void __launch_bounds__(256) e5_score_kernel(const float* __restrict__ xn0, const float* __restrict__ xn1, int n,
                                                       double thr2, const double* __restrict__ hypE, const int* __restrict__ hyp_n,
                                                       const E5State* st, int* __restrict__ inl_out, double* __restrict__ err_out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= E5_BATCH * 10) return;
  const int h = w / 10, k = w - h * 10;
  if (st->done || k >= hyp_n[h]) { if (lane == 0) inl_out[w] = -1; return; }
  double E[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) E[i] = hypE[(size_t)w * 9 + i];
  int cnt = 0;
  double err = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float2 p = __ldg(reinterpret_cast<const float2*>(xn0) + i), q = __ldg(reinterpret_cast<const float2*>(xn1) + i);
    const double a = p.x, b = p.y, c = q.x, d = q.y;
    const double e0 = E[0] * a + E[1] * b + E[2], e1 = E[3] * a + E[4] * b + E[5], e2 = E[6] * a + E[7] * b + E[8];
    const double f0 = E[0] * c + E[3] * d + E[6], f1 = E[1] * c + E[4] * d + E[7];
    const double r = c * e0 + d * e1 + e2;
    const double den = e0 * e0 + e1 * e1 + f0 * f0 + f1 * f1;
    const double s = den > 0 ? r * r / den : 1e300;
    if (s < thr2) { ++cnt; err += s; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    err += __shfl_xor_sync(0xffffffffu, err, o);
  }
  if (lane == 0) { inl_out[w] = cnt; err_out[w] = err; }
}

__global__ void __launch_bounds__(1024) e5_update_kernel(const double* __restrict__ hypE, const int* __restrict__ inl,
                                                         const double* __restrict__ err, int n, double confidence, int max_samples,
                                                         E5State* st) {
  __shared__ int s_inl[32], s_idx[32];
  __shared__ double s_err[32];
  if (st->done) return;
  int bi = -1, bc = -1; double be = 1e300;
  for (int w = threadIdx.x; w < E5_BATCH * 10; w += blockDim.x) {
    const int c = inl[w];
    if (c < 0) continue;
    const double e = err[w];
    if (c > bc || (c == bc && (e < be || (e == be && w < bi)))) { bc = c; be = e; bi = w; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, bc, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const double oe = __shfl_xor_sync(0xffffffffu, be, o);
    if (oc > bc || (oc == bc && oi >= 0 && (oe < be || (oe == be && oi < bi)))) { bc = oc; be = oe; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_inl[threadIdx.x >> 5] = bc; s_idx[threadIdx.x >> 5] = bi; s_err[threadIdx.x >> 5] = be; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 32; ++k)
      if (s_inl[k] > bc || (s_inl[k] == bc && s_idx[k] >= 0 && (s_err[k] < be || (s_err[k] == be && s_idx[k] < bi)))) {
        bc = s_inl[k]; be = s_err[k]; bi = s_idx[k];
      }
    if (bi >= 0 && (bc > st->best_inl || (bc == st->best_inl && be < st->best_err))) {
      st->best_inl = bc; st->best_err = be;
      for (int i = 0; i < 9; ++i) st->bestE[i] = hypE[(size_t)bi * 9 + i];
    }
    st->tested += E5_BATCH;
    // OpenCV's RANSACUpdateNumIters rule with model size 5
    bool enough = false;
    if (st->best_inl >= 5) {
      const double wgt = (double)st->best_inl / (double)n;
      const double p5 = wgt * wgt * wgt * wgt * wgt;
      if (p5 >= 1.0 - 1e-12) enough = true;
      else if (p5 > 1e-12) enough = (double)st->tested >= log(1.0 - confidence) / log(1.0 - p5);
    }
    if (enough || st->tested >= max_samples) st->done = 1;
  }
}

__global__ void e5_finish_kernel(const E5State* st, double* __restrict__ E_out, int* __restrict__ n_inliers) {
  if (threadIdx.x == 0) {
    const bool ok = st->best_inl >= 5;
    for (int i = 0; i < 9; ++i) E_out[i] = ok ? st->bestE[i] : __longlong_as_double(0x7ff8000000000000LL);   // NaN = no model
    *n_inliers = ok ? st->best_inl : 0;
  }
}

extern "C" __attribute__((visibility("default"))) size_t i4d_essential_workspace_bytes(void) {
  return 256 + (size_t)E5_BATCH * 10 * (9 * sizeof(double) + sizeof(int) + sizeof(double)) + (size_t)E5_BATCH * sizeof(int) + 256;
}

extern "C" __attribute__((visibility("default"))) int i4d_essential_ransac(
    const float* xn0, const float* xn1, int n, double threshold_norm, double confidence, int max_iters, unsigned int seed,
    double* E_out, int* n_inliers, void* workspace, size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(xn0 && xn1 && E_out && n_inliers && workspace, "null pointer");
  I4D_CHECK_ARG(n >= 5, "need at least 5 correspondences");
  I4D_CHECK_ARG(threshold_norm > 0 && confidence > 0 && confidence < 1 && max_iters > 0, "bad parameters");
  if (workspace_bytes < i4d_essential_workspace_bytes()) {
    i4d_set_error("i4d_essential_ransac: workspace too small");
    return I4D_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(workspace);
  E5State* state = reinterpret_cast<E5State*>(w); w += 256;
  double* hypE = reinterpret_cast<double*>(w); w += (size_t)E5_BATCH * 10 * 9 * sizeof(double);
  double* err = reinterpret_cast<double*>(w); w += (size_t)E5_BATCH * 10 * sizeof(double);
  int* inl = reinterpret_cast<int*>(w); w += (size_t)E5_BATCH * 10 * sizeof(int);
  int* hyp_n = reinterpret_cast<int*>(w);
  int rounds = i4d_cdiv(max_iters, E5_BATCH);
  if (rounds > E5_MAX_ROUNDS) rounds = E5_MAX_ROUNDS;
  const int max_samples = rounds * E5_BATCH;
  e5_init_kernel<<<1, 32, 0, st>>>(state);
  for (int r = 0; r < rounds; ++r) {
    e5_hypotheses_kernel<<<E5_BATCH / 64, 64, 0, st>>>(xn0, xn1, n, seed, r, state, hypE, hyp_n);
    e5_score_kernel<<<E5_BATCH * 10 * 32 / 256, 256, 0, st>>>(xn0, xn1, n, threshold_norm * threshold_norm, hypE, hyp_n, state, inl, err);
    e5_update_kernel<<<1, 1024, 0, st>>>(hypE, inl, err, n, confidence, max_samples, state);
  }
  e5_finish_kernel<<<1, 32, 0, st>>>(state, E_out, n_inliers);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
