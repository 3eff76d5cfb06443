// Fundamental-matrix robust estimation on sm_100a: batched-hypothesis RANSAC (8-point minimal solves, one thread
// per hypothesis; one warp per hypothesis for consensus scoring) with the MAGSAC++ marginalised (sigma-consensus++)
// quality function, followed by the MAGSAC++ iteratively re-weighted least-squares polisher over all correspondences run
// to its fixed point, all on the device without host syncs.  f64 throughout the solvers, -fmad=false.
//
// MAGSAC++ (Barath et al., CVPR 2020) as OpenCV's USAC implements it (opencv/modules/calib3d/src/usac/quality.cpp,
// not vendored in the reference: OpenCV is an unpinned wheel dependency, 4.13.0 in the build container): residuals are
// squared Sampson errors, noise scale marginalised over sigma in (0, sigma_max], DoF n = 4, k = 3.64 (0.99 quantile of chi_4):
//   weight(r^2) = Gamma((n-1)/2, r^2/(2 sigma_max^2)) - Gamma((n-1)/2, k^2/2)                          for r < k sigma_max, else 0
//   loss(r^2)   = 2^((n+1)/2)/sigma_max * [ sigma_max^2/2 * gamma((n+1)/2, x) + r^2/4 * weight(r^2) ],  x = r^2/(2 sigma_max^2)
//   quality     = sum over r < k sigma_max of (1 - loss / loss(k sigma_max))
// The incomplete gamma functions OpenCV tabulates have closed forms for n = 4 (erfc / exp), evaluated directly here.
// sigma_max is not an argument of cv2.findFundamentalMat: scripts/magsac_probe.py fits the cut-off k*sigma_max = 4.5 px to
// the models OpenCV 4.13 returns (their one-step IRLS displacement is minimal there for every threshold tried).
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d):
//   matching/geometric_verification.py:43-102   pydegensac.findFundamentalMatrix | cv2.findFundamentalMat(USAC_MAGSAC, 0.5, 0.999, 100000)
//   sfm/two_view_geometry.py:127-197            RelativeOrientation.estimate_F_matrix (same two branches)
// The inlier rule matches OpenCV's USAC: sqrt(Sampson error) < threshold on raw pixel coordinates.
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

#define RS_BATCH 4096        // hypotheses per round
#define RS_TOPK 8            // hypotheses re-scored on all points per round
#define RS_SUB 8192          // points in the pre-scoring subsample
#define RS_MAX_ROUNDS 24
#define RS_NPOL 4             // polisher starts (the objective has several local optima on near-degenerate scenes)

struct RansacState {
  double T0[3], T1[3];          // normalisation: x' = s * (x - cx), y' = s * (y - cy)  -> {s, cx, cy}
  double bestF[9];              // best model so far (pixel coordinates, unit Frobenius norm)
  double best_score;
  int best_inliers;
  int done;                     // termination flag (confidence reached)
  int hyp_tested;
  int topk[RS_TOPK];
  double topF[RS_NPOL][9];      // the RS_NPOL best fully-scored hypotheses so far (topF[0] == bestF), starts of the polisher
  double top_score[RS_NPOL];
  double polishF[RS_NPOL][9];
  double polishX[RS_NPOL][9];   // the polisher's last eigenvector (normalised coordinates): start of the next inverse iteration
  int have_x[RS_NPOL];
  double final_score[RS_NPOL + 1];   // quality of {bestF, polishF[0..]} under the selection rule of the polish mode
  int converged[RS_NPOL];       // polisher j reached its fixed point: its remaining polish work is skipped
  int polish_done;              // polish iterations executed (all starts)
  int failed;                   // no valid model (every minimal sample degenerate): F = NaN, mask = all zeros, like cv2's (None, zeros)
  int use_polished;
};
#define RS_POLISH_GRID_MAX 320

__device__ __forceinline__ unsigned int rs_hash(unsigned int a, unsigned int b, unsigned int c) {
  unsigned int x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

// squared Sampson error of correspondence (x0,y0)<->(x1,y1) under F (row-major, pixel coords)
__device__ __forceinline__ double sampson_sq(const double* F, double x0, double y0, double x1, double y1) {
  double a0 = F[0] * x0 + F[1] * y0 + F[2], a1 = F[3] * x0 + F[4] * y0 + F[5], a2 = F[6] * x0 + F[7] * y0 + F[8];
  double b0 = F[0] * x1 + F[3] * y1 + F[6], b1 = F[1] * x1 + F[4] * y1 + F[7];
  double e = x1 * a0 + y1 * a1 + a2;
  double den = a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1;
  return den > 0 ? e * e / den : 1e300;
}
__device__ __forceinline__ float sampson_sq_f(const float* F, float x0, float y0, float x1, float y1) {
  float a0 = F[0] * x0 + F[1] * y0 + F[2], a1 = F[3] * x0 + F[4] * y0 + F[5], a2 = F[6] * x0 + F[7] * y0 + F[8];
  float b0 = F[0] * x1 + F[3] * y1 + F[6], b1 = F[1] * x1 + F[4] * y1 + F[7];
  float e = x1 * a0 + y1 * a1 + a2;
  float den = a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1;
  return den > 0.f ? e * e / den : 3e38f;
}

// ---- MAGSAC++ quality / weight (n = 4 degrees of freedom) ----------------------------------------------------
// Gamma(3/2, x) = sqrt(pi)/2 erfc(sqrt x) + sqrt(x) exp(-x);  Gamma(5/2, x) = 3/2 Gamma(3/2, x) + x^(3/2) exp(-x)
__device__ __forceinline__ double upper_gamma_1p5(double x) {
  double sx = sqrt(x);
  return 0.886226925452758 * erfc(sx) + sx * exp(-x);
}
struct MagsacConst {
  double inv_2s2;      // 1 / (2 sigma_max^2)
  double cut2;         // (k sigma_max)^2
  double g_off;        // Gamma(3/2, k^2/2)
  double half_s2;      // sigma_max^2 / 2
  double inv_max_loss; // 1 / [ sigma_max^2/2 * gamma(5/2, k^2/2) ]   (the common factor 2^(5/2)/sigma_max cancels)
};
__host__ __device__ inline MagsacConst magsac_const(double sigma_max, double kq) {
  MagsacConst c;
  c.inv_2s2 = 1.0 / (2.0 * sigma_max * sigma_max);
  c.cut2 = kq * kq * sigma_max * sigma_max;
  const double xk = kq * kq * 0.5, sxk = sqrt(xk);
  c.g_off = 0.886226925452758 * erfc(sxk) + sxk * exp(-xk);
  c.half_s2 = 0.5 * sigma_max * sigma_max;
  const double up25 = 1.5 * c.g_off + xk * sxk * exp(-xk);
  c.inv_max_loss = 1.0 / (c.half_s2 * (1.329340388179137 - up25));
  return c;
}
// 1 - loss/max_loss for a squared residual below the cut-off
__device__ __forceinline__ double magsac_gain(const MagsacConst& c, double e) {
  const double x = e * c.inv_2s2, sx = sqrt(x), ex = exp(-x);
  const double up15 = 0.886226925452758 * erfc(sx) + sx * ex;
  const double lo25 = 1.329340388179137 - (1.5 * up15 + x * sx * ex);
  return 1.0 - (c.half_s2 * lo25 + 0.25 * e * (up15 - c.g_off)) * c.inv_max_loss;
}
__device__ __forceinline__ float magsac_gain_f(float inv_2s2, float g_off, float half_s2, float inv_max_loss, float e) {
  const float x = e * inv_2s2, sx = sqrtf(x), ex = __expf(-x);
  const float up15 = 0.8862269f * erfcf(sx) + sx * ex;
  const float lo25 = 1.3293404f - (1.5f * up15 + x * sx * ex);
  return 1.f - (half_s2 * lo25 + 0.25f * e * (up15 - g_off)) * inv_max_loss;
}

// ---- normalisation statistics (single CTA; N up to a few 100k) ------------------------------------------
__global__ void __launch_bounds__(1024) rs_norm_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int n,
                                                       RansacState* st) {
  __shared__ double red[4][32];
  __shared__ double mean[4];
  double s[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s[0] += x0[2 * i]; s[1] += x0[2 * i + 1]; s[2] += x1[2 * i]; s[3] += x1[2 * i + 1];
  }
  for (int pass = 0; pass < 2; ++pass) {
    for (int c = 0; c < 4; ++c) {
      double v = s[c];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0;
      for (int w = 0; w < 32; ++w) t += red[threadIdx.x][w];
      if (pass == 0) mean[threadIdx.x] = t / n;
      else if (threadIdx.x < 2) {
        double md = t / n;  // mean distance to the centroid
        double* T = threadIdx.x == 0 ? st->T0 : st->T1;
        T[0] = md > 0 ? 1.4142135623730951 / md : 1.0;
        T[1] = mean[2 * threadIdx.x]; T[2] = mean[2 * threadIdx.x + 1];
      }
    }
    __syncthreads();
    if (pass == 0) {
      s[0] = s[1] = s[2] = s[3] = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double dx = x0[2 * i] - mean[0], dy = x0[2 * i + 1] - mean[1];
        s[0] += sqrt(dx * dx + dy * dy);
        dx = x1[2 * i] - mean[2]; dy = x1[2 * i + 1] - mean[3];
        s[1] += sqrt(dx * dx + dy * dy);
      }
    }
  }
  if (threadIdx.x == 0) {
    st->best_score = -1.0; st->best_inliers = 0; st->done = 0; st->hyp_tested = 0;
    st->polish_done = 0; st->failed = 0; st->use_polished = 0;
    for (int j = 0; j < RS_NPOL; ++j) {
      st->converged[j] = 0; st->top_score[j] = -1.0; st->have_x[j] = 0;
      for (int i = 0; i < 9; ++i) st->topF[j][i] = 0.0;
    }
    for (int i = 0; i < 9; ++i) st->bestF[i] = 0.0;
  }
}

// ---- small dense solvers (one thread) -------------------------------------------------------------------
// rank-2 projection of a 3x3 matrix (row-major) by one-sided Jacobi SVD
__device__ void rs_rank2(double* F) {
  double A[3][3], V[3][3];  // columns
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) { A[j][i] = F[3 * i + j]; V[j][i] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rot = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) { al += A[p][i] * A[p][i]; be += A[q][i] * A[q][i]; ga += A[p][i] * A[q][i]; }
        if (ga != 0.0 && fabs(ga) > 1e-16 * sqrt(al * be)) {
          rot = true;
          double zeta = (be - al) / (2.0 * ga);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          for (int i = 0; i < 3; ++i) {
            double ap = A[p][i], aq = A[q][i];
            A[p][i] = c * ap - s * aq; A[q][i] = s * ap + c * aq;
            double vp = V[p][i], vq = V[q][i];
            V[p][i] = c * vp - s * vq; V[q][i] = s * vp + c * vq;
          }
        }
      }
    if (!rot) break;
  }
  int jm = 0; double best = 1e300;
  for (int j = 0; j < 3; ++j) {
    double s = A[j][0] * A[j][0] + A[j][1] * A[j][1] + A[j][2] * A[j][2];
    if (s < best) { best = s; jm = j; }
  }
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) {
      double v = 0;
      for (int j = 0; j < 3; ++j)
        if (j != jm) v += A[j][i] * V[j][k];   // sum_j (u_j sigma_j)_i (v_j)_k
      F[3 * i + k] = v;
    }
}

// F (normalised coords) -> pixel coords: T1^T F T0, then unit Frobenius norm
__device__ void rs_denormalise(const double* Fn, const double* T0, const double* T1, double* F) {
  // T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]
  double M[9];
  // M = Fn * T0
  for (int i = 0; i < 3; ++i) {
    M[3 * i + 0] = Fn[3 * i + 0] * T0[0];
    M[3 * i + 1] = Fn[3 * i + 1] * T0[0];
    M[3 * i + 2] = -Fn[3 * i + 0] * T0[0] * T0[1] - Fn[3 * i + 1] * T0[0] * T0[2] + Fn[3 * i + 2];
  }
  // F = T1^T * M
  for (int k = 0; k < 3; ++k) {
    F[0 + k] = T1[0] * M[0 + k];
    F[3 + k] = T1[0] * M[3 + k];
    F[6 + k] = -T1[0] * T1[1] * M[0 + k] - T1[0] * T1[2] * M[3 + k] + M[6 + k];
  }
  double nrm = 0;
  for (int i = 0; i < 9; ++i) nrm += F[i] * F[i];
  nrm = sqrt(nrm);
  if (nrm > 0) for (int i = 0; i < 9; ++i) F[i] /= nrm;
}

// ---- hypothesis generation: 8 distinct random correspondences -> null vector by complete-pivot elimination ----
__global__ void __launch_bounds__(128) rs_hypotheses_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                            int n, unsigned int seed, int round, RansacState* st,
                                                            float* __restrict__ hypF, int* __restrict__ hyp_ok) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= RS_BATCH) return;
  if (st->done) { hyp_ok[h] = 0; return; }
  const double s0 = st->T0[0], cx0 = st->T0[1], cy0 = st->T0[2], s1 = st->T1[0], cx1 = st->T1[1], cy1 = st->T1[2];
  int idx[8];
  unsigned int ctr = 0;
  for (int j = 0; j < 8; ++j) {
    for (;;) {
      int c = (int)(rs_hash(seed, (unsigned)(round * RS_BATCH + h), ctr++) % (unsigned)n);
      bool dup = false;
      for (int t = 0; t < j; ++t) dup |= (idx[t] == c);
      if (!dup) { idx[j] = c; break; }
    }
  }
  double A[8][9];
  for (int j = 0; j < 8; ++j) {
    double ax = s0 * ((double)x0[2 * idx[j]] - cx0), ay = s0 * ((double)x0[2 * idx[j] + 1] - cy0);
    double bx = s1 * ((double)x1[2 * idx[j]] - cx1), by = s1 * ((double)x1[2 * idx[j] + 1] - cy1);
    A[j][0] = bx * ax; A[j][1] = bx * ay; A[j][2] = bx; A[j][3] = by * ax; A[j][4] = by * ay; A[j][5] = by;
    A[j][6] = ax; A[j][7] = ay; A[j][8] = 1.0;
  }
  int perm[9];
  for (int c = 0; c < 9; ++c) perm[c] = c;
  bool ok = true;
  for (int k = 0; k < 8; ++k) {
    int pr = k, pc = k; double best = 0;
    for (int r = k; r < 8; ++r)
      for (int c = k; c < 9; ++c) {
        double v = fabs(A[r][c]);
        if (v > best) { best = v; pr = r; pc = c; }
      }
    if (best < 1e-12) { ok = false; break; }
    if (pr != k) for (int c = 0; c < 9; ++c) { double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
    if (pc != k) {
      for (int r = 0; r < 8; ++r) { double t = A[r][k]; A[r][k] = A[r][pc]; A[r][pc] = t; }
      int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
    }
    double inv = 1.0 / A[k][k];
    for (int r = k + 1; r < 8; ++r) {
      double f = A[r][k] * inv;
      if (f != 0.0) for (int c = k; c < 9; ++c) A[r][c] -= f * A[k][c];
    }
  }
  double Fn[9];
  if (ok) {
    double z[9];
    z[8] = 1.0;
    for (int k = 7; k >= 0; --k) {
      double s = A[k][8] * z[8];
      for (int c = k + 1; c < 8; ++c) s += A[k][c] * z[c];
      z[k] = -s / A[k][k];
    }
    for (int c = 0; c < 9; ++c) Fn[perm[c]] = z[c];
    rs_rank2(Fn);
    double F[9];
    rs_denormalise(Fn, st->T0, st->T1, F);
    for (int i = 0; i < 9; ++i) {
      if (!isfinite(F[i])) ok = false;
      hypF[h * 9 + i] = (float)F[i];
    }
  }
  hyp_ok[h] = ok ? 1 : 0;
}

// MAGSAC++ quality of every hypothesis on a strided subsample (f32: this pass only ranks hypotheses for the full scoring)
__global__ void __launch_bounds__(256) rs_prescore_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                          int n, const float* __restrict__ hypF,
                                                          const int* __restrict__ hyp_ok, MagsacConst mc,
                                                          const RansacState* st, float* __restrict__ score) {
  int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (h >= RS_BATCH) return;
  if (st->done || !hyp_ok[h]) { if (lane == 0) score[h] = -1.f; return; }
  float F[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) F[i] = __ldg(hypF + h * 9 + i);
  const int sub = min(n, RS_SUB);
  const long long stride = n / sub;
  const float c2 = (float)mc.cut2, i2s = (float)mc.inv_2s2, goff = (float)mc.g_off, hs2 = (float)mc.half_s2,
              iml = (float)mc.inv_max_loss;
  float acc = 0.f;
  for (int i = lane; i < sub; i += 32) {
    long long p = (long long)i * stride;
    float2 a = __ldg(reinterpret_cast<const float2*>(x0) + p), b = __ldg(reinterpret_cast<const float2*>(x1) + p);
    float e = sampson_sq_f(F, a.x, a.y, b.x, b.y);
    if (e < c2) acc += fmaxf(0.f, magsac_gain_f(i2s, goff, hs2, iml, e));
  }
  acc = warp_sum(acc);
  if (lane == 0) score[h] = acc;
}

__global__ void __launch_bounds__(1024) rs_topk_kernel(const float* __restrict__ score, RansacState* st) {
  __shared__ float sv[1024];
  __shared__ int si[1024];
  __shared__ int taken[RS_TOPK];
  if (st->done) return;
  for (int k = 0; k < RS_TOPK; ++k) {
    float best = -2.f; int bi = -1;
    for (int i = threadIdx.x; i < RS_BATCH; i += blockDim.x) {
      bool skip = false;
      for (int t = 0; t < k; ++t) skip |= (taken[t] == i);
      if (!skip && score[i] > best) { best = score[i]; bi = i; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (threadIdx.x < o && sv[threadIdx.x + o] > sv[threadIdx.x]) { sv[threadIdx.x] = sv[threadIdx.x + o]; si[threadIdx.x] = si[threadIdx.x + o]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) { taken[k] = si[0]; st->topk[k] = (sv[0] >= 0.f) ? si[0] : -1; }
    __syncthreads();
  }
}

// full scoring of the RS_TOPK candidates on all points: one CTA per candidate
__global__ void __launch_bounds__(512) rs_fullscore_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int n,
                                                           const float* __restrict__ hypF, MagsacConst mc, double thr2,
                                                           const RansacState* st, double* __restrict__ cand_score,
                                                           int* __restrict__ cand_inl) {
  __shared__ double rs[16];
  __shared__ int ri[16];
  const int k = blockIdx.x;
  if (st->done) return;
  const int h = st->topk[k];
  if (h < 0) { if (threadIdx.x == 0) { cand_score[k] = -1.0; cand_inl[k] = 0; } return; }
  double F[9];
  for (int i = 0; i < 9; ++i) F[i] = (double)hypF[h * 9 + i];
  double acc = 0; int inl = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 a = __ldg(reinterpret_cast<const float2*>(x0) + i), b = __ldg(reinterpret_cast<const float2*>(x1) + i);
    double e = sampson_sq(F, a.x, a.y, b.x, b.y);
    if (e < mc.cut2) acc += magsac_gain(mc, e);
    inl += e < thr2;
  }
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); inl += __shfl_xor_sync(0xffffffffu, inl, o); }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = acc; ri[threadIdx.x >> 5] = inl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; int ti = 0;
    for (int w = 0; w < 16; ++w) { t += rs[w]; ti += ri[w]; }
    cand_score[k] = t; cand_inl[k] = ti;
  }
}

__global__ void rs_update_kernel(const float* __restrict__ hypF, const double* __restrict__ cand_score,
                                 const int* __restrict__ cand_inl, int n, double confidence, int max_hyp,
                                 RansacState* st) {
  if (threadIdx.x != 0 || st->done) return;
  for (int k = 0; k < RS_TOPK; ++k) {
    if (st->topk[k] < 0) continue;
    if (cand_score[k] > st->best_score) {
      st->best_score = cand_score[k];
      st->best_inliers = cand_inl[k];
      for (int i = 0; i < 9; ++i) st->bestF[i] = (double)hypF[st->topk[k] * 9 + i];
    }
    // sorted insertion into the list of polisher starts
    int pos = RS_NPOL;
    while (pos > 0 && cand_score[k] > st->top_score[pos - 1]) --pos;
    if (pos < RS_NPOL) {
      for (int j = RS_NPOL - 1; j > pos; --j) {
        st->top_score[j] = st->top_score[j - 1];
        for (int i = 0; i < 9; ++i) st->topF[j][i] = st->topF[j - 1][i];
      }
      st->top_score[pos] = cand_score[k];
      for (int i = 0; i < 9; ++i) st->topF[pos][i] = (double)hypF[st->topk[k] * 9 + i];
    }
  }
  st->hyp_tested += RS_BATCH;
  // standard RANSAC stopping rule on the inlier ratio at the user threshold, 8-point samples
  double eps = (double)st->best_inliers / (double)n;
  double p8 = pow(eps, 8.0);
  bool enough = false;
  if (p8 > 1e-12) {
    double need = log(1.0 - confidence) / log(fmax(1.0 - p8, 1e-300));
    enough = (double)st->hyp_tested >= need;
  }
  if (enough || st->hyp_tested >= max_hyp) st->done = 1;
}

// ---- polisher -----------------------------------------------------------------------------------------------
// mode 0 (MAGSAC):     MAGSAC++ weights (see the header), iterated to the fixed point of the re-weighted normalised 8-point fit.
// mode 1 (LO-RANSAC):  unit weights on the inliers at the caller's threshold (the final least squares on inliers of
//                      pydegensac's LO-RANSAC, matching/geometric_verification.py:66-76), iterated until the set is stable.
// Per-CTA partial moment matrices go to a workspace and are summed in a fixed order (no atomics: bit-reproducible).
__global__ void __launch_bounds__(256) rs_polish_accum_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                              int n, MagsacConst mc, double thr2, int mode,
                                                              const RansacState* st, double* __restrict__ partial) {
  __shared__ double red[8][45];
  const int j = blockIdx.y;
  if (st->converged[j] || st->failed) return;
  double F[9];
  for (int i = 0; i < 9; ++i) F[i] = st->polishF[j][i];
  const double s0 = st->T0[0], cx0 = st->T0[1], cy0 = st->T0[2], s1 = st->T1[0], cx1 = st->T1[1], cy1 = st->T1[2];
  double acc[45];
#pragma unroll
  for (int i = 0; i < 45; ++i) acc[i] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float2 a = __ldg(reinterpret_cast<const float2*>(x0) + i), b = __ldg(reinterpret_cast<const float2*>(x1) + i);
    double e = sampson_sq(F, a.x, a.y, b.x, b.y);
    double w;
    if (mode == 0) {
      if (e >= mc.cut2) continue;
      w = upper_gamma_1p5(e * mc.inv_2s2) - mc.g_off;
      if (w <= 0) continue;
    } else {
      if (e >= thr2) continue;
      w = 1.0;
    }
    double ax = s0 * ((double)a.x - cx0), ay = s0 * ((double)a.y - cy0);
    double bx = s1 * ((double)b.x - cx1), by = s1 * ((double)b.y - cy1);
    double v[9] = {bx * ax, bx * ay, bx, by * ax, by * ay, by, ax, ay, 1.0};
    int t = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
      for (int q = p; q < 9; ++q) acc[t++] += w * v[p] * v[q];
  }
#pragma unroll
  for (int i = 0; i < 45; ++i) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 45) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[((size_t)j * gridDim.x + blockIdx.x) * 45 + threadIdx.x] = t;
  }
}

// Smallest eigenvector of the 9x9 moment matrix, rank-2 projection, denormalisation, convergence test.
// The matrix is symmetric positive semi-definite with one eigenvalue far below the rest (the epipolar constraint), so inverse
// iteration on C + eps*I (one Cholesky factorisation, two triangular solves per step, f64) converges in a handful of steps.
__global__ void __launch_bounds__(64) rs_polish_solve_kernel(RansacState* st, const double* __restrict__ partial, int n_partial,
                                                             double tol) {
  __shared__ double cov[45];
  const int j = blockIdx.x;
  if (st->converged[j] || st->failed) return;
  if (threadIdx.x < 45) {
    double t = 0;
    for (int b = 0; b < n_partial; ++b) t += partial[((size_t)j * n_partial + b) * 45 + threadIdx.x];
    cov[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // every loop below has compile-time bounds and is fully unrolled: L, x, y live in registers (no local-memory round trips
  // on the dependent chains of this single-thread solve)
  double L[9][9];
  double tr = 0;
  {
    int t = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
      for (int q = p; q < 9; ++q) { L[q][p] = cov[t]; ++t; }
#pragma unroll
    for (int p = 0; p < 9; ++p) tr += L[p][p];
  }
  if (!(tr > 0)) { st->converged[j] = 1; return; }   // no support: keep the previous model
  const double eps = 1e-13 * tr;
  // Cholesky of C + eps I (lower triangle, in place); reciprocals of the pivots are kept: the 18 divisions per inverse-iteration
  // step become multiplications (the iteration is self-correcting, the fixed point is the same)
  double Linv[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double d = L[j][j] + eps;
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0)) d = eps;            // rank-deficient input: regularise the pivot
    d = sqrt(d);
    L[j][j] = d;
    const double inv = 1.0 / d;
    Linv[j] = inv;
#pragma unroll
    for (int i = j + 1; i < 9; ++i) {
      double v = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  double x[9], y[9];
  // start: the eigenvector of the previous polish iteration (the re-weighted moment matrix changes little from one IRLS step to
  // the next, so 2-3 inverse-iteration steps reach the same 1e-14 agreement that takes ~16 from a generic start), else a generic
  // vector that is not orthogonal to the null direction
  const bool warm = st->have_x[j] != 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] = warm ? st->polishX[j][i] : 1.0 / 3.0 + 0.01 * i;
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {                                  // L y = x
      double v = x[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
      y[i] = v * Linv[i];
    }
#pragma unroll
    for (int i = 8; i >= 0; --i) {                                 // L^T z = y (z overwrites y)
      double v = y[i];
#pragma unroll
      for (int k = i + 1; k < 9; ++k) v -= L[k][i] * y[k];
      y[i] = v * Linv[i];
    }
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) nrm += y[i] * y[i];
    nrm = 1.0 / sqrt(nrm);
    double diff = 0, diffn = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double z = y[i] * nrm;
      diff += (z - x[i]) * (z - x[i]);
      diffn += (z + x[i]) * (z + x[i]);
      x[i] = z;
    }
    if (it > 0 && fmin(diff, diffn) < 1e-28) break;               // converged up to sign
  }
  double Fn[9];
  for (int i = 0; i < 9; ++i) { Fn[i] = x[i]; st->polishX[j][i] = x[i]; }
  st->have_x[j] = 1;
  rs_rank2(Fn);
  double F[9];
  rs_denormalise(Fn, st->T0, st->T1, F);
  bool ok = true;
  for (int i = 0; i < 9; ++i) ok &= isfinite(F[i]);
  if (!ok) { st->converged[j] = 1; return; }
  // fixed point reached?  (unit Frobenius norm on both sides, sign-aligned)
  double dp = 0, dm = 0;
  for (int i = 0; i < 9; ++i) {
    dp = fmax(dp, fabs(F[i] - st->polishF[j][i]));
    dm = fmax(dm, fabs(F[i] + st->polishF[j][i]));
  }
  for (int i = 0; i < 9; ++i) st->polishF[j][i] = F[i];
  atomicAdd(&st->polish_done, 1);
  if (fmin(dp, dm) < tol) st->converged[j] = 1;
}

__global__ void rs_begin_polish_kernel(RansacState* st) {
  if (threadIdx.x == 0) {
    for (int j = 0; j < RS_NPOL; ++j) {
      const bool have = st->top_score[j] >= 0.0;            // fewer valid hypotheses than starts: duplicate the best, already "converged"
      for (int i = 0; i < 9; ++i) st->polishF[j][i] = have ? st->topF[j][i] : st->bestF[i];
      if (!have && j > 0) st->converged[j] = 1;
    }
    if (st->best_inliers < 8) st->failed = 1;
  }
}

// Quality of {bestF, polishF} on all points under the selection rule of the mode (MAGSAC++ quality / inlier count at the
// threshold): the polished model is kept only if it is at least as good as the RANSAC winner it started from.
__global__ void __launch_bounds__(512) rs_final_score_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int n,
                                                             MagsacConst mc, double thr2, int mode, RansacState* st) {
  __shared__ double rs[16];
  if (st->failed) return;
  const double* Fc = blockIdx.x == 0 ? st->bestF : st->polishF[blockIdx.x - 1];
  double F[9];
  for (int i = 0; i < 9; ++i) F[i] = Fc[i];
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 a = __ldg(reinterpret_cast<const float2*>(x0) + i), b = __ldg(reinterpret_cast<const float2*>(x1) + i);
    double e = sampson_sq(F, a.x, a.y, b.x, b.y);
    if (mode == 0) { if (e < mc.cut2) acc += magsac_gain(mc, e); }
    else acc += e < thr2 ? 1.0 : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 16; ++w) t += rs[w];
    st->final_score[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) rs_mask_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int n,
                                                      double thr2, const RansacState* st, unsigned char* __restrict__ mask,
                                                      int* __restrict__ n_inl, double* __restrict__ F_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool failed = st->failed != 0;
  const double* Fc = st->bestF;           // a polished model replaces the RANSAC winner only if it is at least as good
  double fs = st->final_score[0];
  for (int j = 0; j < RS_NPOL; ++j)
    if (st->final_score[j + 1] >= fs) { fs = st->final_score[j + 1]; Fc = st->polishF[j]; }
  double F[9];
  for (int k = 0; k < 9; ++k) F[k] = Fc[k];
  if (i == 0) {
    // scale like OpenCV (F[2][2] = 1) when possible; NaN marks "no model" (the caller returns F = None; cv2 returns (None, zeros))
    double s = fabs(F[8]) > 1e-300 ? 1.0 / F[8] : 1.0;
    for (int k = 0; k < 9; ++k) F_out[k] = failed ? __longlong_as_double(0x7ff8000000000000LL) : F[k] * s;
  }
  int inl = 0;
  if (i < n) {
    float2 a = __ldg(reinterpret_cast<const float2*>(x0) + i), b = __ldg(reinterpret_cast<const float2*>(x1) + i);
    inl = failed ? 0 : (sampson_sq(F, a.x, a.y, b.x, b.y) < thr2);
    mask[i] = (unsigned char)inl;
  }
  unsigned int bal = __ballot_sync(0xffffffffu, inl);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_inl, __popc(bal));
}

extern "C" __attribute__((visibility("default"))) size_t i4d_fundamental_workspace_bytes(void) {
  return sizeof(RansacState) + (size_t)RS_BATCH * (9 * sizeof(float) + sizeof(int) + sizeof(float)) +
         RS_TOPK * (sizeof(double) + sizeof(int)) + (size_t)RS_NPOL * RS_POLISH_GRID_MAX * 45 * sizeof(double) + 512;
}

extern "C" __attribute__((visibility("default"))) int i4d_fundamental_ransac(
    const float* x0, const float* x1, int n, double threshold, double confidence, int max_iters, unsigned int seed,
    double sigma_max, int polish_iters, int polish_mode, double* F_out, unsigned char* mask, int* n_inliers, void* workspace,
    size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(x0 && x1 && F_out && mask && n_inliers && workspace, "null pointer");
  I4D_CHECK_ARG(n >= 8, "need at least 8 correspondences");
  I4D_CHECK_ARG(threshold > 0 && confidence > 0 && confidence < 1 && max_iters > 0 && sigma_max > 0, "bad parameters");
  I4D_CHECK_ARG(polish_mode == 0 || polish_mode == 1, "polish_mode must be 0 (MAGSAC++) or 1 (LO-RANSAC)");
  if (workspace_bytes < i4d_fundamental_workspace_bytes()) {
    i4d_set_error("i4d_fundamental_ransac: workspace too small");
    return I4D_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = reinterpret_cast<char*>(workspace);
  RansacState* state = reinterpret_cast<RansacState*>(w); w += (sizeof(RansacState) + 15) & ~(size_t)15;
  float* hypF = reinterpret_cast<float*>(w); w += (size_t)RS_BATCH * 9 * sizeof(float);
  int* hyp_ok = reinterpret_cast<int*>(w); w += (size_t)RS_BATCH * sizeof(int);
  float* score = reinterpret_cast<float*>(w); w += (size_t)RS_BATCH * sizeof(float);
  double* cand_score = reinterpret_cast<double*>(w); w += RS_TOPK * sizeof(double);
  int* cand_inl = reinterpret_cast<int*>(w); w += RS_TOPK * sizeof(int);
  w = reinterpret_cast<char*>(((uintptr_t)w + 15) & ~(uintptr_t)15);
  double* partial = reinterpret_cast<double*>(w);
  const double kq = 3.64;  // 0.99 quantile of the chi distribution with 4 degrees of freedom
  const MagsacConst mc = magsac_const(sigma_max, kq);
  const double thr2 = threshold * threshold;
  I4D_CUDA_CALL(cudaMemsetAsync(n_inliers, 0, sizeof(int), st));
  rs_norm_kernel<<<1, 1024, 0, st>>>(x0, x1, n, state);
  int rounds = i4d_cdiv(max_iters, RS_BATCH);
  if (rounds > RS_MAX_ROUNDS) rounds = RS_MAX_ROUNDS;
  const int max_hyp = rounds * RS_BATCH;
  for (int r = 0; r < rounds; ++r) {
    rs_hypotheses_kernel<<<RS_BATCH / 128, 128, 0, st>>>(x0, x1, n, seed, r, state, hypF, hyp_ok);
    rs_prescore_kernel<<<RS_BATCH * 32 / 256, 256, 0, st>>>(x0, x1, n, hypF, hyp_ok, mc, state, score);
    rs_topk_kernel<<<1, 1024, 0, st>>>(score, state);
    rs_fullscore_kernel<<<RS_TOPK, 512, 0, st>>>(x0, x1, n, hypF, mc, thr2, state, cand_score, cand_inl);
    rs_update_kernel<<<1, 32, 0, st>>>(hypF, cand_score, cand_inl, n, confidence, max_hyp, state);
  }
  rs_begin_polish_kernel<<<1, 32, 0, st>>>(state);
  int grid = i4d_num_sms();                       // x RS_NPOL starts in grid.y
  if (grid > RS_POLISH_GRID_MAX) grid = RS_POLISH_GRID_MAX;
  if (grid > i4d_cdiv(n, 256)) grid = i4d_cdiv(n, 256);
  for (int it = 0; it < polish_iters; ++it) {
    rs_polish_accum_kernel<<<dim3(grid, RS_NPOL), 256, 0, st>>>(x0, x1, n, mc, thr2, polish_mode, state, partial);
    rs_polish_solve_kernel<<<RS_NPOL, 64, 0, st>>>(state, partial, grid, 1e-13);
  }
  rs_final_score_kernel<<<RS_NPOL + 1, 512, 0, st>>>(x0, x1, n, mc, thr2, polish_mode, state);
  rs_mask_kernel<<<i4d_cdiv(n, 256), 256, 0, st>>>(x0, x1, n, thr2, state, mask, n_inliers, F_out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
