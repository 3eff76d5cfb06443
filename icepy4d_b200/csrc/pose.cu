// Relative pose from an essential matrix on sm_100a: projection onto the essential manifold, the four (R, t) candidates and the
// cheirality vote — the device part of `estimate_pose`.
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d):
//   sfm/geometry.py:31-76         estimate_pose: cv2.findEssentialMat(RANSAC) on K-normalised points + cv2.recoverPose
//   sfm/two_view_geometry.py:52-109  RelativeOrientation.estimate_pose (camera update happens on the host)
// The robust estimate itself comes from the batched-hypothesis RANSAC of ransac.cu run on focal-scaled normalised
// coordinates (host side, icepy4d_b200/sfm/geometry.py); this file does what OpenCV's recoverPose and the inlier rule
// of findEssentialMat do: data-parallel over the correspondences, one warp-reduced counter per candidate.
//
//   pose_decompose_kernel   (1 thread)   E -> U diag(1,1,0) V^T (Jacobi eigen-decomposition of E^T E in f64), R1 = U W V^T,
//                                        R2 = U W^T V^T, t = u3 (unit), det(R) = +1
//   pose_vote_kernel        (1 thread per correspondence) Sampson inlier test (err < thr^2, OpenCV's rule) and, for inliers, the
//                                        depths of the point in both cameras for the four candidates -> 4-bit code + counters
//   pose_select_kernel      (1 thread)   best candidate in OpenCV's order (R1,t), (R2,t), (R1,-t), (R2,-t)
//   pose_mask_kernel        (1 thread per correspondence) mask = inlier AND in front of both cameras for the chosen pose (recoverPose
//                                        updates the mask it is given, geometry.py:69)
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

struct PoseWs {
  double E[9];        // projected essential matrix
  double R[2][9];     // R1, R2
  double t[3];        // unit translation (u3)
  int counts[4];      // cheirality votes per candidate
  int n_inl;          // Sampson inliers
  int best;           // chosen candidate
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// cyclic Jacobi on a symmetric 3x3 matrix (f64): A -> eigenvalues on the diagonal, V = eigenvectors (columns)
__device__ void jacobi3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

__global__ void pose_decompose_kernel(const double* __restrict__ E_in, PoseWs* ws) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double E[3][3], A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) E[i][j] = E_in[i * 3 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = E[0][i] * E[0][j] + E[1][i] * E[1][j] + E[2][i] * E[2][j];   // E^T E
  jacobi3(A, V);
  // the two largest eigenvalues
  int o[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (A[o[b]][o[b]] > A[o[a]][o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
  double v[3][3], u[3][3];                      // rows = vectors v1, v2, v3 / u1, u2, u3
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 3; ++i) v[k][i] = V[i][o[k]];
  cross3(v[0], v[1], v[2]);                     // det(V) = +1
  for (int k = 0; k < 2; ++k) {
    double n = 0;
    for (int i = 0; i < 3; ++i) { u[k][i] = E[i][0] * v[k][0] + E[i][1] * v[k][1] + E[i][2] * v[k][2]; n += u[k][i] * u[k][i]; }
    n = sqrt(n);
    for (int i = 0; i < 3; ++i) u[k][i] /= (n > 0 ? n : 1.0);
  }
  // Gram-Schmidt on u2 (the two singular values of a noisy estimate differ slightly)
  double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
  double n2 = 0;
  for (int i = 0; i < 3; ++i) { u[1][i] -= d * u[0][i]; n2 += u[1][i] * u[1][i]; }
  n2 = sqrt(n2);
  for (int i = 0; i < 3; ++i) u[1][i] /= (n2 > 0 ? n2 : 1.0);
  cross3(u[0], u[1], u[2]);                     // det(U) = +1
  // E = u1 v1^T + u2 v2^T;  R1 = U W V^T, R2 = U W^T V^T with W = [[0,-1,0],[1,0,0],[0,0,1]]
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      ws->E[i * 3 + j] = u[0][i] * v[0][j] + u[1][i] * v[1][j];
      ws->R[0][i * 3 + j] = -u[0][i] * v[1][j] + u[1][i] * v[0][j] + u[2][i] * v[2][j];
      ws->R[1][i * 3 + j] = u[0][i] * v[1][j] - u[1][i] * v[0][j] + u[2][i] * v[2][j];
    }
  for (int i = 0; i < 3; ++i) ws->t[i] = u[2][i];
  for (int c = 0; c < 4; ++c) ws->counts[c] = 0;
  ws->n_inl = 0;
  ws->best = 0;
}

// depths (lambda0, lambda1) of the least-squares intersection  lambda1 * x1 = R (lambda0 * x0) + t
__device__ __forceinline__ bool in_front(const double* R, const double* t, double sgn, double a, double b, double c, double d,
                                         double dist_thresh) {
  const double rx[3] = {R[0] * a + R[1] * b + R[2], R[3] * a + R[4] * b + R[5], R[6] * a + R[7] * b + R[8]};   // R x0
  const double x1[3] = {c, d, 1.0};
  const double tt[3] = {sgn * t[0], sgn * t[1], sgn * t[2]};
  // [rx  -x1] [l0 l1]^T = -t  -> normal equations
  const double a11 = rx[0] * rx[0] + rx[1] * rx[1] + rx[2] * rx[2];
  const double a12 = -(rx[0] * x1[0] + rx[1] * x1[1] + rx[2] * x1[2]);
  const double a22 = x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];
  const double b1 = -(rx[0] * tt[0] + rx[1] * tt[1] + rx[2] * tt[2]);
  const double b2 = x1[0] * tt[0] + x1[1] * tt[1] + x1[2] * tt[2];
  const double det = a11 * a22 - a12 * a12;
  if (!(fabs(det) > 1e-300)) return false;
  const double l0 = (b1 * a22 - a12 * b2) / det, l1 = (a11 * b2 - a12 * b1) / det;
  return l0 > 0 && l1 > 0 && l0 < dist_thresh && l1 < dist_thresh;     // z in camera 0 = l0, z in camera 1 = l1
}

__global__ void __launch_bounds__(256) pose_vote_kernel(const float* __restrict__ xn0, const float* __restrict__ xn1, int n, double thr2,
                                                        double dist_thresh, PoseWs* ws, unsigned char* __restrict__ code) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int bits = 0;
  if (i < n) {
    const double a = xn0[2 * i], b = xn0[2 * i + 1], c = xn1[2 * i], d = xn1[2 * i + 1];
    const double* E = ws->E;
    const double e0 = E[0] * a + E[1] * b + E[2], e1 = E[3] * a + E[4] * b + E[5], e2 = E[6] * a + E[7] * b + E[8];   // E x0
    const double f0 = E[0] * c + E[3] * d + E[6], f1 = E[1] * c + E[4] * d + E[7];                                     // E^T x1
    const double r = c * e0 + d * e1 + e2;
    const double den = e0 * e0 + e1 * e1 + f0 * f0 + f1 * f1;
    if (den > 0 && r * r / den < thr2) {
      bits = 16;
      if (in_front(ws->R[0], ws->t, 1.0, a, b, c, d, dist_thresh)) bits |= 1;
      if (in_front(ws->R[1], ws->t, 1.0, a, b, c, d, dist_thresh)) bits |= 2;
      if (in_front(ws->R[0], ws->t, -1.0, a, b, c, d, dist_thresh)) bits |= 4;
      if (in_front(ws->R[1], ws->t, -1.0, a, b, c, d, dist_thresh)) bits |= 8;
    }
    code[i] = (unsigned char)bits;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, (bits >> k) & 1);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(k < 4 ? &ws->counts[k] : &ws->n_inl, __popc(m));
  }
}

__global__ void pose_select_kernel(PoseWs* ws, double* E_out, double* R_out, double* t_out, int* n_good) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = 0;
  for (int c = 1; c < 4; ++c)
    if (ws->counts[c] > ws->counts[best]) best = c;           // first maximum in OpenCV's candidate order
  ws->best = best;
  const double sgn = best >= 2 ? -1.0 : 1.0;
  for (int i = 0; i < 9; ++i) { R_out[i] = ws->R[best & 1][i]; E_out[i] = ws->E[i]; }
  for (int i = 0; i < 3; ++i) t_out[i] = sgn * ws->t[i];
  n_good[0] = ws->counts[best];
  n_good[1] = ws->n_inl;
}

__global__ void __launch_bounds__(256) pose_mask_kernel(const unsigned char* __restrict__ code, int n, const PoseWs* ws,
                                                        unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = (unsigned char)((code[i] >> ws->best) & 1);
}

extern "C" __attribute__((visibility("default"))) size_t i4d_pose_workspace_bytes(int n) {
  return sizeof(PoseWs) + 64 + (size_t)(n > 0 ? n : 0);
}

extern "C" __attribute__((visibility("default"))) int i4d_essential_pose(const double* E_in, const float* xn0, const float* xn1, int n,
                                                                       double threshold_norm, double distance_threshold,
                                                                       double* E_out, double* R_out, double* t_out,
                                                                       unsigned char* mask, int* n_good, void* workspace,
                                                                       size_t workspace_bytes, void* stream) {
  I4D_CHECK_ARG(E_in && xn0 && xn1 && E_out && R_out && t_out && mask && n_good && workspace, "null pointer");
  I4D_CHECK_ARG(n >= 5, "at least 5 correspondences are needed");
  I4D_CHECK_ARG(threshold_norm > 0 && distance_threshold > 0, "thresholds must be positive");
  if (workspace_bytes < i4d_pose_workspace_bytes(n)) {
    i4d_set_error("i4d_essential_pose: workspace too small (%zu < %zu bytes)", workspace_bytes, i4d_pose_workspace_bytes(n));
    return I4D_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PoseWs* ws = reinterpret_cast<PoseWs*>(workspace);
  unsigned char* code = reinterpret_cast<unsigned char*>(workspace) + ((sizeof(PoseWs) + 63) / 64) * 64;
  pose_decompose_kernel<<<1, 32, 0, st>>>(E_in, ws);
  pose_vote_kernel<<<i4d_cdiv(n, 256), 256, 0, st>>>(xn0, xn1, n, threshold_norm * threshold_norm, distance_threshold, ws, code);
  pose_select_kernel<<<1, 32, 0, st>>>(ws, E_out, R_out, t_out, n_good);
  pose_mask_kernel<<<i4d_cdiv(n, 256), 256, 0, st>>>(code, n, ws, mask);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
