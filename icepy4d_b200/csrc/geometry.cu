// Two-view geometry kernels on sm_100a, all f64, one thread per correspondence (latency/ALU bound — tens of MB
// of traffic at most — so they are reported as points/s, not against a roofline).
// Compiled with -fmad=false so that the f64 arithmetic rounds like the host code it replaces.
//
// Reference behaviour replaced (paths into /root/reference/src/icepy4d):
//   sfm/geometry.py:103-118            undistort_points -> cv2.undistortPoints(pts, K, dist, None, K) (5 fixed-point
//                                      iterations of the Brown model, f64 inside, f32 out)
//   thirdparty/triangulation.py:79-177 iterative_LS_triangulation (Hartley-Sturm, <= 10 iterations, tolerance 3e-5)
//   sfm/triangulation.py:154-183       triangulate_points_linear / triangulate_nviews (6x6 SVD null vector)
//   sfm/interpolate_colors.py:14-87, sfm/geometry.py:78-100  interpolate_point_colors: cv2.projectPoints (Brown model, f64,
//                                      result cast to f32) + bilinear interpolation with clipped neighbours
#include "common.cuh"
#include "../../include/icepy4d_b200.h"

struct Cam9 { double K[9]; double d[5]; };
struct Proj2 { double P1[12]; double P2[12]; };

__global__ void __launch_bounds__(256) undistort_kernel(const float* __restrict__ pts, int n, Cam9 c,
                                                        float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
  const double ifx = 1.0 / fx, ify = 1.0 / fy;
  const double k1 = c.d[0], k2 = c.d[1], p1 = c.d[2], p2 = c.d[3], k3 = c.d[4];
  float2 p = reinterpret_cast<const float2*>(pts)[i];
  const double x0 = ((double)p.x - cx) * ifx, y0 = ((double)p.y - cy) * ify;
  double x = x0, y = y0;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    double r2 = x * x + y * y;
    double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
    if (icdist < 0) { x = x0; y = y0; break; }
    double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
    double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  // new camera matrix P = K (rows of K; K[1] = skew, normally 0)
  double xx = c.K[0] * x + c.K[1] * y + c.K[2];
  double yy = c.K[3] * x + c.K[4] * y + c.K[5];
  double ww = 1.0 / (c.K[6] * x + c.K[7] * y + c.K[8]);
  reinterpret_cast<float2*>(out)[i] = make_float2((float)(xx * ww), (float)(yy * ww));
}

// least-squares solve of a 4x3 system by Householder QR (A and b are overwritten)
__device__ __forceinline__ void lstsq_4x3(double A[4][3], double b[4], double x[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double nrm = 0.0;
#pragma unroll
    for (int i = k; i < 4; ++i) nrm += A[i][k] * A[i][k];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    double alpha = A[k][k] > 0 ? -nrm : nrm;
    double v[4];
#pragma unroll
    for (int i = k; i < 4; ++i) v[i] = A[i][k];
    v[k] -= alpha;
    double vn = 0.0;
#pragma unroll
    for (int i = k; i < 4; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    double inv = 2.0 / vn;
#pragma unroll
    for (int j = k; j < 3; ++j) {
      double dot = 0.0;
#pragma unroll
      for (int i = k; i < 4; ++i) dot += v[i] * A[i][j];
      dot *= inv;
#pragma unroll
      for (int i = k; i < 4; ++i) A[i][j] -= dot * v[i];
    }
    double dot = 0.0;
#pragma unroll
    for (int i = k; i < 4; ++i) dot += v[i] * b[i];
    dot *= inv;
#pragma unroll
    for (int i = k; i < 4; ++i) b[i] -= dot * v[i];
  }
  x[2] = b[2] / A[2][2];
  x[1] = (b[1] - A[1][2] * x[2]) / A[1][1];
  x[0] = (b[0] - A[0][1] * x[1] - A[0][2] * x[2]) / A[0][0];
}

__global__ void __launch_bounds__(128) tri_iterls_kernel(const float* __restrict__ u1, const float* __restrict__ u2,
                                                         int n, Proj2 pr, double tol, double* __restrict__ X,
                                                         int* __restrict__ status) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double* P1 = pr.P1; const double* P2 = pr.P2;
  float2 a = reinterpret_cast<const float2*>(u1)[idx], c = reinterpret_cast<const float2*>(u2)[idx];
  const double ux[4] = {(double)a.x, (double)a.y, (double)c.x, (double)c.y};
  // A0 = [u*P[2,:3] - P[row,:3]],  b0 = -(u*P[2,3] - P[row,3])        (triangulation.py:113-123)
  double A0[4][3], b0[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double* P = r < 2 ? P1 : P2;
    int row = r & 1;
#pragma unroll
    for (int j = 0; j < 3; ++j) A0[r][j] = ux[r] * P[8 + j] - P[4 * row + j];
    b0[r] = -(ux[r] * P[11] - P[4 * row + 3]);
  }
  double d1 = 1.0, d2 = 1.0, d1n = 1.0, d2n = 1.0;
  double x[3] = {0, 0, 0};
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    double A[4][3], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      b[r] = b0[r];
#pragma unroll
      for (int j = 0; j < 3; ++j) A[r][j] = A0[r][j];
    }
    lstsq_4x3(A, b, x);
    d1n = P1[8] * x[0] + P1[9] * x[1] + P1[10] * x[2] + P1[11];
    d2n = P2[8] * x[0] + P2[9] * x[1] + P2[10] * x[2] + P2[11];
    if (fabs(d1n - d1) <= tol && fabs(d2n - d2) <= tol) break;
    // the reference re-weights the *already weighted* system each iteration (triangulation.py:157-160)
    double w1 = 1.0 / d1n, w2 = 1.0 / d2n;
#pragma unroll
    for (int j = 0; j < 3; ++j) { A0[0][j] *= w1; A0[1][j] *= w1; A0[2][j] *= w2; A0[3][j] *= w2; }
    b0[0] *= w1; b0[1] *= w1; b0[2] *= w2; b0[3] *= w2;
    d1 = d1n; d2 = d2n;
  }
  X[3 * (size_t)idx + 0] = x[0]; X[3 * (size_t)idx + 1] = x[1]; X[3 * (size_t)idx + 2] = x[2];
  int st = (d1n > 0 && d2n > 0) ? 1 : 0;   // `i < 10` in the reference is always true (triangulation.py:169)
  if (d1n <= 0) st -= 1;
  if (d2n <= 0) st -= 2;
  status[idx] = st;
}

// DLT: right singular vector of the smallest singular value of the 6x6 matrix [[P1, -x1, 0], [P2, 0, -x2]],
// by one-sided (Hestenes) Jacobi directly on M — no squaring of the condition number.
__global__ void __launch_bounds__(64) tri_dlt_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int n,
                                                     Proj2 pr, double* __restrict__ X) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float2 a = reinterpret_cast<const float2*>(x1)[idx], c = reinterpret_cast<const float2*>(x2)[idx];
  double A[6][6], V[6][6];  // A[col][row]
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) { A[j][i] = 0.0; V[j][i] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) { A[j][i] = pr.P1[4 * i + j]; A[j][3 + i] = pr.P2[4 * i + j]; }
  A[4][0] = -(double)a.x; A[4][1] = -(double)a.y; A[4][2] = -1.0;
  A[5][3] = -(double)c.x; A[5][4] = -(double)c.y; A[5][5] = -1.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < 5; ++p) {
#pragma unroll
      for (int q = p + 1; q < 6; ++q) {
        double al = 0, be = 0, ga = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { al += A[p][i] * A[p][i]; be += A[q][i] * A[q][i]; ga += A[p][i] * A[q][i]; }
        if (fabs(ga) > 1e-15 * sqrt(al * be) && ga != 0.0) {
          rotated = true;
          double zeta = (be - al) / (2.0 * ga);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            double ap = A[p][i], aq = A[q][i];
            A[p][i] = cs * ap - sn * aq; A[q][i] = sn * ap + cs * aq;
            double vp = V[p][i], vq = V[q][i];
            V[p][i] = cs * vp - sn * vq; V[q][i] = sn * vp + cs * vq;
          }
        }
      }
    }
    if (!rotated) break;
  }
  int jmin = 0; double best = INFINITY;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += A[j][i] * A[j][i];
    if (s < best) { best = s; jmin = j; }
  }
  double v0 = 0, v1 = 0, v2 = 0, v3 = 1;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j == jmin) { v0 = V[j][0]; v1 = V[j][1]; v2 = V[j][2]; v3 = V[j][3]; }
  X[3 * (size_t)idx + 0] = v0 / v3; X[3 * (size_t)idx + 1] = v1 / v3; X[3 * (size_t)idx + 2] = v2 / v3;
}

struct CamRt { double R[9]; double t[3]; double K[9]; double d[5]; };

// sfm/interpolate_colors.py:14-87: project every 3-D point with the Brown model (the arithmetic of cv2.projectPoints, f64, result
// rounded to f32 like `project_points` does), then bilinear interpolation of image / 255 with the reference's clipped-neighbour
// weights (x1, y1 are clipped BEFORE the weights are formed, interpolate_colors.py:63-80).  One thread per point.
__global__ void __launch_bounds__(256) project_colors_kernel(const double* __restrict__ X, int n, CamRt c,
                                                             const unsigned char* __restrict__ img, int H, int W, int C,
                                                             int swap_rb, double* __restrict__ col, float* __restrict__ proj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double Xw = X[3 * (size_t)i], Yw = X[3 * (size_t)i + 1], Zw = X[3 * (size_t)i + 2];
  double x = Xw * c.R[0] + Yw * c.R[1] + Zw * c.R[2] + c.t[0];
  double y = Xw * c.R[3] + Yw * c.R[4] + Zw * c.R[5] + c.t[1];
  double z = Xw * c.R[6] + Yw * c.R[7] + Zw * c.R[8] + c.t[2];
  z = z != 0.0 ? 1.0 / z : 1.0;
  x *= z; y *= z;
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
  const double cdist = 1.0 + c.d[0] * r2 + c.d[1] * r4 + c.d[4] * r6;
  const double xd = x * cdist + c.d[2] * a1 + c.d[3] * a2;
  const double yd = y * cdist + c.d[2] * a3 + c.d[3] * a1;
  const float uf = (float)(xd * c.K[0] + c.K[2]), vf = (float)(yd * c.K[4] + c.K[5]);
  if (proj) { proj[2 * (size_t)i] = uf; proj[2 * (size_t)i + 1] = vf; }
  const double u = (double)uf, v = (double)vf;
  int x0 = (int)floor(u), y0 = (int)floor(v);
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
  y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
  const double wa = ((double)x1 - u) * ((double)y1 - v), wb = ((double)x1 - u) * (v - (double)y0);
  const double wc = (u - (double)x0) * ((double)y1 - v), wd = (u - (double)x0) * (v - (double)y0);
  for (int ch = 0; ch < C; ++ch) {
    const int sc = (swap_rb && C == 3) ? 2 - ch : ch;                   // cv2.cvtColor(image, COLOR_BGR2RGB)
    const double Ia = (double)((float)img[((size_t)y0 * W + x0) * C + sc] / 255.0f);
    const double Ib = (double)((float)img[((size_t)y1 * W + x0) * C + sc] / 255.0f);
    const double Ic = (double)((float)img[((size_t)y0 * W + x1) * C + sc] / 255.0f);
    const double Id = (double)((float)img[((size_t)y1 * W + x1) * C + sc] / 255.0f);
    col[(size_t)i * C + ch] = wa * Ia + wb * Ib + wc * Ic + wd * Id;
  }
}

extern "C" __attribute__((visibility("default"))) int i4d_interpolate_point_colors(
    const double* X, int n, const double* R_host, const double* t_host, const double* K_host, const double* dist_host, int n_dist,
    const unsigned char* image, int H, int W, int C, int convert_bgr2rgb, double* colors, float* projections, void* stream) {
  I4D_CHECK_ARG(X && R_host && t_host && K_host && image && colors && n >= 0, "null pointer");
  I4D_CHECK_ARG(H > 0 && W > 0 && C >= 1 && C <= 4, "image must be H x W x C with 1 <= C <= 4");
  I4D_CHECK_ARG(n_dist >= 0 && n_dist <= 5, "only the k1,k2,p1,p2[,k3] Brown model is supported");
  if (n == 0) return I4D_OK;
  CamRt c;
  for (int i = 0; i < 9; ++i) { c.R[i] = R_host[i]; c.K[i] = K_host[i]; }
  for (int i = 0; i < 3; ++i) c.t[i] = t_host[i];
  for (int i = 0; i < 5; ++i) c.d[i] = (dist_host && i < n_dist) ? dist_host[i] : 0.0;
  project_colors_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(X, n, c, image, H, W, C, convert_bgr2rgb, colors, projections);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_undistort_points(const float* pts, int n, const double* K_host, const double* dist_host,
                                    int n_dist, float* out, void* stream) {
  I4D_CHECK_ARG(pts && out && K_host && n >= 0, "null pointer");
  I4D_CHECK_ARG(n_dist >= 0 && n_dist <= 5, "only the k1,k2,p1,p2[,k3] Brown model is supported");
  if (n == 0) return I4D_OK;
  Cam9 c;
  for (int i = 0; i < 9; ++i) c.K[i] = K_host[i];
  for (int i = 0; i < 5; ++i) c.d[i] = (dist_host && i < n_dist) ? dist_host[i] : 0.0;
  undistort_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, c, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_triangulate_iterative_ls(const float* u1, const float* u2, int n, const double* P1_host,
                                            const double* P2_host, double tolerance, double* X, int* status,
                                            void* stream) {
  I4D_CHECK_ARG(u1 && u2 && X && status && P1_host && P2_host && n >= 0, "null pointer");
  if (n == 0) return I4D_OK;
  Proj2 pr;
  for (int i = 0; i < 12; ++i) { pr.P1[i] = P1_host[i]; pr.P2[i] = P2_host[i]; }
  tri_iterls_kernel<<<i4d_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(u1, u2, n, pr, tolerance, X, status);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

extern "C" __attribute__((visibility("default"))) int i4d_triangulate_dlt(const float* x1, const float* x2, int n, const double* P1_host,
                                   const double* P2_host, double* X, void* stream) {
  I4D_CHECK_ARG(x1 && x2 && X && P1_host && P2_host && n >= 0, "null pointer");
  if (n == 0) return I4D_OK;
  Proj2 pr;
  for (int i = 0; i < 12; ++i) { pr.P1[i] = P1_host[i]; pr.P2[i] = P2_host[i]; }
  tri_dlt_kernel<<<i4d_cdiv(n, 64), 64, 0, (cudaStream_t)stream>>>(x1, x2, n, pr, X);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// ---- absolute orientation (Helmert): sfm/absolute_orientation.py:141-154 + thirdparty/transformations.py:889-1020 ---------
// Moments of two corresponding point sets for the closed-form similarity (Horn's quaternion method as the reference's
// affine_matrix_from_points(shear=False, scale=True, usesvd=False) applies it): centroids, the 3x3 cross-covariance of the
// centred sets and their sums of squares.  One CTA, two passes, fixed-order tree reductions (bit-reproducible).
//   out[0..2] = mean(v0), out[3..5] = mean(v1), out[6..14] = sum_i c0_i c1_i^T (row-major: [a][b] = sum c0[a] c1[b]),
//   out[15] = sum |c0|^2, out[16] = sum |c1|^2
__global__ void __launch_bounds__(1024) helmert_moments_kernel(const double* __restrict__ v0, const double* __restrict__ v1, int n,
                                                               double* __restrict__ out) {
  __shared__ double red[32][11];
  __shared__ double mean[6];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto block_sum = [&](double (&acc)[11], int cnt) {
    for (int c = 0; c < cnt; ++c) {
      double v = acc[c];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (tid < cnt) {
      double t = 0;
      for (int w = 0; w < 32; ++w) t += red[w][tid];
      red[0][tid] = t;
    }
    __syncthreads();
  };
  double acc[11];
  for (int c = 0; c < 11; ++c) acc[c] = 0.0;
  for (int i = tid; i < n; i += blockDim.x)
    for (int c = 0; c < 3; ++c) { acc[c] += v0[3 * i + c]; acc[3 + c] += v1[3 * i + c]; }
  block_sum(acc, 6);
  if (tid < 6) { mean[tid] = red[0][tid] / n; out[tid] = mean[tid]; }
  __syncthreads();
  for (int c = 0; c < 11; ++c) acc[c] = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    double a[3], b[3];
    for (int c = 0; c < 3; ++c) { a[c] = v0[3 * i + c] - mean[c]; b[c] = v1[3 * i + c] - mean[3 + c]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) acc[3 * r + c] += a[r] * b[c];
    acc[9] += a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    acc[10] += b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  }
  block_sum(acc, 11);
  if (tid < 11) out[6 + tid] = red[0][tid];
}

extern "C" __attribute__((visibility("default"))) int i4d_helmert_moments(const double* v0, const double* v1, int n, double* out,
                                                                         void* stream) {
  I4D_CHECK_ARG(v0 && v1 && out && n >= 1, "null pointer or empty point set");
  helmert_moments_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(v0, v1, n, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}

// sfm/absolute_orientation.py:269-272: points_out = T @ [x; 1], dehomogenised.  T row-major 4x4 (host), f64 throughout.
struct Mat16 { double m[16]; };
__global__ void __launch_bounds__(256) apply_transform_kernel(const double* __restrict__ X, int n, Mat16 T, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
  double h[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) h[r] = ((T.m[4 * r] * x + T.m[4 * r + 1] * y) + T.m[4 * r + 2] * z) + T.m[4 * r + 3];
  out[3 * i] = h[0] / h[3]; out[3 * i + 1] = h[1] / h[3]; out[3 * i + 2] = h[2] / h[3];
}

extern "C" __attribute__((visibility("default"))) int i4d_apply_transform(const double* X, int n, const double* T_host, double* out,
                                                                         void* stream) {
  I4D_CHECK_ARG(T_host && n >= 0 && (n == 0 || (X && out)), "null pointer");
  if (n == 0) return I4D_OK;
  Mat16 T;
  for (int i = 0; i < 16; ++i) T.m[i] = T_host[i];
  apply_transform_kernel<<<i4d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(X, n, T, out);
  I4D_CUDA_LAUNCH_CHECK();
  return I4D_OK;
}
