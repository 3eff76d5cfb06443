// Minimal (five-point) essential-matrix solver, one sample per thread, f64.  Plain C++ (host + device) so that the same code is
// unit-tested on the CPU (tests/test_five_point_host.py builds it with g++).
//
// Replaces the minimal solver behind cv2.findEssentialMat(..., method=cv2.RANSAC) called by the reference at
// /root/reference/src/icepy4d/sfm/geometry.py:63-65.  The algorithm is Nister's ("An efficient solution to the five-point
// relative pose problem", PAMI 2004), the same formulation OpenCV uses:
//   1. the 4-dimensional right null space {X, Y, Z, W} of the 5 x 9 epipolar constraint matrix,  E = xX + yY + zZ + W;
//   2. the ten cubic constraints det(E) = 0 and 2 E E^T E - trace(E E^T) E = 0 as a 10 x 20 coefficient matrix over the
//      monomials of (x, y, z), Gauss-Jordan elimination of the ten leading monomials;
//   3. three relations free of x^2, y^2, xy -> a 3 x 3 matrix B(z) of polynomials in z with B(z) [x y 1]^T = 0,
//      det B(z) = 0 is a degree-10 polynomial; its real roots (Durand-Kerner + Newton polish) give z, then x, y from B(z).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define FP_HD __host__ __device__
#else
#define FP_HD
#endif

namespace fivept {

// monomial order (Nister): x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x | yz2 yz y | z3 z2 z 1
FP_HD inline int mono_index(int a, int b, int c) {
  // exponents (a, b, c) of (x, y, z), a + b + c <= 3
  const int key = a * 16 + b * 4 + c;
  switch (key) {
    case 3 * 16: return 0;           // x3
    case 3 * 4: return 1;            // y3
    case 2 * 16 + 4: return 2;       // x2y
    case 16 + 2 * 4: return 3;       // xy2
    case 2 * 16 + 1: return 4;       // x2z
    case 2 * 16: return 5;           // x2
    case 2 * 4 + 1: return 6;        // y2z
    case 2 * 4: return 7;            // y2
    case 16 + 4 + 1: return 8;       // xyz
    case 16 + 4: return 9;           // xy
    case 16 + 2: return 10;          // xz2
    case 16 + 1: return 11;          // xz
    case 16: return 12;              // x
    case 4 + 2: return 13;           // yz2
    case 4 + 1: return 14;           // yz
    case 4: return 15;               // y
    case 3: return 16;               // z3
    case 2: return 17;               // z2
    case 1: return 18;               // z
    default: return 19;              // 1
  }
}
FP_HD inline void mono_exp(int i, int& a, int& b, int& c) {
  const int A[20] = {3, 0, 2, 1, 2, 2, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
  const int B[20] = {0, 3, 1, 2, 0, 0, 2, 2, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0};
  const int C[20] = {0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 2, 1, 0, 2, 1, 0, 3, 2, 1, 0};
  a = A[i]; b = B[i]; c = C[i];
}

// r += s * p * l   with p a polynomial of degree <= 2 (20 coefficients, Nister order) and l linear: l = {x, y, z, 1} coefficients
FP_HD inline void pmul_lin_acc(const double* p, const double* l, double s, double* r) {
  for (int i = 0; i < 20; ++i) {
    if (p[i] == 0.0) continue;
    int a, b, c;
    mono_exp(i, a, b, c);
    r[mono_index(a + 1, b, c)] += s * p[i] * l[0];
    r[mono_index(a, b + 1, c)] += s * p[i] * l[1];
    r[mono_index(a, b, c + 1)] += s * p[i] * l[2];
    r[i] += s * p[i] * l[3];
  }
}
FP_HD inline void lin_to_poly(const double* l, double* p) {
  for (int i = 0; i < 20; ++i) p[i] = 0.0;
  p[12] = l[0]; p[15] = l[1]; p[18] = l[2]; p[19] = l[3];
}

// 4 smallest eigenvectors of the symmetric 9 x 9 matrix A (cyclic Jacobi); basis[k][0..8]
FP_HD inline void null_space_9(double A[9][9], double basis[4][9]) {
  double V[9][9];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 9; ++i) {
      diag += A[i][i] * A[i][i];
      for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-30 * diag) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  bool used[9];
  for (int i = 0; i < 9; ++i) used[i] = false;
  for (int k = 0; k < 4; ++k) {
    int bi = -1; double bv = 1e300;
    for (int i = 0; i < 9; ++i)
      if (!used[i] && A[i][i] < bv) { bv = A[i][i]; bi = i; }
    used[bi] = true;
    for (int i = 0; i < 9; ++i) basis[k][i] = V[i][bi];
  }
}

// polynomial in z (ascending coefficients): r[0..da+db] = a * b
FP_HD inline void zmul(const double* a, int da, const double* b, int db, double* r) {
  for (int i = 0; i <= da + db; ++i) r[i] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) r[i + j] += a[i] * b[j];
}
FP_HD inline double zeval(const double* a, int d, double z) {
  double v = a[d];
  for (int i = d - 1; i >= 0; --i) v = v * z + a[i];
  return v;
}

// real roots of c[0] + c[1] z + ... + c[10] z^10 (Durand-Kerner on all complex roots, Newton polish of the real ones)
FP_HD inline int real_roots_10(const double* c_in, double* roots) {
  int deg = 10;
  double cmax = 0.0;
  for (int i = 0; i <= 10; ++i) cmax = fmax(cmax, fabs(c_in[i]));
  if (!(cmax > 0.0) || !isfinite(cmax)) return 0;
  while (deg > 0 && fabs(c_in[deg]) <= 1e-14 * cmax) --deg;
  if (deg == 0) return 0;
  double c[11];
  for (int i = 0; i <= deg; ++i) c[i] = c_in[i] / c_in[deg];
  double R = 0.0;                                   // Cauchy bound
  for (int i = 0; i < deg; ++i) R = fmax(R, fabs(c[i]));
  R = fmin(1.0 + R, 1e6);
  double zr[10], zi[10];
  for (int k = 0; k < deg; ++k) {
    const double ang = 2.0 * 3.14159265358979323846 * k / deg + 0.4;
    const double rad = R * (0.35 + 0.6 * (k + 1) / deg);           // spread the radii: roots of these polynomials span decades
    zr[k] = rad * cos(ang); zi[k] = rad * sin(ang);
  }
  for (int it = 0; it < 400; ++it) {
    double change = 0.0;
    for (int k = 0; k < deg; ++k) {
      double pr = 1.0, pi = 0.0;                                   // p(z_k), monic Horner
      for (int i = deg - 1; i >= 0; --i) {
        const double t = pr * zr[k] - pi * zi[k] + c[i];
        pi = pr * zi[k] + pi * zr[k];
        pr = t;
      }
      double qr = 1.0, qi = 0.0;                                   // prod_{j != k} (z_k - z_j)
      for (int j = 0; j < deg; ++j) {
        if (j == k) continue;
        const double dr = zr[k] - zr[j], di = zi[k] - zi[j];
        const double t = qr * dr - qi * di;
        qi = qr * di + qi * dr;
        qr = t;
      }
      const double den = qr * qr + qi * qi;
      if (!(den > 0.0)) continue;
      const double wr = (pr * qr + pi * qi) / den, wi = (pi * qr - pr * qi) / den;
      zr[k] -= wr; zi[k] -= wi;
      change = fmax(change, (fabs(wr) + fabs(wi)) / (1.0 + fabs(zr[k]) + fabs(zi[k])));
    }
    if (change < 1e-14) break;
  }
  int n = 0;
  for (int k = 0; k < deg; ++k) {
    if (!isfinite(zr[k]) || !isfinite(zi[k])) continue;
    if (fabs(zi[k]) > 1e-5 * (1.0 + fabs(zr[k]))) continue;
    double z = zr[k];
    for (int it = 0; it < 4; ++it) {                               // Newton on the real polynomial
      double p = c_in[deg], d = 0.0;
      for (int i = deg - 1; i >= 0; --i) { d = d * z + p; p = p * z + c_in[i]; }
      if (d == 0.0 || !isfinite(d)) break;
      const double zn = z - p / d;
      if (!isfinite(zn)) break;
      z = zn;
    }
    roots[n++] = z;
  }
  return n;
}

// x0[5][2], x1[5][2]: calibrated image coordinates with x1^T E x0 = 0.  Writes up to 10 row-major E (unit Frobenius norm);
// returns their number.
FP_HD inline int solve(const double (*x0)[2], const double (*x1)[2], double (*E_out)[9]) {
  // ---- null space of the constraint matrix (through Q^T Q: 9 x 9 symmetric) ----
  double QtQ[9][9];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) QtQ[i][j] = 0.0;
  for (int k = 0; k < 5; ++k) {
    const double a = x0[k][0], b = x0[k][1], c = x1[k][0], d = x1[k][1];
    const double q[9] = {c * a, c * b, c, d * a, d * b, d, a, b, 1.0};
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) QtQ[i][j] += q[i] * q[j];
  }
  double basis[4][9];
  null_space_9(QtQ, basis);
  // E_ij as a linear polynomial {x, y, z, 1}
  double El[9][4];
  for (int i = 0; i < 9; ++i)
    for (int k = 0; k < 4; ++k) El[i][k] = basis[k][i];
  // ---- the ten cubic constraints ----
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int cidx = 0; cidx < 20; ++cidx) A[r][cidx] = 0.0;
  double P[20], Q2[20];
  // row 0: det(E) = sum_i E0i * cofactor_0i
  {
    const int c1[3] = {1, 2, 0}, c2[3] = {2, 0, 1};
    for (int i = 0; i < 3; ++i) {
      // minor = E[1][c1] E[2][c2] - E[1][c2] E[2][c1]
      for (int k = 0; k < 20; ++k) Q2[k] = 0.0;
      lin_to_poly(El[3 + c1[i]], P);
      pmul_lin_acc(P, El[6 + c2[i]], 1.0, Q2);
      lin_to_poly(El[3 + c2[i]], P);
      pmul_lin_acc(P, El[6 + c1[i]], -1.0, Q2);
      pmul_lin_acc(Q2, El[i], 1.0, A[0]);
    }
  }
  // rows 1..9: (E E^T - 1/2 trace(E E^T) I) E = 0
  {
    double EEt[3][3][20];
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        for (int k = 0; k < 20; ++k) EEt[i][j][k] = 0.0;
        for (int m = 0; m < 3; ++m) {
          lin_to_poly(El[3 * i + m], P);
          pmul_lin_acc(P, El[3 * j + m], 1.0, EEt[i][j]);
        }
        if (j != i)
          for (int k = 0; k < 20; ++k) EEt[j][i][k] = EEt[i][j][k];
      }
    for (int k = 0; k < 20; ++k) {
      const double tr = 0.5 * (EEt[0][0][k] + EEt[1][1][k] + EEt[2][2][k]);
      EEt[0][0][k] -= tr; EEt[1][1][k] -= tr; EEt[2][2][k] -= tr;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int m = 0; m < 3; ++m) pmul_lin_acc(EEt[i][m], El[3 * m + j], 1.0, A[1 + 3 * i + j]);
  }
  // ---- Gauss-Jordan on the ten leading monomials (partial pivoting) ----
  for (int col = 0; col < 10; ++col) {
    int piv = col; double pv = fabs(A[col][col]);
    for (int r = col + 1; r < 10; ++r)
      if (fabs(A[r][col]) > pv) { pv = fabs(A[r][col]); piv = r; }
    if (!(pv > 1e-300)) return 0;
    if (piv != col)
      for (int k = 0; k < 20; ++k) { const double t = A[col][k]; A[col][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[col][col];
    for (int k = 0; k < 20; ++k) A[col][k] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == col) continue;
      const double f = A[r][col];
      if (f == 0.0) continue;
      for (int k = 0; k < 20; ++k) A[r][k] -= f * A[col][k];
    }
  }
  // ---- B(z): rows k = <x2z> - z <x2>, l = <y2z> - z <y2>, m = <xyz> - z <xy>; columns x (deg 3), y (deg 3), 1 (deg 4) ----
  double B[3][3][5];
  for (int r = 0; r < 3; ++r) {
    const double* hi = A[4 + 2 * r];      // rows 4, 6, 8
    const double* lo = A[5 + 2 * r];      // rows 5, 7, 9
    for (int v = 0; v < 2; ++v) {         // x: columns 10..12 (z2, z, 1), y: 13..15
      const int o = 10 + 3 * v;
      B[r][v][0] = hi[o + 2];
      B[r][v][1] = hi[o + 1] - lo[o + 2];
      B[r][v][2] = hi[o] - lo[o + 1];
      B[r][v][3] = -lo[o];
      B[r][v][4] = 0.0;
    }
    B[r][2][0] = hi[19];
    B[r][2][1] = hi[18] - lo[19];
    B[r][2][2] = hi[17] - lo[18];
    B[r][2][3] = hi[16] - lo[17];
    B[r][2][4] = -lo[16];
  }
  // det B(z) = B00 (B11 B22 - B12 B21) - B01 (B10 B22 - B12 B20) + B02 (B10 B21 - B11 B20)
  double det[11];
  for (int i = 0; i <= 10; ++i) det[i] = 0.0;
  {
    double t1[8], t2[8], t3[11];
    zmul(B[1][1], 3, B[2][2], 4, t1); zmul(B[1][2], 4, B[2][1], 3, t2);
    for (int i = 0; i <= 7; ++i) t1[i] -= t2[i];
    zmul(B[0][0], 3, t1, 7, t3);
    for (int i = 0; i <= 10; ++i) det[i] += t3[i];
    zmul(B[1][0], 3, B[2][2], 4, t1); zmul(B[1][2], 4, B[2][0], 3, t2);
    for (int i = 0; i <= 7; ++i) t1[i] -= t2[i];
    zmul(B[0][1], 3, t1, 7, t3);
    for (int i = 0; i <= 10; ++i) det[i] -= t3[i];
    double u1[7], u2[7];
    zmul(B[1][0], 3, B[2][1], 3, u1); zmul(B[1][1], 3, B[2][0], 3, u2);
    for (int i = 0; i <= 6; ++i) u1[i] -= u2[i];
    zmul(B[0][2], 4, u1, 6, t3);
    for (int i = 0; i <= 10; ++i) det[i] += t3[i];
  }
  double roots[10];
  const int nr = real_roots_10(det, roots);
  int n_out = 0;
  for (int k = 0; k < nr; ++k) {
    const double z = roots[k];
    // x, y from the best-conditioned pair of rows of B(z)
    double b[3][3];
    for (int r = 0; r < 3; ++r) {
      b[r][0] = zeval(B[r][0], 3, z); b[r][1] = zeval(B[r][1], 3, z); b[r][2] = zeval(B[r][2], 4, z);
    }
    int r0 = 0, r1 = 1; double best = -1.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double d = fabs(b[p][0] * b[q][1] - b[p][1] * b[q][0]);
        const double sc = (fabs(b[p][0]) + fabs(b[p][1])) * (fabs(b[q][0]) + fabs(b[q][1]));
        const double cond = sc > 0 ? d / sc : 0.0;
        if (cond > best) { best = cond; r0 = p; r1 = q; }
      }
    const double D = b[r0][0] * b[r1][1] - b[r0][1] * b[r1][0];
    if (!(fabs(D) > 0.0)) continue;
    const double x = (b[r0][1] * b[r1][2] - b[r0][2] * b[r1][1]) / D;
    const double y = (b[r0][2] * b[r1][0] - b[r0][0] * b[r1][2]) / D;
    double nrm = 0.0, E[9];
    for (int i = 0; i < 9; ++i) {
      E[i] = x * El[i][0] + y * El[i][1] + z * El[i][2] + El[i][3];
      nrm += E[i] * E[i];
    }
    if (!(nrm > 0.0) || !isfinite(nrm)) continue;
    nrm = 1.0 / sqrt(nrm);
    for (int i = 0; i < 9; ++i) E_out[n_out][i] = E[i] * nrm;
    ++n_out;
  }
  return n_out;
}

}  // namespace fivept
