// oracle/linalg.hpp -- tiny dependency-free linear algebra for the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under
// anticipated-vins-mono_b200/ may include or link this.
//
// Restates the handful of Eigen operations the reference path uses, with the
// same formulas Eigen uses where the formula is observable in the result
// (quaternion product / rotation / toRotationMatrix, 3x3 cofactor inverse,
// partial-pivot LU inverse, LLT).  Eigen itself is an un-vendored, un-pinned
// dependency of the reference (vins_estimator/CMakeLists.txt:28).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

struct V3 {
  double x, y, z;
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / norm(a); }

struct M3 {
  double m[3][3];
  double* operator[](int i) { return m[i]; }
  const double* operator[](int i) const { return m[i]; }
};
inline M3 zero3() { M3 r; std::memset(&r, 0, sizeof r); return r; }
inline M3 eye3() { M3 r = zero3(); r[0][0] = r[1][1] = r[2][2] = 1; return r; }
inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
  return r;
}
inline V3 operator*(const M3& a, V3 v) {
  return {a[0][0] * v.x + a[0][1] * v.y + a[0][2] * v.z, a[1][0] * v.x + a[1][1] * v.y + a[1][2] * v.z,
          a[2][0] * v.x + a[2][1] * v.y + a[2][2] * v.z};
}
inline M3 operator*(double s, const M3& a) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = s * a[i][j];
  return r;
}
inline M3 operator+(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = a[i][j] + b[i][j];
  return r;
}
inline M3 operator-(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = a[i][j] - b[i][j];
  return r;
}
inline M3 operator-(const M3& a) { return -1.0 * a; }
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = a[j][i];
  return r;
}
// Utility::skewSymmetric, vins_estimator/src/utility/utility.h:26-34
inline M3 skew(V3 q) {
  M3 r = zero3();
  r[0][1] = -q.z; r[0][2] = q.y; r[1][0] = q.z; r[1][2] = -q.x; r[2][0] = -q.y; r[2][1] = q.x;
  return r;
}
// Eigen fixed-size 3x3 inverse (cofactors / determinant)
inline M3 inverse3(const M3& a) {
  M3 c;
  c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  c[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
  c[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
  c[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
  c[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
  c[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  c[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
  c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  double det = a[0][0] * c[0][0] + a[0][1] * c[1][0] + a[0][2] * c[2][0];
  return (1.0 / det) * c;
}

// Quaternion with Eigen semantics (coeffs x y z w)
struct Q4 { double x, y, z, w; };
inline Q4 qmul(Q4 a, Q4 b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q4 qconj(Q4 q) { return {-q.x, -q.y, -q.z, q.w}; }
// Eigen's Quaternion::inverse(): conjugate / squaredNorm
inline Q4 qinv(Q4 q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
inline Q4 qnormalized(Q4 q) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
inline V3 qvec(Q4 q) { return {q.x, q.y, q.z}; }
// QuaternionBase::_transformVector
inline V3 qrot(Q4 q, V3 v) {
  V3 u{q.x, q.y, q.z};
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// QuaternionBase::toRotationMatrix
inline M3 qmat(Q4 q) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r[0][0] = 1 - (tyy + tzz); r[0][1] = txy - twz; r[0][2] = txz + twy;
  r[1][0] = txy + twz; r[1][1] = 1 - (txx + tzz); r[1][2] = tyz - twx;
  r[2][0] = txz - twy; r[2][1] = tyz + twx; r[2][2] = 1 - (txx + tyy);
  return r;
}
// Quaterniond(Matrix3d) constructor
inline Q4 qfrommat(const M3& R) {
  double t = R[0][0] + R[1][1] + R[2][2];
  Q4 q;
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[2][1] - R[1][2]) * t; q.y = (R[0][2] - R[2][0]) * t; q.z = (R[1][0] - R[0][1]) * t;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    double c[4];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (R[k][j] - R[j][k]) * t;
    c[j] = (R[j][i] + R[i][j]) * t;
    c[k] = (R[k][i] + R[i][k]) * t;
    q = {c[0], c[1], c[2], c[3]};
  }
  return q;
}
// QuaternionBase::slerp(t, other)
inline Q4 qslerp(Q4 a, double t, Q4 b) {
  const double one = 1.0 - 2.220446049250313e-16;
  double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  double ad = std::fabs(d);
  double s0, s1;
  if (ad >= one) {
    s0 = 1.0 - t; s1 = t;
  } else {
    double th = std::acos(ad), st = std::sin(th);
    s0 = std::sin((1.0 - t) * th) / st;
    s1 = std::sin(t * th) / st;
  }
  if (d < 0) s1 = -s1;
  return {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}
// Utility::deltaQ, utility.h:11-24  (NOT normalized)
inline Q4 deltaQ(V3 th) { return {th.x / 2, th.y / 2, th.z / 2, 1.0}; }

// ---- dense, row-major, dynamic ------------------------------------------------
// In-place lower Cholesky A = L L^T (lower triangle of a holds L). false if not PD.
inline bool cholesky(double* a, int n) {
  for (int j = 0; j < n; j++) {
    double d = a[j * n + j];
    for (int k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    a[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = a[i * n + j];
      for (int k = 0; k < j; k++) s -= a[i * n + k] * a[j * n + k];
      a[i * n + j] = s / d;
    }
  }
  return true;
}
inline void chol_solve(const double* l, int n, double* b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= l[i * n + k] * b[k];
    b[i] = s / l[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= l[k * n + i] * b[k];
    b[i] = s / l[i * n + i];
  }
}
// Inverse by partial-pivot LU (what Eigen's MatrixBase::inverse() does for N > 4)
inline bool lu_inverse(const double* a_in, int n, double* inv) {
  std::vector<double> a(a_in, a_in + n * n);
  std::vector<int> piv(n);
  for (int i = 0; i < n; i++) piv[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(a[i * n + k]) > best) { best = std::fabs(a[i * n + k]); p = i; }
    if (best == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(a[k * n + j], a[p * n + j]);
      std::swap(piv[k], piv[p]);
    }
    for (int i = k + 1; i < n; i++) {
      a[i * n + k] /= a[k * n + k];
      double f = a[i * n + k];
      for (int j = k + 1; j < n; j++) a[i * n + j] -= f * a[k * n + j];
    }
  }
  // solve L U X = P I
  for (int c = 0; c < n; c++) {
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) {
      double s = (piv[i] == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) s -= a[i * n + k] * y[k];
      y[i] = s;
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < n; k++) s -= a[i * n + k] * y[k];
      y[i] = s / a[i * n + i];
    }
    for (int i = 0; i < n; i++) inv[i * n + c] = y[i];
  }
  return true;
}
// Cyclic Jacobi eigen-decomposition of a symmetric matrix (row-major n x n).
// On return w holds eigenvalues (ascending) and v the eigenvectors as COLUMNS.
// Stands in for Eigen::SelfAdjointEigenSolver (marginalization_factor.cpp:268,283);
// eigenvalues agree to O(eps*||A||), the quantities built from them are basis independent.
inline void jacobi_eigh(const double* a_in, int n, double* w, double* v) {
  std::vector<double> a(a_in, a_in + n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) v[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) {
      diag += a[i * n + i] * a[i * n + i];
      for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    }
    if (off <= 1e-40 * (diag + 1e-300) || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[p * n + q];
        if (apq == 0.0) continue;
        double app = a[p * n + p], aqq = a[q * n + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
  std::vector<double> vv(v, v + n * n);
  for (int j = 0; j < n; j++) {
    w[j] = a[idx[j] * n + idx[j]];
    for (int i = 0; i < n; i++) v[i * n + j] = vv[i * n + idx[j]];
  }
}


// Symmetric eigen-decomposition by Householder tridiagonalisation + implicit QL iterations -- the algorithm family of
// Eigen::SelfAdjointEigenSolver (tridiagonalisation + implicit symmetric QR, marginalization_factor.cpp:268,283), so
// that the CPU baseline's marginalization costs what the reference's does (the cyclic Jacobi above is O(10 n^3) slower).
// Same conventions as jacobi_eigh: w ascending, eigenvectors as COLUMNS of v (row-major n x n).  The two routines agree
// to O(eps ||A||) in the eigenvalues and in every basis-independent quantity (tests/test_oracle_marg.py).
inline void tridiag_ql_eigh(const double* a_in, int n, double* d, double* V) {
  std::vector<double> e(n, 0.0);
  for (int i = 0; i < n * n; i++) V[i] = a_in[i];
  if (n == 0) return;
  // ---- Householder reduction to tridiagonal form (EISPACK tred2)
  for (int j = 0; j < n; j++) d[j] = V[(n - 1) * n + j];
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = V[(i - 1) * n + j]; V[i * n + j] = 0.0; V[j * n + i] = 0.0; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1], g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        V[j * n + i] = f;
        g = e[j] + V[j * n + j] * f;
        for (int k = j + 1; k <= i - 1; k++) { g += V[k * n + j] * d[k]; e[k] += V[k * n + j] * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; k++) V[k * n + j] -= (f * e[k] + g * d[k]);
        d[j] = V[(i - 1) * n + j];
        V[i * n + j] = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    V[(n - 1) * n + i] = V[i * n + i];
    V[i * n + i] = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = V[k * n + i + 1] / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += V[k * n + i + 1] * V[k * n + j];
        for (int k = 0; k <= i; k++) V[k * n + j] -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) V[k * n + i + 1] = 0.0;
  }
  for (int j = 0; j < n; j++) { d[j] = V[(n - 1) * n + j]; V[(n - 1) * n + j] = 0.0; }
  V[(n - 1) * n + n - 1] = 1.0;
  e[0] = 0.0;
  // ---- implicit QL with eigenvector accumulation (EISPACK tql2)
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; l++) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; m++; }
    if (m > l) {
      int iter = 0;
      do {
        iter++;
        double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = V[k * n + i + 1];
            V[k * n + i + 1] = s * V[k * n + i] + c * h;
            V[k * n + i] = c * V[k * n + i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  // ascending order
  for (int i = 0; i < n - 1; i++) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < n; j++) std::swap(V[j * n + i], V[j * n + k]);
    }
  }
}

}  // namespace orc
