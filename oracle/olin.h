// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or call anything under oracle/.
//
// Small dense linear algebra the reference gets from Eigen 3.3 (un-vendored, absent here):
//   * 3x3 inverse by cofactors            (Eigen::Matrix3d::inverse, used at
//                                          include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:355,359)
//   * symmetric 3x3 eigen-decomposition   (Eigen::SelfAdjointEigenSolver<Matrix3d>::compute, :333)
//   * 6x6 SVD least-squares solve         (Eigen::JacobiSVD<6x6>::solve, include/ndt_omp/ndt_omp_impl2.hpp:138-140)
// Eigen's eigen-solver is an iterative tridiagonal QL; only V*L*V^-1 and L are consumed by the
// reference, so a cyclic-Jacobi solver (sqrt and + - * / only) is an equivalent restatement.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace olin {

struct M3 { double a[3][3]; };
struct V3 { double v[3]; };

inline M3 m3_zero() { M3 r; std::memset(&r, 0, sizeof r); return r; }
inline M3 m3_identity() { M3 r = m3_zero(); r.a[0][0] = r.a[1][1] = r.a[2][2] = 1.0; return r; }

inline M3 m3_mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = x.a[i][0] * y.a[0][j];
      s += x.a[i][1] * y.a[1][j];
      s += x.a[i][2] * y.a[2][j];
      r.a[i][j] = s;
    }
  return r;
}
inline V3 m3_mulv(const M3& x, const V3& y) {
  V3 r;
  for (int i = 0; i < 3; i++) {
    double s = x.a[i][0] * y.v[0];
    s += x.a[i][1] * y.v[1];
    s += x.a[i][2] * y.v[2];
    r.v[i] = s;
  }
  return r;
}
inline M3 m3_T(const M3& x) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = x.a[j][i];
  return r;
}

// Eigen compute_inverse_size3: inverse(i,j) = cofactor(j,i) / det, det expanded along column 0.
inline double m3_cofactor(const M3& m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m.a[i1][j1] * m.a[i2][j2] - m.a[i1][j2] * m.a[i2][j1];
}
inline M3 m3_inverse(const M3& m) {
  double c0 = m3_cofactor(m, 0, 0), c1 = m3_cofactor(m, 1, 0), c2 = m3_cofactor(m, 2, 0);
  double det = (c0 * m.a[0][0] + c1 * m.a[1][0]) + c2 * m.a[2][0];
  double invdet = 1.0 / det;
  M3 r;
  r.a[0][0] = c0 * invdet; r.a[0][1] = c1 * invdet; r.a[0][2] = c2 * invdet;
  for (int i = 1; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = m3_cofactor(m, j, i) * invdet;
  return r;
}

// Cyclic Jacobi for a symmetric 3x3.  Output eigenvalues ascending, eigenvectors in columns of V.
inline void sym3_eig(const M3& A_in, double evals[3], M3& V) {
  double a[3][3];
  // Eigen::SelfAdjointEigenSolver reads the lower triangle only
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = (i >= j) ? A_in.a[i][j] : A_in.a[j][i];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = a[p][q];
        if (apq == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        int r = 3 - p - q;
        double app = a[p][p], aqq = a[q][q], arp = a[r][p], arq = a[r][q];
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = a[q][p] = 0.0;
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; k++) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  // stable ascending sort of three values
  if (d[idx[1]] < d[idx[0]]) std::swap(idx[0], idx[1]);
  if (d[idx[2]] < d[idx[1]]) std::swap(idx[1], idx[2]);
  if (d[idx[1]] < d[idx[0]]) std::swap(idx[0], idx[1]);
  for (int j = 0; j < 3; j++) {
    evals[j] = d[idx[j]];
    for (int k = 0; k < 3; k++) V.a[k][j] = v[k][idx[j]];
  }
}

// One-sided (Hestenes) Jacobi SVD of a 6x6, then x = V * diag(1/s_k, k < rank) * U^T * b with
// Eigen's rank rule (SVDBase::rank/threshold): s_k > max(s_0 * 6 * eps, DBL_MIN).
inline void svd6_solve(const double A[36] /*row-major*/, const double b[6], double x[6], double sv_out[6] = nullptr) {
  double U[6][6], V[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { U[i][j] = A[i * 6 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int p = 0; p < 5; p++)
      for (int q = p + 1; q < 6; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 6; k++) { alpha += U[k][p] * U[k][p]; beta += U[k][q] * U[k][q]; gamma += U[k][p] * U[k][q]; }
        if (gamma == 0.0) continue;
        if (std::fabs(gamma) <= 2.220446049250313e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 6; k++) {
          double up = U[k][p], uq = U[k][q];
          U[k][p] = c * up - s * uq; U[k][q] = s * up + c * uq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sv[6];
  for (int j = 0; j < 6; j++) {
    double n = 0;
    for (int k = 0; k < 6; k++) n += U[k][j] * U[k][j];
    sv[j] = std::sqrt(n);
  }
  double smax = 0;
  for (int j = 0; j < 6; j++) smax = std::max(smax, sv[j]);
  double thr = std::max(smax * 6.0 * 2.220446049250313e-16, 2.2250738585072014e-308);
  for (int i = 0; i < 6; i++) x[i] = 0.0;
  for (int j = 0; j < 6; j++) {
    if (!(sv[j] > thr)) continue;
    double ub = 0;
    for (int k = 0; k < 6; k++) ub += U[k][j] * b[k];   // (u_j * s_j)^T b
    double coef = ub / (sv[j] * sv[j]);                 // = (u_j^T b) / s_j
    for (int i = 0; i < 6; i++) x[i] += V[i][j] * coef;
  }
  if (sv_out) { for (int j = 0; j < 6; j++) sv_out[j] = sv[j]; std::sort(sv_out, sv_out + 6, [](double a, double b) { return a > b; }); }
}

}  // namespace olin
