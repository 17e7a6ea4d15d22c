// Double-precision control math shared by host and device code of the product library:
// se(3) <-> SE(3) (Sophus a621ff2 semantics: sophus/so3.cpp:127-202, sophus/se3.cpp:170-220),
// Eigen-style quaternion <-> matrix, 3x3 cofactor inverse, symmetric 3x3 eigen-decomposition,
// 6x6 SVD least-squares solve (replaces Eigen::JacobiSVD<6x6>::solve at ndt_omp_impl2.hpp:138-140).
// All functions are __host__ __device__ and use only + - * / sqrt and libm transcendentals.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LVS_HD __host__ __device__ __forceinline__
#else
#define LVS_HD inline
#endif

namespace lvs {

static constexpr double kSmallEps = 1e-10;   // Sophus SMALL_EPS (so3.h:35)

struct Q4 { double w, x, y, z; };
struct Pose { Q4 q; double t[3]; };          // unit quaternion + translation, like Sophus::SE3

LVS_HD void mat3_mul(const double* A, const double* B, double* C) {   // row-major 3x3
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
}

LVS_HD double cof3(const double* m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
// inverse(i,j) = cofactor(j,i) / det with det expanded along column 0 (Eigen compute_inverse_size3)
LVS_HD void mat3_inverse(const double* m, double* r) {
  double c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  double det = (c0 * m[0] + c1 * m[3]) + c2 * m[6];
  double inv = 1.0 / det;
  r[0] = c0 * inv; r[1] = c1 * inv; r[2] = c2 * inv;
  for (int i = 1; i < 3; i++)
    for (int j = 0; j < 3; j++) r[i * 3 + j] = cof3(m, j, i) * inv;
}

// Cyclic Jacobi, ascending eigenvalues, eigenvectors in the columns of V (row-major).
LVS_HD void sym3_eigen(const double* Ain, double* ev, double* V) {
  // Eigen::SelfAdjointEigenSolver reads the lower triangle only
  double a00 = Ain[0], a11 = Ain[4], a22 = Ain[8];
  double a01 = Ain[3], a02 = Ain[6], a12 = Ain[7];
  double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = a01 * a01 + a02 * a02 + a12 * a12;
    double dg = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-34 * dg || off == 0.0) break;
    // rotation (0,1), third index 2
    for (int pq = 0; pq < 3; pq++) {
      double apq = pq == 0 ? a01 : (pq == 1 ? a02 : a12);
      if (apq == 0.0) continue;
      double app = pq == 2 ? a11 : a00, aqq = pq == 0 ? a11 : a22;
      double theta = (aqq - app) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double npp = app - t * apq, nqq = aqq + t * apq;
      if (pq == 0) {        // p=0,q=1,r=2 : arp=a02, arq=a12
        double arp = a02, arq = a12;
        a00 = npp; a11 = nqq; a01 = 0.0; a02 = c * arp - s * arq; a12 = s * arp + c * arq;
      } else if (pq == 1) { // p=0,q=2,r=1 : arp=a01, arq=a12
        double arp = a01, arq = a12;
        a00 = npp; a22 = nqq; a02 = 0.0; a01 = c * arp - s * arq; a12 = s * arp + c * arq;
      } else {              // p=1,q=2,r=0 : arp=a01, arq=a02
        double arp = a01, arq = a02;
        a11 = npp; a22 = nqq; a12 = 0.0; a01 = c * arp - s * arq; a02 = s * arp + c * arq;
      }
      for (int k = 0; k < 3; k++) {
        double vp = v[k * 3 + p], vq = v[k * 3 + q];
        v[k * 3 + p] = c * vp - s * vq;
        v[k * 3 + q] = s * vp + c * vq;
      }
    }
  }
  double d[3] = {a00, a11, a22};
  int i0 = 0, i1 = 1, i2 = 2, tmp;
  if (d[i1] < d[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
  if (d[i2] < d[i1]) { tmp = i1; i1 = i2; i2 = tmp; }
  if (d[i1] < d[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
  ev[0] = d[i0]; ev[1] = d[i1]; ev[2] = d[i2];
  for (int k = 0; k < 3; k++) { V[k * 3] = v[k * 3 + i0]; V[k * 3 + 1] = v[k * 3 + i1]; V[k * 3 + 2] = v[k * 3 + i2]; }
}

// x = pinv(A) b through a one-sided Jacobi SVD; rank rule of Eigen::SVDBase (s_k > max(s_max*6*eps, DBL_MIN)).
// Every loop over matrix indices is fully unrolled so that on the device U and V live in registers (the solve runs on
// one thread of the CTA that finishes an evaluation; with U/V in local memory it cost more than the evaluation itself).
LVS_HD void svd6_solve(const double* A /*row-major 6x6*/, const double* b, double* x) {
  double U[36], V[36];
#pragma unroll
  for (int i = 0; i < 36; i++) { U[i] = A[i]; V[i] = ((i % 7) == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
#pragma unroll
    for (int p = 0; p < 5; p++) {
#pragma unroll
      for (int q = p + 1; q < 6; q++) {
        double al = 0, be = 0, ga = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) { double up = U[k * 6 + p], uq = U[k * 6 + q]; al += up * up; be += uq * uq; ga += up * uq; }
        if (ga != 0.0 && !(fabs(ga) <= 2.220446049250313e-16 * sqrt(al * be))) {
          rotated = true;
          double zeta = (be - al) / (2.0 * ga);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int k = 0; k < 6; k++) {
            double up = U[k * 6 + p], uq = U[k * 6 + q];
            U[k * 6 + p] = c * up - s * uq; U[k * 6 + q] = s * up + c * uq;
            double vp = V[k * 6 + p], vq = V[k * 6 + q];
            V[k * 6 + p] = c * vp - s * vq; V[k * 6 + q] = s * vp + c * vq;
          }
        }
      }
    }
    if (!rotated) break;
  }
  double sv[6], smax = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double n = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) n += U[k * 6 + j] * U[k * 6 + j];
    sv[j] = sqrt(n);
    if (sv[j] > smax) smax = sv[j];
  }
  double thr = smax * 6.0 * 2.220446049250313e-16;
  if (thr < 2.2250738585072014e-308) thr = 2.2250738585072014e-308;
  double xx[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (sv[j] > thr) {
      double ub = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) ub += U[k * 6 + j] * b[k];
      double coef = ub / (sv[j] * sv[j]);
#pragma unroll
      for (int i = 0; i < 6; i++) xx[i] += V[i * 6 + j] * coef;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) x[i] = xx[i];
}

// Fast path of the Newton solve: Gaussian elimination with partial pivoting, fully unrolled (registers only).
// Returns false when a pivot is tiny relative to the largest entry of A (numerically rank-deficient H): the caller
// then takes the SVD path, which implements the pseudo-inverse semantics of Eigen::JacobiSVD::solve.
LVS_HD bool lu6_solve(const double* A /*row-major 6x6*/, const double* b, double* x) {
  double M[36], r[6];
  double amax = 0.0;
#pragma unroll
  for (int i = 0; i < 36; i++) { M[i] = A[i]; amax = fmax(amax, fabs(A[i])); }
#pragma unroll
  for (int i = 0; i < 6; i++) r[i] = b[i];
  if (!(amax > 0.0) || !(amax < 1.0e300)) return false;
  const double tiny = amax * 1e-10;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    // bring the largest |entry| of column k (rows k..5) to row k by compare-and-swap (static indices only)
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      if (fabs(M[i * 6 + k]) > fabs(M[k * 6 + k])) {
#pragma unroll
        for (int j = k; j < 6; j++) { double t = M[k * 6 + j]; M[k * 6 + j] = M[i * 6 + j]; M[i * 6 + j] = t; }
        double t = r[k]; r[k] = r[i]; r[i] = t;
      }
    }
    const double piv = M[k * 6 + k];
    if (!(fabs(piv) > tiny)) return false;
    const double inv = 1.0 / piv;
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      const double f = M[i * 6 + k] * inv;
#pragma unroll
      for (int j = k + 1; j < 6; j++) M[i * 6 + j] -= f * M[k * 6 + j];
      r[i] -= f * r[k];
    }
  }
#pragma unroll
  for (int k = 5; k >= 0; k--) {
    double t = r[k];
#pragma unroll
    for (int j = k + 1; j < 6; j++) t -= M[k * 6 + j] * x[j];
    x[k] = t / M[k * 6 + k];
  }
  return true;
}

// ---- quaternions (Eigen 3.3 Geometry semantics) ----
LVS_HD Q4 quat_from_mat(const double* m) {   // row-major 3x3
  Q4 q;
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[k * 3 + j] - m[j * 3 + k]) * t;
    v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}

LVS_HD void quat_to_mat(const Q4& q, double* r) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r[0] = 1 - (tyy + tzz); r[1] = txy - twz;       r[2] = txz + twy;
  r[3] = txy + twz;       r[4] = 1 - (txx + tzz); r[5] = tyz - twx;
  r[6] = txz - twy;       r[7] = tyz + twx;       r[8] = 1 - (txx + tyy);
}

LVS_HD Q4 quat_mul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
LVS_HD void quat_normalize(Q4& q) {
  double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
LVS_HD void quat_rotate(const Q4& q, const double* v, double* r) {
  double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  r[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  r[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  r[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}

LVS_HD void hat3(const double* w, double* O) {
  O[0] = 0;     O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2];  O[4] = 0;     O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0];  O[8] = 0;
}

LVS_HD Pose se3_exp(const double* u) {       // u = [upsilon, omega]
  Pose P;
  const double* om = u + 3;
  double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  double half = 0.5 * theta, imag, real = cos(half);
  if (theta < kSmallEps) {
    double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = sin(half) / theta;
  }
  P.q.w = real; P.q.x = imag * om[0]; P.q.y = imag * om[1]; P.q.z = imag * om[2];
  quat_normalize(P.q);
  double O[9], O2[9], V[9];
  hat3(om, O);
  mat3_mul(O, O, O2);
  if (theta < kSmallEps) {
    quat_to_mat(P.q, V);
  } else {
    double t2 = theta * theta;
    double a = (1 - cos(theta)) / t2, b = (theta - sin(theta)) / (t2 * theta);
    for (int i = 0; i < 9; i++) V[i] = (((i % 4) == 0 ? 1.0 : 0.0) + a * O[i]) + b * O2[i];
  }
  for (int i = 0; i < 3; i++) P.t[i] = (V[i * 3] * u[0] + V[i * 3 + 1] * u[1]) + V[i * 3 + 2] * u[2];
  return P;
}

LVS_HD void se3_log(const Pose& P, double* out) {
  double n = sqrt(P.q.x * P.q.x + P.q.y * P.q.y + P.q.z * P.q.z), w = P.q.w, f;
  if (n < kSmallEps) f = 2. / w - 2. * (n * n) / (w * (w * w));
  else f = 2 * atan(n / w) / n;
  double theta = f * n;
  double om[3] = {f * P.q.x, f * P.q.y, f * P.q.z};
  double O[9], O2[9];
  hat3(om, O);
  mat3_mul(O, O, O2);
  double c = (theta < kSmallEps) ? (1. / 12.) : (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  double Vi[9];
  for (int i = 0; i < 9; i++) Vi[i] = (((i % 4) == 0 ? 1.0 : 0.0) - 0.5 * O[i]) + c * O2[i];
  for (int i = 0; i < 3; i++) out[i] = (Vi[i * 3] * P.t[0] + Vi[i * 3 + 1] * P.t[1]) + Vi[i * 3 + 2] * P.t[2];
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

LVS_HD Pose se3_mul(const Pose& a, const Pose& b) {
  Pose r;
  double rt[3];
  quat_rotate(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  r.q = quat_mul(a.q, b.q);
  quat_normalize(r.q);
  return r;
}

// SE3::matrix().cast<float>() as a column-major 4x4
LVS_HD void pose_to_matrix4f(const Pose& P, float* M) {
  double R[9];
  quat_to_mat(P.q, R);
  for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) M[c * 4 + r] = (float)R[r * 3 + c]; M[c * 4 + 3] = 0.f; }
  M[12] = (float)P.t[0]; M[13] = (float)P.t[1]; M[14] = (float)P.t[2]; M[15] = 1.f;
}

// Sophus::SE3(R, t) from a column-major float 4x4 (quaternion taken from R, not re-normalised), then log()
LVS_HD void matrix4f_to_se3_log(const float* M, double* p) {
  double R[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[r * 3 + c] = (double)M[c * 4 + r];
  Pose P;
  P.q = quat_from_mat(R);
  P.t[0] = (double)M[12]; P.t[1] = (double)M[13]; P.t[2] = (double)M[14];
  se3_log(P, p);
}

// ---- expf, as the reference binds it.  updateDerivatives calls unqualified `exp` on a float (include/ndt_omp/ndt_omp_impl2.hpp:581).
// Under PCL's headers (<pcl/pcl_macros.h> includes <math.h>; libstdc++'s <math.h> wrapper does `using std::exp`) overload resolution
// picks std::exp(float) = expf, i.e. glibc's single-precision routine (>= 2.27 on the reference's Ubuntu 18.04: sysdeps/ieee754/
// flt-32/e_expf.c + e_exp2f_data.c, EXP2F_TABLE_BITS = 5, cubic polynomial, evaluated in double and rounded once to float).
// This is that routine restated operation by operation, with the contractions of glibc's FMA build (the ifunc variant every
// x86-64 CPU since Haswell selects): bit-identical to the host's expf over ALL 2.24e9 finite floats in [-104, 88.8]
// (tools/expf_sweep.c; without the contraction of `z - kd` it differs in one input per sign).  T[i] = bits(2^(i/32)) - (i << 47).
#ifdef __CUDACC__
static __constant__ unsigned long long c_exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, 0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL,
    0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, 0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, 0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL,
    0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, 0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, 0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL,
    0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// tab: the 32-entry table above, in shared memory (the index is lane-divergent) or in global memory.
__device__ __forceinline__ float glibc_expf(float x, const unsigned long long* tab) {
  if (!(x >= -0x1.9fe368p6f)) return x != x ? x + x : 0.0f;      // NaN; x < log(2^-150): underflow to +0 (and -inf)
  if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);      // x > log(2^128): overflow
  const double xd = (double)x;
  const double z = __dmul_rn(0x1.71547652b82fep+5, xd);          // x * 32 / ln 2 (intrinsics: never contracted, whatever -fmad says)
  double kd = __dadd_rn(z, 0x1.8p+52);                           // round to integer through the shift constant
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, 0x1.8p+52);
  const double r = fma(0x1.71547652b82fep+5, xd, -kd);
  const unsigned long long t = tab[ki & 31ULL] + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double zz = fma(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
  const double r2 = __dmul_rn(r, r);
  double y = fma(0x1.62e42ff0c52d6p-6, r, 1.0);
  y = fma(zz, r2, y);
  return (float)__dmul_rn(y, s);
}
#endif

}  // namespace lvs
