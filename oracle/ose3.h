// ORACLE — TEST INFRASTRUCTURE ONLY (see olin.h header).
//
// se(3) <-> SE(3) as the reference gets it from Sophus a621ff2 (vendored only as
// /root/reference/3rdtools/Sophus-a621ff2-ubuntu18.04.zip; it needs Eigen, which this image lacks).
// PINNED against those sources: oracle/build_ref.sh compiles Sophus' own so3.cpp / se3.cpp against a stand-in for the Eigen operations
// they use (oracle/ref_stubs/eigen_min.h) into oracle/_ref/libsophus_ref.so, and
// tests/test_oracle_ndt.py::test_se3_restatement_against_the_reference_sophus_sources holds this file to it bit for bit (random, tiny-angle,
// near-pi arguments; Sophus' own test_se3.cpp cases pass on the compiled sources).  What that pins is Sophus; Eigen's rounding is not.
// Restated from  Sophus/sophus/so3.cpp:127-202 (logAndTheta, expAndTheta), so3.h:35 (SMALL_EPS),
// se3.cpp:60-110 (operator*, inverse), se3.cpp:170-220 (exp, log); the Eigen::Quaterniond
// pieces Sophus relies on (matrix->quaternion, quaternion->matrix, product, normalize) follow
// Eigen 3.3 Geometry/Quaternion.h semantics.
#pragma once
#include "olin.h"

namespace ose3 {
using olin::M3;
using olin::V3;

static const double SMALL_EPS = 1e-10;

struct Quat { double w, x, y, z; };
struct SE3 { Quat q; V3 t; };

inline Quat quat_from_matrix(const M3& m) {
  Quat q;
  double t = m.a[0][0] + m.a[1][1] + m.a[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m.a[2][1] - m.a[1][2]) * t;
    q.y = (m.a[0][2] - m.a[2][0]) * t;
    q.z = (m.a[1][0] - m.a[0][1]) * t;
  } else {
    int i = 0;
    if (m.a[1][1] > m.a[0][0]) i = 1;
    if (m.a[2][2] > m.a[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m.a[i][i] - m.a[j][j] - m.a[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m.a[k][j] - m.a[j][k]) * t;
    v[j] = (m.a[j][i] + m.a[i][j]) * t;
    v[k] = (m.a[k][i] + m.a[i][k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}

inline M3 quat_to_matrix(const Quat& q) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.a[0][0] = 1 - (tyy + tzz); r.a[0][1] = txy - twz;       r.a[0][2] = txz + twy;
  r.a[1][0] = txy + twz;       r.a[1][1] = 1 - (txx + tzz); r.a[1][2] = tyz - twx;
  r.a[2][0] = txz - twy;       r.a[2][1] = tyz + twx;       r.a[2][2] = 1 - (txx + tyy);
  return r;
}

inline Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline void quat_normalize(Quat& q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
// Eigen QuaternionBase::_transformVector: v + w*uv + vec x uv, uv = 2*(vec x v)
inline V3 quat_rotate(const Quat& q, const V3& v) {
  double ux = q.y * v.v[2] - q.z * v.v[1], uy = q.z * v.v[0] - q.x * v.v[2], uz = q.x * v.v[1] - q.y * v.v[0];
  ux += ux; uy += uy; uz += uz;
  V3 r;
  r.v[0] = v.v[0] + q.w * ux + (q.y * uz - q.z * uy);
  r.v[1] = v.v[1] + q.w * uy + (q.z * ux - q.x * uz);
  r.v[2] = v.v[2] + q.w * uz + (q.x * uy - q.y * ux);
  return r;
}

inline M3 hat(const V3& w) {
  M3 r = olin::m3_zero();
  r.a[0][1] = -w.v[2]; r.a[0][2] = w.v[1];
  r.a[1][0] = w.v[2];  r.a[1][2] = -w.v[0];
  r.a[2][0] = -w.v[1]; r.a[2][1] = w.v[0];
  return r;
}

// so3.cpp:172-199
inline Quat so3_exp(const V3& omega, double* theta) {
  *theta = std::sqrt(omega.v[0] * omega.v[0] + omega.v[1] * omega.v[1] + omega.v[2] * omega.v[2]);
  double half = 0.5 * (*theta);
  double imag, real = std::cos(half);
  if (*theta < SMALL_EPS) {
    double t2 = (*theta) * (*theta), t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = std::sin(half) / (*theta);
  }
  Quat q = {real, imag * omega.v[0], imag * omega.v[1], imag * omega.v[2]};
  quat_normalize(q);  // SO3(const Quaterniond&) normalises (so3.cpp:43-48)
  return q;
}

// so3.cpp:127-170
inline V3 so3_log(const Quat& q, double* theta) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  double w = q.w, f;
  if (n < SMALL_EPS) {
    f = 2. / w - 2. * (n * n) / (w * (w * w));
  } else {
    f = 2 * std::atan(n / w) / n;  // the |w|<eps branch is overwritten by this line in the source
  }
  *theta = f * n;
  V3 r = {{f * q.x, f * q.y, f * q.z}};
  return r;
}

// se3.cpp:170-198
inline SE3 se3_exp(const double u[6]) {
  V3 ups = {{u[0], u[1], u[2]}}, om = {{u[3], u[4], u[5]}};
  double theta;
  SE3 r;
  r.q = so3_exp(om, &theta);
  M3 Om = hat(om), Om2 = olin::m3_mul(Om, Om), V;
  if (theta < SMALL_EPS) {
    V = quat_to_matrix(r.q);
  } else {
    double t2 = theta * theta;
    double a = (1 - std::cos(theta)) / t2, b = (theta - std::sin(theta)) / (t2 * theta);
    V = olin::m3_identity();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V.a[i][j] = (V.a[i][j] + a * Om.a[i][j]) + b * Om2.a[i][j];
  }
  r.t = olin::m3_mulv(V, ups);
  return r;
}

// se3.cpp:200-220
inline void se3_log(const SE3& T, double out[6]) {
  double theta;
  V3 om = so3_log(T.q, &theta);
  M3 Om = hat(om), Om2 = olin::m3_mul(Om, Om), Vi = olin::m3_identity();
  double c = (theta < SMALL_EPS) ? (1. / 12.) : (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Vi.a[i][j] = (Vi.a[i][j] - 0.5 * Om.a[i][j]) + c * Om2.a[i][j];
  V3 ups = olin::m3_mulv(Vi, T.t);
  out[0] = ups.v[0]; out[1] = ups.v[1]; out[2] = ups.v[2];
  out[3] = om.v[0]; out[4] = om.v[1]; out[5] = om.v[2];
}

// se3.cpp:84-91 / so3.cpp:62-69 (quaternion product re-normalised)
inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  V3 rt = quat_rotate(a.q, b.t);
  for (int i = 0; i < 3; i++) r.t.v[i] = a.t.v[i] + rt.v[i];
  r.q = quat_mul(a.q, b.q);
  quat_normalize(r.q);
  return r;
}

// SE3(const Matrix3d&, const Vector3d&): quaternion from matrix, NOT normalised (so3.cpp:40-41)
inline SE3 se3_from_Rt(const M3& R, const V3& t) { SE3 r; r.q = quat_from_matrix(R); r.t = t; return r; }

// SE3::matrix() cast to float, column-major 4x4 (Eigen::Matrix4f layout)
inline void se3_to_matrix4f(const SE3& T, float M[16]) {
  M3 R = quat_to_matrix(T.q);
  for (int c = 0; c < 3; c++) { for (int r = 0; r < 3; r++) M[c * 4 + r] = (float)R.a[r][c]; M[c * 4 + 3] = 0.f; }
  M[12] = (float)T.t.v[0]; M[13] = (float)T.t.v[1]; M[14] = (float)T.t.v[2]; M[15] = 1.f;
}

}  // namespace ose3
