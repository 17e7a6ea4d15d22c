// Minimal stand-in for the Eigen types the shims use (see README.md in this directory).  The few operations the shims actually
// perform (quaternion <-> rotation matrix, Isometry3d::matrix) are implemented for real, so that the shims can be EXECUTED against these
// stand-ins (tests/test_shim_exec_gpu.py), not only compiled.
#pragma once
#include <cmath>
#include <type_traits>
namespace Eigen {
template <typename T, int R, int C>
struct Matrix {
  typedef T Scalar;
  T v[R * C];                                          // column-major, like Eigen's default
  Matrix() : v{} {}
  Matrix(T a, T b, T c) : v{} { static_assert(R * C == 3, "vector3 constructor"); v[0] = a; v[1] = b; v[2] = c; }
  T* data() { return v; }
  const T* data() const { return v; }
  T& operator()(int r, int c) { return v[c * R + r]; }
  const T& operator()(int r, int c) const { return v[c * R + r]; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { return v[2]; }
  const T& x() const { return v[0]; }
  const T& y() const { return v[1]; }
  const T& z() const { return v[2]; }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; for (int i = 0; i < R * C; i++) m.v[i] = (U)v[i]; return m; }
  static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
};
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
template <typename M>
struct Map {
  typedef typename std::remove_const<M>::type Plain;
  const typename Plain::Scalar* p;
  explicit Map(const typename Plain::Scalar* q) : p(q) {}
  operator Plain() const { Plain m; for (unsigned i = 0; i < sizeof(m.v) / sizeof(m.v[0]); i++) m.v[i] = p[i]; return m; }
};
struct Quaterniond {
  double c[4];                                         // x y z w
  Quaterniond() : c{0, 0, 0, 1} {}
  Quaterniond(double w, double x, double y, double z) : c{x, y, z, w} {}
  explicit Quaterniond(const Matrix3d& m) : c{0, 0, 0, 1} {    // Eigen 3.3 Quaternion = rotation matrix (Geometry/Quaternion.h, quaternionbase_assign_impl)
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      c[3] = 0.5 * t;
      t = 0.5 / t;
      c[0] = (m(2, 1) - m(1, 2)) * t; c[1] = (m(0, 2) - m(2, 0)) * t; c[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
      c[i] = 0.5 * t;
      t = 0.5 / t;
      c[3] = (m(k, j) - m(j, k)) * t;
      c[j] = (m(j, i) + m(i, j)) * t;
      c[k] = (m(k, i) + m(i, k)) * t;
    }
  }
  void normalize() { const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); for (double& v : c) v /= n; }
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  double w() const { return c[3]; }
  Matrix3d toRotationMatrix() const {                   // Eigen 3.3 QuaternionBase::toRotationMatrix
    Matrix3d r;
    const double tx = 2 * c[0], ty = 2 * c[1], tz = 2 * c[2];
    const double twx = tx * c[3], twy = ty * c[3], twz = tz * c[3], txx = tx * c[0], txy = ty * c[0], txz = tz * c[0], tyy = ty * c[1], tyz = tz * c[1], tzz = tz * c[2];
    r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
    return r;
  }
};
struct Isometry3d {
  Matrix3d R;
  Vector3d t;
  Matrix3d& linear() { return R; }
  const Matrix3d& linear() const { return R; }
  Vector3d& translation() { return t; }
  const Vector3d& translation() const { return t; }
  Matrix4d matrix() const {
    Matrix4d m = Matrix4d::Identity();
    for (int r = 0; r < 3; r++) { for (int c2 = 0; c2 < 3; c2++) m(r, c2) = R(r, c2); m(r, 3) = t.v[r]; }
    return m;
  }
  static Isometry3d Identity() { Isometry3d T; T.R = Matrix3d::Identity(); return T; }
};
}  // namespace Eigen
