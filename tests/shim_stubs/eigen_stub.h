// Minimal stand-in for the Eigen types the shims use (see README.md in this directory).
#pragma once
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
template <typename M>
struct Map {
  typedef typename std::remove_const<M>::type Plain;
  const typename Plain::Scalar* p;
  explicit Map(const typename Plain::Scalar* q) : p(q) {}
  operator Plain() const { Plain m; for (unsigned i = 0; i < sizeof(m.v) / sizeof(m.v[0]); i++) m.v[i] = p[i]; return m; }
};
struct Quaterniond {
  double c[4];                                         // x y z w
  Quaterniond(double w, double x, double y, double z) : c{x, y, z, w} {}
  explicit Quaterniond(const Matrix3d&) : c{0, 0, 0, 1} {}
  void normalize() {}
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  double w() const { return c[3]; }
  Matrix3d toRotationMatrix() const { return Matrix3d::Identity(); }
};
struct Isometry3d {
  Matrix3d R;
  Vector3d t;
  Matrix3d& linear() { return R; }
  const Matrix3d& linear() const { return R; }
  Vector3d& translation() { return t; }
  const Vector3d& translation() const { return t; }
  Matrix4d matrix() const { return Matrix4d::Identity(); }
  static Isometry3d Identity() { Isometry3d T; T.R = Matrix3d::Identity(); return T; }
};
}  // namespace Eigen
