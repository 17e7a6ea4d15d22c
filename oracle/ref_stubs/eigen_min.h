// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the Eigen 3 operations that the reference's own code uses on the paths this repository restates, so that this code can be
// compiled in an image that has no Eigen (oracle/build_ref.sh -> oracle/_ref/*.so): Sophus a621ff2's so3.cpp / se3.cpp, the member functions of
// the NDT registration classes and of the voxel grids, information_matrix_calculator.cpp, the prior / plane edges, and g2o's slam3d edge
// math, Huber kernel, quadratic form and numeric Jacobians.  The compiled libraries are the checkers of the restatements in oracle/: they
// pin the reference's formulas, constants, thresholds, branches, index conventions and control flow.  They do NOT pin Eigen's own rounding:
// the kernels below are written here (Eigen 3.3 semantics; evaluation orders as oracle/ndt_oracle.cpp and oracle/pgo_oracle.cpp assume them,
// see DotRule) and the two iterative solvers (3 x 3 eigen-decomposition, 6 x 6 SVD solve) are those of oracle/olin.h.  So an agreement to
// the last bit says "same reference code", not "same Eigen".  tests/test_oracle_ndt.py::test_eigen_stand_in_against_numpy holds the semantics
// of this header against numpy on its own.
//
// Only what those sources touch is provided: fixed-size matrices with eager arithmetic, the comma initialiser, block / corner / head / tail /
// column views, transposes, Map, Quaternion, AngleAxis, Isometry3, and the little of VectorXf / MatrixXi / MatrixXd the voxel code and the
// information matrix use.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <memory>
#include <type_traits>
#include <vector>
#include <ostream>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

template <typename S, int R, int C, int Options = 0>      // Options: Eigen::ColMajor (= 0) is the only value that is ever named
class Matrix;

// Writable view of a fixed RB x CB block of a Matrix<S, R, C>.
template <typename S, int R, int C, int RB, int CB>
class BlockRef {
 public:
  BlockRef(Matrix<S, R, C>& m, int r0, int c0) : m_(m), r0_(r0), c0_(c0) { assert(r0 >= 0 && c0 >= 0 && r0 + RB <= R && c0 + CB <= C); }
  BlockRef& operator=(const Matrix<S, RB, CB>& v) {
    for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) m_(r0_ + r, c0_ + c) = v(r, c);
    return *this;
  }
  BlockRef& operator=(const BlockRef& o) { return *this = static_cast<Matrix<S, RB, CB> >(o); }
  template <int R2, int C2>
  BlockRef& operator=(const BlockRef<S, R2, C2, RB, CB>& o) { return *this = static_cast<Matrix<S, RB, CB> >(o); }
  operator Matrix<S, RB, CB>() const {
    Matrix<S, RB, CB> v;
    for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) v(r, c) = m_(r0_ + r, c0_ + c);
    return v;
  }
  void setIdentity() { for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) m_(r0_ + r, c0_ + c) = r == c ? S(1) : S(0); }
  S norm() const { return static_cast<Matrix<S, RB, CB> >(*this).norm(); }

 private:
  Matrix<S, R, C>& m_;
  int r0_, c0_;
};

// m.block(r0, c0, nr, nc).cast<T>() of a const matrix (run-time sizes): converts to the fixed-size matrix the caller asks for.
template <typename S, int R, int C, typename T>
class DynBlockCast {
 public:
  DynBlockCast(const Matrix<S, R, C>& m, int r0, int c0, int nr, int nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
  template <int RB, int CB>
  operator Matrix<T, RB, CB>() const {
    assert(RB == nr_ && CB == nc_);
    Matrix<T, RB, CB> v;
    for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) v(r, c) = static_cast<T>(m_(r0_ + r, c0_ + c));
    return v;
  }

 private:
  const Matrix<S, R, C>& m_;
  int r0_, c0_, nr_, nc_;
};
// m.block(r0, c0, nr, nc) of a non-const matrix (run-time sizes): assignable from a fixed-size matrix of that shape, setZero()
template <typename S, int R, int C>
class DynBlockRef {
 public:
  DynBlockRef(Matrix<S, R, C>& m, int r0, int c0, int nr, int nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) { assert(r0 >= 0 && c0 >= 0 && r0 + nr <= R && c0 + nc <= C); }
  template <int RB, int CB>
  DynBlockRef& operator=(const Matrix<S, RB, CB>& v) {
    assert(RB == nr_ && CB == nc_);
    for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) m_(r0_ + r, c0_ + c) = v(r, c);
    return *this;
  }
  void setZero() { for (int r = 0; r < nr_; r++) for (int c = 0; c < nc_; c++) m_(r0_ + r, c0_ + c) = S(0); }

 private:
  Matrix<S, R, C>& m_;
  int r0_, c0_, nr_, nc_;
};
template <typename S, int R, int C>
class DynBlock {
 public:
  DynBlock(const Matrix<S, R, C>& m, int r0, int c0, int nr, int nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
  template <typename T>
  DynBlockCast<S, R, C, T> cast() const { return DynBlockCast<S, R, C, T>(m_, r0_, c0_, nr_, nc_); }

 private:
  const Matrix<S, R, C>& m_;
  int r0_, c0_, nr_, nc_;
};

// The float dot of FOUR products reduces pairwise, (t0 + t2) + (t1 + t3), everything else left to right: the evaluation orders the
// restatement in oracle/ndt_oracle.cpp assumes for Eigen 3.3 + SSE (its header says what that inference rests on).  Kept identical here
// on purpose - this file stands in for Eigen's interface, not for its rounding.
template <typename S, int N>
struct DotRule {
  template <typename A, typename B>
  static S run(const A& a, const B& b) { S s = a(0) * b(0); for (int k = 1; k < N; k++) s += a(k) * b(k); return s; }
};
template <typename S>
struct DotRule<S, 4> {
  template <typename A, typename B>
  static S run(const A& a, const B& b) { const S t0 = a(0) * b(0), t1 = a(1) * b(1), t2 = a(2) * b(2), t3 = a(3) * b(3); return (t0 + t2) + (t1 + t3); }
};

// m.transpose(): printable, convertible to a matrix, and as the LEFT factor of a product every coefficient is an inner product (DotRule)
template <typename S, int R, int C>
class Transposed {      // R x C is the shape of the transposed matrix
 public:
  explicit Transposed(const Matrix<S, C, R>& src) { for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) t_(r, c) = src(c, r); }
  operator Matrix<S, R, C>() const { return t_; }
  const Matrix<S, R, C>& eval() const { return t_; }
  template <int C2>
  Matrix<S, R, C2> operator*(const Transposed<S, C, C2>& o) const { return (*this) * o.eval(); }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  friend Transposed operator*(T s, const Transposed& t) { Transposed o(t); o.t_ = static_cast<S>(s) * t.t_; return o; }
  template <int C2>
  Matrix<S, R, C2> operator*(const Matrix<S, C, C2>& o) const {
    Matrix<S, R, C2> m;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C2; c++) {
        struct Row { const Matrix<S, R, C>& t; int r; S operator()(int k) const { return t(r, k); } } a = {t_, r};
        struct Col { const Matrix<S, C, C2>& o; int c; S operator()(int k) const { return o(k, c); } } b = {o, c};
        m(r, c) = DotRule<S, C>::run(a, b);
      }
    return m;
  }

 private:
  Matrix<S, R, C> t_;
};

// Writable view of one column: only head(n) / head<N>() are used on it.
template <typename S, int R, int C>
class ColRef {
 public:
  ColRef(Matrix<S, R, C>& m, int c) : m_(m), c_(c) {}
  template <int N>
  BlockRef<S, R, C, N, 1> head() { return BlockRef<S, R, C, N, 1>(m_, 0, c_); }
  BlockRef<S, R, C, 3, 1> head(int n) { assert(n == 3); (void)n; return BlockRef<S, R, C, 3, 1>(m_, 0, c_); }
  ColRef& operator=(const Matrix<S, R, 1>& v) { for (int r = 0; r < R; r++) m_(r, c_) = v(r); return *this; }

 private:
  Matrix<S, R, C>& m_;
  int c_;
};

template <typename S, int R, int C>
class CommaInit {
 public:
  CommaInit(Matrix<S, R, C>& m, S first) : m_(m), k_(0) { put(first); }
  CommaInit& operator,(S v) { put(v); return *this; }
  Matrix<S, R, C> finished() const { return m_; }
  ~CommaInit() { assert(k_ == R * C); }

 private:
  void put(S v) { assert(k_ < R * C); m_(k_ / C, k_ % C) = v; k_++; }     // row by row, like Eigen
  Matrix<S, R, C>& m_;
  int k_;
};

template <typename S, int R, int C, int Options>
class Matrix {
 public:
  Matrix() {}
  template <typename A, typename B, typename D>
  Matrix(A x, B y, D z) { static_assert(R * C == 3, "three-coefficient constructor"); d_[0] = static_cast<S>(x); d_[1] = static_cast<S>(y); d_[2] = static_cast<S>(z); }
  template <typename A, typename B, typename D, typename E>
  Matrix(A x, B y, D z, E w) {
    static_assert(R * C == 4 && (R == 1 || C == 1), "four-coefficient constructor");
    d_[0] = static_cast<S>(x); d_[1] = static_cast<S>(y); d_[2] = static_cast<S>(z); d_[3] = static_cast<S>(w);
  }
  // a row vector initialises a column vector and vice versa (Eigen transposes vectors on assignment)
  template <int R2, int C2, typename = typename std::enable_if<(R2 == C && C2 == R && R != C && (R == 1 || C == 1))>::type>
  Matrix(const Matrix<S, R2, C2>& o) { for (int i = 0; i < R * C; i++) d_[i] = o.coeff(i); }
  S coeff(int i) const { return d_[i]; }
  int rows() const { return R; }
  int cols() const { return C; }
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };

  S& operator()(int r, int c) { assert(r >= 0 && r < R && c >= 0 && c < C); return d_[r * C + c]; }
  const S& operator()(int r, int c) const { assert(r >= 0 && r < R && c >= 0 && c < C); return d_[r * C + c]; }
  S& operator()(int i) { static_assert(C == 1 || R == 1, "vector access"); return d_[i]; }
  const S& operator()(int i) const { static_assert(C == 1 || R == 1, "vector access"); return d_[i]; }
  S& operator[](int i) { static_assert(C == 1 || R == 1, "vector access"); return d_[i]; }
  const S& operator[](int i) const { static_assert(C == 1 || R == 1, "vector access"); return d_[i]; }
  const S& x() const { return (*this)(0); }
  const S& y() const { return (*this)(1); }
  const S& z() const { return (*this)(2); }

  void setZero() { for (int i = 0; i < R * C; i++) d_[i] = S(0); }
  void setIdentity() { for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) (*this)(r, c) = r == c ? S(1) : S(0); }
  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Zero(int r, int c) { assert(r == R && c == C); (void)r; (void)c; return Zero(); }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }

  CommaInit<S, R, C> operator<<(S first) { return CommaInit<S, R, C>(*this, first); }

  S squaredNorm() const { S s = d_[0] * d_[0]; for (int i = 1; i < R * C; i++) s += d_[i] * d_[i]; return s; }
  S norm() const { return std::sqrt(squaredNorm()); }
  Matrix cross(const Matrix& o) const {
    static_assert(R * C == 3, "cross product");
    return Matrix(d_[1] * o.d_[2] - d_[2] * o.d_[1], d_[2] * o.d_[0] - d_[0] * o.d_[2], d_[0] * o.d_[1] - d_[1] * o.d_[0]);
  }

  Transposed<S, C, R> transpose() const { return Transposed<S, C, R>(*this); }
  template <typename T>
  Matrix<T, R, C> cast() const { Matrix<T, R, C> m; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) m(r, c) = static_cast<T>((*this)(r, c)); return m; }
  Matrix& noalias() { return *this; }
  Matrix& derived() { return *this; }
  const Matrix& derived() const { return *this; }
  void resize(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
  S* data() { return d_; }
  template <int R2, int C2>
  S dot(const Matrix<S, R2, C2>& o) const {
    static_assert((R == 1 || C == 1) && (R2 == 1 || C2 == 1) && R * C == R2 * C2, "dot of two vectors");
    return DotRule<S, R * C>::run(*this, o);
  }
  void normalize() { const S n = norm(); for (int i = 0; i < R * C; i++) d_[i] /= n; }
  Matrix normalized() const { Matrix m(*this); m.normalize(); return m; }
  bool operator==(const Matrix& o) const { for (int i = 0; i < R * C; i++) if (!(d_[i] == o.d_[i])) return false; return true; }
  bool operator!=(const Matrix& o) const { return !(*this == o); }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; i++) d_[i] -= o.d_[i]; return *this; }
  Matrix& operator*=(S s) { for (int i = 0; i < R * C; i++) d_[i] *= s; return *this; }

  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; i++) d_[i] += o.d_[i]; return *this; }
  Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; i++) m.d_[i] = -d_[i]; return m; }
  Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; i++) m.d_[i] = d_[i] + o.d_[i]; return m; }
  Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; i++) m.d_[i] = d_[i] - o.d_[i]; return m; }
  Matrix operator*(S s) const { Matrix m; for (int i = 0; i < R * C; i++) m.d_[i] = d_[i] * s; return m; }
  template <int C2>
  Matrix<S, R, C2> operator*(const Matrix<S, C, C2>& o) const {     // coefficient-wise lazy product of small fixed sizes: k ascending
    Matrix<S, R, C2> m;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C2; c++) {
        if constexpr (R == 1) {      // (row vector) x matrix: every coefficient is an inner product
          struct Col { const Matrix<S, C, C2>& o; int c; S operator()(int k) const { return o(k, c); } } b = {o, c};
          m(r, c) = DotRule<S, C>::run(*this, b);
        } else {
          S s = (*this)(r, 0) * o(0, c);
          for (int k = 1; k < C; k++) s += (*this)(r, k) * o(k, c);
          m(r, c) = s;
        }
      }
    return m;
  }

  template <int C2>
  Matrix<S, R, C2> operator*(const Transposed<S, C, C2>& t) const { return (*this) * t.eval(); }     // column x row^T: outer product (one product per coefficient)
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  Matrix& operator/=(T s) { for (int i = 0; i < R * C; i++) d_[i] /= static_cast<S>(s); return *this; }
  static Matrix UnitX() { static_assert(R * C == 3, "unit vector"); return Matrix(S(1), S(0), S(0)); }
  static Matrix UnitY() { static_assert(R * C == 3, "unit vector"); return Matrix(S(0), S(1), S(0)); }
  static Matrix UnitZ() { static_assert(R * C == 3, "unit vector"); return Matrix(S(0), S(0), S(1)); }
  template <int R0, int C0, int RB, int CB>
  S dot(const BlockRef<S, R0, C0, RB, CB>& b) const { return dot(static_cast<Matrix<S, RB, CB> >(b)); }
  template <int R0, int C0, int CB>
  Matrix<S, R, CB> operator*(const BlockRef<S, R0, C0, C, CB>& b) const { return (*this) * static_cast<Matrix<S, C, CB> >(b); }
  static Matrix Ones() { Matrix m; for (int i = 0; i < R * C; i++) m.d_[i] = S(1); return m; }
  const Matrix& array() const { return *this; }                           // coefficient-wise view: the comparisons below
  struct BoolArray { bool b[R * C]; bool all() const { for (int i = 0; i < R * C; i++) if (!b[i]) return false; return true; } };
  BoolArray operator<=(const Matrix& o) const { BoolArray r; for (int i = 0; i < R * C; i++) r.b[i] = d_[i] <= o.d_[i]; return r; }
  BoolArray operator>=(const Matrix& o) const { BoolArray r; for (int i = 0; i < R * C; i++) r.b[i] = d_[i] >= o.d_[i]; return r; }
  Matrix<S, R, 1> diagonal() const { static_assert(R == C, "diagonal"); Matrix<S, R, 1> v; for (int i = 0; i < R; i++) v(i) = (*this)(i, i); return v; }
  Matrix<S, R, R> asDiagonal() const { static_assert(C == 1, "asDiagonal"); Matrix<S, R, R> m; m.setZero(); for (int i = 0; i < R; i++) m(i, i) = d_[i]; return m; }
  S maxCoeff() const { S v = d_[0]; for (int i = 1; i < R * C; i++) if (d_[i] > v) v = d_[i]; return v; }
  template <typename I> S maxCoeff(I* index) const { int k = 0; for (int i = 1; i < R * C; i++) if (d_[i] > d_[k]) k = i; *index = k; return d_[k]; }      // the first maximum wins
  S minCoeff() const { S v = d_[0]; for (int i = 1; i < R * C; i++) if (d_[i] < v) v = d_[i]; return v; }
  Matrix inverse() const;                                                  // 3 x 3 doubles only (defined after olin.h is included)

  // fixed-size views: a copy from a const object, a writable reference otherwise
  template <int N> Matrix<S, N, 1> head() const { static_assert(C == 1, "head"); return sub<N, 1>(0, 0); }
  template <int N> Matrix<S, N, 1> tail() const { static_assert(C == 1, "tail"); return sub<N, 1>(R - N, 0); }
  template <int N> BlockRef<S, R, C, N, 1> head() { static_assert(C == 1, "head"); return BlockRef<S, R, C, N, 1>(*this, 0, 0); }
  template <int N> BlockRef<S, R, C, N, 1> tail() { static_assert(C == 1, "tail"); return BlockRef<S, R, C, N, 1>(*this, R - N, 0); }
  template <int RB, int CB> Matrix<S, RB, CB> topLeftCorner() const { return sub<RB, CB>(0, 0); }
  template <int RB, int CB> Matrix<S, RB, CB> topRightCorner() const { return sub<RB, CB>(0, C - CB); }
  template <int RB, int CB> Matrix<S, RB, CB> bottomRightCorner() const { return sub<RB, CB>(R - RB, C - CB); }
  template <int RB, int CB> BlockRef<S, R, C, RB, CB> topLeftCorner() { return BlockRef<S, R, C, RB, CB>(*this, 0, 0); }
  template <int RB, int CB> BlockRef<S, R, C, RB, CB> topRightCorner() { return BlockRef<S, R, C, RB, CB>(*this, 0, C - CB); }
  template <int RB, int CB> BlockRef<S, R, C, RB, CB> bottomRightCorner() { return BlockRef<S, R, C, RB, CB>(*this, R - RB, C - CB); }
  DynBlockRef<S, R, C> block(int r0, int c0, int nr, int nc) { return DynBlockRef<S, R, C>(*this, r0, c0, nr, nc); }
  DynBlock<S, R, C> block(int r0, int c0, int nr, int nc) const { return DynBlock<S, R, C>(*this, r0, c0, nr, nc); }
  BlockRef<S, R, C, 3, 3> topLeftCorner(int nr, int nc) { assert(nr == 3 && nc == 3); (void)nr; (void)nc; return BlockRef<S, R, C, 3, 3>(*this, 0, 0); }
  template <int RB, int CB> Matrix<S, RB, CB> block(int r0, int c0) const { return sub<RB, CB>(r0, c0); }
  template <int RB, int CB> BlockRef<S, R, C, RB, CB> block(int r0, int c0) { return BlockRef<S, R, C, RB, CB>(*this, r0, c0); }
  Matrix<S, R, 1> col(int c) const { Matrix<S, R, 1> v; for (int r = 0; r < R; r++) v(r) = (*this)(r, c); return v; }
  ColRef<S, R, C> col(int c) { return ColRef<S, R, C>(*this, c); }

 private:
  template <int RB, int CB>
  Matrix<S, RB, CB> sub(int r0, int c0) const {
    Matrix<S, RB, CB> v;
    for (int r = 0; r < RB; r++) for (int c = 0; c < CB; c++) v(r, c) = (*this)(r0 + r, c0 + c);
    return v;
  }
  S d_[R * C];
};

template <typename T, typename S, int R, int C, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
inline Matrix<S, R, C> operator*(T s, const Matrix<S, R, C>& m) { Matrix<S, R, C> o; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) o(r, c) = static_cast<S>(s) * m(r, c); return o; }
template <typename T, typename S, int R, int C, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
inline Matrix<S, R, C> operator/(const Matrix<S, R, C>& m, T s) { Matrix<S, R, C> o; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) o(r, c) = m(r, c) / static_cast<S>(s); return o; }

// the inline stream operators of so3.h / se3.h (never called by the checker)
template <typename S, int R, int C>
inline std::ostream& operator<<(std::ostream& o, const Matrix<S, R, C>& m) {
  for (int r = 0; r < R; r++) { for (int c = 0; c < C; c++) o << (c ? " " : "") << m(r, c); if (r + 1 < R) o << "\n"; }
  return o;
}

template <typename S, int R, int C>
inline std::ostream& operator<<(std::ostream& o, const Transposed<S, R, C>& t) { return o << t.eval(); }

typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;

// Eigen::Quaternion<S>, coefficients (x, y, z, w) with the (w, x, y, z) constructor; Geometry/Quaternion.h semantics.
template <typename S>
class Quaternion {
 public:
  Quaternion() {}
  Quaternion(S w, S x, S y, S z) : x_(x), y_(y), z_(z), w_(w) {}
  explicit Quaternion(const Matrix<S, 3, 3>& m) {          // quaternionbase_assign_impl<Other, 3, 3>: not normalised
    S t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w_ = S(0.5) * t;
      t = S(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t;
      y_ = (m(0, 2) - m(2, 0)) * t;
      z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + S(1.0));
      S v[3];
      v[i] = S(0.5) * t;
      t = S(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t;
      v[j] = (m(j, i) + m(i, j)) * t;
      v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  S w() const { return w_; }
  S x() const { return x_; }
  S y() const { return y_; }
  S z() const { return z_; }
  S& w() { return w_; }
  S& x() { return x_; }
  S& y() { return y_; }
  S& z() { return z_; }
  Matrix<S, 4, 1> coeffs() const { return Matrix<S, 4, 1>(x_, y_, z_, w_); }      // Eigen's coefficient order: x, y, z, w
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(x_, y_, z_); }
  struct Coeffs {            // q.coeffs() of a non-const quaternion: (x, y, z, w), writable
    Quaternion& q;
    template <typename T> Coeffs& operator*=(T s) { q.x_ *= s; q.y_ *= s; q.z_ *= s; q.w_ *= s; return *this; }
    template <int N> Matrix<S, 3, 1> head() const { static_assert(N == 3, "head<3>"); return Matrix<S, 3, 1>(q.x_, q.y_, q.z_); }
    operator Matrix<S, 4, 1>() const { return Matrix<S, 4, 1>(q.x_, q.y_, q.z_, q.w_); }
    Matrix<S, 4, 1> operator-() const { return Matrix<S, 4, 1>(-q.x_, -q.y_, -q.z_, -q.w_); }
    Coeffs& operator=(const Matrix<S, 4, 1>& v) { q.x_ = v(0); q.y_ = v(1); q.z_ = v(2); q.w_ = v(3); return *this; }
    S dot(const Matrix<S, 4, 1>& o) const { return static_cast<Matrix<S, 4, 1> >(*this).dot(o); }
  };
  Coeffs coeffs() { return Coeffs{*this}; }
  void setIdentity() { x_ = y_ = z_ = S(0); w_ = S(1); }
  S squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
  S norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const S n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                      a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
  // v + w * uv + vec x uv with uv = 2 (vec x v)
  Matrix<S, 3, 1> _transformVector(const Matrix<S, 3, 1>& v) const {
    Matrix<S, 3, 1> uv = vec().cross(v);
    uv += uv;
    const Matrix<S, 3, 1> c = vec().cross(uv);
    return Matrix<S, 3, 1>(v(0) + w_ * uv(0) + c(0), v(1) + w_ * uv(1) + c(1), v(2) + w_ * uv(2) + c(2));
  }
  Matrix<S, 3, 3> toRotationMatrix() const {
    const S tx = S(2) * x_, ty = S(2) * y_, tz = S(2) * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const S txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const S tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    Matrix<S, 3, 3> r;
    r(0, 0) = S(1) - (tyy + tzz); r(0, 1) = txy - twz;          r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;          r(1, 1) = S(1) - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;          r(2, 1) = tyz + twx;          r(2, 2) = S(1) - (txx + tyy);
    return r;
  }

 private:
  S x_, y_, z_, w_;
};
typedef Quaternion<double> Quaterniond;

// Eigen::AngleAxisd: only `AngleAxisd(a, axis) * AngleAxisd(b, axis2)` followed by toRotationMatrix() is used (g2o's Plane3D::rotation):
// the product of two angle-axis rotations is the product of their quaternions (cos(a/2), sin(a/2) axis)
template <typename S>
class AngleAxis {
 public:
  AngleAxis(S angle, const Matrix<S, 3, 1>& axis) : angle_(angle), axis_(axis) {}
  Quaternion<S> toQuaternion() const { const S s = std::sin(S(0.5) * angle_), c = std::cos(S(0.5) * angle_); return Quaternion<S>(c, s * axis_(0), s * axis_(1), s * axis_(2)); }
  Quaternion<S> operator*(const AngleAxis& o) const { return toQuaternion() * o.toQuaternion(); }

 private:
  S angle_;
  Matrix<S, 3, 1> axis_;
};
typedef AngleAxis<double> AngleAxisd;

typedef Matrix<int, 4, 1> Vector4i;
typedef Matrix<int, 4, 1> Array4i;          // only constructed from a Vector4i and compared coefficient-wise

// Eigen::VectorXf as VoxelGridCovariance::Leaf::centroid uses it
class VectorXf {
 public:
  VectorXf() {}
  static VectorXf Zero(int n) { VectorXf v; v.d_.assign((size_t)n, 0.0f); return v; }
  void resize(int n) { d_.resize((size_t)n); }
  void setZero() { for (size_t i = 0; i < d_.size(); i++) d_[i] = 0.0f; }
  float& operator[](int i) { return d_[(size_t)i]; }
  const float& operator[](int i) const { return d_[(size_t)i]; }
  VectorXf& operator+=(const VectorXf& o) { assert(o.d_.size() == d_.size()); for (size_t i = 0; i < d_.size(); i++) d_[i] += o.d_[i]; return *this; }
  VectorXf& operator/=(float s) { for (size_t i = 0; i < d_.size(); i++) d_[i] /= s; return *this; }
  struct Head4 { VectorXf& v; Head4& operator+=(const Matrix<float, 4, 1>& p) { for (int i = 0; i < 4; i++) v.d_[(size_t)i] += p(i); return *this; } };
  template <int N> Head4 head() { static_assert(N == 4, "head<4>"); assert(d_.size() >= 4); return Head4{*this}; }
  int size() const { return (int)d_.size(); }

 private:
  std::vector<float> d_;
};

// Eigen::MatrixXi as the neighbour-offset tables use it (3 x n)
class MatrixXi {
 public:
  struct Col3 { int v[3]; };
  MatrixXi() : r_(0), c_(0) {}
  MatrixXi(int r, int c) : r_(r), c_(c), d_((size_t)r * c) {}
  static MatrixXi Zero(int r, int c) { MatrixXi m(r, c); m.setZero(); return m; }
  void setZero() { for (size_t i = 0; i < d_.size(); i++) d_[i] = 0; }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int& operator()(int r, int c) { return d_[(size_t)c * r_ + r]; }
  int operator()(int r, int c) const { return d_[(size_t)c * r_ + r]; }
  Col3 col(int c) const { assert(r_ == 3); Col3 o; for (int r = 0; r < 3; r++) o.v[r] = (*this)(r, c); return o; }

 private:
  int r_, c_;
  std::vector<int> d_;
};
// (Eigen::Vector4i () << relative_coordinates.col (ni), 0).finished ()
struct CommaInit4i {
  Matrix<int, 4, 1> m;
  int k;
  CommaInit4i& operator,(int v) { assert(k < 4); m(k++) = v; return *this; }
  Matrix<int, 4, 1> finished() const { assert(k == 4); return m; }
};
inline CommaInit4i operator<<(Matrix<int, 4, 1> m, const MatrixXi::Col3& c) { CommaInit4i ci{m, 3}; for (int i = 0; i < 3; i++) ci.m(i) = c.v[i]; return ci; }

// Eigen::Isometry3d / Isometry3f as information_matrix_calculator.cpp and g2o's slam3d code use them.  Product and inverse in the
// Isometry mode of Eigen::Transform: (R_a R_b, R_a t_b + t_a) and (R^T, -(R^T t)), sums left to right.
template <typename S>
class Isometry3 {
 public:
  typedef Matrix<S, 3, 3> ConstLinearPart;
  typedef Matrix<S, 3, 1> ConstTranslationPart;
  Isometry3() { m_.setIdentity(); }
  Matrix<S, 4, 4>& matrix() { return m_; }
  const Matrix<S, 4, 4>& matrix() const { return m_; }
  template <typename T>
  Isometry3<T> cast() const { Isometry3<T> o; o.matrix() = m_.template cast<T>(); return o; }
  Matrix<S, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
  Matrix<S, 3, 3> rotation() const { return linear(); }      // Isometry mode: the linear part is the rotation
  Matrix<S, 3, 1> translation() const { return m_.template block<3, 1>(0, 3); }
  BlockRef<S, 4, 4, 3, 1> translation() { return BlockRef<S, 4, 4, 3, 1>(m_, 0, 3); }
  Isometry3& operator=(const Matrix<S, 3, 3>& R) { m_.setIdentity(); m_.template block<3, 3>(0, 0) = R; return *this; }      // rotation, zero translation
  Isometry3 inverse() const {
    Isometry3 o;
    const Matrix<S, 3, 3> R = linear();
    const Matrix<S, 3, 1> t = translation();
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) o.m_(i, j) = R(j, i);
      o.m_(i, 3) = -((R(0, i) * t(0) + R(1, i) * t(1)) + R(2, i) * t(2));
    }
    return o;
  }
  Isometry3 operator*(const Isometry3& b) const {
    Isometry3 o;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) o.m_(i, j) = (m_(i, 0) * b.m_(0, j) + m_(i, 1) * b.m_(1, j)) + m_(i, 2) * b.m_(2, j);
      o.m_(i, 3) = ((m_(i, 0) * b.m_(0, 3) + m_(i, 1) * b.m_(1, 3)) + m_(i, 2) * b.m_(2, 3)) + m_(i, 3);
    }
    return o;
  }

 private:
  Matrix<S, 4, 4> m_;
};
typedef Isometry3<double> Isometry3d;
typedef Isometry3<float> Isometry3f;

// Eigen::MatrixXd as calc_information_matrix uses it: Identity(6, 6), corner(3, 3).array() /= scalar
class MatrixXd {
 public:
  MatrixXd() : r_(0), c_(0) {}
  static MatrixXd Identity(int r, int c) { MatrixXd m; m.r_ = r; m.c_ = c; m.d_.assign((size_t)r * c, 0.0); for (int i = 0; i < r && i < c; i++) m(i, i) = 1.0; return m; }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int r, int c) { return d_[(size_t)r * c_ + c]; }
  double operator()(int r, int c) const { return d_[(size_t)r * c_ + c]; }
  struct Corner {
    MatrixXd& m; int r0, c0, nr, nc;
    Corner& array() { return *this; }
    template <typename T> Corner& operator/=(T s) { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) m(r0 + r, c0 + c) /= s; return *this; }      // double / T: usual arithmetic conversions
  };
  Corner topLeftCorner(int nr, int nc) { return Corner{*this, 0, 0, nr, nc}; }
  Corner bottomRightCorner(int nr, int nc) { return Corner{*this, r_ - nr, c_ - nc, nr, nc}; }

 private:
  int r_, c_;
  std::vector<double> d_;
};

// Eigen::MatrixBase<Derived> in a parameter list: the matrix itself
template <typename D>
using MatrixBase = D;

// Eigen::Map<Matrix<double, R, C>> over a buffer in Eigen's default (column-major) storage order
template <typename M>
class Map;
template <typename S, int R, int C>
class Map<Matrix<S, R, C> > {
 public:
  explicit Map(S* p) : p_(p) {}
  S& operator()(int r, int c) { return p_[c * R + r]; }
  const S& operator()(int r, int c) const { return p_[c * R + r]; }
  Map& noalias() { return *this; }
  Map& operator=(const Matrix<S, R, C>& m) { for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) (*this)(r, c) = m(r, c); return *this; }
  Matrix<S, R, C> eval() const { Matrix<S, R, C> m; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) m(r, c) = (*this)(r, c); return m; }

 private:
  S* p_;
};
template <typename S, int R, int K, int C>
inline Matrix<S, R, C> operator*(const Matrix<S, R, K>& a, const Map<Matrix<S, K, C> >& b) { return a * b.eval(); }

template <typename T>
using aligned_allocator = std::allocator<T>;

// computeTransformation only stores final_transformation_ into one of these (ndt_omp_impl2.hpp:110-111)
enum { Affine = 2, ColMajor = 0 };
template <typename S, int Dim, int Mode, int Options>
class Transform {
 public:
  Matrix<S, Dim + 1, Dim + 1>& matrix() { return m_; }

 private:
  Matrix<S, Dim + 1, Dim + 1> m_;
};

}  // namespace Eigen

#include "../olin.h"

namespace Eigen {

// Matrix3d::inverse(): Eigen's cofactor inverse of a 3 x 3, as oracle/olin.h restates it
template <>
inline Matrix<double, 3, 3> Matrix<double, 3, 3>::inverse() const {
  olin::M3 a;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a.a[r][c] = (*this)(r, c);
  const olin::M3 i = olin::m3_inverse(a);
  Matrix<double, 3, 3> o;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o(r, c) = i.a[r][c];
  return o;
}

// SelfAdjointEigenSolver<Matrix3d>: ascending eigenvalues, eigenvectors in columns - by the cyclic Jacobi iteration of oracle/olin.h (Eigen's
// tridiagonal QL is not restated anywhere in this repository)
template <typename M>
class SelfAdjointEigenSolver;
template <>
class SelfAdjointEigenSolver<Matrix<double, 3, 3> > {
 public:
  void compute(const Matrix<double, 3, 3>& A) {
    olin::M3 a, v;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a.a[r][c] = A(r, c);
    double ev[3];
    olin::sym3_eig(a, ev, v);
    for (int r = 0; r < 3; r++) { vals_(r) = ev[r]; for (int c = 0; c < 3; c++) vecs_(r, c) = v.a[r][c]; }
  }
  const Matrix<double, 3, 1>& eigenvalues() const { return vals_; }
  const Matrix<double, 3, 3>& eigenvectors() const { return vecs_; }

 private:
  Matrix<double, 3, 1> vals_;
  Matrix<double, 3, 3> vecs_;
};

// JacobiSVD<Matrix<double, 6, 6>>(H, ComputeFullU | ComputeFullV).solve(b): the pseudo-inverse solve with Eigen's rank rule, carried out by the
// one-sided Jacobi iteration of oracle/olin.h (Eigen's own two-sided sweep is not restated anywhere in this repository).
enum { ComputeFullU = 0x04, ComputeFullV = 0x10 };
template <typename M>
class JacobiSVD;
template <>
class JacobiSVD<Matrix<double, 6, 6> > {
 public:
  JacobiSVD(const Matrix<double, 6, 6>& A, unsigned int) { for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) a_[r * 6 + c] = A(r, c); }
  Matrix<double, 6, 1> solve(const Matrix<double, 6, 1>& b) const {
    double bb[6], x[6];
    for (int i = 0; i < 6; i++) bb[i] = b(i);
    olin::svd6_solve(a_, bb, x);
    Matrix<double, 6, 1> r;
    for (int i = 0; i < 6; i++) r(i) = x[i];
    return r;
  }

 private:
  double a_[36];
};

}  // namespace Eigen
