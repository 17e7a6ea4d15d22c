// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's own Sophus (so3.cpp / se3.cpp of Sophus a621ff2, unpacked from
// /root/reference/3rdtools/Sophus-a621ff2-ubuntu18.04.zip by oracle/build_ref.sh and compiled against oracle/ref_stubs/eigen_min.h):
// exactly the three uses the NDT path makes of it (include/ndt_omp/ndt_omp_impl2.hpp:119-120, :161-163, :166) plus what Sophus' own
// test_se3.cpp exercises (inverse, point transform).  tests/test_oracle_ndt.py holds oracle/ose3.h against these.
#include "se3.h"

using Sophus::SE3;
using Sophus::SO3;
using Sophus::Vector6d;

static Vector6d v6(const double* p) { Vector6d v; for (int i = 0; i < 6; i++) v[i] = p[i]; return v; }
static void put(const SE3& T, double* q_wxyz, double* t3) {
  const Eigen::Quaterniond& q = T.unit_quaternion();
  q_wxyz[0] = q.w(); q_wxyz[1] = q.x(); q_wxyz[2] = q.y(); q_wxyz[3] = q.z();
  for (int i = 0; i < 3; i++) t3[i] = T.translation()[i];
}

extern "C" {

// Sophus::SE3::exp(p)
void sref_se3_exp(const double* p6, double* q_wxyz, double* t3) { put(SE3::exp(v6(p6)), q_wxyz, t3); }

// Sophus::SE3::exp(p).matrix(), row-major 4x4 (the reference casts it to float: transformation_ = ...matrix().cast<float>())
void sref_se3_exp_matrix(const double* p6, double* M16_rowmajor) {
  const Eigen::Matrix4d M = SE3::exp(v6(p6)).matrix();
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) M16_rowmajor[r * 4 + c] = M(r, c);
}

// Sophus::SE3(R, t).log() - how computeTransformation turns the initial guess into its parameter vector
void sref_se3_log_of_Rt(const double* R9_rowmajor, const double* t3, double* out6) {
  Eigen::Matrix3d R;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = R9_rowmajor[r * 3 + c];
  const Vector6d v = SE3(R, Eigen::Vector3d(t3[0], t3[1], t3[2])).log();
  for (int i = 0; i < 6; i++) out6[i] = v[i];
}

// (Sophus::SE3::exp(delta) * Sophus::SE3::exp(p)).log() - the parameter update of every Newton iteration
void sref_compose_log(const double* delta6, const double* p6, double* out6) {
  const Vector6d v = (SE3::exp(v6(delta6)) * SE3::exp(v6(p6))).log();
  for (int i = 0; i < 6; i++) out6[i] = v[i];
}

// log(exp(p)) and exp(p)^-1, p' = exp(p) * x: the operations of Sophus' own test_se3.cpp
void sref_log_exp(const double* p6, double* out6) {
  const Vector6d v = SE3::exp(v6(p6)).log();
  for (int i = 0; i < 6; i++) out6[i] = v[i];
}
void sref_inverse(const double* p6, double* q_wxyz, double* t3) { put(SE3::exp(v6(p6)).inverse(), q_wxyz, t3); }
void sref_transform(const double* p6, const double* x3, double* out3) {
  const Eigen::Vector3d y = SE3::exp(v6(p6)) * Eigen::Vector3d(x3[0], x3[1], x3[2]);
  for (int i = 0; i < 3; i++) out3[i] = y[i];
}

// The nine transformations of Sophus' own test (Sophus/sophus/test_se3.cpp, se3explog_tests) pushed through its three checks - T vs
// exp(log(T)), T * p vs its homogeneous matrix applied to p, T * T^-1 vs identity - with the Frobenius norms it uses.  Returns the largest
// deviation (the test's bound is SMALL_EPS = 1e-10); NaN counts as failure.  It exercises the stand-in Eigen as much as Sophus.
double sref_selftest(void) {
  const double pi = 3.14159265;
  typedef Eigen::Vector3d V;
  SE3 cases[9] = {
      SE3(SO3::exp(V(0.2, 0.5, 0.0)), V(0, 0, 0)), SE3(SO3::exp(V(0.2, 0.5, -1.0)), V(10, 0, 0)), SE3(SO3::exp(V(0., 0., 0.)), V(0, 100, 5)),
      SE3(SO3::exp(V(0., 0., 0.00001)), V(0, 0, 0)), SE3(SO3::exp(V(0., 0., 0.00001)), V(0, -0.00000001, 0.0000000001)),
      SE3(SO3::exp(V(0., 0., 0.00001)), V(0.01, 0, 0)), SE3(SO3::exp(V(pi, 0, 0)), V(4, -5, 0)),
      SE3(SO3::exp(V(0.2, 0.5, 0.0)), V(0, 0, 0)) * SE3(SO3::exp(V(pi, 0, 0)), V(0, 0, 0)) * SE3(SO3::exp(V(-0.2, -0.5, -0.0)), V(0, 0, 0)),
      SE3(SO3::exp(V(0.3, 0.5, 0.1)), V(2, 0, -7)) * SE3(SO3::exp(V(pi, 0, 0)), V(0, 0, 0)) * SE3(SO3::exp(V(-0.3, -0.5, -0.1)), V(0, 6, 0))};
  double worst = 0;
  auto take = [&](double nrm) { if (!(nrm == nrm)) worst = 1e300; else if (nrm > worst) worst = nrm; };
  Eigen::Matrix4d I;
  I.setIdentity();
  for (int i = 0; i < 9; i++) {
    const Eigen::Matrix4d T = cases[i].matrix();
    take((T - SE3::exp(cases[i].log()).matrix()).norm());
    const V p(1, 2, 4);
    take((cases[i] * p - (T.topLeftCorner<3, 3>() * p + T.topRightCorner<3, 1>())).norm());
    take((T * cases[i].inverse().matrix() - I).norm());
  }
  return worst;
}

}  // extern "C"
