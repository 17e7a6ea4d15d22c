// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Runs g2o's OWN slam3d edge math, the code behind EdgeSE3::computeError / linearizeOplus and VertexSE3::oplusImpl in the reference's pose-graph
// path.  oracle/build_ref.sh unpacks types/slam3d/{isometry3d_gradients.h, isometry3d_mappings.cpp, dquat2mat.cpp, dquat2mat_maxima_generated.cpp}
// from the reference's 3rdtools/g2o-a48ff8c.zip; oracle/extract_ref_functions.py writes the definitions of
//   skew, skewT (both overloads each), computeEdgeSE3Gradient (both overloads)                       [isometry3d_gradients.h]
//   normalize, toCompactQuaternion, fromCompactQuaternion, toVectorMQT, fromVectorMQT                 [isometry3d_mappings.cpp]
//   RobustKernelHuber::robustify                                                                      [core/robust_kernel_impl.cpp]
// exactly as they stand into temporary files (G2O_GRAD_BODIES, G2O_MAP_BODIES) compiled here, with dquat2mat.cpp as it is, into
// oracle/_ref/libg2o_ref.so.  Written here: the typedefs of g2o/core/eigen_types.h, the one-line extractRotation of isometry3d_mappings.h:46-49,
// the three call sites (edge_se3.cpp:70-75, :92-103, vertex_se3.h:90-99, restated in the entry points below) and the Eigen interface
// (oracle/ref_stubs/eigen_min.h; evaluation orders as oracle/pgo_oracle.cpp assumes them - Eigen's rounding is not pinned).
#include <math.h>
#include <cmath>
#include <Eigen/Core>
#include "dquat2mat.h"

namespace g2o {
typedef Eigen::Isometry3d Isometry3D;                          // g2o/core/eigen_types.h
typedef Eigen::Matrix<double, 3, 3, Eigen::ColMajor> Matrix3D;
typedef Eigen::Matrix<double, 3, 1, Eigen::ColMajor> Vector3D;
typedef Eigen::Matrix<double, 6, 1, Eigen::ColMajor> Vector6d;  // isometry3d_mappings.h:40
namespace internal {
inline Isometry3D::ConstLinearPart extractRotation(const Isometry3D& A) { return A.matrix().topLeftCorner<3, 3>(); }      // isometry3d_mappings.h:46-49
Eigen::Quaterniond& normalize(Eigen::Quaterniond& q);
Vector3D toCompactQuaternion(const Matrix3D& R);
Matrix3D fromCompactQuaternion(const Vector3D& v);
Vector6d toVectorMQT(const Isometry3D& t);
Isometry3D fromVectorMQT(const Vector6d& v);
using namespace std;
#include G2O_MAP_BODIES
#include G2O_GRAD_BODIES
}  // namespace internal
}  // namespace g2o

namespace g2o {
// RobustKernelHuber (core/robust_kernel_impl.{h,cpp}): the one member the pose graph's Huber kernels use, taken from robust_kernel_impl.cpp
class RobustKernelHuber {
 public:
  double _delta = 1.0;
  void robustify(double e, Vector3D& rho) const;
};
#include G2O_HUBER_BODIES
}  // namespace g2o

using g2o::Isometry3D;

// x y z qx qy qz qw -> isometry: fromVectorQT (isometry3d_mappings.cpp:137-142) after the quaternion normalisation EdgeSE3::read / VertexSE3::read apply
static Isometry3D iso_from_qt7(const double* v) {
  Eigen::Quaterniond q(v[6], v[3], v[4], v[5]);
  q.normalize();
  Isometry3D t;
  t = q.toRotationMatrix();
  t.translation() = g2o::Vector3D(v[0], v[1], v[2]);
  return t;
}

extern "C" {

// EdgeSE3::computeError (edge_se3.cpp:70-75): _error = toVectorMQT(_inverseMeasurement * Xi^-1 * Xj)
void gref_edge_error(const double* z7, const double* xi7, const double* xj7, double* e6) {
  const Isometry3D Z = iso_from_qt7(z7), Xi = iso_from_qt7(xi7), Xj = iso_from_qt7(xj7);
  const Isometry3D delta = Z.inverse() * Xi.inverse() * Xj;
  const g2o::Vector6d e = g2o::internal::toVectorMQT(delta);
  for (int i = 0; i < 6; i++) e6[i] = e[i];
}

// EdgeSE3::linearizeOplus (edge_se3.cpp:92-103): computeEdgeSE3Gradient(E, Ji, Jj, Z, Xi, Xj); row-major 6 x 6 out
void gref_edge_jacobians(const double* z7, const double* xi7, const double* xj7, double* Ji36, double* Jj36) {
  const Isometry3D Z = iso_from_qt7(z7), Xi = iso_from_qt7(xi7), Xj = iso_from_qt7(xj7);
  Isometry3D E;
  Eigen::Matrix<double, 6, 6> Ji, Jj;
  g2o::internal::computeEdgeSE3Gradient(E, Ji, Jj, Z, Xi, Xj);
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) { Ji36[r * 6 + c] = Ji(r, c); Jj36[r * 6 + c] = Jj(r, c); }
}

// VertexSE3::oplusImpl (vertex_se3.h:90-99, without the every-1000-calls re-orthogonalisation): X * fromVectorMQT(update); rotation row-major + translation out
void gref_oplus(const double* x7, const double* update6, double* R9, double* t3) {
  const Isometry3D X = iso_from_qt7(x7);
  g2o::Vector6d v;
  for (int i = 0; i < 6; i++) v[i] = update6[i];
  const Isometry3D Y = X * g2o::internal::fromVectorMQT(v);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R9[r * 3 + c] = Y.matrix()(r, c); t3[r] = Y.matrix()(r, 3); }
}

// RobustKernelHuber::robustify (robust_kernel_impl.cpp:65-78): rho, rho', rho'' of the squared error
void gref_huber(double e, double delta, double* rho3) {
  g2o::RobustKernelHuber k;
  k._delta = delta;
  g2o::Vector3D rho;
  k.robustify(e, rho);
  for (int i = 0; i < 3; i++) rho3[i] = rho[i];
}

}  // extern "C"
