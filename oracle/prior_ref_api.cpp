// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry point over the reference's OWN unary edge classes - include/g2o/edge_se3_priorxy.hpp, edge_se3_priorxyz.hpp, edge_se3_priorquat.hpp,
// edge_se3_priorvec.hpp and edge_se3_plane.hpp, included as they are - compiled against stand-ins for the g2o and Eigen headers they include (oracle/ref_stubs/):
// setMeasurement followed by computeError on one VertexSE3, i.e. what GraphSLAM::add_se3_prior_*_edge sets up (src/global_graph/graph_slam.cpp:194-240).
#include <g2o/edge_se3_priorxy.hpp>
#include <g2o/edge_se3_priorxyz.hpp>
#include <g2o/edge_se3_priorquat.hpp>
#include <g2o/edge_se3_priorvec.hpp>
#include <g2o/edge_se3_plane.hpp>      // with g2o's own types/slam3d_addons/plane3d.h (unpacked from the vendored zip) behind the VertexPlane stand-in

// g2o's own numeric Jacobians (BaseUnaryEdge / BaseBinaryEdge::linearizeOplus: central differences of 1e-9 through push / oplus / computeError /
// pop) and the MQT mappings behind VertexSE3::oplus, taken from the vendored zip at build time
namespace g2o {
typedef Eigen::Matrix<double, 3, 3> Matrix3D;
typedef Eigen::Matrix<double, 3, 1> Vector3D;
typedef Eigen::Matrix<double, 6, 1> Vector6d;
namespace internal {
inline Isometry3D::ConstLinearPart extractRotation(const Isometry3D& A) { return A.matrix().topLeftCorner<3, 3>(); }
Eigen::Quaterniond& normalize(Eigen::Quaterniond& q);
Vector3D toCompactQuaternion(const Matrix3D& R);
Matrix3D fromCompactQuaternion(const Vector3D& v);
Vector6d toVectorMQT(const Isometry3D& t);
using namespace std;
#include G2O_MAP_BODIES
}  // namespace internal
#include G2O_UNARY_BODIES
#include G2O_BINARY_BODIES
}  // namespace g2o

static g2o::Isometry3D iso_from_qt7(const double* v) {      // x y z qx qy qz qw, quaternion normalised as VertexSE3::read / fromVectorQT do
  Eigen::Quaterniond q(v[6], v[3], v[4], v[5]);
  q.normalize();
  g2o::Isometry3D t;
  t = q.toRotationMatrix();
  t.translation() = Eigen::Vector3d(v[0], v[1], v[2]);
  return t;
}

// kind: 1 xy (meas x y), 2 xyz (x y z), 3 quat (qx qy qz qw), 4 vec (direction 3, measurement 3), 5 plane (measured plane 4, the fixed
// VertexPlane's coefficients 4); e6 zero-padded
extern "C" void pref_prior_error(int kind, const double* meas, const double* x7, double* e6) {
  g2o::VertexSE3 v;
  v.setEstimate(iso_from_qt7(x7));
  for (int i = 0; i < 6; i++) e6[i] = 0.0;
  if (kind == 1) {
    g2o::EdgeSE3PriorXY e; e.vertices()[0] = &v;
    Eigen::Vector2d m; m(0) = meas[0]; m(1) = meas[1];
    e.setMeasurement(m); e.computeError();
    for (int i = 0; i < 2; i++) e6[i] = e.error()(i);
  } else if (kind == 2) {
    g2o::EdgeSE3PriorXYZ e; e.vertices()[0] = &v;
    e.setMeasurement(Eigen::Vector3d(meas[0], meas[1], meas[2])); e.computeError();
    for (int i = 0; i < 3; i++) e6[i] = e.error()(i);
  } else if (kind == 3) {
    g2o::EdgeSE3PriorQuat e; e.vertices()[0] = &v;
    e.setMeasurement(Eigen::Quaterniond(meas[3], meas[0], meas[1], meas[2])); e.computeError();
    for (int i = 0; i < 3; i++) e6[i] = e.error()(i);
  } else if (kind == 5) {
    g2o::VertexPlane vp;
    Eigen::Vector4d pm, pv;
    for (int i = 0; i < 4; i++) { pm(i) = meas[i]; pv(i) = meas[4 + i]; }
    vp.setEstimate(g2o::Plane3D(pv));
    g2o::EdgeSE3Plane e; e.vertices()[0] = &v; e.vertices()[1] = &vp;
    e.setMeasurement(g2o::Plane3D(pm)); e.computeError();
    for (int i = 0; i < 3; i++) e6[i] = e.error()(i);
  } else if (kind == 4) {
    g2o::EdgeSE3PriorVec e; e.vertices()[0] = &v;
    Eigen::Matrix<double, 6, 1> m;
    for (int i = 0; i < 6; i++) m[i] = meas[i];
    e.setMeasurement(m); e.computeError();
    for (int i = 0; i < 3; i++) e6[i] = e.error()(i);
  }
}

// The same edges' linearizeOplus with respect to the pose vertex: J (D x 6) zero-padded to 6 x 6, row-major.  kind 5: the plane vertex is fixed,
// as the nodelet creates it (global_graph_nodelet.cpp:601-611).
extern "C" void pref_prior_jacobian(int kind, const double* meas, const double* x7, double* J36) {
  g2o::VertexSE3 v;
  v.setEstimate(iso_from_qt7(x7));
  for (int i = 0; i < 36; i++) J36[i] = 0.0;
  auto put = [&](const auto& J, int D) { for (int r = 0; r < D; r++) for (int c = 0; c < 6; c++) J36[r * 6 + c] = J(r, c); };
  if (kind == 1) {
    g2o::EdgeSE3PriorXY e; e.vertices()[0] = &v;
    Eigen::Vector2d m; m(0) = meas[0]; m(1) = meas[1];
    e.setMeasurement(m); e.computeError(); e.linearizeOplus(); put(e.jacobianOplusXi(), 2);
  } else if (kind == 2) {
    g2o::EdgeSE3PriorXYZ e; e.vertices()[0] = &v;
    e.setMeasurement(Eigen::Vector3d(meas[0], meas[1], meas[2])); e.computeError(); e.linearizeOplus(); put(e.jacobianOplusXi(), 3);
  } else if (kind == 3) {
    g2o::EdgeSE3PriorQuat e; e.vertices()[0] = &v;
    e.setMeasurement(Eigen::Quaterniond(meas[3], meas[0], meas[1], meas[2])); e.computeError(); e.linearizeOplus(); put(e.jacobianOplusXi(), 3);
  } else if (kind == 4) {
    g2o::EdgeSE3PriorVec e; e.vertices()[0] = &v;
    Eigen::Matrix<double, 6, 1> m;
    for (int i = 0; i < 6; i++) m[i] = meas[i];
    e.setMeasurement(m); e.computeError(); e.linearizeOplus(); put(e.jacobianOplusXi(), 3);
  } else if (kind == 5) {
    g2o::VertexPlane vp;
    Eigen::Vector4d pm, pv;
    for (int i = 0; i < 4; i++) { pm(i) = meas[i]; pv(i) = meas[4 + i]; }
    vp.setEstimate(g2o::Plane3D(pv));
    vp.setFixed(true);
    g2o::EdgeSE3Plane e; e.vertices()[0] = &v; e.vertices()[1] = &vp;
    e.setMeasurement(g2o::Plane3D(pm)); e.computeError(); e.linearizeOplus(); put(e.jacobianOplusXi(), 3);
  }
}
