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
#include G2O_HUBER_BODIES

// An edge whose Jacobians, error and information are handed in, to run g2o's own constructQuadraticForm on them
class GivenBinaryEdge : public BaseBinaryEdge<6, Isometry3D, VertexSE3, VertexSE3> {
 public:
  void computeError() override {}
  bool read(std::istream&) override { return false; }
  bool write(std::ostream&) const override { return false; }
  void give(const double* Ji, const double* Jj, const double* info36, const double* err6) {
    for (int r = 0; r < 6; r++) { _error(r) = err6[r]; for (int c = 0; c < 6; c++) { _jacobianOplusXi(r, c) = Ji[r * 6 + c]; _jacobianOplusXj(r, c) = Jj[r * 6 + c]; _information(r, c) = info36[r * 6 + c]; } }
    _hessian.setZero(); _hessianTransposed.setZero();
  }
};
template <int D>
class GivenUnaryEdge : public BaseUnaryEdge<D, Eigen::Matrix<double, D, 1>, VertexSE3> {
 public:
  void computeError() override {}
  bool read(std::istream&) override { return false; }
  bool write(std::ostream&) const override { return false; }
  void give(const double* J36, const double* info36, const double* err6) {
    for (int r = 0; r < D; r++) { this->_error(r) = err6[r]; for (int c = 0; c < 6; c++) this->_jacobianOplusXi(r, c) = J36[r * 6 + c]; for (int c = 0; c < D; c++) this->_information(r, c) = info36[r * 6 + c]; }
  }
};
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

// g2o's own BaseBinaryEdge::constructQuadraticForm on given Jacobians (row-major 6 x 6), information, error; huber <= 0: no kernel.
// Out: the two diagonal blocks, the two right-hand sides and the off-diagonal block (i, j), all row-major, starting from zero.
extern "C" void pref_quadratic_form_binary(const double* Ji, const double* Jj, const double* info36, const double* err6, double huber, int fixed_i, int fixed_j,
                                           double* Ai36, double* bi6, double* Aj36, double* bj6, double* Hij36) {
  g2o::VertexSE3 vi, vj;
  vi.setFixed(fixed_i != 0); vj.setFixed(fixed_j != 0);
  vi.A().setZero(); vi.b().setZero(); vj.A().setZero(); vj.b().setZero();
  g2o::GivenBinaryEdge e;
  e.vertices()[0] = &vi; e.vertices()[1] = &vj;
  e.give(Ji, Jj, info36, err6);
  g2o::RobustKernelHuber k;
  if (huber > 0) { k._delta = huber; e.setRobustKernel(&k); }
  e.constructQuadraticForm();
  for (int r = 0; r < 6; r++) { bi6[r] = vi.b()(r); bj6[r] = vj.b()(r); for (int c = 0; c < 6; c++) { Ai36[r * 6 + c] = vi.A()(r, c); Aj36[r * 6 + c] = vj.A()(r, c); Hij36[r * 6 + c] = e._hessian(r, c); } }
}

// g2o's own BaseUnaryEdge::constructQuadraticForm for an edge of dimension D (2 or 3) on one VertexSE3
extern "C" void pref_quadratic_form_unary(int D, const double* J36, const double* info36, const double* err6, double huber, double* A36, double* b6) {
  g2o::VertexSE3 v;
  v.A().setZero(); v.b().setZero();
  g2o::RobustKernelHuber k;
  k._delta = huber;
  if (D == 2) { g2o::GivenUnaryEdge<2> e; e.vertices()[0] = &v; e.give(J36, info36, err6); if (huber > 0) e.setRobustKernel(&k); e.constructQuadraticForm(); }
  else { g2o::GivenUnaryEdge<3> e; e.vertices()[0] = &v; e.give(J36, info36, err6); if (huber > 0) e.setRobustKernel(&k); e.constructQuadraticForm(); }
  for (int r = 0; r < 6; r++) { b6[r] = v.b()(r); for (int c = 0; c < 6; c++) A36[r * 6 + c] = v.A()(r, c); }
}
