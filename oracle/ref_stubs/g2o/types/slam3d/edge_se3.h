// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in: what include/g2o/edge_se3_plane.hpp uses of g2o's binary edge base class.
#pragma once
#include <g2o/types/slam3d/types_slam3d.h>
namespace g2o {
template <int D, typename E, typename V1, typename V2>
class BaseBinaryEdge {
 public:
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  BaseBinaryEdge() : _vertices(2, nullptr) { _information.setIdentity(); _error.setZero(); }
  virtual ~BaseBinaryEdge() {}
  virtual void computeError() = 0;
  virtual void setMeasurement(const E& m) { _measurement = m; }
  virtual bool read(std::istream& is) = 0;
  virtual bool write(std::ostream& os) const = 0;
  InformationType& information() { return _information; }
  const InformationType& information() const { return _information; }
  const ErrorVector& error() const { return _error; }
  std::vector<HyperGraphVertex*>& vertices() { return _vertices; }
  typedef Eigen::Matrix<double, D, V1::Dimension> JacobianXiOplusType;
  typedef Eigen::Matrix<double, D, V2::Dimension> JacobianXjOplusType;
  void linearizeOplus();                      // g2o's own definitions (core/base_binary_edge.hpp), taken at build time: G2O_BINARY_BODIES
  void constructQuadraticForm();
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
  const JacobianXjOplusType& jacobianOplusXj() const { return _jacobianOplusXj; }
  RobustKernelHuber* robustKernel() const { return _kernel; }
  void setRobustKernel(RobustKernelHuber* k) { _kernel = k; }
  double chi2() const { return _error.dot(information() * _error); }                                             // base_edge.h
  InformationType robustInformation(const Vector3D& rho) { InformationType result = rho[1] * _information; return result; }      // base_edge.h
  RobustKernelHuber* _kernel = nullptr;
  Eigen::Matrix<double, V1::Dimension, V2::Dimension> _hessian;                 // the off-diagonal block the edge owns (mapHessianMemory)
  Eigen::Matrix<double, V2::Dimension, V1::Dimension> _hessianTransposed;
  bool _hessianRowMajor = false;
 protected:
  std::vector<HyperGraphVertex*> _vertices;
  E _measurement;
  ErrorVector _error;
  InformationType _information;
  JacobianXiOplusType _jacobianOplusXi;
  JacobianXjOplusType _jacobianOplusXj;
};
}  // namespace g2o
