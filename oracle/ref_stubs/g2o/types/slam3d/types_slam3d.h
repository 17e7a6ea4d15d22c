// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for what the reference's own edge classes (include/g2o/edge_se3_prior*.hpp) use of g2o: VertexSE3 with its
// Isometry3D estimate and BaseUnaryEdge<D, E, VertexSE3> with _vertices, _measurement, _error and information().
#pragma once
#include <istream>
#include <ostream>
#include <algorithm>
#include <vector>
#include <Eigen/Core>
namespace g2o {
typedef Eigen::Isometry3d Isometry3D;
struct HyperGraphVertex { virtual ~HyperGraphVertex() {} };
typedef Eigen::Matrix<double, 3, 1> Vector3D;
// what constructQuadraticForm uses of a robust kernel; robustify is g2o's own RobustKernelHuber::robustify (G2O_HUBER_BODIES)
class RobustKernelHuber {
 public:
  double _delta = 1.0;
  void robustify(double e, Vector3D& rho) const;
};
namespace internal { Isometry3D fromVectorMQT(const Eigen::Matrix<double, 6, 1>& v); }      // g2o's own (isometry3d_mappings.cpp), compiled in prior_ref_api.cpp
class VertexSE3 : public HyperGraphVertex {
 public:
  static const int Dimension = 6;
  const Isometry3D& estimate() const { return _estimate; }
  void setEstimate(const Isometry3D& e) { _estimate = e; }
  bool fixed() const { return _fixed; }
  void setFixed(bool f) { _fixed = f; }
  Eigen::Matrix<double, 6, 6>& A() { return _A; }      // the vertex's diagonal Hessian block and right-hand side (BaseVertex::A(), b())
  Eigen::Matrix<double, 6, 1>& b() { return _b; }
  void push() { _backup.push_back(_estimate); }
  void pop() { _estimate = _backup.back(); _backup.pop_back(); }
  void oplus(const double* update) {      // VertexSE3::oplusImpl (vertex_se3.h:90-99) without the every-1000-calls re-orthogonalisation
    Eigen::Matrix<double, 6, 1> v;
    for (int i = 0; i < 6; i++) v[i] = update[i];
    _estimate = _estimate * internal::fromVectorMQT(v);
  }
 private:
  Isometry3D _estimate;
  std::vector<Isometry3D> _backup;
  bool _fixed = false;
  Eigen::Matrix<double, 6, 6> _A;
  Eigen::Matrix<double, 6, 1> _b;
};
template <int D, typename E, typename V>
class BaseUnaryEdge {
 public:
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  BaseUnaryEdge() : _vertices(1, nullptr) { _information.setIdentity(); _error.setZero(); }
  virtual ~BaseUnaryEdge() {}
  virtual void computeError() = 0;
  virtual void setMeasurement(const E& m) { _measurement = m; }
  virtual bool read(std::istream& is) = 0;
  virtual bool write(std::ostream& os) const = 0;
  InformationType& information() { return _information; }
  const InformationType& information() const { return _information; }
  const ErrorVector& error() const { return _error; }
  const E& measurement() const { return _measurement; }
  std::vector<HyperGraphVertex*>& vertices() { return _vertices; }
  typedef Eigen::Matrix<double, D, V::Dimension> JacobianXiOplusType;
  void linearizeOplus();                      // g2o's own definitions (core/base_unary_edge.hpp), taken at build time: G2O_UNARY_BODIES
  void constructQuadraticForm();
  const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
  RobustKernelHuber* robustKernel() const { return _kernel; }
  void setRobustKernel(RobustKernelHuber* k) { _kernel = k; }
  double chi2() const { return _error.dot(information() * _error); }                                             // base_edge.h
  InformationType robustInformation(const Vector3D& rho) { InformationType result = rho[1] * _information; return result; }      // base_edge.h (the rho[2] term is commented out there)
  RobustKernelHuber* _kernel = nullptr;
 protected:
  std::vector<HyperGraphVertex*> _vertices;
  E _measurement;
  ErrorVector _error;
  InformationType _information;
  JacobianXiOplusType _jacobianOplusXi;
};
}  // namespace g2o
