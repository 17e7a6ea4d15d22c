// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for what the reference's own edge classes (include/g2o/edge_se3_prior*.hpp) use of g2o: VertexSE3 with its
// Isometry3D estimate and BaseUnaryEdge<D, E, VertexSE3> with _vertices, _measurement, _error and information().
#pragma once
#include <istream>
#include <ostream>
#include <vector>
#include <Eigen/Core>
namespace g2o {
typedef Eigen::Isometry3d Isometry3D;
struct HyperGraphVertex { virtual ~HyperGraphVertex() {} };
class VertexSE3 : public HyperGraphVertex {
 public:
  const Isometry3D& estimate() const { return _estimate; }
  void setEstimate(const Isometry3D& e) { _estimate = e; }
 private:
  Isometry3D _estimate;
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
 protected:
  std::vector<HyperGraphVertex*> _vertices;
  E _measurement;
  ErrorVector _error;
  InformationType _information;
};
}  // namespace g2o
