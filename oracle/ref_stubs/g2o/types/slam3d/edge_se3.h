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
 protected:
  std::vector<HyperGraphVertex*> _vertices;
  E _measurement;
  ErrorVector _error;
  InformationType _information;
};
}  // namespace g2o
