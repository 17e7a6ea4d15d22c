// Stand-ins for the reference's unary prior edges (include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp) as far as the shim touches them:
// measurement(), information(), vertices(), robustKernel(), with the setMeasurement semantics of the real classes.
#pragma once
#include <cmath>
#include "types/slam3d/vertex_se3.h"
namespace g2o {
template <int D, typename M>
class PriorEdgeStub : public HyperGraph::Edge {
 public:
  const M& measurement() const { return m_; }
  const Eigen::Matrix<double, D, D>& information() const { return i_; }
  RobustKernel* robustKernel() const { return k_; }
  ~PriorEdgeStub() { delete k_; }
  void setInformation(const Eigen::Matrix<double, D, D>& i) { i_ = i; }
  void setRobustKernel(RobustKernel* k) { delete k_; k_ = k; }
 protected:
  M m_;
 private:
  Eigen::Matrix<double, D, D> i_;
  RobustKernel* k_ = nullptr;
};
class EdgeSE3PriorXY : public PriorEdgeStub<2, Eigen::Matrix<double, 2, 1>> { public: void setMeasurement(const Eigen::Matrix<double, 2, 1>& m) { m_ = m; } };
class EdgeSE3PriorXYZ : public PriorEdgeStub<3, Eigen::Vector3d> { public: void setMeasurement(const Eigen::Vector3d& m) { m_ = m; } };
class EdgeSE3PriorQuat : public PriorEdgeStub<3, Eigen::Quaterniond> {
 public:
  void setMeasurement(const Eigen::Quaterniond& q) { m_ = q.w() < 0.0 ? Eigen::Quaterniond(-q.w(), -q.x(), -q.y(), -q.z()) : q; }      // edge_se3_priorquat.hpp:52-57
};
class EdgeSE3PriorVec : public PriorEdgeStub<3, Eigen::Matrix<double, 6, 1>> {
 public:
  void setMeasurement(const Eigen::Matrix<double, 6, 1>& m) {                                                                          // edge_se3_priorvec.hpp:50-53
    for (int h = 0; h < 2; h++) {
      const double n = std::sqrt(m.v[3 * h] * m.v[3 * h] + m.v[3 * h + 1] * m.v[3 * h + 1] + m.v[3 * h + 2] * m.v[3 * h + 2]);
      for (int a = 0; a < 3; a++) m_.v[3 * h + a] = m.v[3 * h + a] / n;
    }
  }
};
}  // namespace g2o
