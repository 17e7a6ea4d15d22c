#pragma once
#include "vertex_se3.h"
namespace g2o {
class EdgeSE3 : public HyperGraph::Edge {
 public:
  const Eigen::Isometry3d& measurement() const { return m_; }
  const Eigen::Matrix<double, 6, 6>& information() const { return i_; }
  RobustKernel* robustKernel() const { return k_; }
  ~EdgeSE3() { delete k_; }
  void setMeasurement(const Eigen::Isometry3d& m) { m_ = m; }
  void setInformation(const Eigen::Matrix<double, 6, 6>& i) { i_ = i; }
  void setRobustKernel(RobustKernel* k) { delete k_; k_ = k; }
 private:
  Eigen::Isometry3d m_;
  Eigen::Matrix<double, 6, 6> i_;
  RobustKernel* k_ = nullptr;
};
}  // namespace g2o
