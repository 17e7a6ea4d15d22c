#pragma once
#include "vertex_se3.h"
namespace g2o {
class EdgeSE3 : public HyperGraph::Edge {
 public:
  const Eigen::Isometry3d& measurement() const { return m_; }
  const Eigen::Matrix<double, 6, 6>& information() const { return i_; }
  RobustKernel* robustKernel() const { return k_; }
 private:
  Eigen::Isometry3d m_;
  Eigen::Matrix<double, 6, 6> i_;
  RobustKernel* k_ = nullptr;
};
}  // namespace g2o
