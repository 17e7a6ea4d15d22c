// Stand-in for the reference's EdgeSE3Plane (include/g2o/edge_se3_plane.hpp) as far as the shim touches it.
#pragma once
#include "types/slam3d_addons/vertex_plane.h"
namespace g2o {
class EdgeSE3Plane : public HyperGraph::Edge {
 public:
  const Plane3D& measurement() const { return m_; }
  const Eigen::Matrix<double, 3, 3>& information() const { return i_; }
  RobustKernel* robustKernel() const { return k_; }
  ~EdgeSE3Plane() { delete k_; }
  void setMeasurement(const Plane3D& m) { m_ = m; }
  void setInformation(const Eigen::Matrix<double, 3, 3>& i) { i_ = i; }
  void setRobustKernel(RobustKernel* k) { delete k_; k_ = k; }
 private:
  Plane3D m_;
  Eigen::Matrix<double, 3, 3> i_;
  RobustKernel* k_ = nullptr;
};
}  // namespace g2o
