#pragma once
#include "../../core/sparse_optimizer.h"
namespace g2o {
class VertexSE3 : public HyperGraph::Vertex {
 public:
  const Eigen::Isometry3d& estimate() const { return e_; }
  void setEstimate(const Eigen::Isometry3d& e) { e_ = e; }
  bool fixed() const { return f_; }
  void setFixed(bool f) { f_ = f; }
 private:
  Eigen::Isometry3d e_;
  bool f_ = false;
};
}  // namespace g2o
