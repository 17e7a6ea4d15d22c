// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for g2o::VertexPlane (estimate = Plane3D); Plane3D itself is g2o's own plane3d.h, unpacked from
// the reference's g2o zip by oracle/build_ref.sh and found on the include path as "plane3d.h".
#pragma once
#include <g2o/types/slam3d/types_slam3d.h>
#include "plane3d.h"
namespace g2o {
class VertexPlane : public HyperGraphVertex {
 public:
  static const int Dimension = 3;
  const Plane3D& estimate() const { return _estimate; }
  void setEstimate(const Plane3D& p) { _estimate = p; }
  bool fixed() const { return _fixed; }
  void setFixed(bool f) { _fixed = f; }
  Eigen::Matrix<double, 3, 3>& A() { return _A; }
  Eigen::Matrix<double, 3, 1>& b() { return _b; }
  void push() { _backup.push_back(_estimate); }
  void pop() { _estimate = _backup.back(); _backup.pop_back(); }
  void oplus(const double* update) { _estimate.oplus(Eigen::Vector3d(update[0], update[1], update[2])); }      // vertex_plane.h oplusImpl
 private:
  Plane3D _estimate;
  std::vector<Plane3D> _backup;
  bool _fixed = false;
  Eigen::Matrix<double, 3, 3> _A;
  Eigen::Matrix<double, 3, 1> _b;
};
}  // namespace g2o
