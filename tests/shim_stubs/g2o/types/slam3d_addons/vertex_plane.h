// Stand-ins for g2o::Plane3D / VertexPlane as far as the shim touches them (toVector, estimate, fixed).
#pragma once
#include <cmath>
#include "../slam3d/vertex_se3.h"
namespace g2o {
class Plane3D {
 public:
  Plane3D() { c_(2, 0) = 1.0; }
  explicit Plane3D(const Eigen::Vector4d& v) : c_(v) {      // fromVector: scaled to a unit normal (plane3d.h)
    const double n = std::sqrt(v(0, 0) * v(0, 0) + v(1, 0) * v(1, 0) + v(2, 0) * v(2, 0));
    for (int a = 0; a < 4; a++) c_(a, 0) = v(a, 0) * (1. / n);
  }
  Eigen::Vector4d toVector() const { return c_; }
 private:
  Eigen::Vector4d c_;
};
class VertexPlane : public HyperGraph::Vertex {
 public:
  const Plane3D& estimate() const { return e_; }
  void setEstimate(const Plane3D& e) { e_ = e; }
  bool fixed() const { return f_; }
  void setFixed(bool f) { f_ = f; }
 private:
  Plane3D e_;
  bool f_ = false;
};
}  // namespace g2o
