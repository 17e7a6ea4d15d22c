// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for g2o::VertexPlane (estimate = Plane3D); Plane3D itself is g2o's own plane3d.h, unpacked from
// the reference's g2o zip by oracle/build_ref.sh and found on the include path as "plane3d.h".
#pragma once
#include <g2o/types/slam3d/types_slam3d.h>
#include "plane3d.h"
namespace g2o {
class VertexPlane : public HyperGraphVertex {
 public:
  const Plane3D& estimate() const { return _estimate; }
  void setEstimate(const Plane3D& p) { _estimate = p; }
 private:
  Plane3D _estimate;
};
}  // namespace g2o
