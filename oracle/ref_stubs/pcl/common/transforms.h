// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for pcl::transformPointCloud(in, out, Eigen::Affine3f) (PCL 1.8 common/impl/transforms.hpp, dense cloud):
// x' = m00 x + m01 y + m02 z + m03 left to right in float; the other fields are copied.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <typename P>
void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const Eigen::Isometry3f& T) {
  if (&in != &out) out.points.resize(in.points.size());
  const Eigen::Matrix<float, 4, 4>& m = T.matrix();
  for (size_t i = 0; i < in.points.size(); i++) {
    const P p = in.points[i];
    P q = p;
    q.x = m(0, 0) * p.x + m(0, 1) * p.y + m(0, 2) * p.z + m(0, 3);
    q.y = m(1, 0) * p.x + m(1, 1) * p.y + m(1, 2) * p.z + m(1, 3);
    q.z = m(2, 0) * p.x + m(2, 1) * p.y + m(2, 2) * p.z + m(2, 3);
    out.points[i] = q;
  }
}
}  // namespace pcl
