// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for pcl::search::KdTree<PointT>::nearestKSearch(point, 1, indices, squared distances): exhaustive search with
// FLANN's L2_Simple accumulation order, (dx*dx + dy*dy) + dz*dz in float; the first point wins a tie.
#pragma once
#include <limits>
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <typename P>
class KdTree {
 public:
  typedef std::shared_ptr<KdTree<P> > Ptr;
  void setInputCloud(const typename pcl::PointCloud<P>::ConstPtr& c) { cloud_ = c; }
  int nearestKSearch(const P& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    (void)k;
    float best = std::numeric_limits<float>::max();
    int at = -1;
    for (size_t i = 0; i < cloud_->points.size(); i++) {
      const P& p = cloud_->points[i];
      const float ex = q.x - p.x, ey = q.y - p.y, ez = q.z - p.z;
      const float d = (ex * ex + ey * ey) + ez * ez;
      if (d < best) { best = d; at = (int)i; }
    }
    idx[0] = at; d2[0] = best;
    return at >= 0 ? 1 : 0;
  }

 private:
  typename pcl::PointCloud<P>::ConstPtr cloud_;
};
} }  // namespace pcl::search
