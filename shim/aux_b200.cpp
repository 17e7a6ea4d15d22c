// Reference-side bindings for the two stages either side of the NDT path.  Like the other shims this
// file needs PCL / Eigen / ROS headers and is compiled on the lv_slam side only.
//
//  * lv_slam::InformationMatrixCalculator::calc_fitness_score (src/global_graph/information_matrix_calculator.cpp:53-87): the body
//    below replaces the kd-tree loop; calc_information_matrix (:27-51) stays as it is and keeps calling it.
//  * lidar_odometry::PrefilteringNodelet::distance_filter + downsample (src/lidar_odometry/prefiltering_nodelet.cpp:164-181,
//    138-148): lvs_prefilter() below replaces the two calls in cloud_callback (:117-127) when downsample_method is VOXELGRID.
#include <global_graph/information_matrix_calculator.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <stdexcept>
#include <vector>
#include "lvslam_b200.h"

namespace lv_slam {

double InformationMatrixCalculator::calc_fitness_score(const pcl::PointCloud<PointT>::ConstPtr& cloud1, const pcl::PointCloud<PointT>::ConstPtr& cloud2,
                                                       const Eigen::Isometry3d& relpose, double max_range) {
  static thread_local lvs_ndt_t* h = nullptr;          // one registration object per calling thread, reused between calls
  if (!h) {
    lvs_ndt_params prm;
    lvs_ndt_default_params(&prm);
    if (lvs_ndt_create(&prm, 0, nullptr, &h) != LVS_OK) throw std::runtime_error(lvs_last_error());
  }
  const Eigen::Matrix4f T = relpose.matrix().cast<float>();     // pcl::transformPointCloud(*cloud2, ..., relpose.cast<float>())
  double score = 0;
  int rc = lvs_ndt_set_target(h, &cloud1->points.data()->x, cloud1->points.size(), sizeof(PointT), 0);
  if (!rc) rc = lvs_ndt_set_source(h, &cloud2->points.data()->x, cloud2->points.size(), sizeof(PointT), 0);
  if (!rc) rc = lvs_ndt_fitness_score(h, T.data(), max_range, &score, nullptr);
  if (rc) throw std::runtime_error(lvs_last_error());
  return score;
}

}  // namespace lv_slam

// distance_filter + VoxelGrid of the prefiltering nodelet in one call.  PointXYZI is 32 bytes: x y z pad intensity pad pad pad, so
// the intensity is gathered into a packed x y z i buffer first (n_fields = 4).
pcl::PointCloud<pcl::PointXYZI>::Ptr lvs_prefilter(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, bool use_distance_filter, double distance_near,
                                                   double distance_far, float downsample_resolution) {
  static thread_local lvs_prefilter_t* pf = nullptr;
  if (!pf && lvs_prefilter_create(0, nullptr, &pf) != LVS_OK) throw std::runtime_error(lvs_last_error());
  const size_t n = cloud->points.size();
  std::vector<float> in(4 * n), out(4 * n);
  for (size_t i = 0; i < n; i++) { const auto& p = cloud->points[i]; in[4 * i] = p.x; in[4 * i + 1] = p.y; in[4 * i + 2] = p.z; in[4 * i + 3] = p.intensity; }
  size_t m = 0;
  int flags = 0;
  if (lvs_prefilter_run(pf, in.data(), n, 16, 4, 0, distance_near, distance_far, use_distance_filter ? 1 : 0, downsample_resolution, out.data(), n, 0, &m, &flags))
    throw std::runtime_error(lvs_last_error());
  pcl::PointCloud<pcl::PointXYZI>::Ptr filtered(new pcl::PointCloud<pcl::PointXYZI>());
  filtered->points.resize(m);
  for (size_t i = 0; i < m; i++) { auto& p = filtered->points[i]; p.x = out[4 * i]; p.y = out[4 * i + 1]; p.z = out[4 * i + 2]; p.intensity = out[4 * i + 3]; }
  filtered->width = (uint32_t)m; filtered->height = 1; filtered->is_dense = false;
  filtered->header = cloud->header;
  return filtered;
}
