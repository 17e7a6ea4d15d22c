// include/global_graph/information_matrix_calculator.hpp:14,35
#pragma once
#include <limits>
#include "../pcl/point_cloud.h"
namespace lv_slam {
class InformationMatrixCalculator {
 public:
  using PointT = pcl::PointXYZI;
  static double calc_fitness_score(const pcl::PointCloud<PointT>::ConstPtr& cloud1, const pcl::PointCloud<PointT>::ConstPtr& cloud2,
                                   const Eigen::Isometry3d& relpose, double max_range = std::numeric_limits<double>::max());
};
}  // namespace lv_slam
