// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for pcl::PointCloud as information_matrix_calculator.{hpp,cpp} use it.
#pragma once
#include <memory>
#include <vector>
#include <Eigen/Core>
namespace pcl {
template <typename P>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<P> > Ptr;
  typedef std::shared_ptr<const PointCloud<P> > ConstPtr;
  std::vector<P> points;
  bool is_dense = true;
};
}  // namespace pcl
