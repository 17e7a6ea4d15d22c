// One translation unit that holds all three registration classes, like src/lidar_odometry/scan_matching_odom_nodelet.cpp:24-26,327-329
// (reg_s2s / reg_s2k / ground_s2k): shim/ndt_b200.h is included once per class.
#include <ndt_b200.h>
#define LVS_SHIM_PCA
#include <ndt_b200.h>
#undef LVS_SHIM_PCA
#define LVS_SHIM_GROUND
#include <ndt_b200.h>
#undef LVS_SHIM_GROUND
#include <ndt_b200.h>          // and a repeated inclusion is a no-op

typedef pcl::PointXYZI PointT;
struct NodeletLike {
  pclomp::NormalDistributionsTransform<PointT, PointT> s2s;
  pclpca::NormalDistributionsTransform<PointT, PointT> s2k;
  pclomp_ground::NormalDistributionsTransformGround<PointT, PointT> ground_s2k;
  void configure() {                                   // scan_matching_odom_nodelet.cpp:109-126
    s2k.setResolution(1.0f); s2k.setNumThreads(4); s2k.setNeighborhoodSearchMethod(pclpca::DIRECT1);
    s2k.setTransformationEpsilon(0.01); s2k.setMaximumIterations(64);
    ground_s2k.setResolution(10.0f); ground_s2k.setNumThreads(4); ground_s2k.setNeighborhoodSearchMethod(pclomp_ground::DIRECT1);
    ground_s2k.setTransformationEpsilon(0.01); ground_s2k.setMaximumIterations(64);
    s2s.setNeighborhoodSearchMethod(pclomp::DIRECT7);
  }
};
void nodelet_like_probe() { NodeletLike n; n.configure(); }
