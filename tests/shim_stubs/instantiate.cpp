// What src/ndt_omp/ndt_omp.cpp (pclomp and, with -DLVS_SHIM_GROUND, pclomp_ground) and src/ndt_pca/ndt_pca.cpp become: explicit instantiations for the reference's
// three point types, plus one use of the base-pointer helper.
#include <ndt_b200.h>
#if defined(LVS_SHIM_PCA)
namespace ns = pclpca;
#define NDT_CLASS NormalDistributionsTransform
#elif defined(LVS_SHIM_GROUND)
namespace ns = pclomp_ground;          // src/ndt_omp/ndt_omp.cpp:9-14 instantiates this one next to pclomp
#define NDT_CLASS NormalDistributionsTransformGround
#else
namespace ns = pclomp;
#define NDT_CLASS NormalDistributionsTransform
#endif
template class ns::NDT_CLASS<pcl::PointXYZ, pcl::PointXYZ>;
template class ns::NDT_CLASS<pcl::PointXYZI, pcl::PointXYZI>;
template class ns::NDT_CLASS<pcl::PointXYZRGBL, pcl::PointXYZRGBL>;
double shim_probe(const pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>::Ptr& reg) { return ns::lvs_fitness_score<pcl::PointXYZI>(reg, 2.0); }
