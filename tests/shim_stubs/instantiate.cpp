// What src/ndt_omp/ndt_omp.cpp and src/ndt_pca/ndt_pca.cpp become: explicit instantiations for the reference's
// three point types, plus one use of the base-pointer helper.
#include <ndt_b200.h>
#ifdef LVS_SHIM_PCA
namespace ns = pclpca;
#else
namespace ns = pclomp;
#endif
template class ns::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ>;
template class ns::NormalDistributionsTransform<pcl::PointXYZI, pcl::PointXYZI>;
template class ns::NormalDistributionsTransform<pcl::PointXYZRGBL, pcl::PointXYZRGBL>;
double shim_probe(const pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>::Ptr& reg) { return ns::lvs_fitness_score<pcl::PointXYZI>(reg, 2.0); }
