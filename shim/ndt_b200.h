// Reference-side binding for the NDT path: a pcl::Registration subclass with the class name, namespace and setters of
// pclomp::NormalDistributionsTransform (/root/reference/include/ndt_omp/ndt_omp.h:69-551) — under LVS_SHIM_PCA, of
// pclpca::NormalDistributionsTransform (include/ndt_pca/ndt_pca.h); under LVS_SHIM_GROUND, of
// pclomp_ground::NormalDistributionsTransformGround (include/ndt_omp/ndt_ground.h:69, the object scan_matching_odom_nodelet.cpp:121-126,329
// configures as ground_s2k) — whose computeTransformation forwards to liblvslam_b200.
// Header-only; needs PCL 1.8 + Eigen, which are NOT in the build image of this repository, so it is compiled only on the
// lv_slam side.  It replaces src/ndt_omp/ndt_omp.cpp / src/ndt_pca/ndt_pca.cpp in lv_slam's CMake targets.
//
// The header may be included several times in one translation unit, once per class (scan_matching_odom_nodelet.cpp:24-26 includes
// ndt_omp.h, ndt_pca.h and ndt_ground.h together):
//     #include <ndt_b200.h>                                        // pclomp::NormalDistributionsTransform
//     #define LVS_SHIM_PCA
//     #include <ndt_b200.h>                                        // pclpca::NormalDistributionsTransform
//     #undef LVS_SHIM_PCA
//     #define LVS_SHIM_GROUND
//     #include <ndt_b200.h>                                        // pclomp_ground::NormalDistributionsTransformGround
//     #undef LVS_SHIM_GROUND
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include "lvslam_b200.h"

#if defined(LVS_SHIM_PCA)
#ifndef LVS_NDT_B200_PCA_DEFINED
#define LVS_NDT_B200_PCA_DEFINED
#define LVS_SHIM_NS pclpca
#define LVS_SHIM_CLASS NormalDistributionsTransform
#define LVS_SHIM_VARIANT LVS_NDT_PCA
#endif
#elif defined(LVS_SHIM_GROUND)
#ifndef LVS_NDT_B200_GROUND_DEFINED
#define LVS_NDT_B200_GROUND_DEFINED
#define LVS_SHIM_NS pclomp_ground
#define LVS_SHIM_CLASS NormalDistributionsTransformGround
#define LVS_SHIM_VARIANT LVS_NDT_GROUND
#endif
#else
#ifndef LVS_NDT_B200_OMP_DEFINED
#define LVS_NDT_B200_OMP_DEFINED
#define LVS_SHIM_NS pclomp
#define LVS_SHIM_CLASS NormalDistributionsTransform
#define LVS_SHIM_VARIANT LVS_NDT_OMP
#endif
#endif

#ifdef LVS_SHIM_NS
namespace LVS_SHIM_NS {

enum NeighborSearchMethod { KDTREE, DIRECT26, DIRECT7, DIRECT1 };   // ndt_omp.h:61

#ifdef LVS_SHIM_PCA
// What pclpca::NormalDistributionsTransform::getTargetCells() (include/ndt_pca/ndt_pca.h:129-133) hands out.  The reference returns its
// VoxelGridCovariance BY VALUE (a std::map of Leaf objects); no caller in lv_slam uses it.  The voxel grid of this implementation lives
// in device memory, so the binding returns the same per-leaf data as flat arrays, one entry per occupied cell in ascending leaf index:
// Leaf::nr_points / mean_ / icov_ / evals_ / centroid and the integer PCA weight getDimension2d() (voxel_grid_covariance_pca.h:128-265).
struct TargetCells {
  int min_b[3], max_b[3], div_b[3];                 // getMinBoxCoordinates / getMaxBoxCoordinates / getNrDivisions
  std::vector<int32_t> keys, nr_points, weight;     // leaf index; raw point count (-1: invalidated leaf); int(scale * |mean|)
  std::vector<double> mean, icov, evals;            // [n][3], [n][9] row-major, [n][3]
  std::vector<float> centroid;                      // [n][3]
};
#endif

template <typename PointSource, typename PointTarget>
class LVS_SHIM_CLASS : public pcl::Registration<PointSource, PointTarget> {
  typedef pcl::Registration<PointSource, PointTarget> Base;
  typedef typename Base::PointCloudSource PointCloudSource;
  typedef typename Base::PointCloudTarget PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;

 public:
  typedef boost::shared_ptr<LVS_SHIM_CLASS<PointSource, PointTarget> > Ptr;

  LVS_SHIM_CLASS() {
    this->reg_name_ = "NormalDistributionsTransform";
    lvs_ndt_default_params(&prm_);
    prm_.variant = LVS_SHIM_VARIANT;
    this->transformation_epsilon_ = prm_.transformation_epsilon;
    this->max_iterations_ = prm_.max_iterations;
    check(lvs_ndt_create(&prm_, 0, nullptr, &h_));
  }
  virtual ~LVS_SHIM_CLASS() { lvs_ndt_destroy(h_); }

  void setNumThreads(int) {}   // OpenMP team size of the CPU implementation
  inline void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    Base::setInputTarget(cloud);
    const PointTarget* p = cloud->points.data();
    check(lvs_ndt_set_target(h_, &p->x, cloud->points.size(), sizeof(PointTarget), 0));
  }
  inline void setInputSource(const PointCloudSourceConstPtr& cloud) {
    Base::setInputSource(cloud);
    const PointSource* p = cloud->points.data();
    check(lvs_ndt_set_source(h_, &p->x, cloud->points.size(), sizeof(PointSource), 0));
  }
  inline void setResolution(float r) { prm_.resolution = r; push(); }
  inline float getResolution() const { return prm_.resolution; }
  inline double getStepSize() const { return prm_.step_size; }
  inline void setStepSize(double s) { prm_.step_size = s; push(); }
  inline double getOulierRatio() const { return prm_.outlier_ratio; }
  inline void setOulierRatio(double o) { prm_.outlier_ratio = o; push(); }
  inline void setNeighborhoodSearchMethod(NeighborSearchMethod m) { prm_.search_method = (int)m; push(); }
  inline double getTransformationProbability() const { return res_.trans_probability; }
  inline int getFinalNumIteration() const { return res_.iterations; }
  // pcl::Registration::getFitnessScore is not virtual; a class-level overload with the same name and arguments hides it for callers
  // that hold this type, and loop_detector.hpp:176,255 (which holds a pcl::Registration::Ptr) calls lvs_fitness_score() below.
  // Evaluated for the final transformation of the last align(), like the base class.
  inline double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double s = 0;
    check(lvs_ndt_fitness_score(h_, nullptr, max_range, &s, nullptr));
    return s;
  }
#ifdef LVS_SHIM_PCA
  TargetCells getTargetCells() const {
    TargetCells c;
    int n = 0;
    check(lvs_ndt_get_grid(h_, c.min_b, c.max_b, c.div_b));
    check(lvs_ndt_num_cells(h_, &n));
    c.keys.resize(n); c.nr_points.resize(n); c.weight.resize(n); c.mean.resize(3 * (size_t)n); c.icov.resize(9 * (size_t)n); c.evals.resize(3 * (size_t)n);
    c.centroid.resize(3 * (size_t)n);
    if (n) check(lvs_ndt_get_cells(h_, c.keys.data(), c.nr_points.data(), c.mean.data(), c.icov.data(), c.evals.data(), c.centroid.data(), c.weight.data()));
    return c;
  }
#endif
  double calculateScore(const PointCloudSource& cloud) const {   // evaluates an already transformed cloud (ndt_omp_impl2.hpp:1007-1040)
    lvs_ndt_t* tmp = nullptr;
    check(lvs_ndt_create(&prm_, 0, nullptr, &tmp));
    const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double s = 0;
    const PointTarget* t = this->target_->points.data();
    int rc = lvs_ndt_set_target(tmp, &t->x, this->target_->points.size(), sizeof(PointTarget), 0);
    if (!rc) rc = lvs_ndt_set_source(tmp, &cloud.points.data()->x, cloud.points.size(), sizeof(PointSource), 0);
    if (!rc) rc = lvs_ndt_calculate_score(tmp, I, &s);
    lvs_ndt_destroy(tmp);
    check(rc);
    return s;
  }

 protected:
  // pcl::Registration::align() calls this after copying the source into `output`.
  virtual void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) {
    prm_.transformation_epsilon = this->transformation_epsilon_;   // setTransformationEpsilon / setMaximumIterations live in the base class
    prm_.max_iterations = this->max_iterations_;
    push();
    check(lvs_ndt_align(h_, guess.data(), &res_));                 // Eigen::Matrix4f is column-major, like the ABI
    this->nr_iterations_ = res_.iterations;
    this->converged_ = res_.converged != 0;
    this->final_transformation_ = Eigen::Map<const Eigen::Matrix4f>(res_.final_transformation);
    // the reference leaves the cloud of the last line-search point in `output`
    std::vector<float> xyz(3 * output.points.size());
    check(lvs_ndt_get_aligned_cloud(h_, xyz.data(), 0));
    for (size_t i = 0; i < output.points.size(); i++) { output.points[i].x = xyz[3 * i]; output.points[i].y = xyz[3 * i + 1]; output.points[i].z = xyz[3 * i + 2]; }
  }

 private:
  void push() { check(lvs_ndt_set_params(h_, &prm_)); }
  static void check(int rc) { if (rc != LVS_OK) throw std::runtime_error(std::string("lvslam_b200: ") + lvs_status_string(rc) + ": " + lvs_last_error()); }
  lvs_ndt_t* h_ = nullptr;
  lvs_ndt_params prm_;
  lvs_ndt_result res_{};
};

// For callers that only hold the base pointer (include/global_graph/loop_detector.hpp:176,255):
//   double score = lvs_fitness_score(registration, fitness_score_max_range);
template <typename PointT>
inline double lvs_fitness_score(const typename pcl::Registration<PointT, PointT>::Ptr& reg, double max_range) {
  if (auto* n = dynamic_cast<LVS_SHIM_CLASS<PointT, PointT>*>(reg.get())) return n->getFitnessScore(max_range);
  return reg->getFitnessScore(max_range);     // any other registration (GICP, ICP) keeps PCL's kd-tree path
}

}  // namespace
#undef LVS_SHIM_NS
#undef LVS_SHIM_CLASS
#undef LVS_SHIM_VARIANT
#endif  // LVS_SHIM_NS
