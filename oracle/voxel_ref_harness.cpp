// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Runs the REFERENCE'S OWN voxel code: oracle/build_ref.sh lets oracle/extract_ref_functions.py write the definitions of
//   VoxelGridCovariance<PointT>::applyFilter, getNeighborhoodAtPoint (both overloads), getNeighborhoodAtPoint7, getNeighborhoodAtPoint1
// exactly as they stand in /root/reference/include/ndt_omp/voxel_grid_covariance_omp_impl.hpp into a temporary file (REF_VOXEL_BODIES) and
// compiles them here into oracle/_ref/libvoxel_ref.so; with -DREF_PCA the same functions of pclpca::VoxelGridCovariance from
// include/ndt_pca/voxel_grid_covariance_pca_impl.hpp (the per-leaf dimension label and weight, :364-397) into libvoxel_pca_ref.so.
// Supplied by this repository, because PCL and Eigen are not in the image:
//   * oracle/ref_stubs/eigen_min.h   Eigen's interface (the 3 x 3 eigen-solver and inverse are the Jacobi / cofactor routines of oracle/olin.h)
//   * the class below                the member declarations of include/ndt_omp/voxel_grid_covariance_omp.h and of pcl::VoxelGrid that the bodies
//                                    touch - INCLUDING the Leaf constructor's initial values (voxel_grid_covariance_omp.h:98-106: cov_ starts
//                                    as the IDENTITY and applyFilter adds the point products on top of it) and the constructor's settings
//                                    (:203-215: downsample_all_data_ = false, 6 points, 0.01)
//   * pcl::getMinMax3D, pcl::getAllNeighborCellIndices, pcl::PointCloud   PCL 1.8 (not vendored by the reference)
// The kd-tree radius search (pcl::KdTreeFLANN, FLANN) is not part of this library.
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include <Eigen/Core>

#define PCL_WARN(...) do { } while (0)
#define pcl_isfinite(x) std::isfinite(x)

namespace boost { namespace mpl { template <typename T> struct size { static const int value = 4; }; } }

namespace pcl {
struct PointXYZ { float x, y, z; };
struct PCLPointField { uint32_t offset; };
template <typename P>
struct PointCloud {
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = false;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const P& p) { points.push_back(p); }
  P& back() { return points.back(); }
};
// pcl::getMinMax3D (PCL 1.8 common/impl/common.hpp): non-finite points skipped when the cloud is not dense
template <typename P>
void getMinMax3D(const PointCloud<P>& cloud, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt) {
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (size_t i = 0; i < cloud.points.size(); i++) {
    const P& p = cloud.points[i];
    if (!cloud.is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  min_pt = Eigen::Vector4f(mn[0], mn[1], mn[2], 0.0f);
  max_pt = Eigen::Vector4f(mx[0], mx[1], mx[2], 0.0f);
}
template <typename P>
void getMinMax3D(const PointCloud<P>*, const std::string&, float, float, Eigen::Vector4f&, Eigen::Vector4f&, bool) { std::abort(); }   // field filter: never set
template <typename P>
int getFieldIndex(const PointCloud<P>&, const std::string&, std::vector<PCLPointField>&) { return -1; }                               // no rgb field
template <typename P> struct NdCopyPointEigenFunctor { NdCopyPointEigenFunctor(const P&, Eigen::VectorXf&) {} };
template <typename P> struct NdCopyEigenPointFunctor { NdCopyEigenPointFunctor(const Eigen::VectorXf&, P&) {} };
template <typename L, typename F> void for_each_type(F) { std::abort(); }                                                              // downsample_all_data_ is false
// pcl::getAllNeighborCellIndices (PCL 1.8 filters/voxel_grid.h): the 13 "half" offsets, then their negatives; no centre cell
inline Eigen::MatrixXi getAllNeighborCellIndices() {
  Eigen::MatrixXi m(3, 26);
  int k = 0;
  for (int i = -1; i < 2; i++) for (int j = -1; j < 2; j++) { m(0, k) = i; m(1, k) = j; m(2, k) = -1; k++; }
  for (int i = -1; i < 2; i++) { m(0, k) = i; m(1, k) = -1; m(2, k) = 0; k++; }
  m(0, k) = -1; m(1, k) = 0; m(2, k) = 0; k++;
  for (int h = 0; h < 13; h++) for (int a = 0; a < 3; a++) m(a, 13 + h) = -m(a, h);
  return m;
}
}  // namespace pcl

#ifdef REF_PCA
#define REF_NS pclpca
#else
#define REF_NS pclomp
#endif

namespace REF_NS {

// Declarations of include/ndt_omp/voxel_grid_covariance_omp.h:56-604 (and of the pcl::VoxelGrid members it pulls in with `using`).
template <typename PointT>
class VoxelGridCovariance {
 public:
  typedef pcl::PointCloud<PointT> PointCloud;
  typedef int FieldList;
  struct Leaf {          // voxel_grid_covariance_omp.h:92-195, constructor :98-106
    Leaf() : nr_points(0), mean_(Eigen::Vector3d::Zero()), centroid(), cov_(Eigen::Matrix3d::Identity()), icov_(Eigen::Matrix3d::Zero()),
             evecs_(Eigen::Matrix3d::Identity()), evals_(Eigen::Vector3d::Zero()), dimension_features_(Eigen::Vector3d::Zero()), dimension_label_(0),
             dimension_2d_(0) {}
    Eigen::Vector3d dimension_features_;      // pclpca only (voxel_grid_covariance_pca.h:141-143, 252-262)
    int dimension_label_;
    double dimension_2d_;
    int getDimension2d() const { return (dimension_2d_); }      // sic: an int getter over a double member (voxel_grid_covariance_pca.h:222-226)
    int nr_points;
    Eigen::Vector3d mean_;
    Eigen::VectorXf centroid;
    Eigen::Matrix3d cov_, icov_, evecs_;
    Eigen::Vector3d evals_;
  };
  typedef Leaf* LeafPtr;
  typedef const Leaf* LeafConstPtr;

  // pcl::VoxelGrid / pcl::Filter members
  const PointCloud* input_ = nullptr;
  std::string filter_field_name_;
  double filter_limit_min_ = -std::numeric_limits<float>::max(), filter_limit_max_ = std::numeric_limits<float>::max();
  bool filter_limit_negative_ = false;
  Eigen::Vector4f leaf_size_, inverse_leaf_size_;
  Eigen::Vector4i min_b_, max_b_, div_b_, divb_mul_;
  bool downsample_all_data_ = false, save_leaf_layout_ = false;      // voxel_grid_covariance_omp.h:211-212
  std::vector<int> leaf_layout_;
  // VoxelGridCovariance members, constructor values of :203-208
  bool searchable_ = true;
  int min_points_per_voxel_ = 6;
  double min_covar_eigvalue_mult_ = 0.01;
  std::map<size_t, Leaf> leaves_;
  std::vector<int> voxel_centroids_leaf_indices_;
  std::string getClassName() const { return "VoxelGridCovariance"; }

  void setLeafSize(float l) {      // pcl::VoxelGrid::setLeafSize(lx, ly, lz): inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
    leaf_size_ = Eigen::Vector4f(l, l, l, 1.0f);
    inverse_leaf_size_ = Eigen::Vector4f(1.0f / l, 1.0f / l, 1.0f / l, 1.0f);
  }

  void applyFilter(PointCloud& output);
  int getNeighborhoodAtPoint(const Eigen::MatrixXi&, const PointT& reference_point, std::vector<LeafConstPtr>& neighbors) const;
  int getNeighborhoodAtPoint(const PointT& reference_point, std::vector<LeafConstPtr>& neighbors) const;
  int getNeighborhoodAtPoint7(const PointT& reference_point, std::vector<LeafConstPtr>& neighbors) const;
  int getNeighborhoodAtPoint1(const PointT& reference_point, std::vector<LeafConstPtr>& neighbors) const;
};

}  // namespace REF_NS

#include REF_VOXEL_BODIES      // the reference's own definitions of the member functions declared above

typedef REF_NS::VoxelGridCovariance<pcl::PointXYZ> RefGrid;
struct VoxHandle {
  RefGrid grid;
  pcl::PointCloud<pcl::PointXYZ> target, centroids;
};

extern "C" {

void* vref_create(void) { return new VoxHandle(); }
void vref_destroy(void* h) { delete (VoxHandle*)h; }

// setInputCloud + setLeafSize + filter(): returns the number of occupied cells
int vref_build(void* h, const float* xyz, size_t n, size_t stride_floats, float leaf, int min_points, double eig_mult) {
  VoxHandle& H = *(VoxHandle*)h;
  H.target.points.resize(n);
  for (size_t i = 0; i < n; i++) { H.target.points[i].x = xyz[i * stride_floats]; H.target.points[i].y = xyz[i * stride_floats + 1]; H.target.points[i].z = xyz[i * stride_floats + 2]; }
  H.target.is_dense = false;
  H.grid.input_ = &H.target;
  H.grid.setLeafSize(leaf);
  H.grid.min_points_per_voxel_ = min_points;
  H.grid.min_covar_eigvalue_mult_ = eig_mult;
  H.grid.applyFilter(H.centroids);
  return (int)H.grid.leaves_.size();
}

void vref_get_grid(void* h, int32_t* min_b, int32_t* max_b, int32_t* div_b) {
  RefGrid& g = ((VoxHandle*)h)->grid;
  for (int a = 0; a < 3; a++) { min_b[a] = g.min_b_[a]; max_b[a] = g.max_b_[a]; div_b[a] = g.div_b_[a]; }
}

// all occupied cells, ascending key; any pointer may be NULL
void vref_get_pca(void* h, int32_t* label, int32_t* weight) {      // pclpca: dimension label and what getDimension2d() returns, per cell
  RefGrid& g = ((VoxHandle*)h)->grid;
  size_t k = 0;
  for (auto it = g.leaves_.begin(); it != g.leaves_.end(); ++it, ++k) { label[k] = it->second.dimension_label_; weight[k] = it->second.getDimension2d(); }
}

void vref_get_leaves(void* h, int32_t* keys, int32_t* nr_points, double* mean3, double* cov9, double* icov9, double* evecs9, double* evals3, float* centroid3) {
  RefGrid& g = ((VoxHandle*)h)->grid;
  size_t k = 0;
  for (auto it = g.leaves_.begin(); it != g.leaves_.end(); ++it, ++k) {
    const RefGrid::Leaf& l = it->second;
    if (keys) keys[k] = (int32_t)it->first;
    if (nr_points) nr_points[k] = l.nr_points;
    for (int i = 0; i < 3; i++) {
      if (mean3) mean3[k * 3 + i] = l.mean_[i];
      if (evals3) evals3[k * 3 + i] = l.evals_[i];
      if (centroid3) centroid3[k * 3 + i] = l.centroid.size() > i ? l.centroid[i] : 0.0f;
      for (int j = 0; j < 3; j++) {
        if (cov9) cov9[k * 9 + i * 3 + j] = l.cov_(i, j);
        if (icov9) icov9[k * 9 + i * 3 + j] = l.icov_(i, j);
        if (evecs9) evecs9[k * 9 + i * 3 + j] = l.evecs_(i, j);
      }
    }
  }
}

// keys of the cells the direct search returns for a point, in the order the reference pushes them.  mode: 1 = DIRECT26, 2 = DIRECT7, 3 = DIRECT1
int vref_neighbours(void* h, const float* xyz3, int mode, int32_t* keys_out /* >= 26 */) {
  RefGrid& g = ((VoxHandle*)h)->grid;
  pcl::PointXYZ p = {xyz3[0], xyz3[1], xyz3[2]};
  std::vector<RefGrid::LeafConstPtr> nb;
  if (mode == 1) g.getNeighborhoodAtPoint(p, nb); else if (mode == 2) g.getNeighborhoodAtPoint7(p, nb); else g.getNeighborhoodAtPoint1(p, nb);
  for (size_t i = 0; i < nb.size(); i++) {
    int32_t key = -1;
    for (auto it = g.leaves_.begin(); it != g.leaves_.end(); ++it) if (&it->second == nb[i]) { key = (int32_t)it->first; break; }
    keys_out[i] = key;
  }
  return (int)nb.size();
}

}  // extern "C"
