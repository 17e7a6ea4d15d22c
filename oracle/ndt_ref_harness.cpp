// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Runs the REFERENCE'S OWN NDT member functions.  oracle/build_ref.sh lets oracle/extract_ref_functions.py write the definitions of
//   computeTransformation, computeDerivatives, computePointDerivatives_AngleAxisd (float and double), updateDerivatives, computeHessian,
//   updateHessian, updateIntervalMT, trialValueSelectionMT, computeStepLengthMT, calculateScore
// exactly as they stand in /root/reference/include/ndt_omp/ndt_omp_impl2.hpp into a temporary file (REF_NDT_BODIES) and compiles them here,
// together with the reference's own Sophus (so3.cpp / se3.cpp from the vendored zip), into oracle/_ref/libndt_ref.so.  The same source
// builds two more libraries: -DREF_PCA takes the functions of pclpca::NormalDistributionsTransform from include/ndt_pca/ndt_pca_impl2.hpp
// (libndt_pca_ref.so), -DREF_GROUND those of pclomp_ground::NormalDistributionsTransformGround from include/ndt_omp/ndt_ground_impl.hpp,
// computeDerivatives_seg instead of computeDerivatives (libndt_ground_ref.so).  What those definitions
// need from outside is provided by stand-ins written in this repository, because PCL and Eigen are not in the image:
//   * oracle/ref_stubs/eigen_min.h       the Eigen operations they use (interface only; evaluation orders as the restatement assumes them)
//   * the class below                    the member DECLARATIONS of include/ndt_omp/ndt_omp.h:69-551 and of pcl::Registration that the bodies
//                                        touch, the two one-line inline helpers of ndt_omp.h:479-496, and a no-op computeAngleDerivatives
//                                        (ndt_omp_impl2.hpp:309-420 precomputes Euler-angle tables that only the unused Euler path reads)
//   * VoxelGridAdapter                   the voxel structure is NOT the reference's: the cells come from the caller (the restatement's
//                                        leaves) and the three direct searches / the radius search are implemented here
//   * pcl::transformPointCloud           dense branch of PCL 1.8 (not vendored by the reference)
// So this library pins N4-N9 and N12 of SURVEY.md §8a - the Newton loop, the derivative passes, the per-(point, cell) float math, the
// all-double Hessian, the More-Thuente search, calculateScore - against the reference's text, and does not pin the voxel build (N0-N3)
// or Eigen's rounding.  Single-threaded (the OpenMP pragma is ignored: compiled without -fopenmp).
#include <math.h>      // like <pcl/pcl_macros.h>: libstdc++'s <math.h> does `using std::exp`, so the reference's unqualified exp(float) is expf
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <vector>
#include <Eigen/Core>
#include "se3.h"

static int omp_get_thread_num() { return 0; }      // the reference has the same fallback under #ifndef _OPENMP (ndt_omp_impl2.hpp:190-193)

namespace pcl {
struct PointXYZ { float x, y, z; };
template <typename P>
struct PointCloud {
  std::vector<P> points;
  size_t size() const { return points.size(); }
};
// pcl::transformPointCloud (PCL 1.8 common/impl/transforms.hpp), dense cloud: x' = m00 x + m01 y + m02 z + m03, left to right, in float
template <typename P>
void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); i++) {
    const P p = in.points[i];
    P q = p;
    q.x = T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3);
    q.y = T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3);
    q.z = T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3);
    out.points[i] = q;
  }
}
}  // namespace pcl

#if defined(REF_PCA)
#define REF_NS pclpca
#define REF_CLASS NormalDistributionsTransform
#elif defined(REF_GROUND)
#define REF_NS pclomp_ground
#define REF_CLASS NormalDistributionsTransformGround
#else
#define REF_NS pclomp
#define REF_CLASS NormalDistributionsTransform
#endif

namespace REF_NS {

enum NeighborSearchMethod { KDTREE, DIRECT26, DIRECT7, DIRECT1 };     // ndt_omp.h:61

struct RefLeaf {          // what the bodies read of VoxelGridCovariance::Leaf
  int nr_points;
  int in_cloud;           // pushed to the centroid cloud (>= min_points at build time, even if invalidated later)
  Eigen::Vector3d mean_;
  Eigen::Matrix3d icov_;
  float centroid[3];
  int weight_;            // pclpca: what `int getDimension2d()` returns (voxel_grid_covariance_pca.h:222-226)
  Eigen::Matrix3d evecs_; // pclomp_ground reads the leaf normal from these (ndt_ground_impl.hpp:507-511)
  Eigen::Vector3d evals_;
  Eigen::Vector3d getMean() const { return mean_; }
  Eigen::Matrix3d getInverseCov() const { return icov_; }
  int getDimension2d() const { return weight_; }
  Eigen::Matrix3d getEvecs() const { return evecs_; }
  Eigen::Vector3d getEvals() const { return evals_; }
};

// Stand-in for pclomp::VoxelGridCovariance<PointT>: cells handed in by the caller, searches as voxel_grid_covariance_omp_impl.hpp:373-442
// and voxel_grid_covariance_omp.h:506-534 define them (the 27-cell scan replaces the kd-tree over the centroids, nearest first).
template <typename PointT>
class VoxelGridAdapter {
 public:
  typedef const RefLeaf* LeafConstPtr;
  std::map<size_t, RefLeaf> leaves_;
  int min_b_[3], max_b_[3], divb_mul_[3];
  float leaf_size_ = 1.0f;
  int min_points_per_voxel_ = 6;

  void probe(const int ijk[3], int dx, int dy, int dz, std::vector<LeafConstPtr>& out) const {
    const int d[3] = {dx, dy, dz};
    for (int a = 0; a < 3; a++) if (!(min_b_[a] - ijk[a] <= d[a] && max_b_[a] - ijk[a] >= d[a])) return;
    const int key = (ijk[0] + dx - min_b_[0]) * divb_mul_[0] + (ijk[1] + dy - min_b_[1]) * divb_mul_[1] + (ijk[2] + dz - min_b_[2]) * divb_mul_[2];
    auto it = leaves_.find((size_t)key);
    if (it != leaves_.end() && it->second.nr_points >= min_points_per_voxel_) out.push_back(&it->second);
  }
  void cell_of(const PointT& p, int ijk[3]) const {
    ijk[0] = (int)std::floor(p.x / leaf_size_); ijk[1] = (int)std::floor(p.y / leaf_size_); ijk[2] = (int)std::floor(p.z / leaf_size_);
  }
  int getNeighborhoodAtPoint1(const PointT& p, std::vector<LeafConstPtr>& out) const {
    out.clear();
    if (leaves_.empty()) return 0;
    int ijk[3]; cell_of(p, ijk);
    probe(ijk, 0, 0, 0, out);
    return (int)out.size();
  }
  int getNeighborhoodAtPoint7(const PointT& p, std::vector<LeafConstPtr>& out) const {
    out.clear();
    if (leaves_.empty()) return 0;
    int ijk[3]; cell_of(p, ijk);
    static const int o[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int k = 0; k < 7; k++) probe(ijk, o[k][0], o[k][1], o[k][2], out);
    return (int)out.size();
  }
  int getNeighborhoodAtPoint(const PointT& p, std::vector<LeafConstPtr>& out) const {    // pcl::getAllNeighborCellIndices: 13 half offsets, then their negatives
    out.clear();
    if (leaves_.empty()) return 0;
    int ijk[3]; cell_of(p, ijk);
    int o[26][3], k = 0;
    for (int i = -1; i < 2; i++) for (int j = -1; j < 2; j++) { o[k][0] = i; o[k][1] = j; o[k][2] = -1; k++; }
    for (int i = -1; i < 2; i++) { o[k][0] = i; o[k][1] = -1; o[k][2] = 0; k++; }
    o[k][0] = -1; o[k][1] = 0; o[k][2] = 0; k++;
    for (int h = 0; h < 13; h++) for (int a = 0; a < 3; a++) o[13 + h][a] = -o[h][a];
    for (k = 0; k < 26; k++) probe(ijk, o[k][0], o[k][1], o[k][2], out);
    return (int)out.size();
  }
  int radiusSearch(const PointT& p, double radius, std::vector<LeafConstPtr>& out, std::vector<float>& dist, unsigned int = 0) const {
    out.clear(); dist.clear();
    if (leaves_.empty()) return 0;
    int ijk[3]; cell_of(p, ijk);
    const int reach = (int)std::ceil(radius / leaf_size_);
    const float r2 = (float)(radius * radius);
    std::vector<std::pair<float, LeafConstPtr> > hits;
    for (int dz = -reach; dz <= reach; dz++)
      for (int dy = -reach; dy <= reach; dy++)
        for (int dx = -reach; dx <= reach; dx++) {
          const int c[3] = {ijk[0] + dx, ijk[1] + dy, ijk[2] + dz};
          bool in = true;
          for (int a = 0; a < 3; a++) if (c[a] < min_b_[a] || c[a] > max_b_[a]) in = false;
          if (!in) continue;
          const int key = (c[0] - min_b_[0]) * divb_mul_[0] + (c[1] - min_b_[1]) * divb_mul_[1] + (c[2] - min_b_[2]) * divb_mul_[2];
          auto it = leaves_.find((size_t)key);
          if (it == leaves_.end() || !it->second.in_cloud) continue;
          const RefLeaf& l = it->second;
          const float ex = p.x - l.centroid[0], ey = p.y - l.centroid[1], ez = p.z - l.centroid[2];
          const float d2 = (ex * ex + ey * ey) + ez * ez;
          if (d2 < r2) hits.push_back(std::make_pair(d2, &l));
        }
    std::stable_sort(hits.begin(), hits.end(), [](const std::pair<float, LeafConstPtr>& a, const std::pair<float, LeafConstPtr>& b) { return a.first < b.first; });
    for (size_t i = 0; i < hits.size(); i++) { out.push_back(hits[i].second); dist.push_back(hits[i].first); }
    return (int)out.size();
  }
};

// The declarations of include/ndt_omp/ndt_omp.h (class pclomp::NormalDistributionsTransform, :69-551) and of pcl::Registration that the
// extracted definitions refer to - same names, types, default arguments and constness.
template <typename PointSource, typename PointTarget>
class REF_CLASS {
 public:
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef VoxelGridAdapter<PointTarget> TargetGrid;
  typedef const RefLeaf* TargetGridLeafConstPtr;

  // pcl::Registration members
  int nr_iterations_ = 0, max_iterations_ = 35;
  bool converged_ = false;
  double transformation_epsilon_ = 0.1;
  Eigen::Matrix4f final_transformation_, transformation_, previous_transformation_;
  const PointCloudSource* input_ = nullptr;
  const PointCloudTarget* target_ = nullptr;
  std::function<void(const PointCloudSource&, const std::vector<int>&, const PointCloudTarget&, const std::vector<int>&)> update_visualizer_;
  // ndt_omp.h members, constructor defaults of ndt_omp_impl2.hpp:54-83
  TargetGrid target_cells_;
  float resolution_ = 1.0f;
  double step_size_ = 0.1, outlier_ratio_ = 0.55, gauss_d1_ = 0, gauss_d2_ = 0, gauss_d3_ = 0, trans_probability_ = 0;
  NeighborSearchMethod search_method = DIRECT7;
  int num_threads_ = 1;

  REF_CLASS() { final_transformation_.setIdentity(); transformation_.setIdentity(); previous_transformation_.setIdentity(); }

  double calculateScore(const PointCloudSource& cloud) const;
  void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess);
  double computeDerivatives(Eigen::Matrix<double, 6, 1>& score_gradient, Eigen::Matrix<double, 6, 6>& hessian, PointCloudSource& trans_cloud,
                            Eigen::Matrix<double, 6, 1>& p, bool compute_hessian = true);
  double updateDerivatives(Eigen::Matrix<double, 6, 1>& score_gradient, Eigen::Matrix<double, 6, 6>& hessian, const Eigen::Matrix<float, 4, 6>& point_gradient_,
                           const Eigen::Matrix<float, 24, 6>& point_hessian_, const Eigen::Vector3d& x_trans, const Eigen::Matrix3d& c_inv,
                           bool compute_hessian = true) const;
  void computeAngleDerivatives(Eigen::Matrix<double, 6, 1>&, bool = true) {}      // see the header of this file
  void computePointDerivatives_AngleAxisd(Eigen::Vector3d& x, Eigen::Matrix<double, 6, 1>& p, Eigen::Matrix<double, 3, 6>& point_gradient_,
                                          Eigen::Matrix<double, 18, 6>& point_hessian_, bool compute_hessian = true) const;
  void computePointDerivatives_AngleAxisd(Eigen::Vector3d& x, Eigen::Matrix<double, 6, 1>& p, Eigen::Matrix<float, 4, 6>& point_gradient_,
                                          Eigen::Matrix<float, 24, 6>& point_hessian_, bool compute_hessian = true) const;
  void computeHessian(Eigen::Matrix<double, 6, 6>& hessian, PointCloudSource& trans_cloud, Eigen::Matrix<double, 6, 1>& p);
  void updateHessian(Eigen::Matrix<double, 6, 6>& hessian, const Eigen::Matrix<double, 3, 6>& point_gradient_, const Eigen::Matrix<double, 18, 6>& point_hessian_,
                     const Eigen::Vector3d& x_trans, const Eigen::Matrix3d& c_inv) const;
#ifdef REF_GROUND      // ndt_ground.h:303-309, :419-428
  double computeDerivatives_seg(Eigen::Matrix<double, 6, 1>& score_gradient, Eigen::Matrix<double, 6, 6>& hessian, PointCloudSource& trans_cloud,
                                Eigen::Matrix<double, 6, 1>& p, int flag_class, bool compute_hessian = true);
  double computeStepLengthMT(const Eigen::Matrix<double, 6, 1>& x, Eigen::Matrix<double, 6, 1>& step_dir, double step_init, double step_max, double step_min,
                             double& score, Eigen::Matrix<double, 6, 1>& score_gradient, Eigen::Matrix<double, 6, 6>& hessian, PointCloudSource& trans_cloud,
                             int flag_class);
#else
  double computeStepLengthMT(const Eigen::Matrix<double, 6, 1>& x, Eigen::Matrix<double, 6, 1>& step_dir, double step_init, double step_max, double step_min,
                             double& score, Eigen::Matrix<double, 6, 1>& score_gradient, Eigen::Matrix<double, 6, 6>& hessian, PointCloudSource& trans_cloud);
#endif
  bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t);
  double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t);
  inline double auxilaryFunction_PsiMT(double a, double f_a, double f_0, double g_0, double mu = 1.e-4) { return (f_a - f_0 - mu * g_0 * a); }    // ndt_omp.h:479-483
  inline double auxilaryFunction_dPsiMT(double g_a, double g_0, double mu = 1.e-4) { return (g_a - mu * g_0); }                                  // ndt_omp.h:492-496
};

}  // namespace REF_NS

#include REF_NDT_BODIES      // the reference's own definitions of the member functions declared above

// ---------------------------------------------------------------------------------------------------------------------
typedef REF_NS::REF_CLASS<pcl::PointXYZ, pcl::PointXYZ> RefNDT;
using REF_NS::RefLeaf;
struct RefHandle {
  RefNDT ndt;
  pcl::PointCloud<pcl::PointXYZ> source, target;
  long n_eval = 0;
};

static void load(pcl::PointCloud<pcl::PointXYZ>& c, const float* xyz, size_t n, size_t stride) {
  c.points.resize(n);
  for (size_t i = 0; i < n; i++) { c.points[i].x = xyz[i * stride]; c.points[i].y = xyz[i * stride + 1]; c.points[i].z = xyz[i * stride + 2]; }
}
static Eigen::Matrix4f mat4_colmajor(const float* M16) {
  Eigen::Matrix4f T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = M16[c * 4 + r];
  return T;
}
static void gauss(RefNDT& n) {      // computeTransformation recomputes these itself (:96-104); the taps need them beforehand
  const double c1 = 10 * (1 - n.outlier_ratio_), c2 = n.outlier_ratio_ / pow(n.resolution_, 3);
  n.gauss_d3_ = -log(c2);
  n.gauss_d1_ = -log(c1 + c2) - n.gauss_d3_;
  n.gauss_d2_ = -2 * log((-log(c1 * exp(-0.5) + c2) - n.gauss_d3_) / n.gauss_d1_);
}

extern "C" {

void* nref_create(void) { return new RefHandle(); }
void nref_destroy(void* h) { delete (RefHandle*)h; }

void nref_set_params(void* h, float resolution, double step_size, double outlier_ratio, double trans_eps, int max_iter, int search) {
  RefNDT& n = ((RefHandle*)h)->ndt;
  n.resolution_ = resolution; n.step_size_ = step_size; n.outlier_ratio_ = outlier_ratio; n.transformation_epsilon_ = trans_eps;
  n.max_iterations_ = max_iter; n.search_method = (REF_NS::NeighborSearchMethod)search;
  gauss(n);
}

// The target's cells, as the caller's voxel build produced them (ascending key): raw point count (-1 = invalidated), mean, inverse covariance,
// float centroid, membership of the centroid cloud; plus the grid geometry.
void nref_set_target_cells(void* h, int n_cells, const int32_t* keys, const int32_t* nr_points, const double* mean3, const double* icov9, const float* centroid3,
                           const int32_t* in_cloud, const int32_t* min_b, const int32_t* max_b, const int32_t* div_b, float leaf_size, int min_points,
                           const int32_t* weight, const double* evecs9, const double* evals3) {
  RefHandle& H = *(RefHandle*)h;
  auto& G = H.ndt.target_cells_;
  G.leaves_.clear();
  for (int a = 0; a < 3; a++) { G.min_b_[a] = min_b[a]; G.max_b_[a] = max_b[a]; }
  G.divb_mul_[0] = 1; G.divb_mul_[1] = div_b[0]; G.divb_mul_[2] = div_b[0] * div_b[1];
  G.leaf_size_ = leaf_size; G.min_points_per_voxel_ = min_points;
  for (int k = 0; k < n_cells; k++) {
    RefLeaf l;
    l.nr_points = nr_points[k]; l.in_cloud = in_cloud[k];
    l.weight_ = weight ? weight[k] : 1;
    for (int i = 0; i < 3; i++) { l.evals_[i] = evals3 ? evals3[k * 3 + i] : 0.0; for (int j = 0; j < 3; j++) l.evecs_(i, j) = evecs9 ? evecs9[k * 9 + i * 3 + j] : (i == j ? 1.0 : 0.0); }
    for (int i = 0; i < 3; i++) { l.mean_[i] = mean3[k * 3 + i]; l.centroid[i] = centroid3[k * 3 + i]; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) l.icov_(i, j) = icov9[k * 9 + i * 3 + j];
    G.leaves_[(size_t)keys[k]] = l;
  }
  H.ndt.target_ = &H.target;
}

void nref_set_source(void* h, const float* xyz, size_t n, size_t stride_floats) {
  RefHandle& H = *(RefHandle*)h;
  load(H.source, xyz, n, stride_floats);
  H.ndt.input_ = &H.source;
}

// one computeDerivatives call: trans cloud = T16 * source (T16 column-major; NULL = float(SE3::exp(p)))
double nref_eval_derivatives(void* h, const double* p6, const float* T16, int compute_hessian, double* g6, double* H36) {
  RefHandle& H = *(RefHandle*)h;
  gauss(H.ndt);
  Eigen::Matrix<double, 6, 1> p, g;
  for (int i = 0; i < 6; i++) p[i] = p6[i];
  const Eigen::Matrix4f T = T16 ? mat4_colmajor(T16) : Sophus::SE3::exp(p).matrix().cast<float>();
  pcl::PointCloud<pcl::PointXYZ> trans;
  pcl::transformPointCloud(H.source, trans, T);
  Eigen::Matrix<double, 6, 6> Hm;
#ifdef REF_GROUND
  const double s = H.ndt.computeDerivatives_seg(g, Hm, trans, p, 1, compute_hessian != 0);      // flag_class = 1, as computeTransformation calls it (:131-133)
#else
  const double s = H.ndt.computeDerivatives(g, Hm, trans, p, compute_hessian != 0);
#endif
  for (int i = 0; i < 6; i++) { g6[i] = g[i]; for (int j = 0; j < 6; j++) H36[i * 6 + j] = Hm(i, j); }
  return s;
}

void nref_eval_hessian(void* h, const double* p6, const float* T16, double* H36) {
  RefHandle& H = *(RefHandle*)h;
  gauss(H.ndt);
  Eigen::Matrix<double, 6, 1> p;
  for (int i = 0; i < 6; i++) p[i] = p6[i];
  const Eigen::Matrix4f T = T16 ? mat4_colmajor(T16) : Sophus::SE3::exp(p).matrix().cast<float>();
  pcl::PointCloud<pcl::PointXYZ> trans;
  pcl::transformPointCloud(H.source, trans, T);
  Eigen::Matrix<double, 6, 6> Hm;
  H.ndt.computeHessian(Hm, trans, p);
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H36[i * 6 + j] = Hm(i, j);
}

double nref_calculate_score(void* h, const float* T16) {
  RefHandle& H = *(RefHandle*)h;
  gauss(H.ndt);
  pcl::PointCloud<pcl::PointXYZ> trans;
  pcl::transformPointCloud(H.source, trans, mat4_colmajor(T16));
  return H.ndt.calculateScore(trans);
}

// pcl::Registration::align (PCL 1.8 registration/impl/registration.hpp): output = input, state reset, computeTransformation(output, guess)
int nref_align(void* h, const float* guess16, float* final16, int* iterations, double* trans_probability, float* aligned_xyz) {
  RefHandle& H = *(RefHandle*)h;
  RefNDT& n = H.ndt;
  pcl::PointCloud<pcl::PointXYZ> output = H.source;
  n.final_transformation_.setIdentity(); n.transformation_.setIdentity(); n.previous_transformation_.setIdentity();
  n.converged_ = false;
  n.computeTransformation(output, mat4_colmajor(guess16));
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) final16[c * 4 + r] = n.final_transformation_(r, c);
  *iterations = n.nr_iterations_;
  *trans_probability = n.trans_probability_;
  if (aligned_xyz) for (size_t i = 0; i < output.points.size(); i++) { aligned_xyz[3 * i] = output.points[i].x; aligned_xyz[3 * i + 1] = output.points[i].y; aligned_xyz[3 * i + 2] = output.points[i].z; }
  return n.converged_ ? 1 : 0;
}

}  // extern "C"
