// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything under oracle/.
//
// PARITY: the reference (BurryChen/lv_slam) ships no test, fixture or golden vector for the NDT path, and PCL / Eigen are not in this
// image.  Its OWN code is nevertheless the checker of this restatement: oracle/build_ref.sh takes the member functions of
// the three registration classes and of the voxel grid out of the reference's *_impl2.hpp / ndt_ground_impl.hpp /
// voxel_grid_covariance_omp_impl.hpp at build time and compiles them, with the reference's vendored Sophus, against interface stand-ins
// (oracle/ndt_ref_harness.cpp, oracle/voxel_ref_harness.cpp, oracle/ref_stubs/eigen_min.h) into oracle/_ref/*.so; tests/test_oracle_ndt.py
// holds every function below to them BIT FOR BIT (voxel cells, searches, derivative passes, Hessian, score, whole aligns, line search).
// Not pinned by that: Eigen's rounding and its two iterative solvers (the stand-in uses olin.h's), PCL's own getMinMax3D /
// transformPointCloud / getAllNeighborCellIndices / kd-tree (restated on both sides).  The closed-form / finite-difference checks of
// tests/ remain.
//
// CPU restatement of lv_slam's NDT scan matching, the variant that is actually compiled
// (src/ndt_omp/ndt_omp.cpp:2 and src/ndt_pca/ndt_pca.cpp:2 include the *_impl2.hpp Lie-algebra files):
//   VoxelGridCovariance::applyFilter      include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:49-370
//     (pca label + weight)                include/ndt_pca/voxel_grid_covariance_pca_impl.hpp:364-397,
//                                         include/ndt_pca/voxel_grid_covariance_pca.h:222-226
//   getNeighborhoodAtPoint{,7,1}          include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:373-442
//   radiusSearch                          include/ndt_omp/voxel_grid_covariance_omp.h:506-534
//   computeTransformation                 include/ndt_omp/ndt_omp_impl2.hpp:88-188
//   computeDerivatives                    include/ndt_omp/ndt_omp_impl2.hpp:197-305  (pca: ndt_pca_impl2.hpp:201-311)
//   computePointDerivatives_AngleAxisd    include/ndt_omp/ndt_omp_impl2.hpp:504-532 (float), :535-563 (double)
//   updateDerivatives                     include/ndt_omp/ndt_omp_impl2.hpp:567-619
//   computeHessian / updateHessian        include/ndt_omp/ndt_omp_impl2.hpp:623-714
//   updateIntervalMT/trialValueSelectionMT/computeStepLengthMT   :718-755, :759-838, :842-1003
//   calculateScore                        include/ndt_omp/ndt_omp_impl2.hpp:1007-1040
//   pclomp_ground (VAR_GROUND)            include/ndt_omp/ndt_ground_impl.hpp: computeTransformation :88-179 (flag_class = 1),
//                                         computeDerivatives_seg :363-572; everything else is textually ndt_omp_impl2.hpp.
//                                         It is the one consumer of the eigenVECTORS (leaf normal): the cyclic Jacobi of olin.h
//                                         converges to the same eigenvector as Eigen's QL up to sign and rounding, and only |n_z| is read.
// PCL pieces the reference calls but does not vendor (PCL 1.8.1, ros:melodic):
//   pcl::Registration::align              (copies input to output, resets transforms, calls computeTransformation)
//   pcl::transformPointCloud (dense)      x' = ((m00*x + m01*y) + m02*z) + m03 in float, no FMA
//   pcl::getMinMax3D, pcl::VoxelGrid::setLeafSize (inverse_leaf_size = 1/leaf in float)
//   pcl::KdTreeFLANN::radiusSearch        over voxel centroids; restated as a scan of the 27-cell block
//                                         keeping centroids with squared float distance < r*r (FLANN RadiusResultSet), nearest first
//   pcl::getAllNeighborCellIndices        26 offsets, centre EXCLUDED: 13 "half" offsets then their negatives
//   pcl::Registration::getFitnessScore    mean squared NN distance (brute-force grid NN here)
//
// The float dot products follow Eigen 3.3 + SSE evaluation order where it can be inferred:
// a 4-wide row.col dot reduces as (t0+t2)+(t1+t3); column-packet matrix products sum k sequentially.
// That inference is best-effort (Eigen is absent); it only affects the last float ulp per term.
//
// exp(float) at ndt_omp_impl2.hpp:581 is UNQUALIFIED.  Every PCL translation unit includes <math.h> through <pcl/pcl_macros.h>,
// and libstdc++'s <math.h> wrapper (GCC >= 6; the reference's image has GCC 7) does `using std::exp`, so the float overload
// std::exp(float) = __builtin_expf = glibc expf is the best match - not (float)exp((double)x).  glibc >= 2.27 (Ubuntu 18.04) carries
// the table-driven expf of sysdeps/ieee754/flt-32/e_expf.c, unchanged through 2.39 (this image): the oracle simply calls the
// host's expf.  The same routine is restated in oracle_expf_restated() below (what the device path computes) and the CPU tests
// hold the two together; tools/expf_sweep.c proved them bit-identical over every finite float in [-104, 88.8].
//
// Build: g++ -O3 -fopenmp -msse4.2 -ffp-contract=off (mirrors /root/reference/CMakeLists.txt:6,11,41-45).
#include <math.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>
#include <array>
#include <algorithm>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "olin.h"
#include "ose3.h"

using olin::M3;
using olin::V3;

namespace {

enum SearchMethod { KDTREE = 0, DIRECT26 = 1, DIRECT7 = 2, DIRECT1 = 3 };  // ndt_omp.h:61
enum Variant { VAR_OMP = 0, VAR_PCA = 1, VAR_GROUND = 2 };

struct Leaf {                       // voxel_grid_covariance_omp.h:92-195
  int nr_points = 0;
  double mean[3] = {0, 0, 0};
  double cov[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};   // the Leaf constructor starts cov_ at the IDENTITY (voxel_grid_covariance_omp.h:98-106) and applyFilter
                                                           // adds the point products on top of it (:240,:285): every covariance carries + I (n - 1) / n^2
  double icov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double evecs[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double evals[3] = {0, 0, 0};
  float centroid[4] = {0, 0, 0, 0};
  int raw_points = 0;               // count before the nr_points=-1 invalidation (for inspection only)
  int in_centroid_cloud = 0;        // pushed to voxel_centroids_ (kd-tree input)
  int dimension_label = 0;          // pca only
  double dimension_2d = 1.0;        // pca only (double member, read through an int getter)
};

struct Pt { float x, y, z; };

struct Trace {                      // one record per Newton iteration, for per-iteration parity
  double p_before[6], delta_dir[6], step, score, p_after[6];
  int trials, hessian_recomputed;
};

struct NDT {
  int variant = VAR_OMP;
  float resolution = 1.0f;          // ndt_omp_impl2.hpp:56-83 defaults
  double step_size = 0.1, outlier_ratio = 0.55, trans_eps = 0.1;
  int max_iter = 35, search = DIRECT7, num_threads = 1;
  int min_points_per_voxel = 6;     // voxel_grid_covariance_omp.h:204
  double min_covar_eigvalue_mult = 0.01;
  double gauss_d1 = 0, gauss_d2 = 0, gauss_d3 = 0;

  // voxel grid
  float leaf_size = 1.0f, inv_leaf = 1.0f;
  int min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0}, divb_mul[3] = {0, 0, 0};
  std::map<size_t, Leaf> leaves;
  std::vector<Pt> target;

  // registration state
  std::vector<Pt> input;            // source
  float final_T[16];
  int nr_iterations = 0;
  bool converged = false;
  double trans_probability = 0;
  long n_eval = 0, n_hess = 0;      // computeDerivatives / computeHessian call counters
  std::vector<Trace> trace;
};

void compute_gauss(NDT& n) {        // ndt_omp_impl2.hpp:93-100
  double c1 = 10 * (1 - n.outlier_ratio);
  double c2 = n.outlier_ratio / std::pow((double)n.resolution, 3);
  n.gauss_d3 = -std::log(c2);
  n.gauss_d1 = -std::log(c1 + c2) - n.gauss_d3;
  n.gauss_d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - n.gauss_d3) / n.gauss_d1);
}

// ---------------------------------------------------------------- voxel grid
void apply_filter(NDT& n) {         // voxel_grid_covariance_omp_impl.hpp:49-370
  n.leaves.clear();
  n.leaf_size = n.resolution;
  n.inv_leaf = 1.0f / n.leaf_size;  // pcl::VoxelGrid::setLeafSize
  const std::vector<Pt>& P = n.target;
  if (P.empty()) { for (int a = 0; a < 3; a++) n.min_b[a] = n.max_b[a] = n.div_b[a] = n.divb_mul[a] = 0; return; }
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  size_t n_finite = 0;
  for (const Pt& p : P) {           // pcl::getMinMax3D: the is_dense=false branch skips non-finite points; on a dense cloud both branches agree
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    n_finite++;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  if (n_finite == 0) { for (int a = 0; a < 3; a++) n.min_b[a] = n.max_b[a] = n.div_b[a] = n.divb_mul[a] = 0; return; }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * n.inv_leaf) + 1;   // :76-85 overflow guard
  int64_t dy = (int64_t)((mx[1] - mn[1]) * n.inv_leaf) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * n.inv_leaf) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) return;
  for (int a = 0; a < 3; a++) {     // :87-103
    n.min_b[a] = (int)std::floor(mn[a] * n.inv_leaf);
    n.max_b[a] = (int)std::floor(mx[a] * n.inv_leaf);
    n.div_b[a] = n.max_b[a] - n.min_b[a] + 1;
  }
  n.divb_mul[0] = 1; n.divb_mul[1] = n.div_b[0]; n.divb_mul[2] = n.div_b[0] * n.div_b[1];

  for (const Pt& p : P) {           // first pass :213-262
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int ijk0 = (int)(std::floor(p.x * n.inv_leaf) - (float)n.min_b[0]);
    int ijk1 = (int)(std::floor(p.y * n.inv_leaf) - (float)n.min_b[1]);
    int ijk2 = (int)(std::floor(p.z * n.inv_leaf) - (float)n.min_b[2]);
    int idx = ijk0 * n.divb_mul[0] + ijk1 * n.divb_mul[1] + ijk2 * n.divb_mul[2];
    Leaf& leaf = n.leaves[(size_t)idx];
    double pt[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; i++) leaf.mean[i] += pt[i];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) leaf.cov[i][j] += pt[i] * pt[j];
    leaf.centroid[0] += p.x; leaf.centroid[1] += p.y; leaf.centroid[2] += p.z;   // float accumulation
    ++leaf.nr_points;
  }

  for (auto& kv : n.leaves) {       // second pass :281-367
    Leaf& leaf = kv.second;
    leaf.raw_points = leaf.nr_points;
    for (int i = 0; i < 4; i++) leaf.centroid[i] /= (float)leaf.nr_points;
    double pt_sum[3] = {leaf.mean[0], leaf.mean[1], leaf.mean[2]};
    for (int i = 0; i < 3; i++) leaf.mean[i] /= leaf.nr_points;
    if (leaf.nr_points < n.min_points_per_voxel) continue;
    leaf.in_centroid_cloud = 1;
    const double np = leaf.nr_points;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        leaf.cov[i][j] = (leaf.cov[i][j] - 2 * (pt_sum[i] * leaf.mean[j])) / np + leaf.mean[i] * leaf.mean[j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) leaf.cov[i][j] *= (np - 1.0) / np;

    M3 C, V; double ev[3];
    std::memcpy(C.a, leaf.cov, sizeof C.a);
    olin::sym3_eig(C, ev, V);
    std::memcpy(leaf.evecs, V.a, sizeof V.a);
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { leaf.nr_points = -1; continue; }
    double min_ev = n.min_covar_eigvalue_mult * ev[2];
    if (ev[0] < min_ev) {
      ev[0] = min_ev;
      if (ev[1] < min_ev) ev[1] = min_ev;
      M3 L = olin::m3_zero(); L.a[0][0] = ev[0]; L.a[1][1] = ev[1]; L.a[2][2] = ev[2];
      M3 Cn = olin::m3_mul(olin::m3_mul(V, L), olin::m3_inverse(V));
      std::memcpy(leaf.cov, Cn.a, sizeof Cn.a);
    }
    for (int i = 0; i < 3; i++) leaf.evals[i] = ev[i];

    if (n.variant == VAR_PCA) {     // voxel_grid_covariance_pca_impl.hpp:364-397
      double s0 = std::sqrt(ev[0]), s1 = std::sqrt(ev[1]), s2 = std::sqrt(ev[2]);
      double f[3] = {(s2 - s1) / s2, (s1 - s0) / s2, s0 / s2};
      int d = 0;
      if (f[1] > f[d]) d = 1;
      if (f[2] > f[d]) d = 2;     // Eigen maxCoeff(&d): first maximum wins
      leaf.dimension_label = d + 1;
      double scale = 1;
      if (leaf.dimension_label == 2) scale = 1.25;
      else if (leaf.dimension_label == 3) scale = 1;
      else if (leaf.dimension_label == 1) scale = 0.75;
      double nm = std::sqrt(leaf.mean[0] * leaf.mean[0] + leaf.mean[1] * leaf.mean[1] + leaf.mean[2] * leaf.mean[2]);
      leaf.dimension_2d = scale * nm;
    }

    M3 Cc; std::memcpy(Cc.a, leaf.cov, sizeof Cc.a);
    M3 Ic = olin::m3_inverse(Cc);
    std::memcpy(leaf.icov, Ic.a, sizeof Ic.a);
    double mxc = -std::numeric_limits<double>::infinity(), mnc = std::numeric_limits<double>::infinity();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { mxc = std::max(mxc, leaf.icov[i][j]); mnc = std::min(mnc, leaf.icov[i][j]); }
    if (mxc == std::numeric_limits<float>::infinity() || mnc == -std::numeric_limits<float>::infinity()) leaf.nr_points = -1;
  }
}

inline int leaf_weight(const NDT& n, const Leaf& l) {   // int getDimension2d() truncates (voxel_grid_covariance_pca.h:222-226)
  return n.variant == VAR_PCA ? (int)l.dimension_2d : 1;
}

// pclomp_ground: angle between the leaf's normal (eigenvector of the smallest eigenvalue, first column of evecs_) and the z axis, in
// degrees with the reference's constant (include/ndt_omp/ndt_ground_impl.hpp:507-511).  evecs_ is assigned before the eigenvalue
// check (voxel_grid_covariance_omp_impl.hpp:335), so invalidated leaves have one too.
inline double leaf_angle2xy(const Leaf& l) {
  double nx = l.evecs[0][0], ny = l.evecs[1][0], nz = l.evecs[2][0];
  double nrm = std::sqrt((nx * nx + ny * ny) + nz * nz);      // Vector3d::normalize(): divides by norm()
  nz = nz / nrm;
  return std::acos(std::fabs(nz)) * 180 / 3.1415926;
}

// Offsets for the direct searches.  DIRECT7: voxel_grid_covariance_omp_impl.hpp:423-430.
static const int kOff7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};

// pcl::getAllNeighborCellIndices (pcl/filters/voxel_grid.h, PCL 1.8): the 13 "half" neighbours
// {(i,j,-1) i,j in -1..1 (i outer)}, {(i,-1,0) i in -1..1}, (-1,0,0), followed by their negatives.
// The centre cell is NOT part of DIRECT26.
struct Off26 { int o[26][3]; Off26() { int k = 0;
  for (int i = -1; i < 2; i++) for (int j = -1; j < 2; j++) { o[k][0] = i; o[k][1] = j; o[k][2] = -1; k++; }
  for (int i = -1; i < 2; i++) { o[k][0] = i; o[k][1] = -1; o[k][2] = 0; k++; }
  o[k][0] = -1; o[k][1] = 0; o[k][2] = 0; k++;
  for (int h = 0; h < 13; h++) for (int a = 0; a < 3; a++) o[13 + h][a] = -o[h][a]; } };
static const Off26 kOff26;
void neighbours_direct(const NDT& n, const Pt& p, int mode, std::vector<const Leaf*>& out) {
  out.clear();
  if (n.leaves.empty()) return;
  int ijk[3] = {(int)std::floor(p.x / n.leaf_size), (int)std::floor(p.y / n.leaf_size), (int)std::floor(p.z / n.leaf_size)};
  auto probe = [&](int dx, int dy, int dz) {
    int d[3] = {dx, dy, dz};
    for (int a = 0; a < 3; a++) if (!(n.min_b[a] - ijk[a] <= d[a] && n.max_b[a] - ijk[a] >= d[a])) return;
    int key = (ijk[0] + dx - n.min_b[0]) * n.divb_mul[0] + (ijk[1] + dy - n.min_b[1]) * n.divb_mul[1] + (ijk[2] + dz - n.min_b[2]) * n.divb_mul[2];
    auto it = n.leaves.find((size_t)key);
    if (it != n.leaves.end() && it->second.nr_points >= n.min_points_per_voxel) out.push_back(&it->second);
  };
  if (mode == DIRECT1) probe(0, 0, 0);
  else if (mode == DIRECT7) for (int k = 0; k < 7; k++) probe(kOff7[k][0], kOff7[k][1], kOff7[k][2]);
  else for (int k = 0; k < 26; k++) probe(kOff26.o[k][0], kOff26.o[k][1], kOff26.o[k][2]);
}

// Radius search over the centroid cloud (kd-tree replaced by the equivalent 27-cell scan — a centroid
// within `radius == leaf` of p lies in a cell at most one step away along every axis).  Leaves later
// invalidated with nr_points=-1 stay in the cloud (pushed at :302,326 before the `continue` at :337-341)
// and radiusSearch does not re-check nr_points.  Sorted by squared distance like FLANN (sorted_=true).
void neighbours_radius(const NDT& n, const Pt& p, double radius, std::vector<const Leaf*>& out) {
  out.clear();
  if (n.leaves.empty()) return;
  int ijk[3] = {(int)std::floor(p.x / n.leaf_size), (int)std::floor(p.y / n.leaf_size), (int)std::floor(p.z / n.leaf_size)};
  int reach = (int)std::ceil(radius / n.leaf_size);
  float r2 = (float)(radius * radius);
  std::vector<std::pair<float, const Leaf*>> hits;
  for (int dz = -reach; dz <= reach; dz++)
    for (int dy = -reach; dy <= reach; dy++)
      for (int dx = -reach; dx <= reach; dx++) {
        int c[3] = {ijk[0] + dx, ijk[1] + dy, ijk[2] + dz};
        bool in = true;
        for (int a = 0; a < 3; a++) if (c[a] < n.min_b[a] || c[a] > n.max_b[a]) in = false;
        if (!in) continue;
        int key = (c[0] - n.min_b[0]) * n.divb_mul[0] + (c[1] - n.min_b[1]) * n.divb_mul[1] + (c[2] - n.min_b[2]) * n.divb_mul[2];
        auto it = n.leaves.find((size_t)key);
        if (it == n.leaves.end() || !it->second.in_centroid_cloud) continue;
        const Leaf& l = it->second;
        float ex = p.x - l.centroid[0], ey = p.y - l.centroid[1], ez = p.z - l.centroid[2];
        float d2 = (ex * ex + ey * ey) + ez * ez;   // FLANN L2_Simple accumulation order
        if (d2 < r2) hits.push_back({d2, &l});
      }
  std::stable_sort(hits.begin(), hits.end(), [](const std::pair<float, const Leaf*>& a, const std::pair<float, const Leaf*>& b) { return a.first < b.first; });
  for (auto& h : hits) out.push_back(h.second);
}

// ---------------------------------------------------------------- per-point math
inline float dot4_sse(const float a[4], const float b[4]) {   // Eigen 3.3 SSE predux: (t0+t2)+(t1+t3)
  float t0 = a[0] * b[0], t1 = a[1] * b[1], t2 = a[2] * b[2], t3 = a[3] * b[3];
  return (t0 + t2) + (t1 + t3);
}

// computePointDerivatives_AngleAxisd, float overload (ndt_omp_impl2.hpp:504-532).
// J is 4x6 (J[r][c]); Hp is 24x6 (Hp[r][c]).  Both start as "zero + identity block" and only the
// listed entries are ever overwritten.
void point_derivs_f(const double x[3], const double p[6], float J[4][6], float Hp[24][6], bool compute_hessian) {
  ose3::SE3 T = ose3::se3_exp(p);                  // recomputed per point per neighbour in the reference
  float M[16]; ose3::se3_to_matrix4f(T, M);
  float x4[4] = {(float)x[0], (float)x[1], (float)x[2], 0.0f};
  float xt[4];
  for (int r = 0; r < 4; r++)                       // Matrix4f * Vector4f: column-packet, k sequential
    xt[r] = ((M[0 * 4 + r] * x4[0] + M[1 * 4 + r] * x4[1]) + M[2 * 4 + r] * x4[2]) + M[3 * 4 + r] * x4[3];
  J[1][3] = -xt[2]; J[2][3] = xt[1];
  J[0][4] = xt[2];  J[2][4] = -xt[0];
  J[0][5] = -xt[1]; J[1][5] = xt[0];
  if (compute_hessian) {
    auto set4 = [&](int row, int col, float a, float b, float c) { Hp[row][col] = a; Hp[row + 1][col] = b; Hp[row + 2][col] = c; Hp[row + 3][col] = 0.0f; };
    set4(12, 3, 0, -xt[1], -xt[2]); set4(16, 3, xt[1], 0, 0);        set4(20, 3, xt[2], 0, 0);
    set4(12, 4, 0, xt[0], 0);       set4(16, 4, -xt[0], 0, -xt[2]);  set4(20, 4, 0, xt[2], 0);
    set4(12, 5, 0, 0, xt[0]);       set4(16, 5, 0, 0, xt[1]);        set4(20, 5, -xt[0], -xt[1], 0);
  }
}

// updateDerivatives (ndt_omp_impl2.hpp:567-619).  Returns score_inc (0 on the early-out).
double update_derivs(const NDT& n, double g[6], double H[6][6], const float J[4][6], const float Hp[24][6],
                     const double x_trans[3], const double c_inv[3][3], bool compute_hessian) {
  float xt4[4] = {(float)x_trans[0], (float)x_trans[1], (float)x_trans[2], 0.0f};
  float C[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) C[i][j] = (i < 3 && j < 3) ? (float)c_inv[i][j] : 0.0f;
  float gauss_d2 = (float)n.gauss_d2;

  float xC[4];                                       // x_trans4 * c_inv4  (row . column dots)
  for (int j = 0; j < 4; j++) { float col[4] = {C[0][j], C[1][j], C[2][j], C[3][j]}; xC[j] = dot4_sse(xt4, col); }
  float q = dot4_sse(xt4, xC);
  float e = ::expf(-gauss_d2 * q * 0.5f);   // unqualified exp(float) under <math.h>'s `using std::exp` = expf (see the header note)
  float score_inc = (float)(-n.gauss_d1 * (double)e);
  e = gauss_d2 * e;
  if (e > 1 || e < 0 || e != e) return 0;
  e = (float)((double)e * n.gauss_d1);

  float CJ[4][6];                                    // c_inv4 * point_gradient4, k sequential
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 6; j++)
      CJ[i][j] = ((C[i][0] * J[0][j] + C[i][1] * J[1][j]) + C[i][2] * J[2][j]) + C[i][3] * J[3][j];
  float a[6];                                        // x_trans4 * CJ
  for (int j = 0; j < 6; j++) { float col[4] = {CJ[0][j], CJ[1][j], CJ[2][j], CJ[3][j]}; a[j] = dot4_sse(xt4, col); }
  for (int j = 0; j < 6; j++) g[j] += (double)(e * a[j]);

  if (compute_hessian) {
    float Mx[6][6];                                  // point_gradient4^T * CJ : Mx[r][c] = J.col(r) . CJ.col(c)
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) {
        float u[4] = {J[0][r], J[1][r], J[2][r], J[3][r]}, v[4] = {CJ[0][c], CJ[1][c], CJ[2][c], CJ[3][c]};
        Mx[r][c] = dot4_sse(u, v);
      }
    for (int i = 0; i < 6; i++) {
      float hp[6];                                   // x_trans4_x_c_inv4 * point_hessian_.block<4,6>(4i,0)
      for (int j = 0; j < 6; j++) { float col[4] = {Hp[4 * i][j], Hp[4 * i + 1][j], Hp[4 * i + 2][j], Hp[4 * i + 3][j]}; hp[j] = dot4_sse(xC, col); }
      for (int j = 0; j < 6; j++)
        H[i][j] += (double)(e * ((-gauss_d2 * a[i] * a[j] + hp[j]) + Mx[j][i]));
    }
  }
  return score_inc;
}

inline Pt transform_pt(const float M[16], const Pt& p) {   // pcl::transformPointCloud, dense branch
  Pt o;
  o.x = ((M[0] * p.x + M[4] * p.y) + M[8] * p.z) + M[12];
  o.y = ((M[1] * p.x + M[5] * p.y) + M[9] * p.z) + M[13];
  o.z = ((M[2] * p.x + M[6] * p.y) + M[10] * p.z) + M[14];
  return o;
}

void transform_cloud(const std::vector<Pt>& in, std::vector<Pt>& out, const float M[16], int nthreads) {
  out.resize(in.size());
  (void)nthreads;
  for (size_t i = 0; i < in.size(); i++) out[i] = transform_pt(M, in[i]);
}

// computeDerivatives (ndt_omp_impl2.hpp:197-305; pca weight ndt_pca_impl2.hpp:293-296)
double compute_derivatives(NDT& n, double g_out[6], double H_out[6][6], const std::vector<Pt>& trans, const double p[6], bool compute_hessian) {
  n.n_eval++;
  const int T = std::max(1, n.num_threads);
  std::vector<double> scores(T, 0.0);
  std::vector<std::array<double, 6>> gs(T);
  std::vector<std::array<double, 36>> Hs(T);
  for (int t = 0; t < T; t++) { gs[t].fill(0.0); Hs[t].fill(0.0); }
  std::vector<std::vector<const Leaf*>> nbrs(T);
  const long N = (long)n.input.size();

#pragma omp parallel for num_threads(T) schedule(guided, 8)
  for (long idx = 0; idx < N; idx++) {
#ifdef _OPENMP
    int tn = omp_get_thread_num();
#else
    int tn = 0;
#endif
    float J[4][6]; float Hp[24][6];
    std::memset(J, 0, sizeof J); std::memset(Hp, 0, sizeof Hp);
    J[0][0] = J[1][1] = J[2][2] = 1.0f;
    const Pt xt = trans[idx];
    std::vector<const Leaf*>& nb = nbrs[tn];
    switch (n.search) {
      case KDTREE: neighbours_radius(n, xt, n.resolution, nb); break;
      case DIRECT26: neighbours_direct(n, xt, DIRECT26, nb); break;
      default:
      case DIRECT7: neighbours_direct(n, xt, DIRECT7, nb); break;
      case DIRECT1: neighbours_direct(n, xt, DIRECT1, nb); break;
    }
    double score_pt = 0, g_pt[6] = {0, 0, 0, 0, 0, 0}, H_pt[6][6];
    std::memset(H_pt, 0, sizeof H_pt);
    double angle2xy = 0;                             // pclomp_ground: overwritten per cell, so the LAST neighbour decides (ndt_ground_impl.hpp:484,511)
    for (const Leaf* cell : nb) {
      const Pt xo = n.input[idx];
      double x[3] = {xo.x, xo.y, xo.z};
      double x_trans[3] = {(double)xt.x - cell->mean[0], (double)xt.y - cell->mean[1], (double)xt.z - cell->mean[2]};
      point_derivs_f(x, p, J, Hp, compute_hessian);
      if (n.variant == VAR_GROUND) {
        // computeDerivatives_seg (ndt_ground_impl.hpp:363-572) calls updateDerivatives TWICE per cell (:519,522): the first call's
        // score is dropped, but both calls add into the point's gradient and Hessian
        angle2xy = leaf_angle2xy(*cell);
        (void)update_derivs(n, g_pt, H_pt, J, Hp, x_trans, cell->icov, compute_hessian);
      }
      score_pt += update_derivs(n, g_pt, H_pt, J, Hp, x_trans, cell->icov, compute_hessian);
      if (n.variant == VAR_PCA) {
        double w = (double)leaf_weight(n, *cell);
        score_pt *= w;
        for (int i = 0; i < 6; i++) g_pt[i] *= w;
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H_pt[i][j] *= w;
      }
    }
    // flag_class == 1 (the only class align() runs, :130-133): the point counts only when its last neighbour is near-horizontal (:533-538)
    if (n.variant == VAR_GROUND && !(angle2xy < 10)) continue;
    scores[tn] += score_pt;
    for (int i = 0; i < 6; i++) gs[tn][i] += g_pt[i];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Hs[tn][i * 6 + j] += H_pt[i][j];
  }
  double score = 0;
  for (int i = 0; i < 6; i++) { g_out[i] = 0; for (int j = 0; j < 6; j++) H_out[i][j] = 0; }
  for (int t = 0; t < T; t++) {
    score += scores[t];
    for (int i = 0; i < 6; i++) g_out[i] += gs[t][i];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H_out[i][j] += Hs[t][i * 6 + j];
  }
  if (n.variant == VAR_GROUND) {                     // :554-561 - only z, roll, pitch are solved for
    const int off[3] = {0, 1, 5};
    for (int k : off) { g_out[k] = 0; for (int j = 0; j < 6; j++) { H_out[k][j] = 0; H_out[j][k] = 0; } }
  }
  return score;
}

// computePointDerivatives_AngleAxisd double overload (:535-563) + updateHessian (:683-714) + computeHessian (:623-679)
void compute_hessian(NDT& n, double H[6][6], const std::vector<Pt>& trans, const double p[6]) {
  n.n_hess++;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[i][j] = 0;
  ose3::SE3 T = ose3::se3_exp(p);
  M3 R = ose3::quat_to_matrix(T.q);
  std::vector<const Leaf*> nb;
  for (size_t idx = 0; idx < n.input.size(); idx++) {
    const Pt xt = trans[idx];
    neighbours_radius(n, xt, n.resolution, nb);
    for (const Leaf* cell : nb) {
      const Pt xo = n.input[idx];
      double x4[3] = {xo.x, xo.y, xo.z};
      double r[3];                                   // (SE3::exp(p).matrix() * [x;0]).head<3>() = R x
      for (int i = 0; i < 3; i++) r[i] = (R.a[i][0] * x4[0] + R.a[i][1] * x4[1]) + R.a[i][2] * x4[2];
      double J[3][6]; std::memset(J, 0, sizeof J);
      J[0][0] = J[1][1] = J[2][2] = 1.0;
      J[1][3] = -r[2]; J[2][3] = r[1]; J[0][4] = r[2]; J[2][4] = -r[0]; J[0][5] = -r[1]; J[1][5] = r[0];
      double Hp[18][6]; std::memset(Hp, 0, sizeof Hp);
      auto set3 = [&](int row, int col, double a, double b, double c) { Hp[row][col] = a; Hp[row + 1][col] = b; Hp[row + 2][col] = c; };
      set3(9, 3, 0, -r[1], -r[2]); set3(12, 3, r[1], 0, 0);       set3(15, 3, r[2], 0, 0);
      set3(9, 4, 0, r[0], 0);      set3(12, 4, -r[0], 0, -r[2]);  set3(15, 4, 0, r[2], 0);
      set3(9, 5, 0, 0, r[0]);      set3(12, 5, 0, 0, r[1]);       set3(15, 5, -r[0], -r[1], 0);

      double d[3] = {(double)xt.x - cell->mean[0], (double)xt.y - cell->mean[1], (double)xt.z - cell->mean[2]};
      const double (*C)[3] = cell->icov;
      double Cd[3];
      for (int i = 0; i < 3; i++) Cd[i] = (C[i][0] * d[0] + C[i][1] * d[1]) + C[i][2] * d[2];
      double e = n.gauss_d2 * std::exp(-n.gauss_d2 * ((d[0] * Cd[0] + d[1] * Cd[1]) + d[2] * Cd[2]) / 2);
      if (e > 1 || e < 0 || e != e) continue;
      e *= n.gauss_d1;
      double CJ[3][6];
      for (int c = 0; c < 6; c++) for (int i = 0; i < 3; i++) CJ[i][c] = (C[i][0] * J[0][c] + C[i][1] * J[1][c]) + C[i][2] * J[2][c];
      double dCJ[6];
      for (int c = 0; c < 6; c++) dCJ[c] = (d[0] * CJ[0][c] + d[1] * CJ[1][c]) + d[2] * CJ[2][c];
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
          double CH[3];
          for (int k = 0; k < 3; k++) CH[k] = (C[k][0] * Hp[3 * i][j] + C[k][1] * Hp[3 * i + 1][j]) + C[k][2] * Hp[3 * i + 2][j];
          double dCH = (d[0] * CH[0] + d[1] * CH[1]) + d[2] * CH[2];
          double JCJ = (J[0][j] * CJ[0][i] + J[1][j] * CJ[1][i]) + J[2][j] * CJ[2][i];
          H[i][j] += e * ((-n.gauss_d2 * dCJ[i] * dCJ[j] + dCH) + JCJ);
        }
    }
  }
}

// ---------------------------------------------------------------- More–Thuente (ndt_omp_impl2.hpp:718-838; ndt_omp.h:479-496)
inline double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool update_interval_mt(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

double trial_value_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
    return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
  } else {
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}

inline double dot6(const double a[6], const double b[6]) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }

// computeStepLengthMT (ndt_omp_impl2.hpp:842-1003)
double step_length_mt(NDT& n, const double x[6], double step_dir[6], double step_init, double step_max, double step_min,
                      double& score, double g[6], double H[6][6], std::vector<Pt>& trans, int* trials_out, int* hess_out) {
  double phi_0 = -score;
  double d_phi_0 = -dot6(g, step_dir);
  double x_t[6];
  *trials_out = 0; *hess_out = 0;
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int i = 0; i < 6; i++) step_dir[i] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
  double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
  bool interval_converged = (step_max - step_min) > 0, open_interval = true;   // sic (:891)
  double a_t = step_init;
  a_t = std::min(a_t, step_max);
  a_t = std::max(a_t, step_min);
  for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
  ose3::se3_to_matrix4f(ose3::se3_exp(x_t), n.final_T);
  transform_cloud(n.input, trans, n.final_T, n.num_threads);
  score = compute_derivatives(n, g, H, trans, x_t, true);
  double phi_t = -score, d_phi_t = -dot6(g, step_dir);
  double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);

  while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval) a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
    ose3::se3_to_matrix4f(ose3::se3_exp(x_t), n.final_T);
    transform_cloud(n.input, trans, n.final_T, n.num_threads);
    score = compute_derivatives(n, g, H, trans, x_t, false);
    phi_t = -score; d_phi_t = -dot6(g, step_dir);
    psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu); d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
    }
    if (open_interval) interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) { compute_hessian(n, H, trans, x_t); *hess_out = 1; }
  *trials_out = step_iterations;
  return a_t;
}

// pcl::Registration::align + computeTransformation (ndt_omp_impl2.hpp:88-188)
void align(NDT& n, const float guess[16], std::vector<Pt>& output) {
  n.trace.clear();
  n.n_eval = n.n_hess = 0;
  static const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy(n.final_T, I4, sizeof I4);
  output = n.input;
  n.nr_iterations = 0;
  n.converged = false;
  compute_gauss(n);
  bool differs = false;                               // guess != Matrix4f::Identity()
  for (int i = 0; i < 16; i++) if (guess[i] != I4[i]) differs = true;
  if (differs) {
    std::memcpy(n.final_T, guess, sizeof I4);
    std::vector<Pt> tmp; transform_cloud(output, tmp, guess, n.num_threads); output.swap(tmp);
  }
  M3 R; V3 t;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R.a[r][c] = (double)guess[c * 4 + r]; t.v[r] = (double)guess[12 + r]; }
  double p[6], delta_p[6], g[6], H[6][6];
  ose3::se3_log(ose3::se3_from_Rt(R, t), p);
  double score = compute_derivatives(n, g, H, output, p, true);
  const double N = (double)n.input.size();
  while (!n.converged) {
    double Hrow[36], neg_g[6];
    for (int i = 0; i < 6; i++) { neg_g[i] = -g[i]; for (int j = 0; j < 6; j++) Hrow[i * 6 + j] = H[i][j]; }
    olin::svd6_solve(Hrow, neg_g, delta_p);
    double nrm = std::sqrt(dot6(delta_p, delta_p));
    if (nrm == 0 || nrm != nrm) {
      n.trans_probability = score / N;
      n.converged = (nrm == nrm);
      return;
    }
    for (int i = 0; i < 6; i++) delta_p[i] /= nrm;
    Trace tr; std::memset(&tr, 0, sizeof tr);
    std::memcpy(tr.p_before, p, sizeof p);
    double step = step_length_mt(n, p, delta_p, nrm, n.step_size, n.trans_eps / 2, score, g, H, output, &tr.trials, &tr.hessian_recomputed);
    std::memcpy(tr.delta_dir, delta_p, sizeof delta_p);
    for (int i = 0; i < 6; i++) delta_p[i] *= step;
    double pn[6];
    ose3::se3_log(ose3::se3_mul(ose3::se3_exp(delta_p), ose3::se3_exp(p)), pn);
    std::memcpy(p, pn, sizeof p);
    tr.step = step; tr.score = score; std::memcpy(tr.p_after, p, sizeof p);
    n.trace.push_back(tr);
    // pclomp_ground tests the step length in the first iteration as well (ndt_ground_impl.hpp:173)
    if (n.nr_iterations > n.max_iter || ((n.nr_iterations || n.variant == VAR_GROUND) && (std::fabs(step) < n.trans_eps))) n.converged = true;
    n.nr_iterations++;
  }
  n.trans_probability = score / N;
}

// calculateScore (ndt_omp_impl2.hpp:1007-1040)
double calculate_score(NDT& n, const std::vector<Pt>& trans) {
  double score = 0;
  std::vector<const Leaf*> nb;
  for (const Pt& xt : trans) {
    neighbours_radius(n, xt, n.resolution, nb);
    for (const Leaf* cell : nb) {
      double d[3] = {(double)xt.x - cell->mean[0], (double)xt.y - cell->mean[1], (double)xt.z - cell->mean[2]};
      double Cd[3];
      for (int i = 0; i < 3; i++) Cd[i] = (cell->icov[i][0] * d[0] + cell->icov[i][1] * d[1]) + cell->icov[i][2] * d[2];
      double e = std::exp(-n.gauss_d2 * ((d[0] * Cd[0] + d[1] * Cd[1]) + d[2] * Cd[2]) / 2);
      double inc = -n.gauss_d1 * e - n.gauss_d3;
      score += inc / (double)nb.size();
    }
  }
  return score / (double)trans.size();
}

void load_points(std::vector<Pt>& dst, const float* xyz, size_t n, size_t stride_floats) {
  dst.resize(n);
  for (size_t i = 0; i < n; i++) { dst[i].x = xyz[i * stride_floats]; dst[i].y = xyz[i * stride_floats + 1]; dst[i].z = xyz[i * stride_floats + 2]; }
}

}  // namespace

// ---------------------------------------------------------------- C entry points (ctypes)
extern "C" {

void* ondt_create(int variant) { NDT* n = new NDT(); n->variant = variant; compute_gauss(*n); return n; }
void ondt_destroy(void* h) { delete (NDT*)h; }

void ondt_set_params(void* h, float resolution, double step_size, double outlier_ratio, double trans_eps, int max_iter, int search, int num_threads) {
  NDT& n = *(NDT*)h;
  bool revox = (n.resolution != resolution) && !n.target.empty();   // setResolution re-inits (ndt_omp.h:126-136)
  n.resolution = resolution; n.step_size = step_size; n.outlier_ratio = outlier_ratio; n.trans_eps = trans_eps;
  n.max_iter = max_iter; n.search = search; n.num_threads = num_threads;
  compute_gauss(n);
  if (revox) apply_filter(n);
}

void ondt_set_target(void* h, const float* xyz, size_t npts, size_t stride_floats) {
  NDT& n = *(NDT*)h;
  load_points(n.target, xyz, npts, stride_floats);
  apply_filter(n);
}

void ondt_set_source(void* h, const float* xyz, size_t npts, size_t stride_floats) { load_points(((NDT*)h)->input, xyz, npts, stride_floats); }

void ondt_get_grid(void* h, int32_t* min_b, int32_t* max_b, int32_t* div_b) {
  NDT& n = *(NDT*)h;
  for (int a = 0; a < 3; a++) { min_b[a] = n.min_b[a]; max_b[a] = n.max_b[a]; div_b[a] = n.div_b[a]; }
}
void ondt_get_gauss(void* h, double* d) { NDT& n = *(NDT*)h; compute_gauss(n); d[0] = n.gauss_d1; d[1] = n.gauss_d2; d[2] = n.gauss_d3; }
int ondt_num_leaves(void* h) { return (int)((NDT*)h)->leaves.size(); }

// All occupied cells in ascending key order.  Any output pointer may be NULL.
void ondt_get_leaves(void* h, int32_t* keys, int32_t* nr_points, int32_t* raw_points, double* mean3, double* cov9, double* icov9,
                     double* evals3, float* centroid3, int32_t* weight, int32_t* label, int32_t* in_cloud) {
  NDT& n = *(NDT*)h;
  size_t k = 0;
  for (auto& kv : n.leaves) {
    const Leaf& l = kv.second;
    if (keys) keys[k] = (int32_t)kv.first;
    if (nr_points) nr_points[k] = l.nr_points;
    if (raw_points) raw_points[k] = l.raw_points;
    if (mean3) for (int i = 0; i < 3; i++) mean3[k * 3 + i] = l.mean[i];
    if (cov9) for (int i = 0; i < 9; i++) cov9[k * 9 + i] = l.cov[i / 3][i % 3];
    if (icov9) for (int i = 0; i < 9; i++) icov9[k * 9 + i] = l.icov[i / 3][i % 3];
    if (evals3) for (int i = 0; i < 3; i++) evals3[k * 3 + i] = l.evals[i];
    if (centroid3) for (int i = 0; i < 3; i++) centroid3[k * 3 + i] = l.centroid[i];
    if (weight) weight[k] = leaf_weight(n, l);
    if (label) label[k] = l.dimension_label;
    if (in_cloud) in_cloud[k] = l.in_centroid_cloud;
    k++;
  }
}

// pclomp_ground: per occupied cell (ascending key order) the angle of its normal to the z axis in degrees, -1 where the leaf
// has no eigen-decomposition (fewer than min_points_per_voxel points).
void ondt_get_leaf_angles(void* h, double* angle2xy) {
  NDT& n = *(NDT*)h;
  size_t k = 0;
  for (auto& kv : n.leaves) { angle2xy[k++] = kv.second.in_centroid_cloud ? leaf_angle2xy(kv.second) : -1.0; }
}

// Eigenvectors of every occupied cell (ascending key order), row-major, columns in the order of the eigenvalues; identity for leaves
// below min_points_per_voxel (the Leaf constructor's value, voxel_grid_covariance_omp.h:103).
void ondt_get_leaf_evecs(void* h, double* evecs9) {
  NDT& n = *(NDT*)h;
  size_t k = 0;
  for (auto& kv : n.leaves) { for (int i = 0; i < 9; i++) evecs9[k * 9 + i] = kv.second.evecs[i / 3][i % 3]; k++; }
}

// Keys of the cells a direct search returns for one point, in the order they are pushed.  mode: DIRECT26 / DIRECT7 / DIRECT1.
int ondt_neighbours(void* h, const float* xyz3, int mode, int32_t* keys_out /* >= 26 */) {
  NDT& n = *(NDT*)h;
  Pt p = {xyz3[0], xyz3[1], xyz3[2]};
  std::vector<const Leaf*> nb;
  neighbours_direct(n, p, mode, nb);
  for (size_t i = 0; i < nb.size(); i++) {
    int32_t key = -1;
    for (auto& kv : n.leaves) if (&kv.second == nb[i]) { key = (int32_t)kv.first; break; }
    keys_out[i] = key;
  }
  return (int)nb.size();
}

// Voxel key the lookup path computes for each (already transformed) point, or -1 when outside the box.
void ondt_lookup_keys(void* h, const float* xyz, size_t npts, size_t stride_floats, int32_t* keys) {
  NDT& n = *(NDT*)h;
  for (size_t i = 0; i < npts; i++) {
    const float* p = xyz + i * stride_floats;
    int ijk[3] = {(int)std::floor(p[0] / n.leaf_size), (int)std::floor(p[1] / n.leaf_size), (int)std::floor(p[2] / n.leaf_size)};
    bool in = !n.leaves.empty();
    for (int a = 0; a < 3; a++) if (ijk[a] < n.min_b[a] || ijk[a] > n.max_b[a]) in = false;
    keys[i] = in ? (ijk[0] - n.min_b[0]) * n.divb_mul[0] + (ijk[1] - n.min_b[1]) * n.divb_mul[1] + (ijk[2] - n.min_b[2]) * n.divb_mul[2] : -1;
  }
}

void ondt_transform(const float* xyz, size_t npts, size_t stride_floats, const float* T16, float* out_xyz /*packed 3*/) {
  for (size_t i = 0; i < npts; i++) {
    Pt p = {xyz[i * stride_floats], xyz[i * stride_floats + 1], xyz[i * stride_floats + 2]};
    Pt o = transform_pt(T16, p);
    out_xyz[i * 3] = o.x; out_xyz[i * 3 + 1] = o.y; out_xyz[i * 3 + 2] = o.z;
  }
}

// One computeDerivatives call on the current source: trans cloud = T16 * source (float), parameters p.
// If T16 is NULL it is SE3::exp(p) cast to float, which is what the line search uses.
double ondt_eval_derivatives(void* h, const double* p6, const float* T16, int compute_hessian, double* g6, double* H36) {
  NDT& n = *(NDT*)h;
  compute_gauss(n);
  float M[16];
  if (T16) std::memcpy(M, T16, sizeof M); else ose3::se3_to_matrix4f(ose3::se3_exp(p6), M);
  std::vector<Pt> trans; transform_cloud(n.input, trans, M, n.num_threads);
  double g[6], H[6][6];
  double s = compute_derivatives(n, g, H, trans, p6, compute_hessian != 0);
  for (int i = 0; i < 6; i++) { g6[i] = g[i]; for (int j = 0; j < 6; j++) H36[i * 6 + j] = H[i][j]; }
  return s;
}

void ondt_eval_hessian(void* h, const double* p6, const float* T16, double* H36) {
  NDT& n = *(NDT*)h;
  compute_gauss(n);
  float M[16];
  if (T16) std::memcpy(M, T16, sizeof M); else ose3::se3_to_matrix4f(ose3::se3_exp(p6), M);
  std::vector<Pt> trans; transform_cloud(n.input, trans, M, n.num_threads);
  double H[6][6];
  compute_hessian(n, H, trans, p6);
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H36[i * 6 + j] = H[i][j];
}

double ondt_calculate_score(void* h, const float* T16) {
  NDT& n = *(NDT*)h;
  compute_gauss(n);
  std::vector<Pt> trans; transform_cloud(n.input, trans, T16, n.num_threads);
  return calculate_score(n, trans);
}

// pcl::Registration::getFitnessScore(max_range) (PCL 1.8 registration/impl/registration.hpp, un-vendored; called at
// include/global_graph/loop_detector.hpp:176,255) and lv_slam::InformationMatrixCalculator::calc_fitness_score
// (src/global_graph/information_matrix_calculator.cpp:53-87): transform the source with T16 (pcl::transformPointCloud), nearest
// target point of every transformed point (kd-tree nearestKSearch, k = 1, FLANN L2_Simple = ((dx*dx + dy*dy) + dz*dz) in float;
// restated as an exhaustive scan, which returns the same minimum), squared distances <= max_range are summed in double in point
// order; mean, or DBL_MAX without a correspondence.  Non-finite points are skipped on both sides.
double ondt_fitness_score(void* h, const float* T16, double max_range, int* n_corr) {
  NDT& n = *(NDT*)h;
  std::vector<Pt> trans; transform_cloud(n.input, trans, T16, n.num_threads);
  std::vector<float> best(trans.size(), -1.0f);
#pragma omp parallel for num_threads(n.num_threads) schedule(static)
  for (long i = 0; i < (long)trans.size(); i++) {
    const Pt& q = trans[i];
    if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;
    float b = INFINITY;
    for (const Pt& t : n.target) {
      if (!(std::isfinite(t.x) && std::isfinite(t.y) && std::isfinite(t.z))) continue;
      const float dx = q.x - t.x, dy = q.y - t.y, dz = q.z - t.z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (d < b) b = d;
    }
    if (std::isfinite(b)) best[i] = b;
  }
  double sum = 0; int nr = 0;
  for (float b : best) if (b >= 0.0f && (double)b <= max_range) { sum += (double)b; nr++; }
  if (n_corr) *n_corr = nr;
  return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
}

// PrefilteringNodelet: distance_filter (src/lidar_odometry/prefiltering_nodelet.cpp:164-181) then downsample() = pcl::VoxelGrid
// (:41-47, 138-148).  pcl::VoxelGrid::applyFilter is un-vendored (PCL 1.8 filters/impl/voxel_grid.hpp); restated: bounding box of
// the finite points, min_b / max_b = floor(p * inverse_leaf), overflow guard dx*dy*dz > INT32_MAX (output = input), leaf index
// ijk . divb_mul with ijk = int(floor(x * inverse_leaf) - float(min_b)), points grouped by index (PCL: unstable std::sort on the
// index; here stable, i.e. input order inside a leaf), one point per leaf in ascending index = float sums / float(count)
// (downsample_all_data_ = true: the CentroidPoint accumulators average intensity as well).
// in: npts points of n_fields floats, stride_floats apart; out: packed n_fields floats.  Returns the number of output points;
// *flags bit 0 = overflow guard fired.
size_t oprefilter(const float* xyz, size_t npts, size_t stride_floats, int n_fields, double near_t, double far_t, int use_filter, float leaf,
                  float* out, int* flags) {
  struct P4 { float x, y, z, w; };
  std::vector<P4> kept;
  kept.reserve(npts);
  if (flags) *flags = 0;
  for (size_t i = 0; i < npts; i++) {
    const float* p = xyz + i * stride_floats;
    P4 q = {p[0], p[1], p[2], n_fields > 3 ? p[3] : 0.0f};
    if (use_filter) {
      const double d = (double)std::sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);      // Eigen Vector3f::norm(), float
      if (!(d > near_t && d < far_t)) continue;
    }
    kept.push_back(q);
  }
  auto emit = [&](const std::vector<P4>& v) {
    for (size_t i = 0; i < v.size(); i++) { out[i * n_fields] = v[i].x; out[i * n_fields + 1] = v[i].y; out[i * n_fields + 2] = v[i].z; if (n_fields > 3) out[i * n_fields + 3] = v[i].w; }
    return v.size();
  };
  if (!(leaf > 0.0f) || kept.empty()) return emit(kept);
  const float inv = 1.0f / leaf;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool any = false;
  for (const P4& q : kept) {
    if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;
    any = true;
    mn[0] = std::min(mn[0], q.x); mn[1] = std::min(mn[1], q.y); mn[2] = std::min(mn[2], q.z);
    mx[0] = std::max(mx[0], q.x); mx[1] = std::max(mx[1], q.y); mx[2] = std::max(mx[2], q.z);
  }
  if (!any) return 0;
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > 2147483647LL) { if (flags) *flags |= 1; return emit(kept); }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; a++) { min_b[a] = (int)std::floor(mn[a] * inv); max_b[a] = (int)std::floor(mx[a] * inv); div_b[a] = max_b[a] - min_b[a] + 1; }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<int, int>> order;       // (leaf index, point)
  for (size_t i = 0; i < kept.size(); i++) {
    const P4& q = kept[i];
    if (!(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;
    const int i0 = (int)(std::floor(q.x * inv) - (float)min_b[0]), i1 = (int)(std::floor(q.y * inv) - (float)min_b[1]), i2 = (int)(std::floor(q.z * inv) - (float)min_b[2]);
    order.push_back({i0 * mul[0] + i1 * mul[1] + i2 * mul[2], (int)i});
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
  size_t n_out = 0;
  for (size_t k = 0; k < order.size();) {
    size_t e = k;
    float sx = 0, sy = 0, sz = 0, sw = 0;
    while (e < order.size() && order[e].first == order[k].first) { const P4& q = kept[order[e].second]; sx += q.x; sy += q.y; sz += q.z; sw += q.w; e++; }
    const float fn = (float)(e - k);
    out[n_out * n_fields] = sx / fn; out[n_out * n_fields + 1] = sy / fn; out[n_out * n_fields + 2] = sz / fn;
    if (n_fields > 3) out[n_out * n_fields + 3] = sw / fn;
    n_out++;
    k = e;
  }
  return n_out;
}

// pcl::transformPointCloud with a double matrix, as the window map of the global-graph nodelet uses it
// (src/global_graph/global_graph_nodelet.cpp:239-241; PCL 1.8 common/impl/transforms.hpp, un-vendored): every coordinate is
// float(m_r0 * x + m_r1 * y + m_r2 * z + m_r3), left to right in double; other fields are copied.  T16 column-major.
void otransform_double(const float* xyz, size_t npts, size_t stride_floats, int n_fields, const double* T16, float* out) {
  for (size_t i = 0; i < npts; i++) {
    const float* p = xyz + i * stride_floats;
    const double x = p[0], y = p[1], z = p[2];
    float* o = out + i * n_fields;
    o[0] = (float)(T16[0] * x + T16[4] * y + T16[8] * z + T16[12]);
    o[1] = (float)(T16[1] * x + T16[5] * y + T16[9] * z + T16[13]);
    o[2] = (float)(T16[2] * x + T16[6] * y + T16[10] * z + T16[14]);
    if (n_fields > 3) o[3] = p[3];
  }
}

// align(): returns nr_iterations.  out_final16 column-major.  stats = {converged, trans_probability, n_eval, n_hess}
int ondt_align(void* h, const float* guess16, float* out_final16, double* stats4, float* out_cloud_xyz /*nullable, packed*/) {
  NDT& n = *(NDT*)h;
  std::vector<Pt> out;
  align(n, guess16, out);
  std::memcpy(out_final16, n.final_T, sizeof n.final_T);
  if (stats4) { stats4[0] = n.converged ? 1 : 0; stats4[1] = n.trans_probability; stats4[2] = (double)n.n_eval; stats4[3] = (double)n.n_hess; }
  if (out_cloud_xyz) for (size_t i = 0; i < out.size(); i++) { out_cloud_xyz[i * 3] = out[i].x; out_cloud_xyz[i * 3 + 1] = out[i].y; out_cloud_xyz[i * 3 + 2] = out[i].z; }
  return n.nr_iterations;
}

int ondt_trace_len(void* h) { return (int)((NDT*)h)->trace.size(); }
// each record: p_before[6] dir[6] step score p_after[6] trials hess  = 22 doubles
void ondt_get_trace(void* h, double* out) {
  NDT& n = *(NDT*)h;
  for (size_t k = 0; k < n.trace.size(); k++) {
    const Trace& t = n.trace[k];
    double* o = out + k * 22;
    std::memcpy(o, t.p_before, 48); std::memcpy(o + 6, t.delta_dir, 48); o[12] = t.step; o[13] = t.score;
    std::memcpy(o + 14, t.p_after, 48); o[20] = t.trials; o[21] = t.hessian_recomputed;
  }
}

// se(3) helpers exposed for the KATs
void ose3_exp_matrix4f(const double* p6, float* M16) { ose3::se3_to_matrix4f(ose3::se3_exp(p6), M16); }
void ose3_exp(const double* p6, double* q4_wxyz, double* t3) {
  ose3::SE3 T = ose3::se3_exp(p6);
  q4_wxyz[0] = T.q.w; q4_wxyz[1] = T.q.x; q4_wxyz[2] = T.q.y; q4_wxyz[3] = T.q.z;
  for (int i = 0; i < 3; i++) t3[i] = T.t.v[i];
}
void ose3_log_from_matrix4f(const float* M16, double* p6) {
  M3 R; V3 t;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R.a[r][c] = (double)M16[c * 4 + r]; t.v[r] = (double)M16[12 + r]; }
  ose3::se3_log(ose3::se3_from_Rt(R, t), p6);
}
void ose3_compose_log(const double* delta6, const double* p6, double* out6) {
  ose3::se3_log(ose3::se3_mul(ose3::se3_exp(delta6), ose3::se3_exp(p6)), out6);
}
void olin_svd6_solve(const double* A36, const double* b6, double* x6, double* sv6) { olin::svd6_solve(A36, b6, x6, sv6); }
void olin_sym3_eig(const double* A9, double* evals3, double* V9) {
  M3 A, V; std::memcpy(A.a, A9, sizeof A.a);
  olin::sym3_eig(A, evals3, V);
  std::memcpy(V9, V.a, sizeof V.a);
}

}  // extern "C"

// glibc expf (sysdeps/ieee754/flt-32/e_expf.c, e_exp2f_data.c: N = 32, cubic), restated with the contractions of its FMA build.
static const uint64_t kExp2fTab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, 0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL,
    0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, 0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, 0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL,
    0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, 0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, 0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL,
    0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};
static float expf_restated(float x) {
  if (!(x >= -0x1.9fe368p6f)) return x != x ? x + x : 0.0f;
  if (x > 0x1.62e42ep6f) return INFINITY;
  const double xd = (double)x;
  const double z = 0x1.71547652b82fep+5 * xd;
  volatile double kdv = z + 0x1.8p+52;
  double kd = kdv;
  uint64_t ki; std::memcpy(&ki, &kd, 8);
  kd -= 0x1.8p+52;
  const double r = std::fma(0x1.71547652b82fep+5, xd, -kd);
  const uint64_t t = kExp2fTab[ki & 31] + (ki << 47);
  double s; std::memcpy(&s, &t, 8);
  const double zz = std::fma(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
  const double r2 = r * r;
  double y = std::fma(0x1.62e42ff0c52d6p-6, r, 1.0);
  y = std::fma(zz, r2, y);
  return (float)(y * s);
}
extern "C" {
// out[i] = restated expf(x[i]); returns the number of inputs whose result differs from the host libm's expf bit for bit
long long oracle_expf_restated(const float* x, float* out, long long n) {
  long long bad = 0;
  for (long long i = 0; i < n; i++) {
    volatile float xi = x[i];
    const float a = expf_restated(xi), b = ::expf(xi);
    uint32_t ua, ub; std::memcpy(&ua, &a, 4); std::memcpy(&ub, &b, 4);
    if (ua != ub && !(a != a && b != b)) bad++;
    if (out) out[i] = a;
  }
  return bad;
}
}
