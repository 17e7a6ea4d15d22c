// Device-side data layout of the NDT path (see DESIGN.md "Data layout in HBM").
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace lvs {

// One record per occupied target cell, 64 B, 16 B aligned: four LDG.128 per probe.
//   mean  : f64 x3  (the reference subtracts the double mean from the float point in double,
//                    include/ndt_omp/ndt_omp_impl2.hpp:272-275)
//   icov  : f32 x9  = Matrix3d::cast<float>() of the inverse covariance (:572-573); stored full because V*L*V^-1 rebuilt
//                    covariances are not exactly symmetric.  Storage order C00 C01 | C10 C11 | C20 C21 | C02 C12 C22 (kIcovSlot):
//                    the (C[r][0], C[r][1]) pairs land 8-byte aligned for the packed FP32 math of the hot kernel
//   meta  : low 24 bits = ndt_pca integer weight int(scale*|mean|) (1 for ndt_omp),
//           bit 30 = leaf usable by the direct searches (nr_points >= min_points and not invalidated),
//           bit 29 = pclomp_ground only: the leaf's normal is within 10 degrees of the z axis (ndt_ground_impl.hpp:507-511,533)
struct __align__(16) VoxelRec {
  double mean[3];
  float icov[9];
  int32_t meta;
};
static_assert(sizeof(VoxelRec) == 64, "VoxelRec must be 64 bytes");

// Record of the same cell for the tolerance-mode evaluation (lvs_ndt_params::accumulation = LVS_ACC_FAST), 48 B, three LDG.128:
//   mh + ml : the double mean split into two floats (mh = float(mean), ml = float(mean - mh)): (x' - mh) - ml reproduces the
//             reference's double subtraction to ~1e-7 m without a conversion instruction
//   c       : float inverse covariance, upper triangle c00 c01 c02 c11 c12 c22 (V*L*V^-1 rebuilt covariances are symmetric to ~1e-16)
struct __align__(16) FastRec {
  float mh[3], ml[3];
  float c[6];
};
static_assert(sizeof(FastRec) == 48, "FastRec must be 48 bytes");

// storage position of the row-major element a = 3 r + c inside VoxelRec::icov
__host__ __device__ constexpr int icov_slot(int a) { return (a % 3 == 2) ? 6 + a / 3 : 2 * (a / 3) + a % 3; }

constexpr int kMetaValidBit = 1 << 30;
constexpr int kMetaHorizBit = 1 << 29;
constexpr int kMetaWeightMask = 0xFFFFFF;

// Geometry of one target's voxel grid (voxel_grid_covariance_omp_impl.hpp:87-103), produced on the device.
struct GridParams {
  int min_b[3], max_b[3], div_b[3], mul[3];
  float leaf, inv_leaf;
  float min_p[3], max_p[3];
  int n_points;        // finite target points
  int n_cells;         // occupied cells (== leaves_.size())
  int n_valid;         // cells usable by the direct searches
  int status;          // 0 ok, LVS_ERR_GRID_OVERFLOW when dx*dy*dz > INT32_MAX, 1 = empty cloud
  long long total_cells;
};

// Grid cell encoding: -1 empty, v >= 0 record index usable by direct search,
// v <= -2 occupied but not usable by direct search (record index = -2 - v).
__host__ __device__ inline int grid_decode_any(int v) { return v >= 0 ? v : -2 - v; }

constexpr int kAcc = 43;   // score + gradient[6] + full Hessian[36] (the reference's H is not symmetric)
constexpr int kPartialStride = 44;   // doubles between the partial vectors of two CTAs (even: the last CTA reads them as double2)

enum EvalKind { EVAL_NONE = -1, EVAL_DERIV_H = 0, EVAL_DERIV_NOH = 1, EVAL_HESS27 = 2 };
enum Phase { PH_INIT = 0, PH_MT_FIRST = 1, PH_MT_TRIAL = 2, PH_HESS27 = 3, PH_DONE = 9 };

constexpr int kMaxTrace = 72;

struct TraceRec {
  double p_before[6], dir[6], step, score, p_after[6];
  int trials, hessian_recomputed;
};

// Complete state of one align() — lives in device memory, advanced by the last CTA of each evaluation.
struct AlignState {
  // what the next evaluation kernel has to do for this pair
  int eval_kind;
  int phase;
  float T[16];          // column-major transform applied to the source cloud (transformPointCloud)
  float Rj[9];          // row-major float rotation of SE3::exp(x_t) used by the point Jacobian
  double Rd[9];         // the same rotation in double (computeHessian's double point Jacobian)
  // Newton state
  double p[6];          // current parameter vector (group-composed, ndt_omp_impl2.hpp:166)
  double x_t[6];        // evaluation point of the line search
  double score, g[6], H[36];
  int nr_iterations, converged, n_eval, n_hess;
  // More-Thuente state (ndt_omp_impl2.hpp:842-1003)
  double dir[6], delta_norm, phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t;
  int open_interval, interval_converged, step_iterations;
  // results
  float final_T[16];
  double trans_probability;
  int n_trace;
  int trace_on;
};

struct AlignConsts {
  double gauss_d1, gauss_d2, gauss_d3;
  double step_size, trans_eps;
  int max_iter;
  int search;
  int variant;
  float resolution;
  int fast;                 // lvs_ndt_params::accumulation == LVS_ACC_FAST
  int lean_final;           // lvs_ndt_params::lean_final_evaluation
};

// One (source, target) pair as the kernels see it.
struct PairDesc {
  const float4* src;
  int n_src;                // points this device evaluates (its shard of the source when the batch is point-sharded)
  int n_total;              // points of the whole source cloud (trans_probability divisor, ndt_omp_impl2.hpp:187)
  const int* grid;
  const VoxelRec* recs;
  const FastRec* frecs;     // tolerance-mode records (same indices as recs)
  const float4* centroids;
  const double* icov64;     // [n_cells][9] double inverse covariance (computeHessian / calculateScore are all-double)
  const GridParams* gp;
};

}  // namespace lvs
