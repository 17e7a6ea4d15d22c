/* lvslam_b200 — C-ABI of the B200-native NDT scan-matching and pose-graph hot path.
 *
 * Drop-in boundary for BurryChen/lv_slam (reference paths relative to the reference tree):
 *   - the NDT entry points replace the bodies behind pcl::Registration::align() for
 *     pclomp::NormalDistributionsTransform  (include/ndt_omp/ndt_omp.h:69-551, computeTransformation at :256-267)
 *     pclpca::NormalDistributionsTransform  (include/ndt_pca/ndt_pca.h, same layout)
 *   - the pose-graph entry points replace lv_slam::GraphSLAM::optimize()
 *     (include/global_graph/graph_slam.hpp:40-149, src/global_graph/graph_slam.cpp:298-331).
 * INTEGRATION.md shows the C++ shim a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only; every function returns an lvs_status (0 = ok, < 0 = error);
 * no exceptions cross the boundary; 4x4 matrices are COLUMN-major float[16] (Eigen::Matrix4f memory order);
 * a handle is single-owner and not re-entrant, several handles may be used concurrently from different
 * host threads (the two reference nodelets do exactly that).  There is no CPU fallback: every entry point
 * fails with LVS_ERR_CUDA / LVS_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef LVSLAM_B200_H_
#define LVSLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum {
  LVS_OK = 0,
  LVS_ERR_INVALID_ARG = -1,
  LVS_ERR_NO_DEVICE = -2,
  LVS_ERR_CUDA = -3,
  LVS_ERR_OOM = -4,
  LVS_ERR_NO_TARGET = -5,     /* align()/eval before setInputTarget */
  LVS_ERR_NO_SOURCE = -6,     /* align()/eval before setInputSource */
  LVS_ERR_GRID_OVERFLOW = -7, /* dx*dy*dz > INT32_MAX: the reference warns and yields an empty grid
                                 (include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:76-85) */
  LVS_ERR_BAD_SLOT = -8,
  LVS_ERR_NOT_SPD = -9,       /* pose graph: Cholesky failure (g2o LinearSolver returns false) */
  LVS_ERR_EMPTY_GRAPH = -10,  /* GraphSLAM::optimize returns -1 when the graph has no edge */
  LVS_ERR_PEER = -11          /* point-sharded evaluation: a peer rank did not deliver its sums in time */
} lvs_status;

/* pclomp::NeighborSearchMethod, include/ndt_omp/ndt_omp.h:61 (same values in pclpca) */
typedef enum { LVS_KDTREE = 0, LVS_DIRECT26 = 1, LVS_DIRECT7 = 2, LVS_DIRECT1 = 3 } lvs_search_method;
/* which registration class is mirrored: pclomp:: (include/ndt_omp/ndt_omp.h), pclpca:: (include/ndt_pca/ndt_pca.h) or
 * pclomp_ground::NormalDistributionsTransformGround (include/ndt_omp/ndt_ground.h, ndt_ground_impl.hpp) - the horizontal-voxel NDT
 * that solves z / roll / pitch only: a source point counts when the LAST cell of its neighbourhood has a normal within 10 degrees of
 * the z axis (computeDerivatives_seg with flag_class = 1, ndt_ground_impl.hpp:363-572), its gradient and Hessian terms count twice
 * (updateDerivatives is called twice per cell, :519,522), rows and columns x, y, yaw are zeroed before the SVD solve (:554-561) and
 * the step-length test also ends the first iteration (:173).  LVS_NDT_GROUND runs with LVS_ACC_EXACT only. */
typedef enum { LVS_NDT_OMP = 0, LVS_NDT_PCA = 1, LVS_NDT_GROUND = 2 } lvs_ndt_variant;
/* How the per-(point, cell) terms of computeDerivatives (ndt_omp_impl2.hpp:567-619) are formed and summed.
 *   LVS_ACC_EXACT (default): every float32 term is formed in the reference's operation order without contraction (bit-identical to
 *                 the CPU path, expf included) and added in fp64, like `hessian(i,j) += float_expr` does.
 *   LVS_ACC_FAST : tolerance mode.  Terms are formed in float32 with fused multiply-adds and a hardware exp2, the 31 distinct
 *                 sums (score, gradient, symmetric translation blocks of H, full rotation block) are kept per thread in float32
 *                 for at most 32 terms and then added in fp64.  Voxel indices stay bit-exact.  Agrees with LVS_ACC_EXACT to
 *                 ~1e-6 relative on (score, g, H), i.e. <= 1e-4 m / 1e-5 rad per Newton iteration; about 3x faster.
 * KDTREE search and the all-double computeHessian pass always run exact. */
typedef enum { LVS_ACC_EXACT = 0, LVS_ACC_FAST = 1 } lvs_ndt_accumulation;

/* Parameters = the setters of the reference class; defaults = its constructor
 * (include/ndt_omp/ndt_omp_impl2.hpp:54-83). */
typedef struct {
  float resolution;            /* setResolution            (1.0)  */
  double step_size;            /* setStepSize              (0.1)  */
  double outlier_ratio;        /* setOulierRatio [sic]     (0.55) */
  double transformation_epsilon; /* setTransformationEpsilon (0.1) */
  int32_t max_iterations;      /* setMaximumIterations     (35)   */
  int32_t search_method;       /* setNeighborhoodSearchMethod (LVS_DIRECT7) */
  int32_t variant;             /* lvs_ndt_variant (LVS_NDT_OMP) */
  int32_t min_points_per_voxel;  /* VoxelGridCovariance default 6 (voxel_grid_covariance_omp.h:204) */
  double min_covar_eigvalue_mult; /* 0.01 (voxel_grid_covariance_omp.h:205) */
  int32_t accumulation;        /* lvs_ndt_accumulation; not a reference parameter (LVS_ACC_EXACT) */
  int32_t lean_final_evaluation; /* not a reference parameter (0).  1: the derivative pass that ENDS an align computes the score and the
                                * gradient only.  The reference computes that pass's Hessian too and never reads it: whether the step just
                                * taken ends the iteration is known before the pass (its More-Thuente loop is dead code for step_size >
                                * epsilon / 2, ndt_omp_impl2.hpp:888-891, so the step length is fixed beforehand; :175-179), and only the
                                * pass's score survives (trans_probability_, :187).  Every result of align() is unchanged. */
} lvs_ndt_params;

/* What the reference object exposes after align(): getFinalTransformation, hasConverged,
 * getFinalNumIteration, getTransformationProbability, plus evaluation counters. */
typedef struct {
  float final_transformation[16];
  int32_t converged;
  int32_t iterations;
  double trans_probability;
  int32_t n_eval;   /* computeDerivatives calls */
  int32_t n_hess;   /* computeHessian calls */
  double score;
} lvs_ndt_result;

/* One record per Newton iteration (parity tap): parameter vector before the step, unit direction,
 * accepted step length, score after the step, parameter vector after composition, line-search trials,
 * whether computeHessian ran. */
typedef struct {
  double p_before[6], dir[6], step, score, p_after[6];
  int32_t trials, hessian_recomputed;
} lvs_ndt_trace_rec;

typedef struct lvs_ndt lvs_ndt_t;

void lvs_ndt_default_params(lvs_ndt_params* p);
const char* lvs_status_string(int status);
const char* lvs_last_error(void);        /* thread-local detail string of the last failure */
int lvs_device_count(void);

/* One registration object (replaces pclomp::/pclpca:: NormalDistributionsTransform).  `stream` is a
 * cudaStream_t (NULL = the handle creates its own); all device work of the handle is ordered on it. */
int lvs_ndt_create(const lvs_ndt_params* params, int device, void* stream, lvs_ndt_t** out);
int lvs_ndt_destroy(lvs_ndt_t* h);
int lvs_ndt_set_params(lvs_ndt_t* h, const lvs_ndt_params* params);  /* re-voxelises if resolution changed (ndt_omp.h:126-136) */
int lvs_ndt_get_params(const lvs_ndt_t* h, lvs_ndt_params* out);

/* setInputTarget (ndt_omp.h:116-121 -> init() :283-289 -> VoxelGridCovariance::filter(true)).
 * xyz points to the first x; consecutive points are stride_bytes apart (32 for pcl::PointXYZI, 12 packed).
 * on_device != 0: xyz is a device pointer on the handle's device (inputs already resident in HBM). */
int lvs_ndt_set_target(lvs_ndt_t* h, const float* xyz, size_t n, size_t stride_bytes, int on_device);
/* setInputSource */
int lvs_ndt_set_source(lvs_ndt_t* h, const float* xyz, size_t n, size_t stride_bytes, int on_device);

/* align(output, guess): runs computeTransformation (ndt_omp_impl2.hpp:88-188) on the device. */
int lvs_ndt_align(lvs_ndt_t* h, const float guess[16], lvs_ndt_result* out);
/* The `output` cloud of align(): final_transformation * source, packed xyz float, n_source points (on a point-sharded object: the
 * points of this rank's chunk only). */
int lvs_ndt_get_aligned_cloud(lvs_ndt_t* h, float* xyz_out, int on_device);
/* Per-iteration trace of the last align (at most max_iterations + 2 records). */
int lvs_ndt_get_trace(lvs_ndt_t* h, lvs_ndt_trace_rec* recs, int capacity, int* n_out);

/* Parity taps.
 * eval_derivatives = one computeDerivatives call (ndt_omp_impl2.hpp:197-305): transformed cloud =
 * T16 * source (T16 NULL: float(SE3::exp(p))), point Jacobians from p.  H36 row-major, full 6x6. */
int lvs_ndt_eval_derivatives(lvs_ndt_t* h, const double p[6], const float* T16, int compute_hessian,
                             double* score, double g[6], double H36[36]);
/* computeHessian (ndt_omp_impl2.hpp:623-679): all-double, radius-search neighbours, unweighted. */
int lvs_ndt_eval_hessian(lvs_ndt_t* h, const double p[6], const float* T16, double H36[36]);
/* calculateScore (ndt_omp_impl2.hpp:1007-1040) of T16 * source. */
int lvs_ndt_calculate_score(lvs_ndt_t* h, const float T16[16], double* score);

/* getFitnessScore(max_range) as the loop detector calls it right after align() (include/global_graph/loop_detector.hpp:176,255;
 * pcl::Registration::getFitnessScore, PCL 1.8 registration.hpp) and InformationMatrixCalculator::calc_fitness_score
 * (src/global_graph/information_matrix_calculator.cpp:53-87): mean SQUARED distance from every point of T16 * source to its nearest
 * target point over the correspondences whose squared distance is <= max_range; DBL_MAX when there is none.  T16 NULL = the final
 * transformation of the last align().  n_correspondences may be NULL. */
int lvs_ndt_fitness_score(lvs_ndt_t* h, const float* T16, double max_range, double* score, int* n_correspondences);

/* Voxel grid taps.  grid: min_b, max_b, div_b (voxel_grid_covariance_omp_impl.hpp:87-103). */
int lvs_ndt_get_grid(lvs_ndt_t* h, int32_t min_b[3], int32_t max_b[3], int32_t div_b[3]);
int lvs_ndt_num_cells(lvs_ndt_t* h, int* n_cells);   /* every occupied cell, i.e. leaves_.size() */
/* All occupied cells in ascending key order; any pointer may be NULL.  nr_points is the raw count
 * (-1 where the reference invalidates the leaf); mean/cov/icov/evals are zero for cells below
 * min_points_per_voxel except mean, exactly like the reference's Leaf. */
int lvs_ndt_get_cells(lvs_ndt_t* h, int32_t* keys, int32_t* nr_points, double* mean3, double* icov9,
                      double* evals3, float* centroid3, int32_t* weight);
/* LVS_NDT_GROUND: per occupied cell (ascending key order, like lvs_ndt_get_cells) 1 when the cell's normal - the eigenvector of the
 * smallest covariance eigenvalue, Leaf::getEvecs().col(0) - is less than 10 degrees from the z axis (ndt_ground_impl.hpp:507-511,533),
 * else 0; all 0 for the other variants. */
int lvs_ndt_get_cell_horizontal(lvs_ndt_t* h, int32_t* horizontal);
/* Voxel key the lookup path computes for T16 * source point i, -1 outside the bounding box
 * (voxel_grid_covariance_omp_impl.hpp:379-394). */
int lvs_ndt_lookup_keys(lvs_ndt_t* h, const float T16[16], int32_t* keys_out);

/* ---- batched registration: many (source, target, guess) pairs advanced together on one device ----
 * Targets and sources live in slots; a pair names one of each.  This is the throughput path used for
 * loop-closure candidates (include/global_graph/loop_detector.hpp:238-263) and stream replay. */
typedef struct lvs_ndt_batch lvs_ndt_batch_t;
int lvs_ndt_batch_create(const lvs_ndt_params* params, int device, void* stream, int n_target_slots,
                         int n_source_slots, lvs_ndt_batch_t** out);
int lvs_ndt_batch_destroy(lvs_ndt_batch_t* b);
int lvs_ndt_batch_set_target(lvs_ndt_batch_t* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device);
int lvs_ndt_batch_set_source(lvs_ndt_batch_t* b, int slot, const float* xyz, size_t n, size_t stride_bytes, int on_device);
/* Plural forms: n clouds in one call (slots[i], xyz[i], counts[i]; one stride and residency for all).  Same semantics as n
 * single calls; resident source clouds are repacked by one kernel launch instead of n. */
int lvs_ndt_batch_set_targets(lvs_ndt_batch_t* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts,
                              size_t stride_bytes, int on_device);
int lvs_ndt_batch_set_sources(lvs_ndt_batch_t* b, int n, const int32_t* slots, const float* const* xyz, const size_t* counts,
                              size_t stride_bytes, int on_device);
/* set_target / set_source with a HOST pointer queue the copy and the repack on the object's upload stream and return; the
 * consumer (voxelisation, align) waits on the device for exactly the clouds it reads, so the copies of later clouds overlap
 * the aligns of earlier ones.  A pageable host buffer may be reused as soon as the call returns; a PINNED host buffer is read
 * by DMA and must stay untouched until lvs_ndt_batch_wait_uploads (or an align that consumes the slot) has returned. */
int lvs_ndt_batch_wait_uploads(lvs_ndt_batch_t* b);
int lvs_ndt_batch_align(lvs_ndt_batch_t* b, int n_pairs, const int32_t* source_slot, const int32_t* target_slot,
                        const float* guesses16 /* n_pairs x 16 */, lvs_ndt_result* results /* n_pairs */);
/* The same align in two halves, for callers that stream: align_begin uploads the pair states and queues the evaluation launches of
 * a typical align without waiting (the Newton / More-Thuente state machine runs on the device); until align_end collects the
 * results the host is free to hand the NEXT batch to set_target(s) / set_source(s) - in other slots than the pairs in flight use -
 * so that its copies and voxelisations overlap the aligns (a device-resident cloud handed over in between must be complete, or
 * have been produced on the object's stream before align_begin).  One align in flight per object; every other entry point that
 * evaluates (align, fitness score, taps) returns LVS_ERR_INVALID_ARG in between.  lvs_ndt_batch_align = begin + end. */
int lvs_ndt_batch_align_begin(lvs_ndt_batch_t* b, int n_pairs, const int32_t* source_slot, const int32_t* target_slot, const float* guesses16);
int lvs_ndt_batch_align_end(lvs_ndt_batch_t* b, lvs_ndt_result* results /* n_pairs of the begin call */);
/* Device time in ms of the last batch_align and the number of kernel launches it issued. */
int lvs_ndt_batch_last_stats(lvs_ndt_batch_t* b, double* device_ms, int* launches, double* deriv_kernel_ms, int* deriv_launches);
/* Measurement controls.  profiling != 0 brackets every evaluation launch with a CUDA event pair so that
 * last_stats can report the summed duration of the launches that did work (deriv_kernel_ms over deriv_launches). */
int lvs_ndt_batch_set_profiling(lvs_ndt_batch_t* b, int on);
/* blocks_per_pair (0 = automatic), evaluation launches queued before the first / each later completion check. */
int lvs_ndt_batch_set_tuning(lvs_ndt_batch_t* b, int blocks_per_pair, int chunk_first, int chunk_next);
/* Kernel launches issued by this object since creation (voxelisation, evaluation, packing). */
int lvs_ndt_batch_total_launches(lvs_ndt_batch_t* b, long long* launches);
/* Bytes this object has copied host->device and device->host since creation (clouds, pair states, results). */
int lvs_ndt_batch_transfer_bytes(lvs_ndt_batch_t* b, long long* h2d, long long* d2h);
int lvs_ndt_batch_num_cells(lvs_ndt_batch_t* b, int target_slot, int* n_cells, int* n_valid);
/* lvs_ndt_fitness_score for a (source slot, target slot) pair of a batch object (loop-closure candidates). */
int lvs_ndt_batch_fitness_score(lvs_ndt_batch_t* b, int source_slot, int target_slot, const float T16[16], double max_range, double* score,
                                int* n_correspondences);
/* ---- point-sharded evaluation across the GPUs of one node (one process per GPU) ----
 * The sum over source points of computeDerivatives (ndt_omp_impl2.hpp:197-305) is split over `world` ranks: every rank is given
 * the same targets, sources, pairs and guesses; it keeps the contiguous chunk [rank*n/world, (rank+1)*n/world) of every source and
 * the 43 sums of each evaluation are exchanged through NVLink peer memory inside the evaluation kernel (no collective launch, no host
 * round trip).  All ranks return bit-identical results.  Set-up: shard_init on every rank -> exchange the 64-byte handles by any
 * means (torch.distributed all_gather, MPI, a file) -> shard_connect with all of them in rank order.  Every rank must then issue
 * the same sequence of align / eval calls.  world == 1 is allowed (self-exchange; used by the single-GPU tests). */
#define LVS_IPC_HANDLE_BYTES 64
int lvs_ndt_batch_shard_init(lvs_ndt_batch_t* b, int rank, int world, int max_pairs, unsigned char handle_out[LVS_IPC_HANDLE_BYTES]);
int lvs_ndt_batch_shard_connect(lvs_ndt_batch_t* b, const unsigned char* handles /* world x LVS_IPC_HANDLE_BYTES */);
/* The batch object that backs a single registration handle (1 target slot, 1 source slot). */
int lvs_ndt_handle_batch(lvs_ndt_t* h, lvs_ndt_batch_t** out);

/* ===================================================================================================================
 * Scan prefilter — the stage right before the NDT path: PrefilteringNodelet::distance_filter + downsample()
 * (src/lidar_odometry/prefiltering_nodelet.cpp:164-181, 138-148; pcl::VoxelGrid with downsample_resolution, launch/
 * dlo_lfa_ggo_kitti.launch:30-36).  Points are n_fields (3 = x y z, 4 = x y z intensity) floats, stride_bytes apart (32 for
 * pcl::PointXYZI); the result is packed n_fields floats per point, at most `capacity` points.
 *   use_distance_filter != 0: keep a point when its float norm d satisfies d > distance_near && d < distance_far (order kept);
 *   leaf_size > 0: pcl::VoxelGrid — one output point per occupied leaf in ascending leaf index, the float mean of its points;
 *                  flags bit 0 is set when PCL's "leaf size is too small" overflow guard fires and the cloud is passed through.
 */
typedef struct lvs_prefilter lvs_prefilter_t;
int lvs_prefilter_create(int device, void* stream, lvs_prefilter_t** out);
int lvs_prefilter_destroy(lvs_prefilter_t* p);
int lvs_prefilter_run(lvs_prefilter_t* p, const float* xyz, size_t n, size_t stride_bytes, int n_fields, int on_device, double distance_near,
                      double distance_far, int use_distance_filter, float leaf_size, float* out, size_t capacity, int out_on_device, size_t* n_out,
                      int* flags_out);

/* Window map of the global-graph nodelet (src/global_graph/global_graph_nodelet.cpp:199-243): between two keyframes every scan is
 * moved into the window's frame with pcl::transformPointCloud and the DOUBLE matrix w_odom^-1 * odom, appended to w_cloud, and the
 * window is downsampled by a 0.1 m VoxelGrid when the next keyframe is declared.  begin = w_cloud.clear(); add = w_cloud +=
 * transform(cloud, T16) (T16 column-major double[16], NULL = the window's own first scan, no transform); flush = VoxelGrid(leaf). */
int lvs_prefilter_accumulate_begin(lvs_prefilter_t* p);
int lvs_prefilter_accumulate_add(lvs_prefilter_t* p, const float* xyz, size_t n, size_t stride_bytes, int n_fields, int on_device, const double* T16);
int lvs_prefilter_accumulate_flush(lvs_prefilter_t* p, int n_fields, float leaf_size, float* out, size_t capacity, int out_on_device, size_t* n_out,
                                   int* flags_out);

/* ===================================================================================================================
 * Pose graph — replaces lv_slam::GraphSLAM::optimize (src/global_graph/graph_slam.cpp:298-331) and the g2o machinery under
 * it for graphs of VertexSE3 / EdgeSE3 with optional Huber kernels (what global_graph builds with GPS/IMU/floor disabled,
 * launch/dlo_lfa_ggo_kitti.launch:8-11).
 *
 * Solver kinds = the `solver_type` strings GraphSLAM's constructor accepts (graph_slam.cpp:48-70, g2o solver registry):
 *   LVS_PGO_LM_CHOL  "lm_var", "lm_var_cholmod", "lm_fix6_3*"   Levenberg-Marquardt, linear system solved to round-off
 *   LVS_PGO_GN_CHOL  "gn_var", "gn_var_cholmod", ...            Gauss-Newton (needs a fixed vertex: no damping)
 *   LVS_PGO_LM_PCG   "lm_pcg"                                    Levenberg-Marquardt + g2o's block-Jacobi PCG (tolerance 1e-6)
 *   LVS_PGO_GN_PCG   "gn_pcg"
 * Poses and measurements are 7-vectors  x y z qx qy qz qw  (g2o's VERTEX_SE3:QUAT / EDGE_SE3:QUAT order); information
 * matrices are the 21 upper-triangular entries, row-major, as in g2o files.  Edge k connects vertices()[0] = ij[2k] ("from")
 * and vertices()[1] = ij[2k+1] ("to") exactly as GraphSLAM::add_se3_edge(v1, v2, ...) does (graph_slam.cpp:136-146).
 * huber_delta[k] <= 0 (or a NULL array) means no robust kernel on edge k (add_robust_kernel, graph_slam.cpp:278-296).
 * fixed[v] != 0 mirrors VertexSE3::setFixed(true); NULL = nothing fixed (the nodelet's default, fix_first_node = false).
 */
typedef enum { LVS_PGO_LM_CHOL = 0, LVS_PGO_GN_CHOL = 1, LVS_PGO_LM_PCG = 2, LVS_PGO_GN_PCG = 3 } lvs_pgo_solver;

typedef struct {
  int32_t iterations;        /* the return value of GraphSLAM::optimize: iterations run, 0 on solver failure, -1 on an empty graph */
  int32_t status;            /* lvs_status of the run */
  double chi2_before;        /* graph->chi2() before (sum of e^T Omega e, not robustified; graph_slam.cpp:316) */
  double chi2_after;         /* the same sum at the final estimate */
  double robust_chi2_after;  /* activeRobustChi2 at the final estimate (what LM compares) */
  double lambda_final;
  double device_ms, linearize_ms, solve_ms;
  int32_t lm_trials, pcg_iterations, launches, linearize_launches;
} lvs_pgo_stats;

typedef struct { double chi2, lambda; int32_t trials, pcg_iterations; } lvs_pgo_iter_rec;   /* one per outer iteration */

typedef struct lvs_pgo lvs_pgo_t;

int lvs_pgo_create(int solver, int device, void* stream, lvs_pgo_t** out);
int lvs_pgo_destroy(lvs_pgo_t* h);
int lvs_pgo_set_graph(lvs_pgo_t* h, int n_vertices, const double* poses7, const uint8_t* fixed, int n_edges, const int32_t* ij,
                      const double* meas7, const double* info21, const double* huber_delta);
/* The same with the reference's unary priors on a VertexSE3 in the edge list (GPS / IMU constraints of the global-graph nodelet:
 * GraphSLAM::add_se3_prior_{xy,xyz,quat,vec}_edge, src/global_graph/graph_slam.cpp:194-240; edge classes include/g2o/edge_se3_priorxy.hpp,
 * edge_se3_priorxyz.hpp, edge_se3_priorquat.hpp, edge_se3_priorvec.hpp; numeric Jacobians as g2o's BaseUnaryEdge::linearizeOplus takes them).
 * edge_type[k] (NULL: every edge is an EdgeSE3) is one of LVS_PGO_EDGE_*; for a prior, ij[2k] is the vertex (ij[2k+1] is ignored), the
 * meas7 row carries  xy | xyz | qx qy qz qw | direction(3) measurement(3)  and the info21 row the D x D information matrix in the top-left
 * corner of the 6 x 6 upper triangle (D = 2 for XY, 3 otherwise; the other entries must be zero).  Edges keep their order in the list, which is
 * g2o's active-edge order (ascending edge id). */
enum { LVS_PGO_EDGE_SE3 = 0, LVS_PGO_EDGE_PRIOR_XY = 1, LVS_PGO_EDGE_PRIOR_XYZ = 2, LVS_PGO_EDGE_PRIOR_QUAT = 3, LVS_PGO_EDGE_PRIOR_VEC = 4, LVS_PGO_EDGE_SE3_PLANE = 5 };
/* LVS_PGO_EDGE_SE3_PLANE: the floor constraint, EdgeSE3Plane (include/g2o/edge_se3_plane.hpp) between a pose and the ONE plane vertex the nodelet creates
 * and fixes (add_plane_node + setFixed(true), global_graph_nodelet.cpp:601-611; GraphSLAM::add_se3_plane_edge, graph_slam.cpp:148-158): with that vertex fixed
 * the edge constrains the pose alone.  Its meas7 row carries the measured plane's 4 coefficients, info21 the 3 x 3 information; the vertex's plane is
 * set once per handle, before set_graph (default 0 0 1 0, the nodelet's). */
int lvs_pgo_set_floor_plane(lvs_pgo_t* h, const double coeffs[4]);
int lvs_pgo_set_graph_typed(lvs_pgo_t* h, int n_vertices, const double* poses7, const uint8_t* fixed, int n_edges, const int32_t* ij,
                            const double* meas7, const double* info21, const double* huber_delta, const int32_t* edge_type);
int lvs_pgo_set_poses(lvs_pgo_t* h, const double* poses7);          /* VertexSE3::setEstimate for every vertex */
int lvs_pgo_optimize(lvs_pgo_t* h, int max_iterations, lvs_pgo_stats* stats);
int lvs_pgo_get_poses(lvs_pgo_t* h, double* poses7);                /* VertexSE3::estimate() of every vertex */
int lvs_pgo_get_trace(lvs_pgo_t* h, lvs_pgo_iter_rec* recs, int capacity, int* n_out);
/* pcg_tolerance > 0 / pcg_max_iterations > 0 override the per-solver defaults (LinearSolverPCG::setTolerance / setMaxIterations). */
int lvs_pgo_set_solver_options(lvs_pgo_t* h, double pcg_tolerance, int pcg_max_iterations);
/* Parity taps: EdgeSE3::computeError + chi2 per edge and activeRobustChi2; the assembled normal equations of
 * BlockSolver::buildSystem (diagonal blocks [n_free][36], unique upper off-diagonal blocks with their (row, col) block
 * indices, right-hand side [6 n_free]); one linear solve of (H + lambda I) x = b on the last linearisation. */
int lvs_pgo_compute_errors(lvs_pgo_t* h, double* err6, double* chi2, double* robust_total);
int lvs_pgo_system_size(lvs_pgo_t* h, int* n_free, int* n_offdiag);
int lvs_pgo_linearize(lvs_pgo_t* h, double* Hd, int32_t* off_ij, double* Ho, double* b);
int lvs_pgo_solve(lvs_pgo_t* h, double lambda, double tolerance, int max_iterations, double* x, int* iterations);
/* Direct solver of the LVS_PGO_*_CHOL kinds: supernodal multifrontal Cholesky on the 6x6 block pattern (replaces
 * LinearSolverCholmod / LinearSolverCSparse, g2o solvers/cholmod/linear_solver_cholmod.h:115-154, solvers/csparse/
 * linear_solver_csparse.h:126-307; block ordering as solver_cholmod.cpp:45-86 "var_cholmod").  lvs_pgo_solve uses it when
 * tolerance <= 0.  chol_analyze is the host-side symbolic step alone (minimum-degree ordering, elimination tree, supernodes) on
 * a block pattern given as n_off (row < col) pairs; stats = {nnz(L) in 6x6 blocks, supernodes, tree levels, largest frontal
 * dimension, arena bytes, factorisation multiply-adds}.  It needs no device. */
int lvs_pgo_chol_analyze(int n_blocks, int n_off, const int32_t* off_ij, long long stats[6], int32_t* perm_out);
int lvs_pgo_chol_info(lvs_pgo_t* h, long long stats[6]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* LVSLAM_B200_H_ */
