// Internal declarations shared by the NDT translation units of liblvslam_b200.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/lvslam_b200.h"
#include "lvs_math.cuh"
#include "ndt_types.cuh"

namespace lvs {

int fail(int status, const char* fmt, ...);          // records the thread-local error string, returns status
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CUDA_TRY(expr)                                                      \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) return ::lvs::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

// GridParams::status values besides 0 (ok) and LVS_ERR_GRID_OVERFLOW
constexpr int kStatusEmpty = 1, kStatusNeedsGrow = 2;
constexpr int kVoxBatch = 8;             // clouds voxelised by one set of launches

struct BuildScratch {                    // reusable workspace of the voxelisation pipeline
  int capacity = 0;
  unsigned int* d_keys[2] = {nullptr, nullptr};
  int* d_idx[2] = {nullptr, nullptr};
  int *d_flags = nullptr, *d_pos = nullptr, *d_seg_start = nullptr, *d_hist = nullptr, *d_hist_scan = nullptr;
  float* d_bbox_partial = nullptr;
  int* d_tile_tot = nullptr;             // per-tile totals of the two-level scans
  double* d_moments = nullptr;           // [cell][9] S1, upper S2
  float* d_csum = nullptr;               // [cell][3] float centroid sums
  unsigned int* d_ticket = nullptr;
  int *d_nseg = nullptr, *d_nvalidpts = nullptr;
  GridParams* h_gp = nullptr;            // pinned
  cudaError_t reserve(int n);
  void release();
};

struct TargetGrid {                      // one voxelised target resident in HBM
  const float4* pts = nullptr;
  int n_points = 0;
  GridParams gp{};                       // host copy
  GridParams* d_gp = nullptr;
  int* d_grid = nullptr;                 // dense int32 index grid, -1 = empty
  size_t grid_capacity = 0;
  VoxelRec* d_recs = nullptr;
  FastRec* d_frecs = nullptr;           // tolerance-mode twin of d_recs
  float4* d_centroids = nullptr;         // float centroid (kd-tree cloud of the reference), w = 1 if in that cloud
  int* d_cell_keys = nullptr;
  int* d_cell_npts = nullptr;
  double* d_cell_evals = nullptr;
  double* d_icov64 = nullptr;            // [n_cells][9] double inverse covariance
  int* d_sorted_idx = nullptr;           // target point indices grouped by cell (stable: input order inside a cell), cell_capacity
  int* d_cell_start = nullptr;           // [n_cells + 1] first position of every cell in d_sorted_idx
  float4* d_sorted_pts = nullptr;        // target points in cell order, filled lazily by the first fitness-score query of a build
  int* d_slab = nullptr;                 // [n_cells][kFitSlabsPerCell] end of every x slab inside its cell (same lazy fill)
  size_t slab_cells = 0;                 // cells d_slab has room for
  bool sorted_pts_valid = false;
  size_t cell_capacity = 0;
  int n_cells = 0;
  int launches_last_build = 0;
  int prev_points = 0;                   // points of the build whose cells are currently marked in d_grid
  bool pending = false;                  // a build is queued and its geometry has not been read back yet
  bool fetch_queued = false;             // a read-back of d_gp into the batch's pinned array is in flight
  lvs_ndt_params built_with{};
  int build(cudaStream_t st, const float4* d_pts, int n, const lvs_ndt_params& prm, BuildScratch& ws);   // asynchronous
  int finish(cudaStream_t st, BuildScratch& ws);                                                         // lazy completion
  int accept(const GridParams& g, cudaStream_t st, BuildScratch& ws);
  int enqueue(cudaStream_t st, const lvs_ndt_params& prm, BuildScratch& ws);
  int prepare(cudaStream_t st, const float4* d_pts, int n, const lvs_ndt_params& prm, BuildScratch& ws);   // host-side part of build()
  static int enqueue_many(cudaStream_t st, int count, TargetGrid* const* grids, BuildScratch* const* wss, const lvs_ndt_params& prm);
  int grow_grid(cudaStream_t st, long long cells);
  void free_cells();
  void release();
};

// voxel-index machinery shared by the NDT target build and the prefilter's VoxelGrid (ndt_voxel.cu)
void vox_bbox(cudaStream_t st, const float4* pts, int n, BuildScratch& ws, GridParams* d_gp, float leaf, long long grid_capacity);
int vox_radix_passes(long long cells);
void exclusive_scan(cudaStream_t st, const int* in, int* out, int n, int* total, int* tile_tot);   // two-level exclusive scan
int vox_sort_segments(cudaStream_t st, const float4* pts, int n, const GridParams* d_gp, int passes, BuildScratch& ws, int* sorted_out, int* cell_start_out);

int pack_points(cudaStream_t st, const float* d_in, size_t stride_floats, int n, float4* d_out);
// Repack of up to kPackMany resident clouds in one launch (descriptors travel as kernel parameters).
constexpr int kPackMany = 96;
struct PackOne { const float* in; float4* out; int n; int stride_floats; };
struct PackMany { PackOne c[kPackMany]; int count; };
int pack_many(cudaStream_t st, const PackMany& pm, int max_n);

// Evaluation kernels (ndt_eval.cu).  One launch advances every active pair by one evaluation and, in the
// last CTA of each pair, by one step of the Newton / More-Thuente state machine.
// Point-sharded evaluation (sum over independent source points): every rank evaluates its contiguous chunk of the
// source against a replicated voxel grid, and the 43 sums are exchanged through peer memory INSIDE the evaluation kernel: the last
// CTA of a pair stores its rank's sums into the mailbox of every peer (NVLink stores on IPC-mapped memory), publishes a flag,
// waits for the flags of all ranks in its own mailbox and adds the contributions in rank order, so every rank obtains the same
// bits and advances an identical copy of the Newton state machine - no host round trip and no separate collective launch.
constexpr int kMailStride = 48;            // doubles per (rank, pair, parity) record: 43 sums, [43] = serial flag, padded to 384 B
struct ShardView {
  double* const* peers = nullptr;          // device array [world]: mailbox base of every rank as mapped in this process
  double* mine = nullptr;                  // this rank's mailbox: [2 parity][world][cap][kMailStride]
  int rank = 0, world = 1, cap = 0;
  long long serial = 0;                    // base serial of the align / tap in flight, identical on every rank (eval_finish adds the pair's evaluation count)
  long long timeout_cycles = 0;
  int* d_error = nullptr;                  // set to 1 when a peer did not answer in time
};

struct EvalLaunch {
  ShardView shard;
  const PairDesc* d_pairs;
  AlignState* d_states;
  TraceRec* d_trace;          // [n_pairs][kMaxTrace] or null
  double* d_partials;         // [n_pairs][blocks_per_pair][kPartialStride], kAcc used
  unsigned int* d_tickets;    // [n_pairs]
  int* d_done_count;          // incremented once per finished pair
  volatile int* h_done_flag = nullptr;   // host-mapped word: receives align_serial when the last pair of the batch finishes (null: not used)
  int align_serial = 0;
  int stage_bytes = 0;          // tolerance mode: > 0 = stage the pair's voxel records in shared memory by one bulk copy (bytes of dynamic shared memory)
  long long* d_dbg = nullptr;   // diagnostics (LVS_DEBUG_TIMING=1): clock64 stamps of the last CTA's tail, 8 per pair; null in normal operation
  int n_pairs;
  int blocks_per_pair;
  int advance;                // 1: run the align state machine; 0: tap mode, only store score/g/H
  float one = 1.0f;           // run-time 1.0f for the packed adds of the hot kernel (ndt_eval.cu: keeps ptxas from contracting them)
  AlignConsts consts;
};
int launch_eval(cudaStream_t st, const EvalLaunch& L);        // direct-search derivative passes (hot)
int launch_eval_cold(cudaStream_t st, const EvalLaunch& L);   // KDTREE-mode derivatives and the all-double Hessian pass
int launch_eval_fast(cudaStream_t st, const EvalLaunch& L);   // tolerance-mode direct-search derivative passes (ndt_eval_fast.cu)
int eval_fast_max_resident_ctas_per_sm();
int eval_max_resident_ctas_per_sm();
int eval_points_per_cta_iteration();
int launch_calc_score(cudaStream_t st, const PairDesc& pair, const float* d_T16, const AlignConsts& c, double* d_partials, int max_blocks,
                      unsigned int* d_ticket, double* d_out);
int launch_lookup_keys(cudaStream_t st, const PairDesc& pair, const float* d_T16, int* d_keys_out);
int launch_transform(cudaStream_t st, const float4* d_src, int n, const float* d_T16, float* d_out_xyz);

// Nearest-neighbour fitness score (ndt_fitness.cu): pcl::Registration::getFitnessScore / InformationMatrixCalculator::
// calc_fitness_score.  d_best [n_src] floats, d_list and d_list2 [n_src + 1] ints each and d_out [2] doubles are caller-provided scratch.
struct FitnessArgs {
  const float4* src; int n_src;
  const float4* tgt; int n_tgt;
  const int* grid; const GridParams* gp;
  const int* cell_start; const int* sorted_idx;
  const float4* tgt_sorted;          // target points in cell order, slab order inside a cell (TargetGrid::d_sorted_pts)
  const int* slab_end;               // TargetGrid::d_slab
  const float* T16;
  double max_range;
  float* best; int* list; int* list2; double* partials; unsigned int* ticket; double* out;   // out[0] = score, out[1] = correspondences
};
int launch_fitness(cudaStream_t st, const FitnessArgs& a, int* launches);
constexpr int kFitSlabsPerCell = 64;
int launch_fitness_gather(cudaStream_t st, const float4* tgt, int n_tgt, int n_cells, const int* sorted_idx, const int* cell_start, const GridParams* gp,
                          int* slab, float4* out, int* launches);

}  // namespace lvs
