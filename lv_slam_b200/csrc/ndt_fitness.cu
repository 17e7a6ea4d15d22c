// Nearest-neighbour fitness score on the device — replaces pcl::Registration::getFitnessScore(max_range) as the loop detector
// calls it right after align() (include/global_graph/loop_detector.hpp:176, 255) and
// lv_slam::InformationMatrixCalculator::calc_fitness_score (src/global_graph/information_matrix_calculator.cpp:53-87):
//   transform the source by the final transformation (pcl::transformPointCloud, float), find for every point its nearest TARGET
//   POINT (pcl::search::KdTree / FLANN L2_Simple: ((dx*dx + dy*dy) + dz*dz) in float), keep the squared distances that are
//   <= max_range (the reference compares the SQUARED distance with max_range), return their mean in double, or DBL_MAX when
//   there is no correspondence.
// The kd-tree is replaced by the voxel structure the target already has: the dense index grid, and the target points grouped by
// cell (TargetGrid::d_sorted_idx / d_cell_start; copied once per build into cell order, d_sorted_pts, so that a cell is one
// contiguous run).  A query scans its own cell, then the 26 cells around it, then the shells of radius 2 and 3, skipping every
// cell whose box is farther away than the best distance so far; the best distance found is exact as soon as it is below the
// distance to the outside of the scanned block.  The few
// queries that stay undecided (far from every target point) are finished by a brute-force pass, one warp per query.  The float
// distances are bit-identical to the CPU path; the double sum is a fixed-shape reduction (run-to-run deterministic).
#include <cfloat>
#include "ndt_eval_common.cuh"

namespace lvs {

constexpr int kFitThreads = 256;
constexpr int kFitMaxRing = 3;

__device__ __forceinline__ float dist2(float qx, float qy, float qz, const float4& p) {
  const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
  return (dx * dx + dy * dy) + dz * dz;
}

// target points copied into cell order (one contiguous run per cell): the search then streams a cell instead of gathering it
__global__ void __launch_bounds__(kFitThreads) fitness_gather_kernel(const float4* __restrict__ tgt, const int* __restrict__ sorted_idx,
                                                                     const int* __restrict__ cell_start, const GridParams* __restrict__ gp,
                                                                     float4* __restrict__ out) {
  const int k = blockIdx.x * kFitThreads + threadIdx.x;
  if (gp->status != 0 || gp->n_cells <= 0) return;
  if (k < cell_start[gp->n_cells]) out[k] = tgt[sorted_idx[k]];
}

__global__ void __launch_bounds__(kFitThreads) fitness_search_kernel(FitnessArgs a) {
  const int i = blockIdx.x * kFitThreads + threadIdx.x;
  if (i >= a.n_src) return;
  const GridParams* gp = a.gp;
  const float4 s = a.src[i];
  float qx, qy, qz;
  transform_point(a.T16, s.x, s.y, s.z, qx, qy, qz);
  if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) { a.best[i] = -2.0f; return; }   // no neighbour: dropped from the mean
  float best = INFINITY;
  bool decided = false;
  if (gp->status == 0 && gp->n_cells > 0) {
    const float inv = gp->inv_leaf, leaf = gp->leaf;
    // the query's cell with the arithmetic the target points were binned with (key_kernel): one monotone map for both
    const int cx = (int)floorf(qx * inv) - gp->min_b[0], cy = (int)floorf(qy * inv) - gp->min_b[1], cz = (int)floorf(qz * inv) - gp->min_b[2];
    const int d0 = gp->div_b[0], d1 = gp->div_b[1], d2 = gp->div_b[2];
    const int m1 = gp->mul[1], m2 = gp->mul[2];
    // position of the query inside its own cell, in metres from the cell's low corner: the distance to a neighbouring cell's box
    // is a lower bound for every point in it, so most of the 26 neighbours are skipped once the own cell has produced a hit
    const float fx = qx - (float)(cx + gp->min_b[0]) * leaf, fy = qy - (float)(cy + gp->min_b[1]) * leaf, fz = qz - (float)(cz + gp->min_b[2]) * leaf;
    const float slack = 1e-3f * leaf;        // rounding of the binning at cell faces
    for (int r = 0; r <= kFitMaxRing && !decided; r++) {
      for (int oz = -r; oz <= r; oz++) {
        const int z = cz + oz;
        if ((unsigned)z >= (unsigned)d2) continue;
        const float gz = oz > 0 ? (float)oz * leaf - fz : (oz < 0 ? fz - (float)(oz + 1) * leaf : 0.0f);
        for (int oy = -r; oy <= r; oy++) {
          const int y = cy + oy;
          if ((unsigned)y >= (unsigned)d1) continue;
          const float gy = oy > 0 ? (float)oy * leaf - fy : (oy < 0 ? fy - (float)(oy + 1) * leaf : 0.0f);
          const bool face = (oz == -r || oz == r || oy == -r || oy == r);
          // inside the shell only the two end cells of the x run are new; on a face the whole run is
          const int step = (face || r <= 1) ? 1 : 2 * r;
          for (int ox = -r; ox <= r; ox += step) {
            const int x = cx + ox;
            if ((unsigned)x >= (unsigned)d0) continue;
            if (r == 1 && ox == 0 && oy == 0 && oz == 0) continue;      // the own cell was shell 0
            const float gx = ox > 0 ? (float)ox * leaf - fx : (ox < 0 ? fx - (float)(ox + 1) * leaf : 0.0f);
            const float hx = fmaxf(gx - slack, 0.0f), hy = fmaxf(gy - slack, 0.0f), hz = fmaxf(gz - slack, 0.0f);
            if (hx * hx + hy * hy + hz * hz >= best) continue;          // nothing in that cell can beat the current best
            const int v = __ldg(a.grid + (x + y * m1 + z * m2));
            if (v == -1) continue;
            const int rec = grid_decode_any(v);
            const int k0 = __ldg(a.cell_start + rec), k1 = __ldg(a.cell_start + rec + 1);
            for (int k = k0; k < k1; k++) best = fminf(best, dist2(qx, qy, qz, __ldg(a.tgt_sorted + k)));
          }
        }
      }
      // every target point outside the (2r+1)^3 block is at least r cells away (minus the rounding of the binning)
      const float safe = (float)r * leaf * 0.999f;
      decided = r >= 1 && best <= safe * safe;
    }
  }
  if (decided) a.best[i] = best;
  else {
    a.best[i] = -1.0f;
    a.list[1 + atomicAdd(a.list, 1)] = i;
  }
}

// Undecided queries: exact scan of every target point, one warp per query.
__global__ void __launch_bounds__(kFitThreads) fitness_brute_kernel(FitnessArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kFitThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kFitThreads) >> 5;
  const int n_list = a.list[0];
  for (int w = warp; w < n_list; w += n_warps) {
    const int i = a.list[1 + w];
    const float4 s = a.src[i];
    float qx, qy, qz;
    transform_point(a.T16, s.x, s.y, s.z, qx, qy, qz);
    float best = INFINITY;
    for (int k = lane; k < a.n_tgt; k += 32) {
      const float4 p = __ldg(a.tgt + k);
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) best = fminf(best, dist2(qx, qy, qz, p));
    }
    for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) a.best[i] = isfinite(best) ? best : -2.0f;
  }
}

// sum and count of the accepted squared distances, fixed shape: CTA partials, the last CTA adds them in order
__global__ void __launch_bounds__(kFitThreads) fitness_reduce_kernel(FitnessArgs a) {
  __shared__ double s_sum[kFitThreads / 32], s_cnt[kFitThreads / 32];
  __shared__ bool s_last;
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * kFitThreads + threadIdx.x; i < a.n_src; i += gridDim.x * kFitThreads) {
    const float d = a.best[i];
    if (d >= 0.0f && (double)d <= a.max_range) { sum += (double)d; cnt += 1.0; }
  }
  for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = 0.0; cnt = 0.0;
    for (int w = 0; w < kFitThreads / 32; w++) { sum += s_sum[w]; cnt += s_cnt[w]; }
    a.partials[2 * blockIdx.x] = sum; a.partials[2 * blockIdx.x + 1] = cnt;
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  sum = 0.0; cnt = 0.0;
  for (unsigned b = 0; b < gridDim.x; b++) { sum += __ldcg(a.partials + 2 * b); cnt += __ldcg(a.partials + 2 * b + 1); }
  a.out[0] = cnt > 0.0 ? sum / cnt : DBL_MAX;
  a.out[1] = cnt;
  *a.ticket = 0;
}

int launch_fitness_gather(cudaStream_t st, const float4* tgt, int n_tgt, const int* sorted_idx, const int* cell_start, const GridParams* gp, float4* out) {
  if (n_tgt <= 0) return LVS_OK;
  fitness_gather_kernel<<<(n_tgt + kFitThreads - 1) / kFitThreads, kFitThreads, 0, st>>>(tgt, sorted_idx, cell_start, gp, out);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

int launch_fitness(cudaStream_t st, const FitnessArgs& a, int* launches) {
  CUDA_TRY(cudaMemsetAsync(a.list, 0, sizeof(int), st));
  if (a.n_src > 0) {
    fitness_search_kernel<<<(a.n_src + kFitThreads - 1) / kFitThreads, kFitThreads, 0, st>>>(a);
    fitness_brute_kernel<<<148 * 4, kFitThreads, 0, st>>>(a);
  }
  const int nb = std::max(1, std::min(64, (a.n_src + kFitThreads * 8 - 1) / (kFitThreads * 8)));   // partials: 2 doubles per CTA
  fitness_reduce_kernel<<<nb, kFitThreads, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  if (launches) *launches += a.n_src > 0 ? 3 : 1;
  return LVS_OK;
}

}  // namespace lvs
