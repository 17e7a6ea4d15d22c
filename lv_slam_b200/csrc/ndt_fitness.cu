// Nearest-neighbour fitness score on the device — replaces pcl::Registration::getFitnessScore(max_range) as the loop detector
// calls it right after align() (include/global_graph/loop_detector.hpp:176, 255) and
// lv_slam::InformationMatrixCalculator::calc_fitness_score (src/global_graph/information_matrix_calculator.cpp:53-87):
//   transform the source by the final transformation (pcl::transformPointCloud, float), find for every point its nearest TARGET
//   POINT (pcl::search::KdTree / FLANN L2_Simple: ((dx*dx + dy*dy) + dz*dz) in float), keep the squared distances that are
//   <= max_range (the reference compares the SQUARED distance with max_range), return their mean in double, or DBL_MAX when
//   there is no correspondence.
// The kd-tree is replaced by the voxel structure the target already has: the dense index grid, and the target points grouped by
// cell (TargetGrid::d_sorted_idx / d_cell_start; copied once per build into cell order — and inside a cell into 64 slabs along x —
// d_sorted_pts, so that a cell is one contiguous run that can be entered at the query's own x and left as soon as x alone is too
// far).  Pass 1, a thread per query: its own cell and the 26 around it, skipping every cell whose box is farther away than the
// best distance so far, within a fixed work budget; the best distance is exact as soon as it is below the distance to the outside
// of the scanned block.  Pass 2, a warp per query: the queries that are still undecided (compacted into a list) walk the shells up
// to radius 6, each cell read cooperatively.  Pass 3: what is left (far from every target point) is finished by brute force over
// shared-memory tiles.  The float distances are bit-identical to the CPU
// path; the double sum is a fixed-shape reduction (run-to-run deterministic).
// Every pruning bound (cell boxes, slab extents) carries a slack of 1e-3 leaf for the rounding of the binning and of the float
// differences: that covers coordinates up to a few kilometres (ulp(2048 m) = 2.4e-4 m against 5e-4 m at a 0.5 m leaf) — scans in
// the sensor or keyframe frame, which is what the callers pass; clouds in a global (UTM-sized) frame would need the slack scaled
// with the coordinate magnitude.
#include <cfloat>
#include "ndt_eval_common.cuh"

namespace lvs {

constexpr int kFitThreads = 256;
constexpr int kFitMaxRing = 6;                // shells the warp-per-query pass walks before brute force takes over
constexpr int kFitSlabs = kFitSlabsPerCell;

__device__ __forceinline__ float dist2(float qx, float qy, float qz, const float4& p) {
  const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
  return (dx * dx + dy * dy) + dz * dz;
}

// Slab of a coordinate inside its cell: kFitSlabs slices along x.  Truncation and clamping keep the map monotone in x, which is
// all the search relies on (the slab a query starts from is only a heuristic).
__device__ __forceinline__ int fit_slab(float x, float xlow, float sscale) { return min(max((int)((x - xlow) * sscale), 0), kFitSlabs - 1); }

// cell (record index) of position k of the cell-ordered point list: the last cell_start entry <= k
__device__ __forceinline__ int cell_of_position(const int* __restrict__ cell_start, int n_cells, int k) {
  int lo = 0, hi = n_cells;                 // invariant: cell_start[lo] <= k < cell_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cell_start + mid) <= k) lo = mid; else hi = mid;
  }
  return lo;
}

// Target points copied into cell order (one contiguous run per cell) and, inside a cell, into slab order along x — a counting
// sort in three passes: slab histogram per cell, exclusive scan of each cell's 64 counters, scatter.  The order inside a slab is
// whatever the atomics give; the minimum distance does not depend on it.  After the scatter slab[cell][s] holds the END of slab s
// relative to the cell's first position.
template <bool SCATTER>
__global__ void __launch_bounds__(kFitThreads) fitness_slab_kernel(const float4* __restrict__ tgt, const int* __restrict__ sorted_idx,
                                                                   const int* __restrict__ cell_start, const GridParams* __restrict__ gp,
                                                                   int* __restrict__ slab, float4* __restrict__ out) {
  const int k = blockIdx.x * kFitThreads + threadIdx.x;
  if (gp->status != 0 || gp->n_cells <= 0) return;
  const int n_cells = gp->n_cells;
  if (k >= cell_start[n_cells]) return;
  const float4 p = tgt[sorted_idx[k]];
  const int rec = cell_of_position(cell_start, n_cells, k);
  // low x face of the cell from the binning arithmetic itself (key_kernel): the same value for every point of the cell
  const float xlow = floorf(p.x * gp->inv_leaf) * gp->leaf;
  int* cnt = slab + (size_t)rec * kFitSlabs + fit_slab(p.x, xlow, (float)kFitSlabs * gp->inv_leaf);
  if (!SCATTER) atomicAdd(cnt, 1);
  else out[__ldg(cell_start + rec) + atomicAdd(cnt, 1)] = p;
}

// exclusive scan of the 64 counters of every cell, one warp per cell (two counters per lane)
__global__ void __launch_bounds__(kFitThreads) fitness_slab_scan_kernel(const GridParams* __restrict__ gp, int* __restrict__ slab) {
  const int lane = threadIdx.x & 31;
  const int cell = (blockIdx.x * kFitThreads + threadIdx.x) >> 5;
  if (gp->status != 0 || cell >= gp->n_cells) return;
  int2* row = reinterpret_cast<int2*>(slab + (size_t)cell * kFitSlabs) + lane;
  const int2 c = *row;
  int incl = c.x + c.y;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const int excl = incl - (c.x + c.y);
  *row = make_int2(excl, excl + c.x);
}

// One cell of the search, thread per query: its points in slab order [k0, k1), the slab table of the cell, the cell's low x face.
// The scan starts at the query's own slab and walks outwards in both directions; a direction ends when the points still ahead —
// none of them can lie more than one slab width behind the current point along x — are farther away along x alone than the best
// distance so far.  `budget` counts the steps (of kU points) the thread may still spend; the scan stops when it is used up and
// the query is handed to the warp-per-query pass.
__device__ __forceinline__ float scan_cell(const float4* __restrict__ pts, const int* __restrict__ slab_end, int k0, int k1, float xlow, float sscale,
                                           float wslack, float qx, float qy, float qz, float best, int& budget) {
  const int sq = fit_slab(qx, xlow, sscale);
  int kr = k0 + (sq ? __ldg(slab_end + sq - 1) : 0);
  int kl = kr - 1;
  constexpr int kU = 4;                      // loads of a step are issued together, from clamped positions
  while (kr < k1) {
    if (--budget < 0) return best;
    float4 p[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) p[u] = __ldg(pts + min(kr + u, k1 - 1));
    bool stop = false;
#pragma unroll
    for (int u = 0; u < kU; u++) {
      if (stop || kr + u >= k1) continue;
      const float b = (p[u].x - wslack) - qx;
      if (b > 0.0f && b * b >= best) stop = true;
      else best = fminf(best, dist2(qx, qy, qz, p[u]));
    }
    if (stop) break;
    kr += kU;
  }
  while (kl >= k0) {
    if (--budget < 0) return best;
    float4 p[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) p[u] = __ldg(pts + max(kl - u, k0));
    bool stop = false;
#pragma unroll
    for (int u = 0; u < kU; u++) {
      if (stop || kl - u < k0) continue;
      const float b = qx - (p[u].x + wslack);
      if (b > 0.0f && b * b >= best) stop = true;
      else best = fminf(best, dist2(qx, qy, qz, p[u]));
    }
    if (stop) break;
    kl -= kU;
  }
  return best;
}

// The same cell scanned by a whole warp for ONE query (all lanes hold the same query and the same `best`): the slabs that can
// hold a point closer along x than the best distance so far form one contiguous run of positions, which the lanes stride through
// with coalesced loads; the lanes' minima are then combined.  A point left out has |dx| > sqrt(best): the slab map is monotone.
__device__ __forceinline__ float scan_cell_warp(const float4* __restrict__ pts, const int* __restrict__ slab_end, int k0, float xlow, float sscale,
                                                float wslack, float qx, float qy, float qz, float best, int lane) {
  int s_lo = 0, s_hi = kFitSlabs - 1;
  if (best < INFINITY) {
    const float rb = sqrtf(best) * 1.0001f + wslack;
    s_lo = fit_slab(qx - rb, xlow, sscale);
    s_hi = fit_slab(qx + rb, xlow, sscale);
  }
  const int lo = k0 + (s_lo ? __ldg(slab_end + s_lo - 1) : 0), hi = k0 + __ldg(slab_end + s_hi);
  float b = best;
#pragma unroll 4
  for (int k = lo + lane; k < hi; k += 32) b = fminf(b, dist2(qx, qy, qz, __ldg(pts + k)));
  for (int o = 16; o; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
  return b;
}

// Shells R0..R1 of the ring search around the query's cell; returns whether `best` is proven to be the nearest distance.
// One thread per query, `budget` steps of work at most (then undecided).
template <int R0, int R1>
__device__ __forceinline__ bool search_rings(const FitnessArgs& a, const GridParams* gp, float qx, float qy, float qz, float& best, int budget) {
  const float inv = gp->inv_leaf, leaf = gp->leaf;
  // the query's cell with the arithmetic the target points were binned with (key_kernel): one monotone map for both
  const int cx = (int)floorf(qx * inv) - gp->min_b[0], cy = (int)floorf(qy * inv) - gp->min_b[1], cz = (int)floorf(qz * inv) - gp->min_b[2];
  const int d0 = gp->div_b[0], d1 = gp->div_b[1], d2 = gp->div_b[2];
  const int m1 = gp->mul[1], m2 = gp->mul[2];
  // position of the query inside its own cell, in metres from the cell's low corner: the distance to a neighbouring cell's box
  // is a lower bound for every point in it, so most of the 26 neighbours are skipped once the own cell has produced a hit
  const float fx = qx - (float)(cx + gp->min_b[0]) * leaf, fy = qy - (float)(cy + gp->min_b[1]) * leaf, fz = qz - (float)(cz + gp->min_b[2]) * leaf;
  const float slack = 1e-3f * leaf;        // rounding of the binning at cell faces
  const float sscale = (float)kFitSlabs * inv;
  const float wslack = leaf * (1.01f / (float)kFitSlabs) + slack;     // x extent of a slab, rounding included
  bool decided = false;
  for (int r = R0; r <= R1 && !decided; r++) {
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
          const int* se = a.slab_end + (size_t)rec * kFitSlabs;
          const float xlow = (float)(x + gp->min_b[0]) * leaf;
          best = scan_cell(a.tgt_sorted, se, k0, k1, xlow, sscale, wslack, qx, qy, qz, best, budget);
          if (budget < 0) return false;
        }
      }
    }
    // every target point outside the (2r+1)^3 block is at least r cells away (minus the rounding of the binning)
    const float safe = (float)r * leaf * 0.999f;
    decided = r >= 1 && best <= safe * safe;
  }
  return decided;
}

// The same search by a whole warp for ONE query.  The cells of a shell are spread over the lanes (box test and index-grid lookup
// of 32 cells at a time: the lookups are dependent loads and would otherwise be paid one after the other); the occupied ones are
// then scanned cooperatively, one cell at a time, each against the best distance the previous ones left.
template <int R0, int R1>
__device__ __forceinline__ bool search_rings_warp(const FitnessArgs& a, const GridParams* gp, float qx, float qy, float qz, float& best, int lane) {
  const float inv = gp->inv_leaf, leaf = gp->leaf;
  const int cx = (int)floorf(qx * inv) - gp->min_b[0], cy = (int)floorf(qy * inv) - gp->min_b[1], cz = (int)floorf(qz * inv) - gp->min_b[2];
  const int d0 = gp->div_b[0], d1 = gp->div_b[1], d2 = gp->div_b[2];
  const int m1 = gp->mul[1], m2 = gp->mul[2];
  const float fx = qx - (float)(cx + gp->min_b[0]) * leaf, fy = qy - (float)(cy + gp->min_b[1]) * leaf, fz = qz - (float)(cz + gp->min_b[2]) * leaf;
  const float slack = 1e-3f * leaf;
  const float sscale = (float)kFitSlabs * inv;
  const float wslack = leaf * (1.01f / (float)kFitSlabs) + slack;
  bool decided = false;
  for (int r = R0; r <= R1 && !decided; r++) {
    // the cells of shell r, enumerated without the inner cube: the two z faces (n x n each), then for every z in between the two
    // y edges (n each) and the two x ends of the rows in between
    const int n = 2 * r + 1, per_mid = 4 * n - 4;
    const int n_shell = r == 0 ? 1 : 2 * n * n + (n - 2) * per_mid;
    for (int c0 = 0; c0 < n_shell; c0 += 32) {
      const int c = c0 + lane;
      int v = -1, xcell = 0;
      float h2 = 0.0f;
      if (c < n_shell) {
        int ox = 0, oy = 0, oz = 0;
        if (r > 0) {
          if (c < 2 * n * n) {
            const int rem = c % (n * n);
            oz = c < n * n ? -r : r; oy = rem / n - r; ox = rem % n - r;
          } else {
            const int u = c - 2 * n * n, w = u % per_mid;
            oz = u / per_mid - r + 1;
            if (w < 2 * n) { oy = w < n ? -r : r; ox = w % n - r; }
            else { oy = (w - 2 * n) / 2 - r + 1; ox = ((w - 2 * n) & 1) ? r : -r; }
          }
        }
        const int x = cx + ox, y = cy + oy, z = cz + oz;
        if ((unsigned)x < (unsigned)d0 && (unsigned)y < (unsigned)d1 && (unsigned)z < (unsigned)d2) {
          const float gx = ox > 0 ? (float)ox * leaf - fx : (ox < 0 ? fx - (float)(ox + 1) * leaf : 0.0f);
          const float gy = oy > 0 ? (float)oy * leaf - fy : (oy < 0 ? fy - (float)(oy + 1) * leaf : 0.0f);
          const float gz = oz > 0 ? (float)oz * leaf - fz : (oz < 0 ? fz - (float)(oz + 1) * leaf : 0.0f);
          const float hx = fmaxf(gx - slack, 0.0f), hy = fmaxf(gy - slack, 0.0f), hz = fmaxf(gz - slack, 0.0f);
          h2 = hx * hx + hy * hy + hz * hz;
          if (h2 < best) { v = __ldg(a.grid + (x + y * m1 + z * m2)); xcell = x; }
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, v != -1);
      while (m) {
        const int from = __ffs(m) - 1;
        m &= m - 1;
        const int vv = __shfl_sync(0xffffffffu, v, from), xx = __shfl_sync(0xffffffffu, xcell, from);
        const float hh = __shfl_sync(0xffffffffu, h2, from);
        if (hh >= best) continue;                                            // an earlier cell of the batch got closer than this box
        const int rec = grid_decode_any(vv);
        best = scan_cell_warp(a.tgt_sorted, a.slab_end + (size_t)rec * kFitSlabs, __ldg(a.cell_start + rec), (float)(xx + gp->min_b[0]) * leaf, sscale, wslack,
                              qx, qy, qz, best, lane);
      }
    }
    const float safe = (float)r * leaf * 0.999f;
    decided = r >= 1 && best <= safe * safe;
  }
  return decided;
}

// Pass 1, every query, one thread each: its own cell and the 26 around it, within a budget of kFitBudget steps.  A query that is
// still undecided (far from every target point, or next to dense cells it could not prune) goes on list 1 with the best distance
// found so far (a real distance, or +inf).
constexpr int kFitBudget = 24;               // x 4 points
__global__ void __launch_bounds__(kFitThreads) fitness_search_kernel(FitnessArgs a) {
  const int i = blockIdx.x * kFitThreads + threadIdx.x;
  if (i >= a.n_src) return;
  const GridParams* gp = a.gp;
  const float4 s = a.src[i];
  float qx, qy, qz;
  transform_point(a.T16, s.x, s.y, s.z, qx, qy, qz);
  if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) { a.best[i] = -2.0f; return; }   // no neighbour: dropped from the mean
  float best = INFINITY;
  const bool grid_ok = gp->status == 0 && gp->n_cells > 0;
  const bool decided = grid_ok && search_rings<0, 1>(a, gp, qx, qy, qz, best, kFitBudget);
  a.best[i] = best;
  if (!decided) {
    int* list = grid_ok ? a.list : a.list2;                   // without a grid there is nothing to walk
    list[1 + atomicAdd(list, 1)] = i;
  }
}

// Pass 2, the queries of list 1, one WARP each: all shells again, starting from the best distance of pass 1 (cells that were
// already scanned prune to nothing).  The heavy and the far queries no longer hold up a warp of easy ones, and a long cell is
// read with coalesced loads.  Still undecided after the last shell -> list 2.
__global__ void __launch_bounds__(kFitThreads) fitness_far_kernel(FitnessArgs a) {
  const int n_list = a.list[0];
  const GridParams* gp = a.gp;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kFitThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kFitThreads) >> 5;
  for (int e = warp; e < n_list; e += n_warps) {
    const int i = a.list[1 + e];
    const float4 s = a.src[i];
    float qx, qy, qz;
    transform_point(a.T16, s.x, s.y, s.z, qx, qy, qz);
    float best = a.best[i];
    const bool decided = search_rings_warp<0, kFitMaxRing>(a, gp, qx, qy, qz, best, lane);
    if (lane == 0) {
      a.best[i] = best;
      if (!decided) a.list2[1 + atomicAdd(a.list2, 1)] = i;
    }
  }
}

// Pass 3, the queries of list 2: exact scan of every target point.  A CTA stages a tile of target points in shared memory and its
// warps run a group of queries against it; the tiles' minima meet in an atomic min on the bit pattern (non-negative floats order
// like integers; a minimum does not depend on the order).  Non-finite target points produce NaN or +inf distances, which fminf
// drops.
constexpr int kBruteTile = 2048, kBruteGroup = 64;
__global__ void __launch_bounds__(kFitThreads) fitness_brute_kernel(FitnessArgs a) {
  __shared__ float4 s_pts[kBruteTile];
  const int n_list = a.list2[0];
  if (n_list == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kW = kFitThreads / 32;
  const int n_tiles = (a.n_tgt + kBruteTile - 1) / kBruteTile, n_groups = (n_list + kBruteGroup - 1) / kBruteGroup;
  for (long long item = blockIdx.x; item < (long long)n_tiles * n_groups; item += gridDim.x) {
    const int tile = (int)(item % n_tiles), group = (int)(item / n_tiles);
    const int c0 = tile * kBruteTile, cn = min(kBruteTile, a.n_tgt - c0);
    __syncthreads();
    for (int t = threadIdx.x; t < kBruteTile; t += kFitThreads) s_pts[t] = t < cn ? __ldg(a.tgt + c0 + t) : make_float4(NAN, NAN, NAN, 0.0f);
    __syncthreads();
    const int q_end = min(n_list, (group + 1) * kBruteGroup);
    for (int e = group * kBruteGroup + warp; e < q_end; e += kW) {
      const int i = a.list2[1 + e];
      const float4 s = a.src[i];
      float qx, qy, qz;
      transform_point(a.T16, s.x, s.y, s.z, qx, qy, qz);
      float best = INFINITY;
#pragma unroll 8
      for (int k = lane; k < kBruteTile; k += 32) best = fminf(best, dist2(qx, qy, qz, s_pts[k]));
      for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (lane == 0 && best < INFINITY) atomicMin(reinterpret_cast<int*>(a.best + i), __float_as_int(best));
    }
  }
}

// sum and count of the accepted squared distances, fixed shape: CTA partials, the last CTA adds them in order
__global__ void __launch_bounds__(kFitThreads) fitness_reduce_kernel(FitnessArgs a) {
  __shared__ double s_sum[kFitThreads / 32], s_cnt[kFitThreads / 32];
  __shared__ bool s_last;
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * kFitThreads + threadIdx.x; i < a.n_src; i += gridDim.x * kFitThreads) {
    const float d = a.best[i];
    if (d >= 0.0f && d < INFINITY && (double)d <= a.max_range) { sum += (double)d; cnt += 1.0; }     // +inf: no target point at all
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

int launch_fitness_gather(cudaStream_t st, const float4* tgt, int n_tgt, int n_cells, const int* sorted_idx, const int* cell_start, const GridParams* gp,
                          int* slab, float4* out, int* launches) {
  if (n_tgt <= 0 || n_cells <= 0) return LVS_OK;
  const int nb = (n_tgt + kFitThreads - 1) / kFitThreads;
  CUDA_TRY(cudaMemsetAsync(slab, 0, (size_t)n_cells * kFitSlabs * sizeof(int), st));
  fitness_slab_kernel<false><<<nb, kFitThreads, 0, st>>>(tgt, sorted_idx, cell_start, gp, slab, out);
  fitness_slab_scan_kernel<<<(n_cells * 32 + kFitThreads - 1) / kFitThreads, kFitThreads, 0, st>>>(gp, slab);
  fitness_slab_kernel<true><<<nb, kFitThreads, 0, st>>>(tgt, sorted_idx, cell_start, gp, slab, out);
  CUDA_TRY(cudaGetLastError());
  if (launches) *launches += 3;
  return LVS_OK;
}

int launch_fitness(cudaStream_t st, const FitnessArgs& a, int* launches) {
  CUDA_TRY(cudaMemsetAsync(a.list, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(a.list2, 0, sizeof(int), st));
  if (a.n_src > 0) {
    fitness_search_kernel<<<(a.n_src + kFitThreads - 1) / kFitThreads, kFitThreads, 0, st>>>(a);
    fitness_far_kernel<<<148 * 4, kFitThreads, 0, st>>>(a);
    fitness_brute_kernel<<<148 * 4, kFitThreads, 0, st>>>(a);
  }
  const int nb = std::max(1, std::min(64, (a.n_src + kFitThreads * 8 - 1) / (kFitThreads * 8)));   // partials: 2 doubles per CTA
  fitness_reduce_kernel<<<nb, kFitThreads, 0, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  if (launches) *launches += a.n_src > 0 ? 4 : 1;
  return LVS_OK;
}

}  // namespace lvs
