// Target voxelisation on the device — replaces VoxelGridCovariance::applyFilter
// (include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:49-370; pca label/weight
// include/ndt_pca/voxel_grid_covariance_pca_impl.hpp:364-397).
//
// Pipeline (all stream-ordered and asynchronous: no host synchronisation once the dense index grid has its capacity —
// the geometry is checked on the device and read back lazily, see TargetGrid::finish):
//   pack_points          strided host layout -> float4
//   bbox_kernel          min/max reduce, last CTA derives min_b/max_b/div_b/mul  (:72-103) and checks the grid capacity
//   key_kernel           int32 voxel key per point, float math exactly as :218-223
//   radix sort           stable LSD sort of (key, point index): 8-bit digits, hist / two-level scan / scatter
//   head + scan          unique keys -> one segment per occupied cell, ascending key (= std::map order)
//   leaf_moments_kernel  one warp per cell: lanes 0..8 own the nine f64 moment accumulators, lanes 9..11 the f32 centroid
//                        sums, each summed SEQUENTIALLY IN INPUT ORDER (bit-identical to the reference's serial accumulation);
//                        addends staged two chunks ahead in shared memory so that only the dependent add is on the chain
//   leaf_finalize_kernel one thread per cell: mean, covariance, eigen-decomposition, inflation, inverse, PCA weight (:281-367)
// Compiled with -fmad=false: no mul/add contraction anywhere in this file.
#include "ndt_internal.cuh"

namespace lvs {

// ------------------------------------------------------------------ pack
__global__ void pack_points_kernel(const float* __restrict__ in, size_t stride_floats, int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = in + (size_t)i * stride_floats;
  out[i] = make_float4(p[0], p[1], p[2], 0.0f);
}

// Many clouds in one launch (batched setInputSource of resident scans): blockIdx.y = cloud.
__global__ void pack_many_kernel(PackMany pm) {
  const PackOne c = pm.c[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const float* p = c.in + (size_t)i * c.stride_floats;
    c.out[i] = make_float4(p[0], p[1], p[2], 0.0f);
  }
}

// ------------------------------------------------------------------ batched pipeline
// Every kernel below works on up to kVoxBatch clouds at once: blockIdx.y selects the job, blockIdx.x covers the largest job of the
// launch and the CTAs a smaller job does not need leave at once.  A keyframe batch therefore costs ~20 launches in total instead of
// ~20 per cloud (the setters were bound by the host's launch rate, not by the device).  The job descriptors travel by value.
struct VoxJob {
  const float4* pts; int n;
  float leaf; long long grid_capacity;
  GridParams* gp;
  int* grid; const int* old_keys; int prev_points;          // cells of the previous build to wipe
  float* bbox_partial; unsigned int* ticket;
  unsigned int* keys[2]; int* idx[2];                        // radix ping-pong; idx[] already arranged so that the last pass lands in `sorted`
  int *hist, *hist_scan, *tile_tot, *flags, *pos, *nseg, *nvalid;
  int* sorted; int* cell_start;
  double* moments; float* csum;
  VoxelRec* recs; FastRec* frecs; float4* centroids; int* cell_keys; int* cell_npts; double* cell_evals; double* icov64;
  int nblk;                                                  // radix tiles of this job
};
struct VoxBatch {
  VoxJob j[kVoxBatch];
  int count;
  int min_points; double eig_mult; int variant;
};
struct ScanJob { const int* in; int* out; int n; int* tile_tot; int* total; };
struct ScanBatch { ScanJob j[kVoxBatch]; int count; };

// ------------------------------------------------------------------ bbox
constexpr int kBboxBlocks = 148;

__global__ void __launch_bounds__(256) bbox_kernel(VoxBatch B) {
  const VoxJob& J = B.j[blockIdx.y];
  const float4* __restrict__ pts = J.pts;
  const int n = J.n;
  float* __restrict__ partial = J.bbox_partial;      // [grid][8]
  unsigned int* __restrict__ ticket = J.ticket;
  GridParams* __restrict__ gp = J.gp;
  const float leaf = J.leaf;
  const long long grid_capacity = J.grid_capacity;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      cnt++;
    }
  }
  __shared__ float s_mn[3][8], s_mx[3][8];
  __shared__ int s_cnt[8];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_reduce = [&]() {
    for (int o = 16; o; o >>= 1) {
      for (int a = 0; a < 3; a++) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
      }
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; } s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int w = 1; w < 8; w++) {
        for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], s_mn[a][w]); mx[a] = fmaxf(mx[a], s_mx[a][w]); }
        cnt += s_cnt[w];
      }
  };
  block_reduce();
  if (threadIdx.x == 0) {
    float* o = partial + blockIdx.x * 8;
    o[0] = mn[0]; o[1] = mn[1]; o[2] = mn[2]; o[3] = mx[0]; o[4] = mx[1]; o[5] = mx[2]; o[6] = __int_as_float(cnt);
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int a = 0; a < 3; a++) { mn[a] = INFINITY; mx[a] = -INFINITY; }
  cnt = 0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
    const float* o = partial + b * 8;
    for (int a = 0; a < 3; a++) { mn[a] = fminf(mn[a], __ldcg(o + a)); mx[a] = fmaxf(mx[a], __ldcg(o + 3 + a)); }
    cnt += __float_as_int(__ldcg(o + 6));
  }
  __syncthreads();
  block_reduce();
  if (threadIdx.x != 0) return;
  *ticket = 0;
  GridParams g;
  g.leaf = leaf;
  g.inv_leaf = __fdiv_rn(1.0f, leaf);               // pcl::VoxelGrid::setLeafSize
  g.n_points = cnt; g.n_cells = 0; g.n_valid = 0; g.status = 0; g.total_cells = 0;
  for (int a = 0; a < 3; a++) { g.min_p[a] = mn[a]; g.max_p[a] = mx[a]; g.min_b[a] = g.max_b[a] = g.div_b[a] = g.mul[a] = 0; }
  if (cnt == 0) { g.status = kStatusEmpty; *gp = g; return; }
  // overflow guard (:76-85)
  long long dx = (long long)__fmul_rn(__fsub_rn(mx[0], mn[0]), g.inv_leaf) + 1;
  long long dy = (long long)__fmul_rn(__fsub_rn(mx[1], mn[1]), g.inv_leaf) + 1;
  long long dz = (long long)__fmul_rn(__fsub_rn(mx[2], mn[2]), g.inv_leaf) + 1;
  if (dx * dy * dz > 2147483647LL) { g.status = LVS_ERR_GRID_OVERFLOW; *gp = g; return; }
  for (int a = 0; a < 3; a++) {
    g.min_b[a] = (int)floorf(__fmul_rn(mn[a], g.inv_leaf));
    g.max_b[a] = (int)floorf(__fmul_rn(mx[a], g.inv_leaf));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  g.mul[0] = 1; g.mul[1] = g.div_b[0]; g.mul[2] = g.div_b[0] * g.div_b[1];
  g.total_cells = (long long)g.div_b[0] * g.div_b[1] * g.div_b[2];
  if (g.total_cells > grid_capacity) g.status = kStatusNeedsGrow;   // the host grows the index grid and rebuilds (TargetGrid::finish)
  *gp = g;
}

// ------------------------------------------------------------------ keys
constexpr unsigned int kInvalidKey = 0xFFFFFFFFu;

__global__ void key_kernel(VoxBatch B) {
  const VoxJob& J = B.j[blockIdx.y];
  const float4* __restrict__ pts = J.pts;
  const int n = J.n;
  const GridParams* __restrict__ gp = J.gp;
  unsigned int* __restrict__ keys = J.keys[0];
  int* __restrict__ idx = J.idx[0];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  unsigned int k = kInvalidKey;
  if (gp->status == 0 && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    float inv = gp->inv_leaf;
    int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)gp->min_b[0]);
    int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)gp->min_b[1]);
    int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)gp->min_b[2]);
    k = (unsigned int)(i0 * gp->mul[0] + i1 * gp->mul[1] + i2 * gp->mul[2]);
  }
  keys[i] = k;
  idx[i] = i;
}

// ------------------------------------------------------------------ two-level exclusive scan
// scan_tiles_kernel: exclusive scan inside tiles of 2048 ints + the tile totals; scan_add_kernel: every CTA sums the totals
// of the tiles before it (they are few) and adds that offset to its tile.  total (may be null) receives the grand total.
constexpr int kScanThreads = 256, kScanIpt = 8, kScanTile = kScanThreads * kScanIpt;

__global__ void __launch_bounds__(kScanThreads) scan_tiles_kernel(ScanBatch S) {
  const ScanJob& J = S.j[blockIdx.y];
  const int n = J.n;
  if ((long long)blockIdx.x * kScanTile >= n) return;
  const int* __restrict__ in = J.in;
  int* __restrict__ out = J.out;
  int* __restrict__ tile_tot = J.tile_tot;
  __shared__ int s_warp[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanIpt;
  int v[kScanIpt], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanIpt; k++) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
  int x = sum;
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < kScanThreads / 32) ? s_warp[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
    if (lane < kScanThreads / 32) s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  int run = (warp ? s_warp[warp - 1] : 0) + x - sum;
#pragma unroll
  for (int k = 0; k < kScanIpt; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
  if (threadIdx.x == kScanThreads - 1) tile_tot[blockIdx.x] = s_warp[kScanThreads / 32 - 1];
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(ScanBatch S) {
  const ScanJob& J = S.j[blockIdx.y];
  const int n = J.n;
  if ((long long)blockIdx.x * kScanTile >= n) return;
  int* __restrict__ out = J.out;
  const int* __restrict__ tile_tot = J.tile_tot;
  int* __restrict__ total = J.total;
  const unsigned last_block = (unsigned)((n + kScanTile - 1) / kScanTile) - 1u;
  __shared__ int s_warp[kScanThreads / 32];
  __shared__ int s_off;
  int acc = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanThreads) acc += tile_tot[b];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < kScanThreads / 32; w++) t += s_warp[w]; s_off = t; }
  __syncthreads();
  const int off = s_off;
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanIpt;
#pragma unroll
  for (int k = 0; k < kScanIpt; k++) if (base + k < n) out[base + k] += off;
  if (total && blockIdx.x == last_block && threadIdx.x == 0) *total = off + tile_tot[blockIdx.x];
}

static void exclusive_scan_many(cudaStream_t st, const ScanBatch& S) {
  int nb = 0;
  for (int k = 0; k < S.count; k++) nb = std::max(nb, (S.j[k].n + kScanTile - 1) / kScanTile);
  if (nb == 0) return;
  scan_tiles_kernel<<<dim3(nb, S.count), kScanThreads, 0, st>>>(S);
  scan_add_kernel<<<dim3(nb, S.count), kScanThreads, 0, st>>>(S);
}

void exclusive_scan(cudaStream_t st, const int* in, int* out, int n, int* total, int* tile_tot) {
  ScanBatch S;
  S.count = 1;
  S.j[0] = ScanJob{in, out, n, tile_tot, total};
  exclusive_scan_many(st, S);
}

// ------------------------------------------------------------------ stable LSD radix sort, 8-bit digits
constexpr int kRsThreads = 256, kRsIpt = 8, kRsTile = kRsThreads * kRsIpt;

__global__ void rs_hist_kernel(VoxBatch B, int cur, int shift) {
  const VoxJob& J = B.j[blockIdx.y];
  const int nblk = J.nblk, n = J.n;
  if ((int)blockIdx.x >= nblk) return;
  const unsigned int* __restrict__ keys = J.keys[cur];
  int* __restrict__ hist = J.hist;                   // [256][nblk]
  __shared__ int s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * kRsTile;
  for (int r = 0; r < kRsIpt; r++) {
    int i = base + r * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void rs_scatter_kernel(VoxBatch B, int cur, int shift) {
  const VoxJob& J = B.j[blockIdx.y];
  const int nblk = J.nblk, n = J.n;
  if ((int)blockIdx.x >= nblk) return;
  const unsigned int* __restrict__ keys_in = J.keys[cur];
  const int* __restrict__ idx_in = J.idx[cur];
  const int* __restrict__ offs = J.hist_scan;        // [256][nblk] exclusive
  unsigned int* __restrict__ keys_out = J.keys[cur ^ 1];
  int* __restrict__ idx_out = J.idx[cur ^ 1];
  __shared__ int s_cnt[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < 8 * 256; k += kRsThreads) (&s_cnt[0][0])[k] = 0;
  __syncthreads();
  const int wbase = blockIdx.x * kRsTile + warp * (32 * kRsIpt);
  unsigned int key[kRsIpt];
  int val[kRsIpt], rank[kRsIpt], dig[kRsIpt];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRsIpt; r++) {
    int i = wbase + r * 32 + lane;
    bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0u;
    val[r] = ok ? idx_in[i] : 0;
    int d = ok ? (int)((key[r] >> shift) & 255) : (256 + lane);
    dig[r] = ok ? d : -1;
    unsigned m = __match_any_sync(0xffffffffu, d);
    int before = ok ? s_cnt[warp][d & 255] : 0;
    rank[r] = before + __popc(m & lt);
    __syncwarp();
    if (ok && (m & lt) == 0) s_cnt[warp][d] = before + __popc(m);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over the 8 warps for every digit
  {
    int d = threadIdx.x, run = 0;
    for (int w = 0; w < 8; w++) { int c = s_cnt[w][d]; s_cnt[w][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsIpt; r++) {
    if (dig[r] >= 0) {
      int pos = offs[dig[r] * nblk + blockIdx.x] + s_cnt[warp][dig[r]] + rank[r];
      keys_out[pos] = key[r];
      idx_out[pos] = val[r];
    }
  }
}

// ------------------------------------------------------------------ segments
__global__ void head_flag_kernel(VoxBatch B, int cur) {
  const VoxJob& J = B.j[blockIdx.y];
  const unsigned int* __restrict__ keys = J.keys[cur];
  const int n = J.n;
  int* __restrict__ flags = J.flags;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { J.nseg[0] = 0; J.nvalid[0] = 0; }      // the scan that follows deposits the segment count here
  if (i >= n) return;
  unsigned k = keys[i];
  flags[i] = (k != kInvalidKey && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

__global__ void seg_start_kernel(VoxBatch B, int cur) {
  const VoxJob& J = B.j[blockIdx.y];
  const unsigned int* __restrict__ keys = J.keys[cur];
  const int* __restrict__ flags = J.flags;
  const int* __restrict__ pos = J.pos;
  const int n = J.n;
  int* __restrict__ seg_start = J.cell_start;
  const int* __restrict__ n_seg = J.nseg;
  int* __restrict__ n_valid_pts = J.nvalid;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) seg_start[pos[i]] = i;
  // end sentinel: first invalid key or n
  if (keys[i] != kInvalidKey && (i == n - 1 || keys[i + 1] == kInvalidKey)) { seg_start[*n_seg] = i + 1; *n_valid_pts = i + 1; }
}

__global__ void clear_cells_kernel(VoxBatch B) {
  const VoxJob& J = B.j[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (J.prev_points > 0 && J.gp->status == 0 && i < J.gp->n_cells) J.grid[J.old_keys[i]] = -1;
}

// ------------------------------------------------------------------ per-cell moments
// One warp per occupied cell.  The reference adds each point to the leaf's Sx (3 doubles), Sxx^T (double) and float centroid in
// INPUT ORDER (voxel_grid_covariance_omp_impl.hpp:233-262); its one-pass covariance cancels heavily, so the sums are reproduced
// with exactly that order and rounding: twelve serial chains per cell (lanes 0..8 the f64 moments, lanes 9..11 the f32 centroid
// sums).  The only serial resource is the dependent add, so everything else is taken off the chain:
//   * all 32 lanes gather one point each and stage its twelve addends (products already formed, exact in f64) in shared memory,
//     double-buffered, two chunks ahead of the chain (the index and the point of later chunks are already in flight);
//   * the chain runs 32 fully unrolled steps per chunk, every lane issuing one DADD and one FADD per step (the lanes that do not
//     own a chain of that type add into a dead register), so there is no divergence and the tail is padded with +0.0, which is
//     exact: an accumulator that starts at +0.0 can never become -0.0.
constexpr int kMomWarps = 4;                 // warps per CTA
constexpr int kMomCtasPerSm = 9;                // 22.5 KB of staging per CTA
struct __align__(16) MomStage {
  double d[32][9];                           // Sx addends (3), upper triangle of x x^T (6)
  float f[32][4];                            // float centroid addends (+ pad)
};

__device__ __forceinline__ void mom_stage(MomStage& s, int lane, bool live, const float4& p) {
  double x = 0.0, y = 0.0, z = 0.0;
  float fx = 0.0f, fy = 0.0f, fz = 0.0f;
  if (live) { x = (double)p.x; y = (double)p.y; z = (double)p.z; fx = p.x; fy = p.y; fz = p.z; }
  double* d = s.d[lane];
  d[0] = x; d[1] = y; d[2] = z;
  d[3] = __dmul_rn(x, x); d[4] = __dmul_rn(x, y); d[5] = __dmul_rn(x, z);
  d[6] = __dmul_rn(y, y); d[7] = __dmul_rn(y, z); d[8] = __dmul_rn(z, z);
  *reinterpret_cast<float4*>(s.f[lane]) = make_float4(fx, fy, fz, 0.0f);
}

__global__ void __launch_bounds__(kMomWarps * 32) leaf_moments_kernel(VoxBatch B) {
  const VoxJob& J = B.j[blockIdx.y];
  const float4* __restrict__ pts = J.pts;
  const int* __restrict__ sidx = J.sorted;
  const int* __restrict__ seg_start = J.cell_start;
  const int* __restrict__ n_seg_p = J.nseg;
  const GridParams* __restrict__ gp = J.gp;
  double* __restrict__ moments = J.moments;          // [cell][9]
  float* __restrict__ csum = J.csum;                 // [cell][3]
  if (gp->status != 0) return;
  __shared__ MomStage s_stage[kMomWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_seg = *n_seg_p;
  const int warps_total = gridDim.x * kMomWarps;
  const int dcol = lane < 9 ? lane : 8;                  // chain owned by this lane (lanes >= 9 shadow chain 8 into a dead register)
  const int fcol = (lane >= 9 && lane < 12) ? lane - 9 : 3;   // column 3 is the zero pad
  MomStage* st = s_stage[warp];
  for (int seg = blockIdx.x * kMomWarps + warp; seg < n_seg; seg += warps_total) {
    const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
    const int n_chunks = (s1 - s0 + 31) >> 5;
    // Leaf::cov_ starts as the IDENTITY (voxel_grid_covariance_omp.h:98-106) and applyFilter adds the point products on top of it (:240,:285):
    // the xx, yy, zz chains start at 1.0, so every covariance carries + I (n - 1) / n^2 like the reference's
    double acc = (dcol == 3 || dcol == 6 || dcol == 8) ? 1.0 : 0.0;
    float facc = 0.0f;
    // software pipeline: while chunk c is summed, the points of chunks c+1..c+3 and the index of chunk c+4 are already in
    // registers or in flight (a gather is ~2 dependent L2 round trips, one chunk's chain only ~300 cycles)
    auto load_idx = [&](int chunk) { const int i = s0 + chunk * 32 + lane; return i < s1 ? __ldg(sidx + i) : -1; };
    auto load_pt = [&](int idx) { float4 p = make_float4(0.f, 0.f, 0.f, 0.f); if (idx >= 0) { p = __ldg(pts + idx); p.w = 1.0f; } return p; };
    const int j0 = load_idx(0), j1 = load_idx(1), j2 = load_idx(2), j3 = load_idx(3);
    int idn = load_idx(4);
    const float4 p0 = load_pt(j0);
    float4 pa = load_pt(j1), pb = load_pt(j2), pc = load_pt(j3);
    __syncwarp();
    mom_stage(st[0], lane, p0.w != 0.0f, p0);
    __syncwarp();
    for (int c = 0; c < n_chunks; c++) {
      const MomStage& cur = st[c & 1];
      if (c + 1 < n_chunks) mom_stage(st[(c + 1) & 1], lane, pa.w != 0.0f, pa);
      pa = pb; pb = pc;
      pc = load_pt(idn);
      idn = load_idx(c + 5);
#pragma unroll
      for (int k = 0; k < 32; k++) {
        acc = __dadd_rn(acc, cur.d[k][dcol]);
        facc = __fadd_rn(facc, cur.f[k][fcol]);
      }
      __syncwarp();
    }
    if (lane < 9) moments[(size_t)seg * 9 + lane] = acc;
    else if (lane < 12) csum[(size_t)seg * 3 + (lane - 9)] = facc;
  }
}

// ------------------------------------------------------------------ leaf finalisation, one thread per cell (:281-367)
__global__ void __launch_bounds__(128) leaf_finalize_kernel(VoxBatch B, int cur) {
  const VoxJob& J = B.j[blockIdx.y];
  const unsigned int* __restrict__ keys = J.keys[cur];
  const int* __restrict__ seg_start = J.cell_start;
  const int* __restrict__ n_seg_p = J.nseg;
  const double* __restrict__ moments = J.moments;
  const float* __restrict__ csum = J.csum;
  VoxelRec* __restrict__ recs = J.recs;
  FastRec* __restrict__ frecs = J.frecs;
  float4* __restrict__ centroids = J.centroids;
  int* __restrict__ cell_keys = J.cell_keys;
  int* __restrict__ cell_npts = J.cell_npts;
  double* __restrict__ cell_evals = J.cell_evals;
  double* __restrict__ icov64 = J.icov64;
  int* __restrict__ grid = J.grid;
  GridParams* __restrict__ gp = J.gp;
  const int min_points = B.min_points, variant = B.variant;
  const double eig_mult = B.eig_mult;
  if (gp->status != 0) return;
  const int n_seg = *n_seg_p;
  int valid_cells = 0;
  for (int seg = blockIdx.x * blockDim.x + threadIdx.x; seg < n_seg; seg += gridDim.x * blockDim.x) {
    const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
    const double* mo = moments + (size_t)seg * 9;
    const double S1[3] = {mo[0], mo[1], mo[2]};
    const double S2[6] = {mo[3], mo[4], mo[5], mo[6], mo[7], mo[8]};
    const int npts = s1 - s0;
    const int key = (int)keys[s0];
    const double np = (double)npts;
    double mean[3] = {S1[0] / np, S1[1] / np, S1[2] / np};
    VoxelRec rec;
    double ic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < 3; a++) rec.mean[a] = mean[a];
    for (int a = 0; a < 9; a++) rec.icov[a] = 0.0f;
    rec.meta = 1;
    const float fn = (float)npts;
    centroids[seg] = make_float4(__fdiv_rn(csum[(size_t)seg * 3], fn), __fdiv_rn(csum[(size_t)seg * 3 + 1], fn), __fdiv_rn(csum[(size_t)seg * 3 + 2], fn),
                                 npts >= min_points ? 1.0f : 0.0f);
    int out_npts = npts;
    double ev[3] = {0, 0, 0};
    bool usable = false;
    if (npts >= min_points) {
      // single-pass covariance (:329-330): full 3x3, element (i,j) uses pt_sum[i]*mean[j]
      double cov[9];
      const int sidx2[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cov[i * 3 + j] = (S2[sidx2[i * 3 + j]] - 2.0 * (S1[i] * mean[j])) / np + mean[i] * mean[j];
      const double sc = (np - 1.0) / np;
      for (int a = 0; a < 9; a++) cov[a] *= sc;
      double V[9];
      sym3_eigen(cov, ev, V);                 // reads the lower triangle, like SelfAdjointEigenSolver
      // pclomp_ground: is the leaf's normal (first column of evecs_, assigned before the eigenvalue check below, :335) within 10 degrees of
      // the z axis?  acos(|n_z| / |n|) * 180 / 3.1415926 < 10 as ndt_ground_impl.hpp:507-511,533 writes it
      bool horiz = false;
      if (variant == LVS_NDT_GROUND) {
        const double nrm = sqrt((V[0] * V[0] + V[3] * V[3]) + V[6] * V[6]);
        horiz = acos(fabs(V[6] / nrm)) * 180 / 3.1415926 < 10;
      }
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        out_npts = -1;
        ev[0] = ev[1] = ev[2] = 0.0;          // evals_ is assigned only after the check (:357)
      } else {
        const double min_ev = eig_mult * ev[2];
        if (ev[0] < min_ev) {
          ev[0] = min_ev;
          if (ev[1] < min_ev) ev[1] = min_ev;
          double Vi[9], VL[9];
          mat3_inverse(V, Vi);
          for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) VL[i * 3 + j] = V[i * 3 + j] * ev[j];
          mat3_mul(VL, Vi, cov);
        }
        int weight = 1;
        if (variant == LVS_NDT_PCA) {
          const double s0_ = sqrt(ev[0]), s1_ = sqrt(ev[1]), s2_ = sqrt(ev[2]);
          const double f0 = (s2_ - s1_) / s2_, f1 = (s1_ - s0_) / s2_, f2 = s0_ / s2_;
          int d = 0; double fm = f0;
          if (f1 > fm) { d = 1; fm = f1; }
          if (f2 > fm) { d = 2; }
          const double scale = d == 1 ? 1.25 : (d == 2 ? 1.0 : 0.75);
          const double nm = sqrt(mean[0] * mean[0] + mean[1] * mean[1] + mean[2] * mean[2]);
          weight = (int)(scale * nm);           // int getDimension2d() truncation (voxel_grid_covariance_pca.h:222-226)
        }
        mat3_inverse(cov, ic);
        double mxc = ic[0], mnc = ic[0];
        for (int a = 1; a < 9; a++) { mxc = fmax(mxc, ic[a]); mnc = fmin(mnc, ic[a]); }
        for (int a = 0; a < 9; a++) rec.icov[icov_slot(a)] = (float)ic[a];
        if (mxc == (double)INFINITY || mnc == -(double)INFINITY) out_npts = -1;
        else usable = true;
        rec.meta = (weight & kMetaWeightMask) | (usable ? kMetaValidBit : 0);
      }
      if (horiz) rec.meta |= kMetaHorizBit;
    }
    recs[seg] = rec;
    {
      FastRec fr;
      for (int a = 0; a < 3; a++) { fr.mh[a] = (float)mean[a]; fr.ml[a] = (float)(mean[a] - (double)fr.mh[a]); }
      fr.c[0] = (float)ic[0]; fr.c[1] = (float)ic[1]; fr.c[2] = (float)ic[2]; fr.c[3] = (float)ic[4]; fr.c[4] = (float)ic[5]; fr.c[5] = (float)ic[8];
      frecs[seg] = fr;
    }
    for (int a = 0; a < 9; a++) icov64[(size_t)seg * 9 + a] = ic[a];
    cell_keys[seg] = key;
    cell_npts[seg] = out_npts;
    cell_evals[seg * 3] = ev[0]; cell_evals[seg * 3 + 1] = ev[1]; cell_evals[seg * 3 + 2] = ev[2];
    grid[key] = usable ? seg : (-2 - seg);
    if (usable) valid_cells++;
  }
  if (valid_cells) atomicAdd(&gp->n_valid, valid_cells);
  if (blockIdx.x == 0 && threadIdx.x == 0) gp->n_cells = n_seg;
}

// ------------------------------------------------------------------ host side
static int radix_passes_for(long long cells) {   // keys < cells, invalid keys 0xFFFFFFFF sort last in every pass
  int bits = 1;
  while (bits < 32 && (1LL << bits) <= cells) bits++;
  return (bits + 7) / 8;
}

// ---- launch helpers over a batch of jobs
static void job_scratch(VoxJob& J, BuildScratch& ws, int passes, int* sorted_out, int* cell_start_out) {
  J.bbox_partial = ws.d_bbox_partial; J.ticket = ws.d_ticket;
  J.keys[0] = ws.d_keys[0]; J.keys[1] = ws.d_keys[1];
  // the index buffer the LAST pass writes is the caller's sorted_out, so the sorted order needs no extra copy; pass 0 reads the
  // identity permutation key_kernel leaves in idx[0]
  J.idx[0] = ws.d_idx[0]; J.idx[1] = ws.d_idx[1];
  J.idx[passes & 1] = sorted_out;
  J.hist = ws.d_hist; J.hist_scan = ws.d_hist_scan; J.tile_tot = ws.d_tile_tot; J.flags = ws.d_flags; J.pos = ws.d_pos;
  J.nseg = ws.d_nseg; J.nvalid = ws.d_nvalidpts;
  J.sorted = sorted_out; J.cell_start = cell_start_out;
  J.moments = ws.d_moments; J.csum = ws.d_csum;
  J.nblk = (J.n + kRsTile - 1) / kRsTile;
}

static int max_n(const VoxBatch& B) { int m = 0; for (int k = 0; k < B.count; k++) m = std::max(m, B.j[k].n); return m; }

static void launch_bbox(cudaStream_t st, const VoxBatch& B) {
  const int gb = (max_n(B) + 255) / 256;
  bbox_kernel<<<dim3(std::max(1, std::min(kBboxBlocks, gb)), B.count), 256, 0, st>>>(B);
}

// key -> stable radix sort -> segment heads -> scan -> segment starts.  Returns the index of the key buffer that holds the sorted keys.
static int launch_sort_segments(cudaStream_t st, const VoxBatch& B, int passes) {
  const int n = max_n(B), tb = 256, gb = (n + tb - 1) / tb;
  const int nblk = (n + kRsTile - 1) / kRsTile;
  key_kernel<<<dim3(gb, B.count), tb, 0, st>>>(B);
  int cur = 0;
  ScanBatch S;
  S.count = B.count;
  for (int p = 0; p < passes; p++) {
    rs_hist_kernel<<<dim3(nblk, B.count), kRsThreads, 0, st>>>(B, cur, p * 8);
    for (int k = 0; k < B.count; k++) S.j[k] = ScanJob{B.j[k].hist, B.j[k].hist_scan, 256 * B.j[k].nblk, B.j[k].tile_tot, nullptr};
    exclusive_scan_many(st, S);
    // pass 0 reads the identity permutation from the scratch buffer even when idx[0] was redirected to the caller's array
    rs_scatter_kernel<<<dim3(nblk, B.count), kRsThreads, 0, st>>>(B, cur, p * 8);
    cur ^= 1;
  }
  head_flag_kernel<<<dim3(gb, B.count), tb, 0, st>>>(B, cur);
  for (int k = 0; k < B.count; k++) S.j[k] = ScanJob{B.j[k].flags, B.j[k].pos, B.j[k].n, B.j[k].tile_tot, B.j[k].nseg};
  exclusive_scan_many(st, S);
  seg_start_kernel<<<dim3(gb, B.count), tb, 0, st>>>(B, cur);
  return cur;
}

// Bounding box and grid geometry of a cloud (getMinMax3D + the min_b / max_b / div_b arithmetic) into *d_gp.
void vox_bbox(cudaStream_t st, const float4* pts, int n, BuildScratch& ws, GridParams* d_gp, float leaf, long long grid_capacity) {
  VoxBatch B;
  memset(&B, 0, sizeof B);
  B.count = 1;
  VoxJob& J = B.j[0];
  J.pts = pts; J.n = n; J.leaf = leaf; J.grid_capacity = grid_capacity; J.gp = d_gp;
  J.bbox_partial = ws.d_bbox_partial; J.ticket = ws.d_ticket;
  launch_bbox(st, B);
}

int vox_radix_passes(long long cells) { return radix_passes_for(cells); }

// Voxel key of every point, stable sort of the point indices by key, one segment per occupied cell: sorted_out [n] receives the
// point indices grouped by cell in ascending key order (input order inside a cell), cell_start_out [cells + 1] the segment
// starts, ws.d_nseg the number of cells.
int vox_sort_segments(cudaStream_t st, const float4* pts, int n, const GridParams* d_gp, int passes, BuildScratch& ws, int* sorted_out,
                      int* cell_start_out) {
  VoxBatch B;
  memset(&B, 0, sizeof B);
  B.count = 1;
  VoxJob& J = B.j[0];
  J.pts = pts; J.n = n; J.gp = const_cast<GridParams*>(d_gp);
  job_scratch(J, ws, passes, sorted_out, cell_start_out);
  launch_sort_segments(st, B, passes);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

// Queues the voxelisation of `count` targets on `st` in batched launches.  Every grid needs d_grid and its cell arrays, and its
// own scratch; never synchronises.
int TargetGrid::enqueue_many(cudaStream_t st, int count, TargetGrid* const* grids, BuildScratch* const* wss, const lvs_ndt_params& prm) {
  for (int base = 0; base < count; base += kVoxBatch) {
    VoxBatch B;
    memset(&B, 0, sizeof B);
    B.count = std::min(kVoxBatch, count - base);
    B.min_points = prm.min_points_per_voxel; B.eig_mult = prm.min_covar_eigvalue_mult; B.variant = prm.variant;
    int passes = 1, max_prev = 0;
    for (int k = 0; k < B.count; k++) passes = std::max(passes, radix_passes_for((long long)grids[base + k]->grid_capacity));
    for (int k = 0; k < B.count; k++) {
      TargetGrid& g = *grids[base + k];
      VoxJob& J = B.j[k];
      J.pts = g.pts; J.n = g.n_points; J.leaf = prm.resolution; J.grid_capacity = (long long)g.grid_capacity; J.gp = g.d_gp;
      J.grid = g.d_grid; J.old_keys = g.d_cell_keys; J.prev_points = g.prev_points;
      J.recs = g.d_recs; J.frecs = g.d_frecs; J.centroids = g.d_centroids; J.cell_keys = g.d_cell_keys; J.cell_npts = g.d_cell_npts; J.cell_evals = g.d_cell_evals;
      J.icov64 = g.d_icov64;
      job_scratch(J, *wss[base + k], passes, g.d_sorted_idx, g.d_cell_start);
      max_prev = std::max(max_prev, g.prev_points);
    }
    // wipe the cells of the previous builds while the old keys and the old geometry still describe the index grids
    if (max_prev > 0) clear_cells_kernel<<<dim3((max_prev + 255) / 256, B.count), 256, 0, st>>>(B);
    launch_bbox(st, B);
    const int cur = launch_sort_segments(st, B, passes);
    const int n = max_n(B);
    leaf_moments_kernel<<<dim3(std::min(148 * kMomCtasPerSm, (n + kMomWarps - 1) / kMomWarps), B.count), kMomWarps * 32, 0, st>>>(B);
    leaf_finalize_kernel<<<dim3(std::min(148 * 4, (n + 127) / 128), B.count), 128, 0, st>>>(B, cur);
    CUDA_TRY(cudaGetLastError());
    for (int k = 0; k < B.count; k++) {
      TargetGrid& g = *grids[base + k];
      g.launches_last_build = k == 0 ? (max_prev > 0 ? 1 : 0) + 2 + passes * 4 + 1 + 2 + 1 + 2 : 0;   // the batch's launches, booked once
      g.prev_points = g.n_points;
      g.pending = true;
      g.sorted_pts_valid = false;
    }
  }
  return LVS_OK;
}

int TargetGrid::enqueue(cudaStream_t st, const lvs_ndt_params& prm, BuildScratch& ws) {
  TargetGrid* g = this;
  BuildScratch* w = &ws;
  return enqueue_many(st, 1, &g, &w, prm);
}

// Host-side preparation of a build (allocations, first-build sizing of the index grid).  Returns 1 when the target is ready for
// enqueue / enqueue_many, 0 when the build is already complete (empty cloud), < 0 on error.
int TargetGrid::prepare(cudaStream_t st, const float4* d_pts, int n, const lvs_ndt_params& prm, BuildScratch& ws) {
  n_points = n;
  pts = d_pts;
  built_with = prm;
  CUDA_TRY(ws.reserve(n));
  if (!d_gp) {
    CUDA_TRY(cudaMalloc(&d_gp, sizeof(GridParams)));
    GridParams g; memset(&g, 0, sizeof g); g.status = kStatusEmpty;
    CUDA_TRY(cudaMemcpyAsync(d_gp, &g, sizeof g, cudaMemcpyHostToDevice, st));
  }
  if (n == 0) {
    if (prev_points > 0 && d_grid) {
      VoxBatch B;
      memset(&B, 0, sizeof B);
      B.count = 1;
      B.j[0].gp = d_gp; B.j[0].grid = d_grid; B.j[0].old_keys = d_cell_keys; B.j[0].prev_points = prev_points;
      clear_cells_kernel<<<dim3((prev_points + 255) / 256, 1), 256, 0, st>>>(B);
    }
    prev_points = 0;
    GridParams g; memset(&g, 0, sizeof g); g.status = kStatusEmpty; g.leaf = prm.resolution; g.inv_leaf = 1.0f / prm.resolution;
    CUDA_TRY(cudaMemcpyAsync(d_gp, &g, sizeof g, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // g lives on this stack frame
    gp = g; n_cells = 0; pending = false; launches_last_build = 0;
    return 0;
  }
  // cell arrays sized for the worst case of one cell per point
  if ((size_t)n > cell_capacity) {
    CUDA_TRY(cudaStreamSynchronize(st));
    free_cells();
    prev_points = 0;                        // the old keys are gone; the grid is re-filled below
    if (d_grid) CUDA_TRY(cudaMemsetAsync(d_grid, 0xFF, grid_capacity * sizeof(int), st));
    size_t cap = (size_t)n + n / 8 + 64;
    CUDA_TRY(cudaMalloc(&d_recs, cap * sizeof(VoxelRec)));
    CUDA_TRY(cudaMalloc(&d_frecs, cap * sizeof(FastRec)));
    CUDA_TRY(cudaMalloc(&d_centroids, cap * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&d_cell_keys, cap * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cell_npts, cap * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cell_evals, cap * 3 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_icov64, cap * 9 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_sorted_idx, cap * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cell_start, (cap + 2) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_sorted_pts, cap * sizeof(float4)));
    cell_capacity = cap;
  }
  if (!d_grid) {
    // first build of this slot: the index grid has no capacity yet, so size it from the geometry (one synchronisation, once)
    vox_bbox(st, d_pts, n, ws, d_gp, prm.resolution, 0LL);
    CUDA_TRY(cudaMemcpyAsync(ws.h_gp, d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    gp = *ws.h_gp;
    if (gp.status == LVS_ERR_GRID_OVERFLOW) { pending = false; n_cells = 0; return fail(LVS_ERR_GRID_OVERFLOW, "leaf size too small for the target extent: dx*dy*dz > INT32_MAX"); }
    if (gp.status == kStatusEmpty) { pending = false; n_cells = 0; return 0; }   // no finite point
    int rc = grow_grid(st, gp.total_cells);
    if (rc) return rc;
  }
  return 1;
}

int TargetGrid::build(cudaStream_t st, const float4* d_pts, int n, const lvs_ndt_params& prm, BuildScratch& ws) {
  const int rc = prepare(st, d_pts, n, prm, ws);
  if (rc <= 0) return rc;
  return enqueue(st, prm, ws);
}

int TargetGrid::grow_grid(cudaStream_t st, long long cells) {
  if (d_grid) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(d_grid); }
  d_grid = nullptr; grid_capacity = 0;
  size_t cap = (size_t)cells + (size_t)cells / 4 + 1024;
  cudaError_t e = cudaMalloc(&d_grid, cap * sizeof(int));
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(LVS_ERR_OOM, "dense voxel index grid allocation failed (%lld cells)", cells); }
  grid_capacity = cap;
  CUDA_TRY(cudaMemsetAsync(d_grid, 0xFF, cap * sizeof(int), st));
  prev_points = 0;
  return LVS_OK;
}

// Completes a queued build: reads the geometry back (one small copy + synchronisation) and, if the bounding box outgrew the
// index grid, grows it and rebuilds.  Called lazily by whoever needs the grid (align, taps).
int TargetGrid::finish(cudaStream_t st, BuildScratch& ws) {
  for (int attempt = 0; pending && attempt < 3; attempt++) {
    CUDA_TRY(cudaMemcpyAsync(ws.h_gp, d_gp, sizeof(GridParams), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    int rc = accept(*ws.h_gp, st, ws);
    if (rc) return rc;
  }
  return LVS_OK;
}

// Takes a geometry record that has just been read back for the queued build.  Leaves `pending` set when it had to re-queue.
int TargetGrid::accept(const GridParams& g, cudaStream_t st, BuildScratch& ws) {
  gp = g;
  pending = false;
  if (gp.status == LVS_ERR_GRID_OVERFLOW) { n_cells = 0; return fail(LVS_ERR_GRID_OVERFLOW, "leaf size too small for the target extent: dx*dy*dz > INT32_MAX"); }
  if (gp.status == kStatusNeedsGrow) {
    int rc = grow_grid(st, gp.total_cells);
    if (rc) return rc;
    return enqueue(st, built_with, ws);     // pending again
  }
  n_cells = gp.status == 0 ? gp.n_cells : 0;
  return LVS_OK;
}

void TargetGrid::free_cells() {
  if (d_recs) cudaFree(d_recs);
  if (d_frecs) cudaFree(d_frecs);
  d_frecs = nullptr;
  if (d_centroids) cudaFree(d_centroids);
  if (d_cell_keys) cudaFree(d_cell_keys);
  if (d_cell_npts) cudaFree(d_cell_npts);
  if (d_cell_evals) cudaFree(d_cell_evals);
  if (d_icov64) cudaFree(d_icov64);
  if (d_sorted_idx) cudaFree(d_sorted_idx);
  if (d_cell_start) cudaFree(d_cell_start);
  if (d_sorted_pts) cudaFree(d_sorted_pts);
  if (d_slab) cudaFree(d_slab);
  d_slab = nullptr; slab_cells = 0;
  d_icov64 = nullptr; d_sorted_idx = nullptr; d_cell_start = nullptr; d_sorted_pts = nullptr; sorted_pts_valid = false;
  d_recs = nullptr; d_centroids = nullptr; d_cell_keys = nullptr; d_cell_npts = nullptr; d_cell_evals = nullptr;
  cell_capacity = 0;
}

void TargetGrid::release() {
  free_cells();
  if (d_grid) cudaFree(d_grid);
  if (d_gp) cudaFree(d_gp);
  d_grid = nullptr; d_gp = nullptr; grid_capacity = 0; n_cells = 0;
}

cudaError_t BuildScratch::reserve(int n) {
  if (n <= capacity) return cudaSuccess;
  release();
  int cap = n + n / 8 + 1024;
  int nblk = (cap + kRsTile - 1) / kRsTile;
  cudaError_t e;
  for (int k = 0; k < 2; k++) {
    if ((e = cudaMalloc(&d_keys[k], (size_t)cap * sizeof(unsigned int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d_idx[k], (size_t)cap * sizeof(int))) != cudaSuccess) return e;
  }
  if ((e = cudaMalloc(&d_flags, (size_t)cap * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_pos, (size_t)cap * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_seg_start, ((size_t)cap + 2) * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_hist, (size_t)256 * nblk * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_hist_scan, (size_t)256 * nblk * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_bbox_partial, 1024 * 8 * sizeof(float))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_tile_tot, (size_t)(256 * nblk / kScanTile + cap / kScanTile + 16) * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_moments, (size_t)cap * 9 * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_csum, (size_t)cap * 3 * sizeof(float))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_ticket, 4 * sizeof(unsigned int))) != cudaSuccess) return e;
  if ((e = cudaMemset(d_ticket, 0, 4 * sizeof(unsigned int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_nseg, 4 * sizeof(int))) != cudaSuccess) return e;
  d_nvalidpts = d_nseg + 1;
  if (!h_gp && (e = cudaMallocHost(&h_gp, sizeof(GridParams))) != cudaSuccess) return e;
  capacity = cap;
  return cudaSuccess;
}

void BuildScratch::release() {
  for (int k = 0; k < 2; k++) { if (d_keys[k]) cudaFree(d_keys[k]); if (d_idx[k]) cudaFree(d_idx[k]); d_keys[k] = nullptr; d_idx[k] = nullptr; }
  if (d_flags) cudaFree(d_flags);
  if (d_pos) cudaFree(d_pos);
  if (d_seg_start) cudaFree(d_seg_start);
  if (d_hist) cudaFree(d_hist);
  if (d_hist_scan) cudaFree(d_hist_scan);
  if (d_bbox_partial) cudaFree(d_bbox_partial);
  if (d_tile_tot) cudaFree(d_tile_tot);
  if (d_moments) cudaFree(d_moments);
  if (d_csum) cudaFree(d_csum);
  d_tile_tot = nullptr; d_moments = nullptr; d_csum = nullptr;
  if (d_ticket) cudaFree(d_ticket);
  if (d_nseg) cudaFree(d_nseg);
  d_flags = d_pos = d_seg_start = d_hist = d_hist_scan = nullptr; d_bbox_partial = nullptr; d_ticket = nullptr; d_nseg = nullptr;
  capacity = 0;
}

int pack_many(cudaStream_t st, const PackMany& pm, int max_n) {
  if (pm.count <= 0 || max_n <= 0) return LVS_OK;
  dim3 grid((unsigned)std::min((max_n + 255) / 256, 148 * 4), (unsigned)pm.count);
  pack_many_kernel<<<grid, 256, 0, st>>>(pm);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

int pack_points(cudaStream_t st, const float* d_in, size_t stride_floats, int n, float4* d_out) {
  if (n == 0) return LVS_OK;
  pack_points_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, stride_floats, n, d_out);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

}  // namespace lvs
