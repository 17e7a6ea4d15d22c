// Radius-search evaluation passes (cold paths) and taps: KDTREE-mode computeDerivatives, computeHessian / updateHessian
// (include/ndt_omp/ndt_omp_impl2.hpp:623-714 — all fp64, radius neighbours, unweighted), calculateScore (:1007-1040),
// radiusSearch (include/ndt_omp/voxel_grid_covariance_omp.h:506-534), lookup-key and output-cloud taps.
// The reference only reaches computeHessian when its More-Thuente loop runs, i.e. when step_size <= transformation_epsilon / 2.
#include "ndt_eval_common.cuh"

namespace lvs {

// Accumulators of one thread.
template <bool HESS>
struct Acc {
  double score;
  double g[6];
  double H[HESS ? 36 : 1];
  __device__ __forceinline__ void zero() {
    score = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) g[i] = 0;
#pragma unroll
    for (int i = 0; i < (HESS ? 36 : 1); i++) H[i] = 0;
  }
};

// updateDerivatives + computePointDerivatives_AngleAxisd for one (point, cell) pair.  xr = R*x (float), d = x' - mean,
// C = float inverse covariance (row-major), w = weight multiplier applied to this cell's contribution.
// The zero / identity entries of the reference's 4x6 and 24x6 matrices are folded away by hand; every remaining
// product and sum is in the reference's (Eigen SSE) order, so each float contribution is bit-identical to the CPU path.
template <bool HESS>
__device__ __forceinline__ void contribute(Acc<HESS>& A, float xr, float yr, float zr, float d0, float d1, float d2, const float* C,
                                           float gd2, double gauss_d1, double w) {
  // xC = d^T C  : per column (t0 + t2) + t1
  const float xC0 = (d0 * C[0] + d2 * C[6]) + d1 * C[3];
  const float xC1 = (d0 * C[1] + d2 * C[7]) + d1 * C[4];
  const float xC2 = (d0 * C[2] + d2 * C[8]) + d1 * C[5];
  const float q = (d0 * xC0 + d2 * xC2) + d1 * xC1;
  float e = glibc_expf((-gd2 * q) * 0.5f, c_exp2f_tab);      // the reference's exp(float) is expf (lvs_math.cuh)
  const float score_inc = (float)(-gauss_d1 * (double)e);
  e = gd2 * e;
  if (e > kOne || e < 0.0f || e != e) return;
  e = (float)((double)e * gauss_d1);

  const float nx = -xr, ny = -yr, nz = -zr;
  // CJ = C * J, columns 3..5 (columns 0..2 are C itself)
  float CJ3[3], CJ4[3], CJ5[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    CJ3[i] = C[i * 3 + 1] * nz + C[i * 3 + 2] * yr;
    CJ4[i] = C[i * 3 + 0] * zr + C[i * 3 + 2] * nx;
    CJ5[i] = C[i * 3 + 0] * ny + C[i * 3 + 1] * xr;
  }
  float a[6];
  a[0] = xC0; a[1] = xC1; a[2] = xC2;
  a[3] = (d0 * CJ3[0] + d2 * CJ3[2]) + d1 * CJ3[1];
  a[4] = (d0 * CJ4[0] + d2 * CJ4[2]) + d1 * CJ4[1];
  a[5] = (d0 * CJ5[0] + d2 * CJ5[2]) + d1 * CJ5[1];

  A.score += (double)score_inc * w;
#pragma unroll
  for (int j = 0; j < 6; j++) A.g[j] += (double)(e * a[j]) * w;

  if (HESS) {
    // CJ as a 3x6 (row i, column c)
    float CJ[3][6];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      CJ[i][0] = C[i * 3]; CJ[i][1] = C[i * 3 + 1]; CJ[i][2] = C[i * 3 + 2];
      CJ[i][3] = CJ3[i]; CJ[i][4] = CJ4[i]; CJ[i][5] = CJ5[i];
    }
    // Mx[r][c] = J.col(r) . CJ.col(c)
    float Mx[6][6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      Mx[0][c] = CJ[0][c]; Mx[1][c] = CJ[1][c]; Mx[2][c] = CJ[2][c];
      Mx[3][c] = yr * CJ[2][c] + nz * CJ[1][c];
      Mx[4][c] = zr * CJ[0][c] + nx * CJ[2][c];
      Mx[5][c] = ny * CJ[0][c] + xr * CJ[1][c];
    }
    // hp[i][j] = (d^T C) . Hp_ij  (non-zero only in the rotation block)
    float hp[3][3];
    hp[0][0] = xC2 * nz + xC1 * ny; hp[0][1] = xC1 * xr;            hp[0][2] = xC2 * xr;
    hp[1][0] = xC0 * yr;            hp[1][1] = xC0 * nx + xC2 * nz; hp[1][2] = xC2 * yr;
    hp[2][0] = xC0 * zr;            hp[2][1] = xC1 * zr;            hp[2][2] = xC0 * nx + xC1 * ny;
    const float ngd2 = -gd2;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const float ai = ngd2 * a[i];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        float t = ai * a[j];
        if (i >= 3 && j >= 3) t = t + hp[i - 3][j - 3];
        A.H[i * 6 + j] += (double)(e * (t + Mx[j][i])) * w;
      }
    }
  }
}

// updateHessian + the double overload of computePointDerivatives_AngleAxisd (:535-563, :683-714), all fp64.
__device__ __forceinline__ void contribute_hess64(double* H, const double* r, const double* d, const double* C, double gd1, double gd2) {
  double Cd[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Cd[i] = (C[i * 3] * d[0] + C[i * 3 + 1] * d[1]) + C[i * 3 + 2] * d[2];
  double e = gd2 * exp(-gd2 * ((d[0] * Cd[0] + d[1] * Cd[1]) + d[2] * Cd[2]) / 2);
  if (e > 1 || e < 0 || e != e) return;
  e *= gd1;
  const double x = r[0], y = r[1], z = r[2];
  double CJ[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    CJ[i][0] = C[i * 3]; CJ[i][1] = C[i * 3 + 1]; CJ[i][2] = C[i * 3 + 2];
    CJ[i][3] = C[i * 3 + 1] * (-z) + C[i * 3 + 2] * y;
    CJ[i][4] = C[i * 3] * z + C[i * 3 + 2] * (-x);
    CJ[i][5] = C[i * 3] * (-y) + C[i * 3 + 1] * x;
  }
  double dCJ[6];
#pragma unroll
  for (int c = 0; c < 6; c++) dCJ[c] = (d[0] * CJ[0][c] + d[1] * CJ[1][c]) + d[2] * CJ[2][c];
  // second-derivative vectors Hp[i][j] (i, j in 3..5), Appendix A.1 of SURVEY.md
  const double hpv[3][3][3] = {{{0, -y, -z}, {0, x, 0}, {0, 0, x}}, {{y, 0, 0}, {-x, 0, -z}, {0, 0, y}}, {{z, 0, 0}, {0, z, 0}, {-x, -y, 0}}};
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) {
      double dCH = 0.0;
      if (i >= 3 && j >= 3) {
        const double* h = hpv[i - 3][j - 3];
        double CH[3];
#pragma unroll
        for (int k = 0; k < 3; k++) CH[k] = (C[k * 3] * h[0] + C[k * 3 + 1] * h[1]) + C[k * 3 + 2] * h[2];
        dCH = (d[0] * CH[0] + d[1] * CH[1]) + d[2] * CH[2];
      }
      // J.col(j) . CJ.col(i)
      double JCJ;
      if (j < 3) JCJ = CJ[j][i];
      else if (j == 3) JCJ = (-z) * CJ[1][i] + y * CJ[2][i];
      else if (j == 4) JCJ = z * CJ[0][i] + (-x) * CJ[2][i];
      else JCJ = (-y) * CJ[0][i] + x * CJ[1][i];
      H[i * 6 + j] += e * ((-gd2 * dCJ[i] * dCJ[j] + dCH) + JCJ);
    }
}

// radiusSearch over the centroid cloud == scan of the 27-cell block around the point's cell, keeping cells that are in
// the centroid cloud (>= min_points at build time, INCLUDING leaves invalidated later) with float squared distance < r^2.
// Results are ordered nearest first (FLANN sorted result set); ties keep scan order.
struct RadiusHits {
  int rec[27];
  float d2[27];
  int n;
};

__device__ __forceinline__ void radius_neighbours(RadiusHits& Hh, const PairDesc& P, const GridView& G, float tx, float ty, float tz, float radius,
                                                  bool sorted) {
  Hh.n = 0;
  const int cx = (int)floorf(tx / G.leaf), cy = (int)floorf(ty / G.leaf), cz = (int)floorf(tz / G.leaf);
  const float r2 = (float)((double)radius * (double)radius);
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int ix = cx + dx, iy = cy + dy, iz = cz + dz;
        if (ix < G.min_b[0] || ix > G.max_b[0] || iy < G.min_b[1] || iy > G.max_b[1] || iz < G.min_b[2] || iz > G.max_b[2]) continue;
        int v = __ldg(P.grid + ((ix - G.min_b[0]) * G.mul[0] + (iy - G.min_b[1]) * G.mul[1] + (iz - G.min_b[2]) * G.mul[2]));
        if (v == -1) continue;
        const int ri = grid_decode_any(v);
        const float4 c = __ldg(P.centroids + ri);
        if (c.w == 0.0f) continue;
        const float ex = tx - c.x, ey = ty - c.y, ez = tz - c.z;
        const float dd = (ex * ex + ey * ey) + ez * ez;
        if (dd < r2) {
          int pos = Hh.n;
          if (sorted) { while (pos > 0 && Hh.d2[pos - 1] > dd) { Hh.d2[pos] = Hh.d2[pos - 1]; Hh.rec[pos] = Hh.rec[pos - 1]; pos--; } }
          Hh.d2[pos] = dd; Hh.rec[pos] = ri;
          Hh.n++;
        }
      }
}

// KDTREE search mode of computeDerivatives: float contributions over radius neighbours.
template <bool HESS>
__device__ __noinline__ void point_kdtree(Acc<HESS>& A, const PairDesc& P, const GridView& G, const float* T, const float* R, float4 s, float gd2,
                                          double gd1, bool pca, bool ground, float radius) {
  float tx, ty, tz;
  transform_point(T, s.x, s.y, s.z, tx, ty, tz);
  if (!(isfinite(tx) && isfinite(ty) && isfinite(tz))) return;
  RadiusHits Hh;
  radius_neighbours(Hh, P, G, tx, ty, tz, radius, true);
  if (Hh.n == 0) return;
  // pclomp_ground: the point counts only when the LAST neighbour's normal is near the z axis (ndt_ground_impl.hpp:484,511,533)
  if (ground && !(P.recs[Hh.rec[Hh.n - 1]].meta & kMetaHorizBit)) return;
  const float xr = (R[0] * s.x + R[1] * s.y) + R[2] * s.z;
  const float yr = (R[3] * s.x + R[4] * s.y) + R[5] * s.z;
  const float zr = (R[6] * s.x + R[7] * s.y) + R[8] * s.z;
  double wsuf[27];
  double run = 1.0;
  for (int k = Hh.n - 1; k >= 0; k--) {
    if (pca) run *= (double)(P.recs[Hh.rec[k]].meta & kMetaWeightMask);
    wsuf[k] = run;
  }
  for (int k = 0; k < Hh.n; k++) {
    const VoxelRec* vr = P.recs + Hh.rec[k];
    float C[9];
    for (int a = 0; a < 9; a++) C[a] = vr->icov[icov_slot(a)];
    const float d0 = (float)((double)tx - vr->mean[0]), d1 = (float)((double)ty - vr->mean[1]), d2 = (float)((double)tz - vr->mean[2]);
    contribute<HESS>(A, xr, yr, zr, d0, d1, d2, C, gd2, gd1, wsuf[k]);
  }
}

__device__ __noinline__ void point_hess27(double* H, const PairDesc& P, const GridView& G, const float* T, const double* Rd, float4 s, double gd1,
                                          double gd2, float radius) {
  float tx, ty, tz;
  transform_point(T, s.x, s.y, s.z, tx, ty, tz);
  if (!(isfinite(tx) && isfinite(ty) && isfinite(tz))) return;
  RadiusHits Hh;
  radius_neighbours(Hh, P, G, tx, ty, tz, radius, false);
  if (Hh.n == 0) return;
  const double x[3] = {(double)s.x, (double)s.y, (double)s.z};
  double r[3];
  for (int i = 0; i < 3; i++) r[i] = (Rd[i * 3] * x[0] + Rd[i * 3 + 1] * x[1]) + Rd[i * 3 + 2] * x[2];
  for (int k = 0; k < Hh.n; k++) {
    const VoxelRec* vr = P.recs + Hh.rec[k];
    const double* C = P.icov64 + (size_t)Hh.rec[k] * 9;
    double Cl[9];
    for (int a = 0; a < 9; a++) Cl[a] = C[a];
    const double d[3] = {(double)tx - vr->mean[0], (double)ty - vr->mean[1], (double)tz - vr->mean[2]};
    contribute_hess64(H, r, d, Cl, gd1, gd2);
  }
}

// CTA-level reduction of NV doubles per thread into partial[NV] (fixed shape => deterministic).
template <int NV>
__device__ __forceinline__ void block_reduce_store(const double* v, double* s_red /*[8][NV]*/, double* partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[warp * NV + k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0;
#pragma unroll
    for (int w = 0; w < kEvalThreads / 32; w++) x += s_red[w * NV + threadIdx.x];
    partial[threadIdx.x] = x;
  }
}

template <bool HESS>
__device__ __noinline__ void run_kdtree(const PairDesc& P, const GridView& G, const float* T, const float* R, int blk, int bpp, float gd2, double gd1,
                                        bool pca, bool ground, float radius, double* s_red, double* partial) {
  Acc<HESS> A;
  A.zero();
  if (!G.empty)
    for (int i = blk * kEvalThreads + threadIdx.x; i < P.n_src; i += bpp * kEvalThreads)
      point_kdtree<HESS>(A, P, G, T, R, __ldg(P.src + i), gd2, gd1, pca, ground, radius);
  double v[kAcc];
  v[0] = A.score;
  for (int i = 0; i < 6; i++) v[1 + i] = A.g[i];
  for (int i = 0; i < 36; i++) v[7 + i] = HESS ? A.H[i] : 0.0;
  block_reduce_store<kAcc>(v, s_red, partial);
}

__device__ __noinline__ void run_hess27(const PairDesc& P, const GridView& G, const float* T, const double* Rd, int blk, int bpp, double gd1, double gd2,
                                        float radius, double* s_red, double* partial) {
  double v[kAcc];
  for (int i = 0; i < kAcc; i++) v[i] = 0;
  if (!G.empty)
    for (int i = blk * kEvalThreads + threadIdx.x; i < P.n_src; i += bpp * kEvalThreads)
      point_hess27(v + 7, P, G, T, Rd, __ldg(P.src + i), gd1, gd2, radius);
  block_reduce_store<kAcc>(v, s_red, partial);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEvalThreads) ndt_eval_cold_kernel(EvalLaunch L) {
  __shared__ double s_red[704];            // CTA reduction (8 warps x 43) and eval_finish's scratch (11 x 64)
  static_assert((kEvalThreads / 32) * kAcc <= 704, "s_red");
  __shared__ float s_T[16], s_R[9];
  __shared__ double s_Rd[9];
  __shared__ int s_last;
  pdl_wait();
  const int pair = blockIdx.x / L.blocks_per_pair, blk = blockIdx.x % L.blocks_per_pair;
  AlignState& S = L.d_states[pair];
  const int kind = S.eval_kind;
  const AlignConsts& c = L.consts;
  if (kind == EVAL_NONE) return;
  if (kind != EVAL_HESS27 && c.search != LVS_KDTREE) return;     // direct-search derivative passes belong to the hot kernel
  const PairDesc P = L.d_pairs[pair];
  if (threadIdx.x < 16) s_T[threadIdx.x] = S.T[threadIdx.x];
  if (threadIdx.x < 9) { s_R[threadIdx.x] = S.Rj[threadIdx.x]; s_Rd[threadIdx.x] = S.Rd[threadIdx.x]; }
  __syncthreads();
  const GridView G = load_grid_view(P.gp);
  const float gd2 = (float)c.gauss_d2;
  const bool pca = c.variant == LVS_NDT_PCA, ground = c.variant == LVS_NDT_GROUND;
  double* partial = L.d_partials + ((size_t)pair * L.blocks_per_pair + blk) * kPartialStride;
  const int bpp = L.blocks_per_pair;
  if (kind == EVAL_HESS27) run_hess27(P, G, s_T, s_Rd, blk, bpp, c.gauss_d1, c.gauss_d2, c.resolution, s_red, partial);
  else if (kind == EVAL_DERIV_H) run_kdtree<true>(P, G, s_T, s_R, blk, bpp, gd2, c.gauss_d1, pca, ground, c.resolution, s_red, partial);
  else run_kdtree<false>(P, G, s_T, s_R, blk, bpp, gd2, c.gauss_d1, pca, ground, c.resolution, s_red, partial);
  pdl_trigger();
  eval_finish(L, pair, kind, kAcc, P.n_total, s_red, &s_last);
}

int launch_eval_cold(cudaStream_t st, const EvalLaunch& L) {
  if (L.n_pairs <= 0) return LVS_OK;
  return launch_pdl(ndt_eval_cold_kernel, (unsigned)(L.n_pairs * L.blocks_per_pair), kEvalThreads, 0, st, L);
}

// ---------------------------------------------------------------------------------------------------------------
// calculateScore (:1007-1040): mean over points of sum over radius neighbours of (-d1*exp(-d2 q/2) - d3)/|nbrs|, all fp64.
__global__ void __launch_bounds__(kEvalThreads) calc_score_kernel(PairDesc P, const float* __restrict__ T16, AlignConsts c, double* partials,
                                                                    unsigned int* ticket, double* out) {
  __shared__ double s_red[(kEvalThreads / 32)];
  __shared__ float s_T[16];
  __shared__ int s_last;
  if (threadIdx.x < 16) s_T[threadIdx.x] = T16[threadIdx.x];
  __syncthreads();
  const GridView G = load_grid_view(P.gp);
  double acc = 0;
  if (!G.empty)
    for (int i = blockIdx.x * kEvalThreads + threadIdx.x; i < P.n_src; i += gridDim.x * kEvalThreads) {
      const float4 s = __ldg(P.src + i);
      float tx, ty, tz;
      transform_point(s_T, s.x, s.y, s.z, tx, ty, tz);
      if (!(isfinite(tx) && isfinite(ty) && isfinite(tz))) continue;
      RadiusHits Hh;
      radius_neighbours(Hh, P, G, tx, ty, tz, c.resolution, true);
      for (int k = 0; k < Hh.n; k++) {
        const VoxelRec* vr = P.recs + Hh.rec[k];
        const double* C = P.icov64 + (size_t)Hh.rec[k] * 9;
        const double d[3] = {(double)tx - vr->mean[0], (double)ty - vr->mean[1], (double)tz - vr->mean[2]};
        double Cd[3];
        for (int a = 0; a < 3; a++) Cd[a] = (C[a * 3] * d[0] + C[a * 3 + 1] * d[1]) + C[a * 3 + 2] * d[2];
        const double e = exp(-c.gauss_d2 * ((d[0] * Cd[0] + d[1] * Cd[1]) + d[2] * Cd[2]) / 2);
        const double inc = -c.gauss_d1 * e - c.gauss_d3;
        acc += inc / (double)Hh.n;
      }
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0;
    for (int w = 0; w < kEvalThreads / 32; w++) x += s_red[w];
    partials[blockIdx.x] = x;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  double x = 0;
  for (unsigned b = 0; b < gridDim.x; b++) x += ((volatile double*)partials)[b];
  *out = x / (double)P.n_src;
  *ticket = 0;
}

int launch_calc_score(cudaStream_t st, const PairDesc& pair, const float* d_T16, const AlignConsts& c, double* d_partials, int max_blocks,
                      unsigned int* d_ticket, double* d_out) {
  int nb = std::max(1, std::min(max_blocks, (pair.n_src + kEvalThreads - 1) / kEvalThreads));
  calc_score_kernel<<<nb, kEvalThreads, 0, st>>>(pair, d_T16, c, d_partials, d_ticket, d_out);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Parity tap: the voxel key the lookup path computes for T * source[i] (-1 outside the bounding box).
__global__ void lookup_keys_kernel(PairDesc P, const float* __restrict__ T16, int* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_src) return;
  const GridView G = load_grid_view(P.gp);
  const float4 s = P.src[i];
  float tx, ty, tz;
  transform_point(T16, s.x, s.y, s.z, tx, ty, tz);
  int key = -1;
  if (!G.empty && isfinite(tx) && isfinite(ty) && isfinite(tz)) {
    const int ix = (int)floorf(tx / G.leaf), iy = (int)floorf(ty / G.leaf), iz = (int)floorf(tz / G.leaf);
    if (ix >= G.min_b[0] && ix <= G.max_b[0] && iy >= G.min_b[1] && iy <= G.max_b[1] && iz >= G.min_b[2] && iz <= G.max_b[2])
      key = (ix - G.min_b[0]) * G.mul[0] + (iy - G.min_b[1]) * G.mul[1] + (iz - G.min_b[2]) * G.mul[2];
  }
  keys[i] = key;
}

int launch_lookup_keys(cudaStream_t st, const PairDesc& pair, const float* d_T16, int* d_keys_out) {
  if (pair.n_src == 0) return LVS_OK;
  lookup_keys_kernel<<<(pair.n_src + 255) / 256, 256, 0, st>>>(pair, d_T16, d_keys_out);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

// The `output` cloud of align(): T * source, packed xyz.
__global__ void transform_kernel(const float4* __restrict__ src, int n, const float* __restrict__ T16, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 s = src[i];
  float tx, ty, tz;
  transform_point(T16, s.x, s.y, s.z, tx, ty, tz);
  out[3 * i] = tx; out[3 * i + 1] = ty; out[3 * i + 2] = tz;
}

int launch_transform(cudaStream_t st, const float4* d_src, int n, const float* d_T16, float* d_out_xyz) {
  if (n == 0) return LVS_OK;
  transform_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_src, n, d_T16, d_out_xyz);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

}  // namespace lvs
