// NDT derivative pass with direct voxel lookup — the hot kernel.  Replaces computeDerivatives / updateDerivatives /
// computePointDerivatives_AngleAxisd (include/ndt_omp/ndt_omp_impl2.hpp:197-305, 504-532, 567-619; pca weight
// include/ndt_pca/ndt_pca_impl2.hpp:293-296), getNeighborhoodAtPoint{,7,1}
// (include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:373-442) and pcl::transformPointCloud (:107,903,949).
//
// One launch = one evaluation of every active (source, target) pair; grid = n_pairs x blocks_per_pair CTAs.
//   * a lane owns a source point: coalesced 16 B load, float transform (the reference's operation order, this TU is
//     compiled with -fmad=false), voxel index, K probes of the dense int32 index grid, 64 B record fetch (4 x LDG.128);
//   * the float32 contribution of a (point, cell) pair — score, 6 gradient terms, all 36 Hessian terms (the reference's H is
//     not symmetric) — is computed exactly as the CPU path rounds it and staged in a per-warp shared-memory tile [43][32];
//   * the fp64 accumulation the reference does per thread is done by the warp as a whole: the 43 x 4 (output, 8-lane group)
//     sums are spread over the 32 lanes, so a lane keeps 6 fp64 accumulators instead of 43 — that is what lets three CTAs
//     of 256 threads live on an SM instead of one;
//   * CTA partial -> ticket -> the last CTA of the pair adds the partials in fixed order and advances the pair's
//     Newton / More-Thuente state machine (eval_finish, ndt_state.cuh).
#include "ndt_eval_common.cuh"

namespace lvs {

// ---- packed FP32 (sm_100 FMUL2 / FFMA2): one issue slot carries two IEEE float operations, each rounded exactly like the scalar
// instruction, so every term stays bit-identical to the CPU path while the float math of a contribution needs about half the
// issue slots.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false (seen with CUDA 12.9), which
// would change the rounding; the packed add is therefore written as fma(x, ONE, y) with ONE = 1.0f taken from a kernel
// parameter — x * 1.0f is exact, the instruction cannot absorb a preceding multiply, and it costs the same FFMA2.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi_of(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 c; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ f32x2 mul2s(float a, f32x2 b) { return mul2(pk(a, a), b); }                 // scalar broadcast operand
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b, f32x2 one) { f32x2 c; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(one), "l"(b)); return c; }

// Staging tile of one warp: 22 rows of output PAIRS, each row holds the float2 of the 32 lanes (+ 4 floats of padding so that the
// LDS.128 reads of the summing lanes fall on distinct bank groups).  Pair p of a contribution (see kSlotOut for the outputs):
//   p = 0        (score, unused)
//   p = 1..3     (g[2k], g[2k+1]),            k = p - 1
//   p = 4..21    (H[2k][j], H[2k+1][j]),      j = (p - 4) / 3, k = (p - 4) % 3
constexpr int kPairsH = 22, kPairsNoH = 4;
constexpr int kTileStride = 68;            // floats per pair row: 32 lanes x 2 + 4 pad (stride / 4 odd)
constexpr int kWarps = kEvalThreads / 32;

// slot (pair * 2 + half) -> index in the canonical output vector (0 score, 1..6 gradient, 7 + 6 i + j Hessian), -1 = padding
__device__ __forceinline__ int slot_to_out(int slot) {
  const int p = slot >> 1, h = slot & 1;
  if (p == 0) return h ? -1 : 0;
  if (p < 4) return 1 + 2 * (p - 1) + h;
  const int j = (p - 4) / 3, k = (p - 4) % 3;
  return 7 + 6 * (2 * k + h) + j;
}

template <bool HESS>
struct Shape {
  static constexpr int NV = HESS ? kAcc : 7;                       // outputs per contribution
  static constexpr int NP = HESS ? kPairsH : kPairsNoH;            // output pairs per contribution
  static constexpr int TPL = (NP + 7) / 8;                         // summation passes: a pass covers 8 pairs x 4 lane groups
  static constexpr int NSLOT = NP * 2;
};

// updateDerivatives + computePointDerivatives_AngleAxisd for one (point, cell) pair, written into column `lane` of the
// staging tile.  xr = R*x (float), d = x' - mean, Cp/C2 = float inverse covariance.  The zero / identity entries of the
// reference's 4x6 and 24x6 matrices are folded away by hand; every remaining product and sum keeps the reference's (Eigen SSE)
// order and rounding, so each float term is bit-identical to the CPU path.  Rows of the 6x6 Hessian are processed two at a
// time in packed registers.  Returns false on the reference's early-out (d2*e > 1, < 0 or NaN, :588-589): nothing is written then.
//   Cp[r] = (C[r][0], C[r][1]),  C2[r] = C[r][2]
template <bool HESS>
__device__ __forceinline__ bool contribute_tile(float2* __restrict__ col /* tile + lane */, float xr, float yr, float zr, float d0, float d1,
                                                float d2, const f32x2* Cp, const float* C2, float gd2, double gauss_d1, f32x2 one, const unsigned long long* etab) {
  constexpr int RS = kTileStride / 2;      // row stride in float2
  const float nx = -xr, ny = -yr, nz = -zr;
  // P[r][k], k = 0..2: row r of [C | C * dR-columns] = (CJ[r][2k], CJ[r][2k+1])
  f32x2 P[3][3];
  const f32x2 ZN = pk(zr, ny), NX = pk(nx, xr);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const float c0 = lo_of(Cp[r]), c1 = hi_of(Cp[r]), c2 = C2[r];
    P[r][0] = Cp[r];
    P[r][1] = pk(c2, c1 * nz + c2 * yr);                                        // CJ[r][3]
    P[r][2] = add2(mul2s(c0, ZN), mul2(pk(c2, c1), NX), one);                   // CJ[r][4] = c0*zr + c2*nx, CJ[r][5] = c0*ny + c1*xr
  }
  // a[j] = (d0*CJ[0][j] + d2*CJ[2][j]) + d1*CJ[1][j];  a[0..2] = d^T C
  f32x2 A[3];
#pragma unroll
  for (int k = 0; k < 3; k++) A[k] = add2(add2(mul2s(d0, P[0][k]), mul2s(d2, P[2][k]), one), mul2s(d1, P[1][k]), one);
  const float xC0 = lo_of(A[0]), xC1 = hi_of(A[0]), xC2 = lo_of(A[1]);
  const float q = (d0 * xC0 + d2 * xC2) + d1 * xC1;
  float e = glibc_expf((-gd2 * q) * 0.5f, etab);      // the reference's exp(float) is expf (lvs_math.cuh)
  const float score_inc = (float)(-gauss_d1 * (double)e);
  e = gd2 * e;
  if (e > kOne || e < 0.0f || e != e) return false;
  e = (float)((double)e * gauss_d1);

  col[0] = make_float2(score_inc, 0.0f);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const f32x2 g = mul2s(e, A[k]);
    col[(1 + k) * RS] = make_float2(lo_of(g), hi_of(g));
  }

  if (HESS) {
    const float a[6] = {xC0, xC1, xC2, hi_of(A[1]), lo_of(A[2]), hi_of(A[2])};
    // hp[i][j] = (d^T C) . Hp_ij  (non-zero only in the rotation block); rows 1 and 2 are kept packed
    const float hp0[3] = {xC2 * nz + xC1 * ny, xC1 * xr, xC2 * xr};
    f32x2 HP[3];
    HP[0] = mul2s(xC0, pk(yr, zr));                                             // (hp[1][0], hp[2][0])
    HP[1] = pk(xC0 * nx + xC2 * nz, xC1 * zr);                                  // (hp[1][1], hp[2][1])
    HP[2] = pk(xC2 * yr, xC0 * nx + xC1 * ny);                                  // (hp[1][2], hp[2][2])
    const float ngd2 = -gd2;
    f32x2 AI[3];
#pragma unroll
    for (int k = 0; k < 3; k++) AI[k] = mul2s(ngd2, A[k]);
#pragma unroll
    for (int j = 0; j < 6; j++) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        // rows i = 2k, 2k+1:  t = (-d2 a_i) a_j  [+ hp[i-3][j-3]],  m = J.col(j) . CJ.col(i),  H_ij = e (t + m)
        f32x2 t = mul2s(a[j], AI[k]);
        if (j >= 3) {
          if (k == 1) t = pk(lo_of(t), hi_of(t) + hp0[j - 3]);
          if (k == 2) t = add2(t, HP[j - 3], one);
        }
        f32x2 m;
        if (j < 3) m = P[j][k];
        else if (j == 3) m = add2(mul2s(yr, P[2][k]), mul2s(nz, P[1][k]), one);
        else if (j == 4) m = add2(mul2s(zr, P[0][k]), mul2s(nx, P[2][k]), one);
        else m = add2(mul2s(ny, P[0][k]), mul2s(xr, P[1][k]), one);
        const f32x2 o = mul2s(e, add2(t, m, one));
        col[(4 + j * 3 + k) * RS] = make_float2(lo_of(o), hi_of(o));
      }
    }
  }
  return true;
}

// Per-warp staging in shared memory (dynamic, carved up in SmemLayout):
//   tile [22][68] f32   float contributions of one round of 32 (point, cell) pairs, as output pairs
//   pts  [6][64]  f32   x', y', z' (transformed point) and R*x of the 64 points of the current warp iteration
//   q    [256]    i32   queue of hits: record index * 64 + point slot
//   qw   [256]    f64   ndt_pca only: weight that multiplies the entry's contribution
constexpr int kPtsPerLane = 2;
constexpr int kPtsPerIter = 32 * kPtsPerLane;
constexpr int kQueueCap = 256;
constexpr int kTileFloats = kPairsH * kTileStride;

// Everything a warp stages lives in ONE contiguous block of shared memory, so tile, points and queue are constant offsets from a
// single per-warp base register (the compiler otherwise rebuilds three warp-indexed pointers in every round).
template <bool PCA>
struct SmemLayout {
  static constexpr size_t tile_off = 0;
  static constexpr size_t pts_off = tile_off + sizeof(float) * kTileFloats;
  static constexpr size_t q_off = pts_off + sizeof(float) * 6 * kPtsPerIter;
  static constexpr size_t qw_off = q_off + sizeof(int) * kQueueCap;
  static constexpr size_t warp_bytes = qw_off + (PCA ? sizeof(double) * kQueueCap : 0);
  static constexpr size_t bytes = warp_bytes * kWarps;
};
static_assert(SmemLayout<false>::qw_off % 8 == 0, "qw must be 8-byte aligned");
static_assert(SmemLayout<false>::warp_bytes % 16 == 0 && SmemLayout<true>::warp_bytes % 16 == 0, "per-warp blocks must stay 16-byte aligned (LDS.128 on the tile)");
static_assert(sizeof(double) * kWarps * kPairsH * 2 * 4 <= SmemLayout<false>::bytes, "the CTA reduction scratch aliases the staging blocks");

// One round: lanes e < n_round take queue entries q[head + lane], stage their float contributions in the tile, then the
// warp adds the round into its fp64 accumulators: in pass t lane l owns pair 8t + (l & 7) and the lane group l >> 3 (8 lanes),
// i.e. two accumulators (the two outputs of the pair), fed by four LDS.128.  Columns of lanes without a contribution (short
// round, or the reference's early-out) are zero-filled and summed like the others: adding +0.0 is exact and an accumulator that
// starts at +0.0 never becomes -0.0, so no per-group predicate is needed in the summation.
template <bool HESS, bool PCA>
__device__ __forceinline__ void process_round(double (*acc)[2], const VoxelRec* __restrict__ recs, const float* pts, const int* q, const double* qw,
                                              float* tile, int lane, int head, int n_round, float gd2, double gd1, f32x2 one, const unsigned long long* etab) {
  using Sh = Shape<HESS>;
  bool used = false;
  double w = 0.0;
  float2* col = reinterpret_cast<float2*>(tile) + lane;
  if (lane < n_round) {
    const int ent = q[head + lane];
    const int rec = ent / kPtsPerIter, slot = ent % kPtsPerIter;
    const VoxelRec* vr = recs + rec;
    const double2 m01 = __ldg(reinterpret_cast<const double2*>(vr));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(vr) + 1);   // mean[2] (8 B) + C00 C01
    const float4 q2 = __ldg(reinterpret_cast<const float4*>(vr) + 2);   // C10 C11 C20 C21
    const float4 q3 = __ldg(reinterpret_cast<const float4*>(vr) + 3);   // C02 C12 C22 + meta
    const double m2 = __hiloint2double(__float_as_int(q1.y), __float_as_int(q1.x));
    const f32x2 Cp[3] = {pk(q1.z, q1.w), pk(q2.x, q2.y), pk(q2.z, q2.w)};
    const float C2[3] = {q3.x, q3.y, q3.z};
    const float tx = pts[0 * kPtsPerIter + slot], ty = pts[1 * kPtsPerIter + slot], tz = pts[2 * kPtsPerIter + slot];
    const float d0 = (float)((double)tx - m01.x), d1 = (float)((double)ty - m01.y), d2 = (float)((double)tz - m2);
    used = contribute_tile<HESS>(col, pts[3 * kPtsPerIter + slot], pts[4 * kPtsPerIter + slot], pts[5 * kPtsPerIter + slot], d0, d1, d2, Cp, C2,
                                 gd2, gd1, one, etab);
    if (PCA) w = qw[head + lane];
  }
  if (!used) {
#pragma unroll
    for (int p = 0; p < Sh::NP; p++) col[p * (kTileStride / 2)] = make_float2(0.0f, 0.0f);
  }
  __syncwarp();
  const int g = lane >> 3;
#pragma unroll
  for (int t = 0; t < Sh::TPL; t++) {
    const int p = 8 * t + (lane & 7);
    const bool live = (Sh::NP % 8 == 0) || t + 1 < Sh::TPL || p < Sh::NP;
    // dead lanes of the last pass read a row whose bank group no live lane of their quarter-warp uses (rows 8 apart share banks)
    const float4* row = reinterpret_cast<const float4*>(tile + (live ? p : p - (HESS ? 16 : 4)) * kTileStride + 16 * g);
    double a0 = acc[t][0], a1 = acc[t][1];
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const float4 v = row[jj];          // lanes 8g + 2jj and 8g + 2jj + 1
      if (PCA) {
        const double w0 = __shfl_sync(0xffffffffu, w, 8 * g + 2 * jj), w1 = __shfl_sync(0xffffffffu, w, 8 * g + 2 * jj + 1);
        a0 += (double)v.x * w0; a1 += (double)v.y * w0;      // zero columns carry w = 0 or a finite weight: exact either way
        a0 += (double)v.z * w1; a1 += (double)v.w * w1;
      } else {
        a0 += (double)v.x; a1 += (double)v.y;
        a0 += (double)v.z; a1 += (double)v.w;
      }
    }
    acc[t][0] = a0; acc[t][1] = a1;      // lanes past NP in the last pass sum a live row into accumulators that are never read
  }
  __syncwarp();
}

template <int MODE, bool HESS, bool PCA, bool GROUND>
__device__ __forceinline__ void run_direct(const PairDesc& P, const GridView& G, const float* T, const float* R, int blk, int bpp, float gd2,
                                           double gd1, unsigned char* s_dyn, double* partial, float one_f, const unsigned long long* etab) {
  using Sh = Shape<HESS>;
  constexpr int K = Probes<MODE>::K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  using SL = SmemLayout<PCA>;
  unsigned char* wbase = s_dyn + (size_t)warp * SL::warp_bytes;
  float* tile = reinterpret_cast<float*>(wbase + SL::tile_off);
  float* pts = reinterpret_cast<float*>(wbase + SL::pts_off);
  int* q = reinterpret_cast<int*>(wbase + SL::q_off);
  double* qw = reinterpret_cast<double*>(wbase + SL::qw_off);   // only mapped for ndt_pca launches
  const VoxelRec* __restrict__ recs = P.recs;
  const int* __restrict__ grid = P.grid;
  const f32x2 one = pk(one_f, one_f);
  double acc[Sh::TPL][2];
#pragma unroll
  for (int t = 0; t < Sh::TPL; t++) acc[t][0] = acc[t][1] = 0.0;
  // floor(x / leaf) in float (voxel_grid_covariance_omp_impl.hpp:379-381).  For a power-of-two leaf the quotient equals the
  // product with the (exact) reciprocal bit for bit, so the division is only issued for other leaf sizes.
  const float inv_leaf = 1.0f / G.leaf;
  const bool pow2 = (__float_as_uint(G.leaf) & 0x007fffffu) == 0u && isfinite(inv_leaf) && inv_leaf >= 1.1754944e-38f;
  const int div0 = G.max_b[0] - G.min_b[0] + 1, div1 = G.max_b[1] - G.min_b[1] + 1, div2 = G.max_b[2] - G.min_b[2] + 1;

  if (!G.empty) {
    // every warp of the pair's CTAs owns one contiguous range of the source, so that any number of CTAs splits the cloud evenly
    // (a single pair is spread over every SM: 35 points per warp at 125 k points instead of 64 on half of the SMs)
    const long long n_warps = (long long)bpp * kWarps, wid = (long long)blk * kWarps + warp;
    const int lo = (int)(wid * P.n_src / n_warps), hi = (int)((wid + 1) * P.n_src / n_warps);
    for (int base = lo; base < hi; base += kPtsPerIter) {
      // ---- phase A: transform, probe the index grid, queue every (point, cell) hit of the warp's 64 points
      int nq = 0;
      // consumes whole rounds from the queue and moves the remainder (< 32 entries) to its front; returns the new length
      auto drain = [&]() {
        __syncwarp();
        int head = 0;
        for (; nq - head >= 32; head += 32) process_round<HESS, PCA>(acc, recs, pts, q, qw, tile, lane, head, 32, gd2, gd1, one, etab);
        const int rem = nq - head;
        int ent = 0; double we = 0.0;
        if (lane < rem) { ent = q[head + lane]; if (PCA) we = qw[head + lane]; }
        __syncwarp();
        if (lane < rem) { q[lane] = ent; if (PCA) qw[lane] = we; }
        return rem;
      };
      for (int h = 0; h < kPtsPerLane; h++) {
        // one half of the warp's points adds at most 32 K entries: with K = 7 that fits behind a remainder of < 32 entries,
        // so the queue is only checked here; DIRECT26 checks before every probe
        if (MODE == LVS_DIRECT7 && h > 0 && nq >= 32) nq = drain();
        const int slot = h * 32 + lane;
        const int i = base + slot;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        bool ok = i < hi;
        if (ok) {
          const float4 s = __ldg(P.src + i);
          transform_point(T, s.x, s.y, s.z, tx, ty, tz);
          ok = isfinite(tx) && isfinite(ty) && isfinite(tz);
          pts[0 * kPtsPerIter + slot] = tx; pts[1 * kPtsPerIter + slot] = ty; pts[2 * kPtsPerIter + slot] = tz;
          // x_t = float(SE3::exp(p).matrix()) * [x, 0]: rotation only (ndt_omp_impl2.hpp:507-508)
          pts[3 * kPtsPerIter + slot] = (R[0] * s.x + R[1] * s.y) + R[2] * s.z;
          pts[4 * kPtsPerIter + slot] = (R[3] * s.x + R[4] * s.y) + R[5] * s.z;
          pts[5 * kPtsPerIter + slot] = (R[6] * s.x + R[7] * s.y) + R[8] * s.z;
        }
        int cx, cy, cz;
        if (pow2) { cx = (int)floorf(tx * inv_leaf); cy = (int)floorf(ty * inv_leaf); cz = (int)floorf(tz * inv_leaf); }
        else { cx = (int)floorf(tx / G.leaf); cy = (int)floorf(ty / G.leaf); cz = (int)floorf(tz / G.leaf); }
        // cell coordinates relative to the grid origin; one unsigned compare per axis and offset tests the bounds
        const int rx = cx - G.min_b[0], ry = cy - G.min_b[1], rz = cz - G.min_b[2];
        const int cell0 = rx * G.mul[0] + ry * G.mul[1] + rz * G.mul[2];
        // ndt_pca multiplies the RUNNING per-point sums by each cell's weight (ndt_pca_impl2.hpp:293-296): the contribution of
        // cell k ends up scaled by the product of the weights of cells k..last, hence the probes run last-to-first.
        double run = 1.0;
        // pclomp_ground counts a point only when the LAST cell of its neighbourhood is near-horizontal (ndt_ground_impl.hpp:484,511,533):
        // the probes run last-to-first, so the first hit decides for all of the point's cells
        bool gate_known = false, gate = true;
#pragma unroll(MODE == LVS_DIRECT26 ? 1 : K)
        for (int k = K - 1; k >= 0; k--) {
          int ox = 0, oy = 0, oz = 0;
          if (MODE == LVS_DIRECT7) { ox = k == 1 ? 1 : k == 2 ? -1 : 0; oy = k == 3 ? 1 : k == 4 ? -1 : 0; oz = k == 5 ? 1 : k == 6 ? -1 : 0; }
          else if (MODE == LVS_DIRECT26) { ox = c_off26[k][0]; oy = c_off26[k][1]; oz = c_off26[k][2]; }
          int v = -1;
          if (ok && (unsigned)(rx + ox) < (unsigned)div0 && (unsigned)(ry + oy) < (unsigned)div1 && (unsigned)(rz + oz) < (unsigned)div2)
            v = __ldg(grid + (cell0 + ox * G.mul[0] + oy * G.mul[1] + oz * G.mul[2]));
          bool hit = v >= 0;
          if (GROUND && hit) {
            if (!gate_known) { gate = (__ldg(&recs[v].meta) & kMetaHorizBit) != 0; gate_known = true; }
            hit = gate;
          }
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (MODE == LVS_DIRECT26) {
            if (m == 0u) continue;
            if (nq > kQueueCap - 32) {
              // queue nearly full (dense DIRECT26 neighbourhoods): drain whole rounds, keep the remainder at the front
              nq = drain();
            }
          }
          if (hit) {
            const int pos = nq + __popc(m & lt_mask);
            q[pos] = v * kPtsPerIter + slot;
            if (PCA) { run *= (double)(__ldg(&recs[v].meta) & kMetaWeightMask); qw[pos] = run; }
          }
          nq += __popc(m);
        }
      }
      static_assert(MODE == LVS_DIRECT26 || 31 + 32 * Probes<MODE>::K <= kQueueCap, "queue capacity per half iteration");
      static_assert(MODE != LVS_DIRECT1 || kPtsPerLane * 32 <= kQueueCap, "queue capacity");
      __syncwarp();
      // ---- phase B: rounds of 32 queued (point, cell) contributions
      for (int head = 0; head < nq; head += 32) process_round<HESS, PCA>(acc, recs, pts, q, qw, tile, lane, head, min(32, nq - head), gd2, gd1, one, etab);
    }
  }
  // CTA partial: s_red[warp][slot][group] (aliases the tile region) -> every output slot sums its 4 lane groups over the 8 warps in
  // fixed order and lands at its place in the canonical output vector
  __syncthreads();
  double* s_red = reinterpret_cast<double*>(s_dyn);
  {
    const int g = lane >> 3;
#pragma unroll
    for (int t = 0; t < Sh::TPL; t++) {
      const int p = 8 * t + (lane & 7);
      if (p < Sh::NP) {
        s_red[(warp * Sh::NSLOT + 2 * p) * 4 + g] = acc[t][0];
        s_red[(warp * Sh::NSLOT + 2 * p + 1) * 4 + g] = acc[t][1];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < Sh::NSLOT) {
    const int out = slot_to_out(threadIdx.x);
    if (out >= 0) {
      double x = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++)
#pragma unroll
        for (int g = 0; g < 4; g++) x += s_red[(w * Sh::NSLOT + threadIdx.x) * 4 + g];
      partial[out] = x;
    }
  }
}

template <int MODE, bool PCA, bool GROUND>
__global__ void __launch_bounds__(kEvalThreads, 3) ndt_eval_kernel(EvalLaunch L) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ float s_T[16], s_R[9];
  __shared__ unsigned long long s_etab[32];
  __shared__ int s_last;
  const long long t_entry = clock64();
  pdl_wait();
  const int pair = blockIdx.x / L.blocks_per_pair, blk = blockIdx.x % L.blocks_per_pair;
  AlignState& S = L.d_states[pair];
  const int kind = S.eval_kind;
  const AlignConsts& c = L.consts;
  if (kind != EVAL_DERIV_H && kind != EVAL_DERIV_NOH) return;
  const PairDesc P = L.d_pairs[pair];
  if (threadIdx.x < 16) s_T[threadIdx.x] = S.T[threadIdx.x];
  if (threadIdx.x < 9) s_R[threadIdx.x] = S.Rj[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 64) s_etab[threadIdx.x - 32] = c_exp2f_tab[threadIdx.x - 32];
  __syncthreads();
  const GridView G = load_grid_view(P.gp);
  const float gd2 = (float)c.gauss_d2;
  double* partial = L.d_partials + ((size_t)pair * L.blocks_per_pair + blk) * kPartialStride;
  const int bpp = L.blocks_per_pair;
  if (kind == EVAL_DERIV_H) run_direct<MODE, true, PCA, GROUND>(P, G, s_T, s_R, blk, bpp, gd2, c.gauss_d1, s_dyn, partial, L.one, s_etab);
  else run_direct<MODE, false, PCA, GROUND>(P, G, s_T, s_R, blk, bpp, gd2, c.gauss_d1, s_dyn, partial, L.one, s_etab);
  pdl_trigger();
  eval_finish(L, pair, kind, kind == EVAL_DERIV_H ? kAcc : 7, P.n_total, reinterpret_cast<double*>(s_dyn), &s_last, t_entry);
}

template <int MODE, bool PCA, bool GROUND = false>
static int launch_eval_as(cudaStream_t st, const EvalLaunch& L) {
  static int attr_dev = -1;     // the opt-in shared-memory size is a per-device function attribute
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  constexpr size_t smem = SmemLayout<PCA>::bytes;
  if (attr_dev != dev) {
    CUDA_TRY(cudaFuncSetAttribute(ndt_eval_kernel<MODE, PCA, GROUND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev = dev;
  }
  return launch_pdl(ndt_eval_kernel<MODE, PCA, GROUND>, (unsigned)(L.n_pairs * L.blocks_per_pair), kEvalThreads, smem, st, L);
}

// The search mode and the registration variant are launch constants, so each combination is its own kernel (unrolled probe
// loop, no weight staging for ndt_omp); whether the Hessian is wanted is per-pair state and is decided inside.
int launch_eval(cudaStream_t st, const EvalLaunch& L) {
  if (L.n_pairs <= 0) return LVS_OK;
  if (L.consts.fast) return launch_eval_fast(st, L);       // tolerance mode: ndt_eval_fast.cu
  const bool pca = L.consts.variant == LVS_NDT_PCA;
  if (L.consts.variant == LVS_NDT_GROUND) {                 // pclomp_ground: the ndt_omp pass with the last-neighbour gate
    switch (L.consts.search) {
      case LVS_DIRECT1: return launch_eval_as<LVS_DIRECT1, false, true>(st, L);
      case LVS_DIRECT7: return launch_eval_as<LVS_DIRECT7, false, true>(st, L);
      case LVS_DIRECT26: return launch_eval_as<LVS_DIRECT26, false, true>(st, L);
      default: return LVS_OK;
    }
  }
  switch (L.consts.search) {
    case LVS_DIRECT1: return pca ? launch_eval_as<LVS_DIRECT1, true>(st, L) : launch_eval_as<LVS_DIRECT1, false>(st, L);
    case LVS_DIRECT7: return pca ? launch_eval_as<LVS_DIRECT7, true>(st, L) : launch_eval_as<LVS_DIRECT7, false>(st, L);
    case LVS_DIRECT26: return pca ? launch_eval_as<LVS_DIRECT26, true>(st, L) : launch_eval_as<LVS_DIRECT26, false>(st, L);
    default: return LVS_OK;                                  // KDTREE: radius-search derivatives live in the cold kernel
  }
}

int eval_max_resident_ctas_per_sm() { return 3; }
int eval_points_per_cta_iteration() { return kWarps * kPtsPerIter; }

}  // namespace lvs
