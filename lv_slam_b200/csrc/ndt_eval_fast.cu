// Tolerance-mode NDT derivative pass with direct voxel lookup (lvs_ndt_params::accumulation = LVS_ACC_FAST).
// Same job as ndt_eval.cu - computeDerivatives / updateDerivatives / computePointDerivatives_AngleAxisd
// (include/ndt_omp/ndt_omp_impl2.hpp:197-305, 504-532, 567-619; pca weight include/ndt_pca/ndt_pca_impl2.hpp:293-296) over
// getNeighborhoodAtPoint{,7,1} (include/ndt_omp/voxel_grid_covariance_omp_impl.hpp:373-442) - with the arithmetic re-derived
// instead of mirrored:
//   * the voxel lookup is untouched: float transform in the reference's operation order (this TU is compiled with -fmad=false,
//     every fused operation below is an explicit fmaf), floor(x / leaf), same probes - voxel indices stay bit-exact;
//   * d = x' - mean comes from a two-float mean ((x' - mh) - ml, ~1e-7 m) instead of a double subtraction, exp(-d2 q / 2) from
//     one MUFU.EX2, the inverse covariance is taken symmetric (6 floats);
//   * only the 31 distinct sums are formed: score, g[6], the symmetric translation block of H (6), the translation-rotation block
//     (9, mirrored afterwards: the reference's two copies differ in the last float ulp of each term only), and the full rotation
//     block (9; the second-derivative term c.Hp_ij of :522-530 is not symmetric);
//   * a lane keeps the 31 sums of ITS contributions in float32 registers (one FFMA per product: no staging tile, no conversion),
//     for at most kFlushRounds = 32 contributions; then the warp transposes-and-adds them (31 SHFL + 31 FADD: lane l ends up
//     with the warp's sum of output l), converts once and adds into ONE fp64 accumulator per lane.
// About 150 issued instructions per (point, cell) contribution against ~500 in the exact kernel.  Sums agree with the exact mode
// to ~1e-6 relative (tests/test_ndt_gpu.py::test_fast_mode_*), far inside the 1e-4 m / 1e-5 rad per-iteration bar.
// Work distribution, queue compaction of the (point, cell) hits, CTA partial -> ticket -> last-CTA reduction and the device-side
// Newton state machine are those of ndt_eval.cu (eval_finish, ndt_eval_common.cuh).
#include "ndt_eval_common.cuh"

namespace lvs {

constexpr int kFWarps = kEvalThreads / 32;
constexpr int kFPtsPerLane = 2;
constexpr int kFPtsPerIter = 32 * kFPtsPerLane;
constexpr int kFQueueCap = 256;
constexpr int kFlushRounds = 32;       // contributions a lane sums in float32 before the warp folds them into fp64
constexpr int kFSlots = 32;            // 31 sums + 1 unused

// slot -> canonical output index (0 score, 1..6 gradient, 7 + 6 i + j Hessian) and its mirror (-1: none)
//   0 score | 1..6 g | 7..12 H_tt upper (00 01 02 11 12 22) | 13..21 H_tr[i][3+j] row-major | 22..30 H_rr[3+i][3+j] row-major
__device__ __forceinline__ void fast_slot_outputs(int s, int& a, int& b) {
  b = -1;
  if (s < 7) { a = s; return; }
  int i, j;
  if (s < 13) {
    const int t = s - 7;
    i = t < 3 ? 0 : (t < 5 ? 1 : 2);
    j = t < 3 ? t : (t < 5 ? t - 2 : 2);
  } else if (s < 22) { i = (s - 13) / 3; j = 3 + (s - 13) % 3; }
  else if (s < 31) { a = 7 + 6 * (3 + (s - 22) / 3) + 3 + (s - 22) % 3; return; }
  else { a = -1; return; }
  a = 7 + 6 * i + j;
  if (i != j) b = 7 + 6 * j + i;
}

// ---- staging of a pair's voxel records in shared memory by ONE bulk copy (TMA, cp.async.bulk) per CTA: north_star names "TMA / shared-memory
// staging of voxel covariances".  The records of one target are ~1.6 k x 48 B = 77 KB: they fit, but cost the third resident CTA of an SM.
// Measured: the L1-cached __ldg path is as fast, so staging is an option (LVS_STAGE_RECORDS=1), not the default.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_stage(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
  }
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
  } while (!ok);
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One (point, cell) contribution added into the lane's float sums.  pa = (x', y', z', xr), pb = (yr, zr): transformed point and
// rotated point R*x; w: ndt_pca weight of this contribution (running product of the cell weights, ndt_pca_impl2.hpp:293-296).
// HESS = false (line-search trials, and the last pass of an align under lean_final_evaluation): score and gradient only.
template <bool PCA, bool STAGE, bool HESS>
__device__ __forceinline__ void fast_contribute(float (&A)[kFSlots], const FastRec* __restrict__ fr, const float4 pa, const float2 pb, float w,
                                                float kexp, float gd1, float gd2) {
  float4 r0, r1, r2;
  if (STAGE) {                                                         // records staged in shared memory
    r0 = reinterpret_cast<const float4*>(fr)[0]; r1 = reinterpret_cast<const float4*>(fr)[1]; r2 = reinterpret_cast<const float4*>(fr)[2];
  } else {
    r0 = __ldg(reinterpret_cast<const float4*>(fr));                   // mh0 mh1 mh2 ml0
    r1 = __ldg(reinterpret_cast<const float4*>(fr) + 1);               // ml1 ml2 c00 c01
    r2 = __ldg(reinterpret_cast<const float4*>(fr) + 2);               // c02 c11 c12 c22
  }
  const float d0 = (pa.x - r0.x) - r0.w, d1 = (pa.y - r0.y) - r1.x, d2 = (pa.z - r0.z) - r1.y;
  const float c00 = r1.z, c01 = r1.w, c02 = r2.x, c11 = r2.y, c12 = r2.z, c22 = r2.w;
  // c = C d (= d^T C), q = d^T C d
  const float c0 = fmaf(c02, d2, fmaf(c01, d1, c00 * d0));
  const float c1 = fmaf(c12, d2, fmaf(c11, d1, c01 * d0));
  const float c2 = fmaf(c22, d2, fmaf(c12, d1, c02 * d0));
  const float q = fmaf(d2, c2, fmaf(d1, c1, d0 * c0));
  const float e0 = ex2_approx(q * kexp);                               // exp(-d2 q / 2)
  float sc = -gd1 * e0;
  const float e1 = gd2 * e0;
  if (!(e1 <= 1.0f && e1 >= 0.0f)) return;                             // the reference's early-out (:588-589), NaN included
  float e2 = e1 * gd1;
  if (PCA) { e2 *= w; sc *= w; }
  const float x = pa.w, y = pb.x, z = pb.y;
  // the nine products x_k c_l feed both a_rot = x_t x c and the second-derivative terms c . Hp_ij
  const float xc0 = x * c0, xc1 = x * c1, xc2 = x * c2, yc0 = y * c0, yc1 = y * c1, yc2 = y * c2, zc0 = z * c0, zc1 = z * c1, zc2 = z * c2;
  const float a3 = yc2 - zc1, a4 = zc0 - xc2, a5 = xc1 - yc0;
  A[0] += sc;
  A[1] = fmaf(e2, c0, A[1]); A[2] = fmaf(e2, c1, A[2]); A[3] = fmaf(e2, c2, A[3]);
  A[4] = fmaf(e2, a3, A[4]); A[5] = fmaf(e2, a4, A[5]); A[6] = fmaf(e2, a5, A[6]);
  if (!HESS) return;
  // H_ij += e2 (-d2 a_i a_j + c.Hp_ij + (J^T C J)_ij)  =  b_i a_j + e2 N_ij  with b = -d2 e2 a
  const float k = -gd2 * e2;
  const float b0 = k * c0, b1 = k * c1, b2 = k * c2, b3 = k * a3, b4 = k * a4, b5 = k * a5;
  A[7] = fmaf(b0, c0, fmaf(e2, c00, A[7]));
  A[8] = fmaf(b0, c1, fmaf(e2, c01, A[8]));
  A[9] = fmaf(b0, c2, fmaf(e2, c02, A[9]));
  A[10] = fmaf(b1, c1, fmaf(e2, c11, A[10]));
  A[11] = fmaf(b1, c2, fmaf(e2, c12, A[11]));
  A[12] = fmaf(b2, c2, fmaf(e2, c22, A[12]));
  // U = C Jr, Jr = [(0,-z,y) (z,0,-x) (-y,x,0)]
  const float u00 = fmaf(y, c02, -(z * c01)), u01 = fmaf(z, c00, -(x * c02)), u02 = fmaf(x, c01, -(y * c00));
  const float u10 = fmaf(y, c12, -(z * c11)), u11 = fmaf(z, c01, -(x * c12)), u12 = fmaf(x, c11, -(y * c01));
  const float u20 = fmaf(y, c22, -(z * c12)), u21 = fmaf(z, c02, -(x * c22)), u22 = fmaf(x, c12, -(y * c02));
  A[13] = fmaf(b0, a3, fmaf(e2, u00, A[13])); A[14] = fmaf(b0, a4, fmaf(e2, u01, A[14])); A[15] = fmaf(b0, a5, fmaf(e2, u02, A[15]));
  A[16] = fmaf(b1, a3, fmaf(e2, u10, A[16])); A[17] = fmaf(b1, a4, fmaf(e2, u11, A[17])); A[18] = fmaf(b1, a5, fmaf(e2, u12, A[18]));
  A[19] = fmaf(b2, a3, fmaf(e2, u20, A[19])); A[20] = fmaf(b2, a4, fmaf(e2, u21, A[20])); A[21] = fmaf(b2, a5, fmaf(e2, u22, A[21]));
  // M_rr = Jr^T U (symmetric), N_rr = M_rr + c.Hp
  const float m00 = fmaf(y, u20, -(z * u10)), m01 = fmaf(y, u21, -(z * u11)), m02 = fmaf(y, u22, -(z * u12));
  const float m11 = fmaf(z, u01, -(x * u21)), m12 = fmaf(z, u02, -(x * u22));
  const float m22 = fmaf(x, u12, -(y * u02));
  A[22] = fmaf(b3, a3, fmaf(e2, m00 - (yc1 + zc2), A[22]));
  A[23] = fmaf(b3, a4, fmaf(e2, m01 + xc1, A[23]));
  A[24] = fmaf(b3, a5, fmaf(e2, m02 + xc2, A[24]));
  A[25] = fmaf(b4, a3, fmaf(e2, m01 + yc0, A[25]));
  A[26] = fmaf(b4, a4, fmaf(e2, m11 - (xc0 + zc2), A[26]));
  A[27] = fmaf(b4, a5, fmaf(e2, m12 + yc2, A[27]));
  A[28] = fmaf(b5, a3, fmaf(e2, m02 + zc0, A[28]));
  A[29] = fmaf(b5, a4, fmaf(e2, m12 + zc1, A[29]));
  A[30] = fmaf(b5, a5, fmaf(e2, m22 - (xc0 + yc1), A[30]));
}

// Transpose-and-add of the warp's float sums: afterwards lane l holds sum over lanes of A[l]; the result is added in fp64 and the
// float sums restart from zero.  Step s: the lanes with bit s set keep the upper half of the remaining slots and hand the lower half to
// their partner (and vice versa), so 16 + 8 + 4 + 2 + 1 = 31 exchanges reduce 32 x 32 values.
__device__ __forceinline__ void fast_flush(float (&A)[kFSlots], double& accd, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; k++) {
      const float keep = up ? A[k + s] : A[k];
      const float send = up ? A[k] : A[k + s];
      A[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  accd += (double)A[0];
#pragma unroll
  for (int k = 0; k < kFSlots; k++) A[k] = 0.0f;
}

// Per-warp staging in shared memory: the 64 points of the current warp iteration and the queue of (record, point) hits.
template <bool PCA>
struct FastSmem {
  static constexpr size_t pa_off = 0;                                            // float4 [64]  x' y' z' xr
  static constexpr size_t pb_off = pa_off + sizeof(float4) * kFPtsPerIter;       // float2 [64]  yr zr
  static constexpr size_t q_off = pb_off + sizeof(float2) * kFPtsPerIter;        // int [256]    record * 64 + point slot
  static constexpr size_t qw_off = q_off + sizeof(int) * kFQueueCap;             // float [256]  ndt_pca weight of the entry
  static constexpr size_t warp_bytes = qw_off + (PCA ? sizeof(float) * kFQueueCap : 0);
  static constexpr size_t bytes = warp_bytes * kFWarps;
};

template <int MODE, bool PCA, bool STAGE>
__global__ void __launch_bounds__(kEvalThreads, STAGE ? 2 : 3) ndt_eval_fast_kernel(EvalLaunch L) {
  extern __shared__ __align__(16) unsigned char s_recs[];          // STAGE: the target's FastRec array
  __shared__ unsigned long long s_bar;
  constexpr int K = Probes<MODE>::K;
  using SM = FastSmem<PCA>;
  __shared__ __align__(16) unsigned char s_stage[SM::bytes];
  __shared__ double s_red[kFWarps][kFSlots];
  __shared__ float s_T[16], s_R[9];
  __shared__ int s_last;
  pdl_wait();
  const int pair = blockIdx.x / L.blocks_per_pair, blk = blockIdx.x % L.blocks_per_pair;
  AlignState& S = L.d_states[pair];
  const int kind = S.eval_kind;
  const AlignConsts& c = L.consts;
  if (kind != EVAL_DERIV_H && kind != EVAL_DERIV_NOH) return;
  const PairDesc P = L.d_pairs[pair];
  if (threadIdx.x < 16) s_T[threadIdx.x] = S.T[threadIdx.x];
  if (threadIdx.x < 9) s_R[threadIdx.x] = S.Rj[threadIdx.x];
  __syncthreads();
  const GridView G = load_grid_view(P.gp);
  const float gd1 = (float)c.gauss_d1, gd2 = (float)c.gauss_d2;
  const float kexp = (float)(-0.5 * c.gauss_d2 * 1.4426950408889634);
  const int bpp = L.blocks_per_pair;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned char* wbase = s_stage + (size_t)warp * SM::warp_bytes;
  float4* pa = reinterpret_cast<float4*>(wbase + SM::pa_off);
  float2* pb = reinterpret_cast<float2*>(wbase + SM::pb_off);
  int* q = reinterpret_cast<int*>(wbase + SM::q_off);
  float* qw = reinterpret_cast<float*>(wbase + SM::qw_off);        // only mapped for ndt_pca launches
  const FastRec* __restrict__ frecs = P.frecs;
  if (STAGE) {
    const unsigned bytes = (unsigned)P.gp->n_cells * (unsigned)sizeof(FastRec);
    if (bytes > 0 && bytes <= (unsigned)L.stage_bytes) { bulk_stage(s_recs, P.frecs, bytes, &s_bar); frecs = reinterpret_cast<const FastRec*>(s_recs); }
  }
  const int* __restrict__ grid = P.grid;
  const bool hess = kind == EVAL_DERIV_H;          // CTA-uniform
  const float* T = s_T;
  const float* R = s_R;

  float A[kFSlots];
#pragma unroll
  for (int k = 0; k < kFSlots; k++) A[k] = 0.0f;
  double accd = 0.0;
  int since_flush = 0;

  const float inv_leaf = 1.0f / G.leaf;
  // for a power-of-two leaf x / leaf equals x * (1 / leaf) bit for bit (see ndt_eval.cu)
  const bool pow2 = (__float_as_uint(G.leaf) & 0x007fffffu) == 0u && isfinite(inv_leaf) && inv_leaf >= 1.1754944e-38f;
  const int div0 = G.max_b[0] - G.min_b[0] + 1, div1 = G.max_b[1] - G.min_b[1] + 1, div2 = G.max_b[2] - G.min_b[2] + 1;

  // rounds of 32 queued contributions starting at `head`; lanes past n_round idle
  auto round = [&](int head, int n_round) {
    if (lane < n_round) {
      const int ent = q[head + lane];
      const int rec = ent / kFPtsPerIter, slot = ent % kFPtsPerIter;
      if (hess) fast_contribute<PCA, STAGE, true>(A, frecs + rec, pa[slot], pb[slot], PCA ? qw[head + lane] : 1.0f, kexp, gd1, gd2);
      else fast_contribute<PCA, STAGE, false>(A, frecs + rec, pa[slot], pb[slot], PCA ? qw[head + lane] : 1.0f, kexp, gd1, gd2);
    }
    if (++since_flush == kFlushRounds) { fast_flush(A, accd, lane); since_flush = 0; }
  };

  if (!G.empty) {
    // every warp of the pair's CTAs owns one contiguous range of the source (see ndt_eval.cu)
    const long long n_warps = (long long)bpp * kFWarps, wid = (long long)blk * kFWarps + warp;
    const int lo = (int)(wid * P.n_src / n_warps), hi = (int)((wid + 1) * P.n_src / n_warps);
    for (int base = lo; base < hi; base += kFPtsPerIter) {
      int nq = 0;
      // consumes whole rounds from the queue and moves the remainder (< 32 entries) to its front; returns the new length
      auto drain = [&]() {
        __syncwarp();
        int head = 0;
        for (; nq - head >= 32; head += 32) round(head, 32);
        const int rem = nq - head;
        int ent = 0; float we = 0.0f;
        if (lane < rem) { ent = q[head + lane]; if (PCA) we = qw[head + lane]; }
        __syncwarp();
        if (lane < rem) { q[lane] = ent; if (PCA) qw[lane] = we; }
        return rem;
      };
#pragma unroll
      for (int h = 0; h < kFPtsPerLane; h++) {
        if (MODE == LVS_DIRECT7 && h > 0 && nq >= 32) nq = drain();
        const int slot = h * 32 + lane;
        const int i = base + slot;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        bool ok = i < hi;
        if (ok) {
          const float4 s = __ldg(P.src + i);
          transform_point(T, s.x, s.y, s.z, tx, ty, tz);
          ok = isfinite(tx) && isfinite(ty) && isfinite(tz);
          // x_t = float(SE3::exp(p).matrix()) * [x, 0]: rotation only (ndt_omp_impl2.hpp:507-508)
          const float xr = (R[0] * s.x + R[1] * s.y) + R[2] * s.z;
          const float yr = (R[3] * s.x + R[4] * s.y) + R[5] * s.z;
          const float zr = (R[6] * s.x + R[7] * s.y) + R[8] * s.z;
          pa[slot] = make_float4(tx, ty, tz, xr);
          pb[slot] = make_float2(yr, zr);
        }
        int cx, cy, cz;
        if (pow2) { cx = (int)floorf(tx * inv_leaf); cy = (int)floorf(ty * inv_leaf); cz = (int)floorf(tz * inv_leaf); }
        else { cx = (int)floorf(tx / G.leaf); cy = (int)floorf(ty / G.leaf); cz = (int)floorf(tz / G.leaf); }
        const int rx = cx - G.min_b[0], ry = cy - G.min_b[1], rz = cz - G.min_b[2];
        const int cell0 = rx * G.mul[0] + ry * G.mul[1] + rz * G.mul[2];
        // ndt_pca scales the RUNNING per-point sums by each cell's weight: the contribution of cell k carries the product of the
        // weights of cells k..last, hence the probes run last-to-first (as in ndt_eval.cu)
        float run = 1.0f;
#pragma unroll(MODE == LVS_DIRECT26 ? 1 : K)
        for (int k = K - 1; k >= 0; k--) {
          int ox = 0, oy = 0, oz = 0;
          if (MODE == LVS_DIRECT7) { ox = k == 1 ? 1 : k == 2 ? -1 : 0; oy = k == 3 ? 1 : k == 4 ? -1 : 0; oz = k == 5 ? 1 : k == 6 ? -1 : 0; }
          else if (MODE == LVS_DIRECT26) { ox = c_off26[k][0]; oy = c_off26[k][1]; oz = c_off26[k][2]; }
          int v = -1;
          if (ok && (unsigned)(rx + ox) < (unsigned)div0 && (unsigned)(ry + oy) < (unsigned)div1 && (unsigned)(rz + oz) < (unsigned)div2)
            v = __ldg(grid + (cell0 + ox * G.mul[0] + oy * G.mul[1] + oz * G.mul[2]));
          const bool hit = v >= 0;
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (MODE == LVS_DIRECT26) {
            if (m == 0u) continue;
            if (nq > kFQueueCap - 32) nq = drain();
          }
          if (hit) {
            const int pos = nq + __popc(m & lt_mask);
            q[pos] = v * kFPtsPerIter + slot;
            if (PCA) { run *= (float)(__ldg(&P.recs[v].meta) & kMetaWeightMask); qw[pos] = run; }
          }
          nq += __popc(m);
        }
      }
      static_assert(MODE == LVS_DIRECT26 || 31 + 32 * Probes<MODE>::K <= kFQueueCap, "queue capacity per half iteration");
      __syncwarp();
      for (int head = 0; head < nq; head += 32) round(head, min(32, nq - head));
      __syncwarp();
    }
  }
  pdl_trigger();
  fast_flush(A, accd, lane);

  // CTA partial: every output slot sums the 8 warps in fixed order and lands at its place(s) in the canonical 43-vector
  s_red[warp][lane] = accd;
  __syncthreads();
  double* partial = L.d_partials + ((size_t)pair * bpp + blk) * kPartialStride;
  if (threadIdx.x < kFSlots) {
    double x = 0;
#pragma unroll
    for (int w = 0; w < kFWarps; w++) x += s_red[w][threadIdx.x];
    int oa, ob;
    fast_slot_outputs(threadIdx.x, oa, ob);
    if (oa >= 0) partial[oa] = x;
    if (ob >= 0) partial[ob] = x;
  }
  eval_finish(L, pair, kind, kind == EVAL_DERIV_H ? kAcc : 7, P.n_total, reinterpret_cast<double*>(s_stage), &s_last);
}

template <int MODE, bool PCA>
static int launch_fast_as(cudaStream_t st, const EvalLaunch& L) {
  if (L.stage_bytes > 0) {
    static int attr_dev = -1;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (attr_dev != dev) {
      CUDA_TRY(cudaFuncSetAttribute(ndt_eval_fast_kernel<MODE, PCA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_dev = dev;
    }
    return launch_pdl(ndt_eval_fast_kernel<MODE, PCA, true>, (unsigned)(L.n_pairs * L.blocks_per_pair), kEvalThreads, (size_t)L.stage_bytes, st, L);
  }
  return launch_pdl(ndt_eval_fast_kernel<MODE, PCA, false>, (unsigned)(L.n_pairs * L.blocks_per_pair), kEvalThreads, 0, st, L);
}

int launch_eval_fast(cudaStream_t st, const EvalLaunch& L) {
  if (L.n_pairs <= 0) return LVS_OK;
  const bool pca = L.consts.variant == LVS_NDT_PCA;
  switch (L.consts.search) {
    case LVS_DIRECT1: return pca ? launch_fast_as<LVS_DIRECT1, true>(st, L) : launch_fast_as<LVS_DIRECT1, false>(st, L);
    case LVS_DIRECT7: return pca ? launch_fast_as<LVS_DIRECT7, true>(st, L) : launch_fast_as<LVS_DIRECT7, false>(st, L);
    case LVS_DIRECT26: return pca ? launch_fast_as<LVS_DIRECT26, true>(st, L) : launch_fast_as<LVS_DIRECT26, false>(st, L);
    default: return LVS_OK;                                  // KDTREE: radius-search derivatives live in the cold kernel
  }
}

int eval_fast_max_resident_ctas_per_sm() { return 3; }

}  // namespace lvs
