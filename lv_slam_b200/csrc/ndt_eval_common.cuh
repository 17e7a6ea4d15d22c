// Shared pieces of the NDT evaluation kernels (ndt_eval.cu = direct-search derivative pass, the hot kernel;
// ndt_eval_cold.cu = radius-search passes: KDTREE-mode derivatives, computeHessian, calculateScore, taps).
#pragma once
#include "ndt_internal.cuh"
#include "ndt_state.cuh"

namespace lvs {

constexpr int kEvalThreads = 256;
constexpr float kOne = 1.0f;

// DIRECT7 probes, in the reference's order (voxel_grid_covariance_omp_impl.hpp:423-430): centre, +x, -x, +y, -y, +z, -z.
// pcl::getAllNeighborCellIndices (PCL 1.8 voxel_grid.h): 13 half-offsets, then their negatives; no centre cell.
static __constant__ int c_off26[26][3] = {
    {-1, -1, -1}, {-1, 0, -1}, {-1, 1, -1}, {0, -1, -1}, {0, 0, -1}, {0, 1, -1}, {1, -1, -1}, {1, 0, -1}, {1, 1, -1},
    {-1, -1, 0},  {0, -1, 0},  {1, -1, 0},  {-1, 0, 0},
    {1, 1, 1},    {1, 0, 1},   {1, -1, 1},  {0, 1, 1},   {0, 0, 1},  {0, -1, 1},  {-1, 1, 1},  {-1, 0, 1},  {-1, -1, 1},
    {1, 1, 0},    {0, 1, 0},   {-1, 1, 0},  {1, 0, 0}};

template <int MODE> struct Probes;
template <> struct Probes<LVS_DIRECT1> { static constexpr int K = 1; };
template <> struct Probes<LVS_DIRECT7> { static constexpr int K = 7; };
template <> struct Probes<LVS_DIRECT26> { static constexpr int K = 26; };


// Programmatic dependent launch: consecutive evaluation launches of an align are queued back to back, each reading the state the previous one's
// last CTA left behind.  Launched with programmatic stream serialisation, a kernel's CTAs may be scheduled as soon as every CTA of the previous
// launch has passed pdl_trigger() (after its share of the points, before the reduction tail) or exited; they then block in pdl_wait() until that
// launch has completed and its writes are visible.  The launch latency and the prologue hide under the previous tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename Kernel>
static int launch_pdl(Kernel kernel, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const EvalLaunch& L) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, L));
  return LVS_OK;
}

struct GridView {
  int min_b[3], max_b[3], mul[3];
  float leaf;
  bool empty;
};

__device__ __forceinline__ GridView load_grid_view(const GridParams* gp) {
  GridView g;
  for (int a = 0; a < 3; a++) { g.min_b[a] = gp->min_b[a]; g.max_b[a] = gp->max_b[a]; g.mul[a] = gp->mul[a]; }
  g.leaf = gp->leaf;
  g.empty = gp->status != 0 || gp->n_cells == 0;
  return g;
}

// pcl::transformPointCloud, dense branch: ((m00*x + m01*y) + m02*z) + m03, float, no contraction.
__device__ __forceinline__ void transform_point(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = ((T[0] * x + T[4] * y) + T[8] * z) + T[12];
  oy = ((T[1] * x + T[5] * y) + T[9] * z) + T[13];
  oz = ((T[2] * x + T[6] * y) + T[10] * z) + T[14];
}

// Completion of one evaluation of one pair: every CTA has stored its partial[nv]; the LAST CTA to take a ticket sums the
// bpp partials in a fixed order (run-to-run deterministic), deposits (score, g, H) in the pair's AlignState and advances the
// Newton / More-Thuente state machine so that the next launch knows what to evaluate.  s_scratch: >= 704 doubles of shared memory.
// The tail is what a single-pair align waits for between two evaluations, so it is kept short: the partials are fetched with every
// load of a thread in flight at once (10 row groups x 22 column pairs of threads, double2 loads), and the ~0.9 KB state is copied to shared memory once,
// advanced there by one thread (6x6 solve, SE(3) exp / log) and copied back once, instead of being walked field by field in L2.
// The pose composition log(exp(dir a_t) exp(p)) the state machine asks for at the end of a line search (~4.6 us of serial fp64 sin / cos /
// atan2 on one thread) depends on nothing this evaluation computes: warp 7 takes no part in the reduction (warps 0-6 synchronise among
// themselves on named barrier 1) and works it out meanwhile, joining the others just before the state machine runs.
__device__ __forceinline__ void eval_finish(const EvalLaunch& L, int pair, int kind, int nv, int n_src, double* s_scratch, int* s_last, long long t_entry = 0) {
  AlignState& S = L.d_states[pair];
  const bool dbg = L.d_dbg != nullptr && threadIdx.x == 0;
  long long stamp[8];
  if (dbg) { stamp[0] = t_entry; stamp[1] = clock64(); }
  const AlignConsts& c = L.consts;
  const int bpp = L.blocks_per_pair;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&L.d_tickets[pair], 1u);
    *s_last = (t == (unsigned)bpp - 1u);
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (dbg) stamp[2] = clock64();
  __shared__ __align__(16) AlignState s_state;
  __shared__ double s_lu[6][7], s_dp[6], s_pn[6], s_pn_at;
  __shared__ int s_perm[6], s_lu_ok, s_pn_ok;
  static_assert(sizeof(AlignState) % 4 == 0, "AlignState is copied word by word");
  {
    const int* src = reinterpret_cast<const int*>(&S);
    int* dst = reinterpret_cast<int*>(&s_state);
    for (int i = threadIdx.x; i < (int)(sizeof(AlignState) / 4); i += kEvalThreads) dst[i] = __ldcg(src + i);
  }
  __syncthreads();
  constexpr int kSumThreads = kEvalThreads - 32;          // warps 0-6 reduce, warp 7 composes the pose
  const bool composer = threadIdx.x >= kSumThreads;
#define LVS_SUB_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(kSumThreads) : "memory")
  if (composer) {
    if (L.advance && threadIdx.x == kSumThreads) {
      const bool useful = s_state.phase == PH_MT_FIRST || s_state.phase == PH_MT_TRIAL || s_state.phase == PH_HESS27;
      if (useful) {
        double delta[6], pn[6];
        for (int i = 0; i < 6; i++) delta[i] = s_state.dir[i] * s_state.a_t;
        se3_log(se3_mul(se3_exp(delta), se3_exp(s_state.p)), pn);
        for (int i = 0; i < 6; i++) s_pn[i] = pn[i];
        s_pn_at = s_state.a_t;
      }
      s_pn_ok = useful ? 1 : 0;
    }
  } else {
  constexpr int R = 10;                  // row groups x 22 column pairs = 220 of the 224 reducing threads, one double2 per load
  static_assert(R * (kPartialStride / 2) <= kSumThreads && R * 64 <= 1024, "reduction layout");
  {
    const int kp = threadIdx.x % (kPartialStride / 2), r = threadIdx.x / (kPartialStride / 2);
    double x0 = 0, x1 = 0;
    if (r < R && 2 * kp < nv) {
      const double2* base = reinterpret_cast<const double2*>(L.d_partials + (size_t)pair * bpp * kPartialStride) + kp;
      constexpr int kIn = 15;            // loads in flight per thread (444 CTAs of a single pair: three passes)
      for (int b0 = r; b0 < bpp; b0 += R * kIn) {
        double2 v[kIn];
#pragma unroll
        for (int u = 0; u < kIn; u++) { const int b = b0 + R * u; v[u] = b < bpp ? __ldcg(base + (size_t)b * (kPartialStride / 2)) : make_double2(0.0, 0.0); }
#pragma unroll
        for (int u = 0; u < kIn; u++) { x0 += v[u].x; x1 += v[u].y; }      // adding +0.0 past the end is exact
      }
    }
    if (r < R) { s_scratch[r * 64 + 2 * kp] = x0; s_scratch[r * 64 + 2 * kp + 1] = x1; }
  }
  LVS_SUB_SYNC();
  if (dbg) stamp[3] = clock64();
  double x = 0;
  if (threadIdx.x < nv)
    for (int r = 0; r < R; r++) x += s_scratch[r * 64 + threadIdx.x];
  if (L.shard.world > 1 || L.shard.mine) {
    // ---- exchange of the rank sums through peer memory (see ShardView)
    const ShardView& sh = L.shard;
    // serial of THIS evaluation of THIS pair: the align's base plus the pair's own evaluation count - the same on every rank whatever
    // number of launches (some of them idle for this pair) each host has issued
    const long long serial = sh.serial + s_state.n_eval + s_state.n_hess + 1;
    const size_t rec = (((size_t)(serial & 1) * sh.world + sh.rank) * sh.cap + pair) * kMailStride;
    if (threadIdx.x < nv)
      for (int p = 0; p < sh.world; p++) *reinterpret_cast<volatile double*>(sh.peers[p] + rec + threadIdx.x) = x;
    __threadfence_system();
    LVS_SUB_SYNC();
    if (threadIdx.x < sh.world) {    // one flag store per peer, after every value store of this CTA is visible system-wide
      __threadfence_system();
      *reinterpret_cast<volatile long long*>(sh.peers[threadIdx.x] + rec + 43) = serial;
    }
    __shared__ int s_peer_ok;
    if (threadIdx.x == 0) s_peer_ok = 1;
    LVS_SUB_SYNC();
    if (threadIdx.x < sh.world) {
      const volatile long long* flag =
          reinterpret_cast<const volatile long long*>(sh.mine + (((size_t)(serial & 1) * sh.world + threadIdx.x) * sh.cap + pair) * kMailStride + 43);
      const long long t0 = clock64();
      while (*flag != serial) {
        if (clock64() - t0 > sh.timeout_cycles) { s_peer_ok = 0; break; }
        __nanosleep(64);
      }
    }
    __threadfence_system();
    LVS_SUB_SYNC();
    if (!s_peer_ok && threadIdx.x == 0) *sh.d_error = 1;
    if (threadIdx.x < nv) {
      x = 0;
      for (int r = 0; r < sh.world; r++)
        x += *reinterpret_cast<const volatile double*>(sh.mine + (((size_t)(serial & 1) * sh.world + r) * sh.cap + pair) * kMailStride + threadIdx.x);
    }
  }
  if (threadIdx.x < nv) {
    const int k = threadIdx.x;
    if (c.variant == LVS_NDT_GROUND && kind != EVAL_HESS27 && k >= 1) {
      // computeDerivatives_seg: updateDerivatives runs twice per cell and both calls add into the gradient and the Hessian, only the
      // second one's score is kept (ndt_ground_impl.hpp:519,522); rows and columns x, y, yaw are zeroed afterwards (:554-561)
      const int i = k < 7 ? k - 1 : (k - 7) / 6, j = k < 7 ? k - 1 : (k - 7) % 6;
      const bool off = i == 0 || i == 1 || i == 5 || j == 0 || j == 1 || j == 5;
      x = off ? 0.0 : 2.0 * x;
    }
    if (kind == EVAL_HESS27) { if (k >= 7) s_state.H[k - 7] = x; }
    else {
      if (k == 0) s_state.score = x;
      else if (k < 7) s_state.g[k - 1] = x;
      else if (kind == EVAL_DERIV_H) s_state.H[k - 7] = x;
    }
  }
  // computeDerivatives zeroes the Hessian even when it does not fill it (ndt_omp_impl2.hpp:204)
  if (kind == EVAL_DERIV_NOH && threadIdx.x < 36) s_state.H[threadIdx.x] = 0.0;
  LVS_SUB_SYNC();
  if (dbg) stamp[4] = clock64();
  // Newton direction H^-1 (-g) of this evaluation, by warp 1 (the state machine asks for it in almost every pass)
  if (L.advance && threadIdx.x >= 32 && threadIdx.x < 64) {
    const int lane = threadIdx.x - 32;
    for (int e = lane; e < 42; e += 32) s_lu[e / 7][e % 7] = (e % 7 < 6) ? s_state.H[(e / 7) * 6 + e % 7] : -s_state.g[e / 7];
    __syncwarp();
    const bool ok = lu6_solve_warp(s_lu, s_perm, s_dp, lane);
    if (lane == 0) s_lu_ok = ok ? 1 : 0;
  }
  }      // warps 0-6
#undef LVS_SUB_SYNC
  __syncthreads();
  if (dbg) stamp[5] = clock64();
  bool fin = false;
  if (threadIdx.x == 0) {
    if (L.advance) fin = align_state_advance(s_state, c, n_src, L.d_trace ? L.d_trace + (size_t)pair * kMaxTrace : nullptr, s_lu_ok ? s_dp : nullptr,
                                             s_pn_ok ? s_pn : nullptr, s_pn_at);
    else {
      s_state.eval_kind = EVAL_NONE;
      if (kind != EVAL_HESS27) s_state.n_eval++; else s_state.n_hess++;
    }
  }
  if (dbg) stamp[6] = clock64();
  __syncthreads();
  {
    int* dst = reinterpret_cast<int*>(&S);
    const int* src = reinterpret_cast<const int*>(&s_state);
    for (int i = threadIdx.x; i < (int)(sizeof(AlignState) / 4); i += kEvalThreads) dst[i] = src[i];
  }
  __threadfence();
  __syncthreads();
  if (dbg) { stamp[7] = clock64(); for (int i = 0; i < 8; i++) L.d_dbg[pair * 8 + i] = stamp[i]; }
  if (threadIdx.x == 0) {
    L.d_tickets[pair] = 0;
    if (fin) {
      // the pair that completes the batch publishes the align's serial in host-mapped memory: the host polls that word instead of
      // synchronising the stream (every state of the batch is in device memory and fenced by then)
      const int done = atomicAdd(L.d_done_count, 1) + 1;
      if (done == L.n_pairs && L.h_done_flag) { __threadfence_system(); *L.h_done_flag = L.align_serial; }
    }
    __threadfence();
  }
}

}  // namespace lvs
