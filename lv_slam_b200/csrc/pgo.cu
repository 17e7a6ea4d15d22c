// Pose-graph optimisation on the device — replaces the body of lv_slam::GraphSLAM::optimize
// (src/global_graph/graph_slam.cpp:298-331), i.e. g2o's SparseOptimizer::optimize with OptimizationAlgorithmLevenberg /
// GaussNewton over a BlockSolver of 6x6 pose blocks (3rdtools/g2o-a48ff8c.zip!g2o/g2o/core/{sparse_optimizer.cpp:366-431,
// optimization_algorithm_levenberg.cpp:58-175, optimization_algorithm_gauss_newton.cpp:50-92, block_solver.hpp:463-566,
// base_binary_edge.hpp:63-129}) for graphs of VertexSE3 / EdgeSE3 with optional Huber kernels.
//
// Data layout in HBM (all fp64): poses as 96 B Rt records; per edge Z, Z^-1 (96 B each), the full 6x6 information (288 B),
// Huber delta; a 960 B per-edge workspace (J_i^T W J_i, J_j^T W J_j, the off-diagonal block already in (row < col)
// orientation, and the two right-hand-side pieces); H as block-diagonal [n][36] + unique upper off-diagonal blocks [m][36];
// CSR incidence lists (vertex -> edges, block -> edges, block row -> blocks) built once per graph on the host.
//
// Kernels:
//   pgo_errors_kernel     per edge: e = toVectorMQT(Z^-1 Xi^-1 Xj), chi2, Huber rho -> fixed-shape two-stage sums
//   pgo_linearize_kernel  per edge: analytic Jacobians, Huber weight, the five products into the edge workspace
//   pgo_assemble_kernel   per H entry: gather the incident edge pieces in ascending edge order (deterministic, no atomics)
//   pgo_pcg_kernel        ONE cooperative launch per linear solve: block-Jacobi preconditioner, every PCG iteration with
//                         grid-wide barriers instead of launches, the LM gain-ratio denominator in the epilogue
//   pgo_update_kernel     X <- X * fromVectorMQT(dx) with the previous poses kept for LM's pop()
// Linear solvers: the *_pcg solver names restate g2o's own LinearSolverPCG (solvers/pcg/linear_solver_pcg.hpp:80-158); the *_var /
// *_cholmod names (direct Cholesky in the reference) run the supernodal multifrontal block Cholesky of pgo_chol.cu, analysed once per
// set_graph on the host (on a second thread, beside the uploads) and factorised on the device at every LM trial.
#include <future>
#include <cooperative_groups.h>
#include <cmath>
#include <map>
#include <new>
#include "ndt_internal.cuh"
#include "pgo_math.cuh"
#include "pgo_chol.cuh"

namespace cg = cooperative_groups;

namespace lvs {

constexpr int kEdgeWs = 120;       // doubles per edge workspace
constexpr int kPgoThreads = 256;
constexpr int kMaxPartials = 2048;

struct PgoScalars {
  double chi2_robust, chi2_plain;
  double dn, scale, max_diag;
  int pcg_iters, pcg_ok;
  unsigned long long max_diag_bits;
};

struct PgoDev {
  int nv, ne, nfree, noff;
  Rt* pose; Rt* pose_bak;
  const int* hidx;              // vertex -> block row or -1
  const int* free_vertex;       // block row -> vertex
  const int2* edge_ij; const int2* edge_h;
  const Rt* Z; const Rt* Zinv;
  const double* info; const double* huber;
  const unsigned char* edge_tr;
  const int* edge_type;         // null: every edge is an EdgeSE3; else 0 EdgeSE3, 1-4 unary priors (vertex i == j)
  const double* pm;             // [ne][8] measurements of the unary priors (edge_type != null)
  double* ws; double* err; double* chi;
  const int* vptr; const int* vinc;     // block row -> (edge * 2 + role) ascending
  const int* optr; const int* oinc;     // off block -> edges ascending
  const int* rptr; const int* rcol; const int* rslot;   // SpMV rows: rslot = -1 diag, else off slot * 2 + transposed
  double* Hd; double* Ho; double* b; double* x;
  double* J; double* r; double* d; double* q; double* s; double* q2;
  double* partials;             // [2][kMaxPartials]
  unsigned int* ticket;
  PgoScalars* sc;
};

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPgoThreads) pgo_errors_kernel(PgoDev D) {
  double rob = 0, pln = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D.ne; k += gridDim.x * blockDim.x) {
    const int2 ij = D.edge_ij[k];
    double e[6];
    const int ty = D.edge_type ? D.edge_type[k] : 0;
    if (ty) prior_error(ty, D.pm + (size_t)k * 8, D.pose[ij.x], e);
    else edge_error(D.Zinv[k], D.pose[ij.x], D.pose[ij.y], e);
    const double c = chi2_of(D.info + (size_t)k * 36, e);
#pragma unroll
    for (int a = 0; a < 6; a++) D.err[(size_t)k * 6 + a] = e[a];
    D.chi[k] = c;
    double r0 = c, r1 = 1.0;
    const double hd = D.huber[k];
    if (hd > 0) huber_rho(c, hd, &r0, &r1);
    rob += r0; pln += c;
  }
  __shared__ double s_a[kPgoThreads / 32], s_b[kPgoThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) { rob += __shfl_xor_sync(0xffffffffu, rob, o); pln += __shfl_xor_sync(0xffffffffu, pln, o); }
  if (lane == 0) { s_a[warp] = rob; s_b[warp] = pln; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < kPgoThreads / 32; w++) { a += s_a[w]; b += s_b[w]; }
    D.partials[blockIdx.x] = a; D.partials[kMaxPartials + blockIdx.x] = b;
    __threadfence();
    s_last = atomicAdd(D.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  double a = 0, b = 0;
  for (unsigned i = 0; i < gridDim.x; i++) { a += __ldcg(D.partials + i); b += __ldcg(D.partials + kMaxPartials + i); }
  D.sc->chi2_robust = a; D.sc->chi2_plain = b;
  *D.ticket = 0;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pgo_linearize_kernel(PgoDev D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D.ne) return;
  const int2 ij = D.edge_ij[k];
  const int2 h = D.edge_h[k];
  double* ws = D.ws + (size_t)k * kEdgeWs;
  if (h.x < 0 && h.y < 0) return;
  double A[36], B[36];
  const int ty = D.edge_type ? D.edge_type[k] : 0;
  if (ty) prior_jacobian(ty, D.pm + (size_t)k * 8, D.pose[ij.x], A);      // unary: only the (i, i) block and b_i exist
  else edge_gradient(D.Z[k], D.pose[ij.x], D.pose[ij.y], A, B);
  const double* info = D.info + (size_t)k * 36;
  double e[6];
#pragma unroll
  for (int a = 0; a < 6; a++) e[a] = D.err[(size_t)k * 6 + a];
  double w = 1.0, r0;
  const double hd = D.huber[k];
  if (hd > 0) huber_rho(D.chi[k], hd, &r0, &w);
  double W[36], omr[6];
#pragma unroll
  for (int a = 0; a < 36; a++) W[a] = w * info[a];
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double s = 0;
#pragma unroll
    for (int c = 0; c < 6; c++) s += info[r * 6 + c] * e[c];
    omr[r] = -s * w;
  }
  double C[36];
  if (h.x >= 0) {
    atwb(A, W, A, C);
#pragma unroll
    for (int a = 0; a < 36; a++) ws[a] = C[a];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < 6; c++) s += A[c * 6 + r] * omr[c];
      ws[108 + r] = s;
    }
  }
  if (h.y >= 0 && !ty) {
    atwb(B, W, B, C);
#pragma unroll
    for (int a = 0; a < 36; a++) ws[36 + a] = C[a];
#pragma unroll
    for (int r = 0; r < 6; r++) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < 6; c++) s += B[c * 6 + r] * omr[c];
      ws[114 + r] = s;
    }
  }
  if (h.x >= 0 && h.y >= 0 && h.x != h.y && !ty) {
    if (D.edge_tr[k]) atwb(B, W, A, C); else atwb(A, W, B, C);   // block (min, max): transposed when the edge runs high -> low
#pragma unroll
    for (int a = 0; a < 36; a++) ws[72 + a] = C[a];
  }
}

// One thread per stored H entry (diagonal blocks, off-diagonal blocks) and per right-hand-side entry.
__global__ void __launch_bounds__(kPgoThreads) pgo_assemble_kernel(PgoDev D) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nd = (long long)D.nfree * 36, no = (long long)D.noff * 36, nb = (long long)D.nfree * 6;
  if (t < nd) {
    const unsigned tu = (unsigned)t;                // 42 nfree + 36 noff < 2^32 (checked in set_graph): 32-bit divisions by constants
    const int v = (int)(tu / 36u), a = (int)(tu % 36u);
    double s = 0;
    for (int p = D.vptr[v]; p < D.vptr[v + 1]; p++) { const int inc = D.vinc[p]; s += D.ws[(size_t)(inc >> 1) * kEdgeWs + (inc & 1) * 36 + a]; }
    D.Hd[t] = s;
    if ((a % 7) == 0) atomicMax(&D.sc->max_diag_bits, (unsigned long long)__double_as_longlong(fabs(s)));   // exact: max is order independent
  } else if (t < nd + no) {
    const unsigned u = (unsigned)(t - nd);
    const int o = (int)(u / 36u), a = (int)(u % 36u);
    double s = 0;
    for (int p = D.optr[o]; p < D.optr[o + 1]; p++) s += D.ws[(size_t)D.oinc[p] * kEdgeWs + 72 + a];
    D.Ho[u] = s;
  } else if (t < nd + no + nb) {
    const unsigned u = (unsigned)(t - nd - no);
    const int v = (int)(u / 6u), a = (int)(u % 6u);
    double s = 0;
    for (int p = D.vptr[v]; p < D.vptr[v + 1]; p++) { const int inc = D.vinc[p]; s += D.ws[(size_t)(inc >> 1) * kEdgeWs + 108 + (inc & 1) * 6 + a]; }
    D.b[u] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Grid-wide deterministic sum: CTA partials in a double-buffered array, one grid barrier, every CTA re-adds them in the same order.
__device__ __forceinline__ double grid_sum(cg::grid_group& grid, double v, double* partials, int& phase, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double* buf = partials + (phase & 1) * kMaxPartials;
  if (threadIdx.x == 0) {
    double a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) a += s_red[w];
    buf[blockIdx.x] = a;
  }
  grid.sync();
  if (warp == 0) {
    double a = 0;
    for (int i = lane; i < (int)gridDim.x; i += 32) a += __ldcg(buf + i);
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_red[32] = a;
  }
  __syncthreads();
  const double out = s_red[32];
  __syncthreads();
  phase++;
  return out;
}

__global__ void __launch_bounds__(kPgoThreads) pgo_pcg_kernel(PgoDev D, double lambda, double tol, double prev_residual, int max_iter) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_red[40];
  const int n = D.nfree * 6;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  int phase = 0;
  // preconditioner J = (H_vv + lambda I)^-1, x = 0, r = b, d = J r, dn = r . d
  for (int v = gtid; v < D.nfree; v += gstride) {
    double M[36], Ji[36];
#pragma unroll
    for (int a = 0; a < 36; a++) M[a] = D.Hd[(size_t)v * 36 + a] + ((a % 7) == 0 ? lambda : 0.0);
    inv6(M, Ji);
#pragma unroll
    for (int a = 0; a < 36; a++) D.J[(size_t)v * 36 + a] = Ji[a];
  }
  grid.sync();
  double loc = 0;
  for (int t = gtid; t < n; t += gstride) {
    const int v = t / 6, c = t % 6;
    const double* Jr = D.J + (size_t)v * 36 + c * 6;
    double s = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) s += Jr[a] * D.b[v * 6 + a];
    D.x[t] = 0.0; D.r[t] = D.b[t]; D.d[t] = s;
    loc += D.b[t] * s;
  }
  double dn = grid_sum(grid, loc, D.partials, phase, s_red);
  double d0 = tol * dn;
  if (prev_residual > 0.0 && prev_residual > d0) d0 = prev_residual;
  // Two grid barriers per iteration (the two dot products).  The search direction is double-buffered: the owner of entry t
  // writes d_new[t] = s[t] + beta d_old[t] while every reader of a neighbour's entry forms the same expression itself, so no
  // barrier is needed between "update d" and the next matrix-vector product.  Iteration 0 uses beta = 0 with s = d.
  double* dcur = D.d;      // holds d_old (iteration 0: the initial direction, also copied to s below)
  double* dnxt = D.q2;
  for (int t = gtid; t < n; t += gstride) D.s[t] = D.d[t];
  double beta = 0.0;
  grid.sync();
  int it = 0;
  for (; it < max_iter; ++it) {
    if (dn <= d0) break;
    // q = (H + lambda I) d with d = s + beta d_old ; dq = d . q
    loc = 0;
    for (int t = gtid; t < n; t += gstride) {
      const int v = t / 6, c = t % 6;
      double acc = 0;
      for (int p = D.rptr[v]; p < D.rptr[v + 1]; p++) {
        const int col = D.rcol[p], slot = D.rslot[p];
        double dv[6];
#pragma unroll
        for (int a = 0; a < 6; a++) dv[a] = D.s[col * 6 + a] + beta * dcur[col * 6 + a];
        if (slot < 0) {
          const double* row = D.Hd + (size_t)v * 36 + c * 6;
#pragma unroll
          for (int a = 0; a < 6; a++) acc += row[a] * dv[a];
          acc += lambda * dv[c];
        } else if (slot & 1) {      // transposed: this row is the block's column
          const double* blk = D.Ho + (size_t)(slot >> 1) * 36;
#pragma unroll
          for (int a = 0; a < 6; a++) acc += blk[a * 6 + c] * dv[a];
        } else {
          const double* row = D.Ho + (size_t)(slot >> 1) * 36 + c * 6;
#pragma unroll
          for (int a = 0; a < 6; a++) acc += row[a] * dv[a];
        }
      }
      const double dt = D.s[t] + beta * dcur[t];
      dnxt[t] = dt;
      D.q[t] = acc;
      loc += dt * acc;
    }
    const double dq = grid_sum(grid, loc, D.partials, phase, s_red);
    const double alpha = dn / dq;
    // per vertex: x += alpha d ; r -= alpha q ; s = J r ; dn' = r . s
    loc = 0;
    for (int v = gtid; v < D.nfree; v += gstride) {
      double rv[6];
#pragma unroll
      for (int a = 0; a < 6; a++) {
        D.x[v * 6 + a] += alpha * dnxt[v * 6 + a];
        rv[a] = D.r[v * 6 + a] - alpha * D.q[v * 6 + a];
        D.r[v * 6 + a] = rv[a];
      }
      const double* Jv = D.J + (size_t)v * 36;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        double sv = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) sv += Jv[c * 6 + a] * rv[a];
        D.s[v * 6 + c] = sv;
        loc += rv[c] * sv;
      }
    }
    const double dn_new = grid_sum(grid, loc, D.partials, phase, s_red);
    beta = dn_new / dn;
    dn = dn_new;
    double* tmp = dcur; dcur = dnxt; dnxt = tmp;
  }
  // LM gain-ratio denominator: sum_j x_j (lambda x_j + b_j)  (OptimizationAlgorithmLevenberg::computeScale)
  loc = 0;
  for (int t = gtid; t < n; t += gstride) loc += D.x[t] * (lambda * D.x[t] + D.b[t]);
  const double scale = grid_sum(grid, loc, D.partials, phase, s_red);
  if (gtid == 0) { D.sc->dn = dn; D.sc->scale = scale; D.sc->pcg_iters = it; D.sc->pcg_ok = (dn == dn) ? 1 : 0; }
}

__global__ void __launch_bounds__(kPgoThreads) pgo_update_kernel(PgoDev D) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= D.nfree) return;
  const int vid = D.free_vertex[v];
  const Rt X = D.pose[vid];
  D.pose_bak[vid] = X;
  double dx[6];
#pragma unroll
  for (int a = 0; a < 6; a++) dx[a] = D.x[v * 6 + a];
  D.pose[vid] = rt_mul(X, rt_from_vector_mqt(dx));
}

__global__ void __launch_bounds__(kPgoThreads) pgo_restore_kernel(PgoDev D) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= D.nfree) return;
  const int vid = D.free_vertex[v];
  D.pose[vid] = D.pose_bak[vid];
}

__global__ void pgo_pack_poses_kernel(const double* __restrict__ p7, int n, Rt* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) out[v] = rt_from_qt7(p7 + (size_t)v * 7);
}
__global__ void pgo_unpack_poses_kernel(const Rt* __restrict__ in, int n, double* __restrict__ p7) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) rt_to_qt7(in[v], p7 + (size_t)v * 7);
}
__global__ void pgo_pack_edges_kernel(const double* __restrict__ m7, const double* __restrict__ info21, int n, Rt* __restrict__ Z, Rt* __restrict__ Zinv,
                                      double* __restrict__ info, const int* __restrict__ edge_type, double* __restrict__ pm, double fp0, double fp1, double fp2,
                                      double fp3) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int ty = edge_type ? edge_type[k] : 0;
  const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
  const double floor_plane[4] = {fp0, fp1, fp2, fp3};
  if (ty) prior_set_measurement(ty, m7 + (size_t)k * 7, pm + (size_t)k * 8, floor_plane);
  const Rt z = rt_from_qt7(ty ? ident : m7 + (size_t)k * 7);
  Z[k] = z; Zinv[k] = rt_inv(z);
  const double* u = info21 + (size_t)k * 21;
  int p = 0;
  for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { info[(size_t)k * 36 + r * 6 + c] = u[p]; info[(size_t)k * 36 + c * 6 + r] = u[p]; p++; }
}

}  // namespace lvs

using namespace lvs;

struct lvs_pgo {
  int device = 0, solver = 0;
  cudaStream_t st = nullptr;
  bool own_stream = false;
  PgoDev D{};
  std::vector<void*> allocs;
  PgoScalars* h_sc = nullptr;
  int pcg_grid = 0;
  bool has_graph = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double pcg_tol_override = 0.0;   // 0: by solver kind
  int pcg_max_iter_override = 0;
  lvs::CholDevice chol;            // direct solver of the *_CHOL kinds (pgo_chol.cu)
  bool use_chol = false;
  long long chol_nnz_l = 0; double chol_flops = 0; int chol_fronts = 0, chol_levels = 0, chol_max_front = 0;
  int solve_launches = 0;          // kernel launches of the linear solves since the last optimize() started
  std::vector<lvs_pgo_iter_rec> trace;
  double floor_plane[4] = {0.0, 0.0, 1.0, 0.0};   // the fixed VertexPlane of the LVS_PGO_EDGE_SE3_PLANE rows (lvs_pgo_set_floor_plane)
};

namespace lvs {

template <typename T>
static int dev_alloc(lvs_pgo* h, T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  h->allocs.push_back((void*)*p);
  return LVS_OK;
}
template <typename T>
static int dev_upload(lvs_pgo* h, T** p, const std::vector<T>& v) {
  int rc = dev_alloc(h, p, v.size());
  if (rc) return rc;
  if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->st));
  return LVS_OK;
}

static void free_graph(lvs_pgo* h) {
  chol_free(h->chol);
  h->use_chol = false;
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->has_graph = false;
  memset(&h->D, 0, sizeof h->D);
}

static double __longlong_as_double_host(unsigned long long b) { double d; memcpy(&d, &b, sizeof d); return d; }

static int blocks_for(long long n) { return (int)std::max<long long>(1, (n + kPgoThreads - 1) / kPgoThreads); }

static int run_errors(lvs_pgo* h) {
  int nb = std::min(kMaxPartials, blocks_for(h->D.ne));
  pgo_errors_kernel<<<nb, kPgoThreads, 0, h->st>>>(h->D);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

static int run_linearize(lvs_pgo* h) {
  CUDA_TRY(cudaMemsetAsync(&h->D.sc->max_diag_bits, 0, sizeof(unsigned long long), h->st));
  pgo_linearize_kernel<<<(h->D.ne + 127) / 128, 128, 0, h->st>>>(h->D);
  CUDA_TRY(cudaGetLastError());
  const long long total = (long long)h->D.nfree * 36 + (long long)h->D.noff * 36 + (long long)h->D.nfree * 6;
  pgo_assemble_kernel<<<blocks_for(total), kPgoThreads, 0, h->st>>>(h->D);
  CUDA_TRY(cudaGetLastError());
  return LVS_OK;
}

static int run_pcg(lvs_pgo* h, double lambda, double tol, double prev_residual, int max_iter) {
  if (h->use_chol && tol <= 1e-20) {
    // direct solve: sparse block Cholesky (LinearSolverCholmod / LinearSolverCSparse in the reference)
    h->solve_launches--;             // callers count one launch per solve
    return chol_solve(h->chol, h->st, h->D.Hd, h->D.Ho, h->D.b, lambda, h->D.x, &h->D.sc->scale, &h->D.sc->pcg_ok, &h->solve_launches);
  }
  void* args[] = {(void*)&h->D, (void*)&lambda, (void*)&tol, (void*)&prev_residual, (void*)&max_iter};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)pgo_pcg_kernel, dim3(h->pcg_grid), dim3(kPgoThreads), args, 0, h->st));
  return LVS_OK;
}

static int fetch_scalars(lvs_pgo* h) {
  CUDA_TRY(cudaMemcpyAsync(h->h_sc, h->D.sc, sizeof(PgoScalars), cudaMemcpyDeviceToHost, h->st));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  return LVS_OK;
}

static void solver_tolerance(const lvs_pgo* h, double* tol, bool* carry_residual, int* max_iter) {
  const bool pcg_kind = h->solver == LVS_PGO_LM_PCG || h->solver == LVS_PGO_GN_PCG;
  *tol = pcg_kind ? 1e-6 : 1e-24;        // LinearSolverPCG::_tolerance ; "direct" kinds run the same iteration to round-off
  *carry_residual = pcg_kind;            // _absoluteTolerance = true with the previous solve's residual
  *max_iter = h->D.nfree * 6;            // _maxIter = -1 -> A.rows()
  if (h->pcg_tol_override > 0) *tol = h->pcg_tol_override;
  if (h->pcg_max_iter_override > 0) *max_iter = h->pcg_max_iter_override;
}

}  // namespace lvs

extern "C" {

int lvs_pgo_create(int solver, int device, void* stream, lvs_pgo_t** out) {
  if (!out) return fail(LVS_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (solver < LVS_PGO_LM_CHOL || solver > LVS_PGO_GN_PCG) return fail(LVS_ERR_INVALID_ARG, "unknown solver %d", solver);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { (void)cudaGetLastError(); return fail(LVS_ERR_NO_DEVICE, "no CUDA device"); }
  if (device < 0 || device >= ndev) return fail(LVS_ERR_NO_DEVICE, "device %d out of range (have %d)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(LVS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  lvs_pgo* h = new (std::nothrow) lvs_pgo();
  if (!h) return fail(LVS_ERR_OOM, "host allocation failed");
  h->device = device; h->solver = solver;
  if (stream) h->st = (cudaStream_t)stream;
  else {
    if ((e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    h->own_stream = true;
  }
  if ((e = cudaMallocHost(&h->h_sc, sizeof(PgoScalars))) != cudaSuccess) { lvs_pgo_destroy(h); return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__); }
  for (int i = 0; i < 4; i++) if ((e = cudaEventCreate(&h->ev[i])) != cudaSuccess) { lvs_pgo_destroy(h); return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pgo_pcg_kernel, kPgoThreads, 0));
  h->pcg_grid = std::max(1, std::min(kMaxPartials, prop.multiProcessorCount * std::max(1, std::min(per_sm, 2))));
  *out = h;
  return LVS_OK;
}

int lvs_pgo_destroy(lvs_pgo_t* h) {
  if (!h) return LVS_OK;
  cudaSetDevice(h->device);
  if (h->st) cudaStreamSynchronize(h->st);
  free_graph(h);
  if (h->h_sc) cudaFreeHost(h->h_sc);
  for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->own_stream && h->st) cudaStreamDestroy(h->st);
  (void)cudaGetLastError();
  delete h;
  return LVS_OK;
}

int lvs_pgo_set_graph(lvs_pgo_t* h, int n_vertices, const double* poses7, const uint8_t* fixed, int n_edges, const int32_t* ij, const double* meas7,
                      const double* info21, const double* huber_delta) {
  return lvs_pgo_set_graph_typed(h, n_vertices, poses7, fixed, n_edges, ij, meas7, info21, huber_delta, nullptr);
}

int lvs_pgo_set_graph_typed(lvs_pgo_t* h, int n_vertices, const double* poses7, const uint8_t* fixed, int n_edges, const int32_t* ij_in, const double* meas7,
                            const double* info21, const double* huber_delta, const int32_t* edge_type) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  const int32_t* ij = ij_in;
  std::vector<int32_t> ij_fixed;
  if (edge_type && n_edges > 0 && ij_in) {
    // a unary prior is carried as a self-edge (i, i): no off-diagonal block, one incidence on vertex i
    ij_fixed.assign(ij_in, ij_in + (size_t)n_edges * 2);
    for (int k = 0; k < n_edges; k++) {
      if (edge_type[k] < LVS_PGO_EDGE_SE3 || edge_type[k] > LVS_PGO_EDGE_SE3_PLANE) return fail(LVS_ERR_INVALID_ARG, "edge %d has unknown kind %d", k, edge_type[k]);
      if (edge_type[k] != LVS_PGO_EDGE_SE3) ij_fixed[2 * k + 1] = ij_fixed[2 * k];
    }
    ij = ij_fixed.data();
  }
  if (n_vertices < 0 || n_edges < 0 || (n_vertices > 0 && !poses7) || (n_edges > 0 && (!ij || !meas7 || !info21)))
    return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  free_graph(h);
  const int nv = n_vertices, ne = n_edges;
  for (int k = 0; k < ne; k++)
    if (ij[2 * k] < 0 || ij[2 * k] >= nv || ij[2 * k + 1] < 0 || ij[2 * k + 1] >= nv) return fail(LVS_ERR_INVALID_ARG, "edge %d references a vertex out of range", k);
  // ---- BlockSolver::buildStructure on the host: free vertices ascending id, unique upper off-diagonal blocks, incidence lists
  std::vector<int> hidx(nv, -1), free_vertex;
  for (int v = 0; v < nv; v++) if (!(fixed && fixed[v])) { hidx[v] = (int)free_vertex.size(); free_vertex.push_back(v); }
  const int nfree = (int)free_vertex.size();
  std::vector<std::pair<int, int>> blocks;   // (col, row) so that sorting groups by column like g2o's block columns
  for (int k = 0; k < ne; k++) {
    int a = hidx[ij[2 * k]], c = hidx[ij[2 * k + 1]];
    if (a < 0 || c < 0 || a == c) continue;
    blocks.push_back({std::max(a, c), std::min(a, c)});
  }
  std::sort(blocks.begin(), blocks.end());
  blocks.erase(std::unique(blocks.begin(), blocks.end()), blocks.end());
  const int noff = (int)blocks.size();
  if ((long long)nfree * 42 + (long long)noff * 36 >= (1ll << 32)) return fail(LVS_ERR_INVALID_ARG, "graph too large (the assembly kernels index its entries with 32 bits)");
  // cholmod_analyze_p / cs_schol once per structure (linear_solver_cholmod.h:271-338): ordering + symbolic factorisation.  Pure host work on
  // the block pattern alone, so it runs on a second host thread beside the incidence lists, allocations and uploads below and is collected
  // just before its result is uploaded (7 of set_graph's 14 ms at 5 000 vertices, 0.13 of 0.3 s at 50 000).
  const bool want_chol = (h->solver == LVS_PGO_LM_CHOL || h->solver == LVS_PGO_GN_CHOL) && nfree > 0;
  std::vector<int> off_ij;
  CholSymbolic sym;
  std::future<void> analysis;          // declared after what it reads and writes: an early return waits for the thread before those go away
  if (want_chol) {
    off_ij.resize((size_t)noff * 2);
    for (int o = 0; o < noff; o++) { off_ij[2 * o] = blocks[o].second; off_ij[2 * o + 1] = blocks[o].first; }
    const int* oij = off_ij.data();
    CholSymbolic* sp = &sym;
    analysis = std::async(std::launch::async, [nfree, noff, oij, sp]() { chol_analyze(nfree, noff, oij, *sp); });
  }
  std::vector<int2> edge_ij(ne), edge_h(ne);
  std::vector<unsigned char> edge_tr(ne, 0);
  std::vector<int> edge_off(ne, -1);
  std::vector<int> vcount(nfree + 1, 0), ocount(noff + 1, 0), rcount(nfree + 1, 0);
  for (int k = 0; k < ne; k++) {
    const int a = hidx[ij[2 * k]], c = hidx[ij[2 * k + 1]];
    edge_ij[k] = make_int2(ij[2 * k], ij[2 * k + 1]);
    edge_h[k] = make_int2(a, c);
    if (a >= 0) vcount[a + 1]++;
    if (c >= 0 && c != a) vcount[c + 1]++;
    if (a >= 0 && c >= 0 && a != c) {
      auto it = std::lower_bound(blocks.begin(), blocks.end(), std::make_pair(std::max(a, c), std::min(a, c)));
      edge_off[k] = (int)(it - blocks.begin());
      edge_tr[k] = a > c;
      ocount[edge_off[k] + 1]++;
    }
  }
  for (int v = 0; v < nfree; v++) vcount[v + 1] += vcount[v];
  for (int o = 0; o < noff; o++) ocount[o + 1] += ocount[o];
  std::vector<int> vinc(vcount[nfree]), oinc(ocount[noff]), vfill(vcount.begin(), vcount.end() - 1), ofill(ocount.begin(), ocount.end() - 1);
  for (int k = 0; k < ne; k++) {   // ascending edge order = g2o's active-edge order
    const int a = edge_h[k].x, c = edge_h[k].y;
    if (a >= 0) vinc[vfill[a]++] = k * 2;
    if (c >= 0 && c != a) vinc[vfill[c]++] = k * 2 + 1;
    if (edge_off[k] >= 0) oinc[ofill[edge_off[k]]++] = k;
  }
  // SpMV rows: lower (transposed) blocks by ascending column, the diagonal, upper blocks by ascending column
  for (int v = 0; v < nfree; v++) rcount[v + 1] = 1;
  for (auto& b : blocks) { rcount[b.second + 1]++; rcount[b.first + 1]++; }
  for (int v = 0; v < nfree; v++) rcount[v + 1] += rcount[v];
  std::vector<int> rcol(rcount[nfree]), rslot(rcount[nfree]);
  {
    std::vector<std::vector<std::pair<int, int>>> rows(nfree);
    for (int o = 0; o < noff; o++) {
      rows[blocks[o].second].push_back({blocks[o].first, o * 2});        // row = block row, col = block col, as stored
      rows[blocks[o].first].push_back({blocks[o].second, o * 2 + 1});    // transposed use
    }
    for (int v = 0; v < nfree; v++) {
      rows[v].push_back({v, -1});
      std::sort(rows[v].begin(), rows[v].end());
      int p = rcount[v];
      for (auto& e : rows[v]) { rcol[p] = e.first; rslot[p] = e.second; p++; }
    }
  }
  // ---- device arrays
  PgoDev& D = h->D;
  D.nv = nv; D.ne = ne; D.nfree = nfree; D.noff = noff;
  int rc;
  int *d_hidx, *d_free, *d_vptr, *d_vinc, *d_optr, *d_oinc, *d_rptr, *d_rcol, *d_rslot;
  int2 *d_eij, *d_eh;
  unsigned char* d_tr;
  if ((rc = dev_upload(h, &d_hidx, hidx)) || (rc = dev_upload(h, &d_free, free_vertex)) || (rc = dev_upload(h, &d_vptr, vcount)) ||
      (rc = dev_upload(h, &d_vinc, vinc)) || (rc = dev_upload(h, &d_optr, ocount)) || (rc = dev_upload(h, &d_oinc, oinc)) ||
      (rc = dev_upload(h, &d_rptr, rcount)) || (rc = dev_upload(h, &d_rcol, rcol)) || (rc = dev_upload(h, &d_rslot, rslot)) ||
      (rc = dev_upload(h, &d_eij, edge_ij)) || (rc = dev_upload(h, &d_eh, edge_h)) || (rc = dev_upload(h, &d_tr, edge_tr))) { free_graph(h); return rc; }
  D.hidx = d_hidx; D.free_vertex = d_free; D.vptr = d_vptr; D.vinc = d_vinc; D.optr = d_optr; D.oinc = d_oinc;
  D.rptr = d_rptr; D.rcol = d_rcol; D.rslot = d_rslot; D.edge_ij = d_eij; D.edge_h = d_eh; D.edge_tr = d_tr;
  Rt *d_Z, *d_Zinv;
  double *d_info, *d_huber, *d_stage_p = nullptr, *d_stage_m = nullptr, *d_stage_i = nullptr, *d_pm = nullptr;
  int* d_type = nullptr;
  std::vector<double> hub(ne, 0.0);
  if (huber_delta) for (int k = 0; k < ne; k++) hub[k] = huber_delta[k];
  if ((rc = dev_alloc(h, &D.pose, nv)) || (rc = dev_alloc(h, &D.pose_bak, nv)) || (rc = dev_alloc(h, &d_Z, ne)) || (rc = dev_alloc(h, &d_Zinv, ne)) ||
      (rc = dev_alloc(h, &d_info, (size_t)ne * 36)) || (rc = dev_upload(h, &d_huber, hub)) || (rc = dev_alloc(h, &D.ws, (size_t)ne * kEdgeWs)) ||
      (rc = dev_alloc(h, &D.err, (size_t)ne * 6)) || (rc = dev_alloc(h, &D.chi, ne)) || (rc = dev_alloc(h, &D.Hd, (size_t)nfree * 36)) ||
      (rc = dev_alloc(h, &D.Ho, (size_t)noff * 36)) || (rc = dev_alloc(h, &D.b, (size_t)nfree * 6)) || (rc = dev_alloc(h, &D.x, (size_t)nfree * 6)) ||
      (rc = dev_alloc(h, &D.J, (size_t)nfree * 36)) || (rc = dev_alloc(h, &D.r, (size_t)nfree * 6)) || (rc = dev_alloc(h, &D.d, (size_t)nfree * 6)) ||
      (rc = dev_alloc(h, &D.q, (size_t)nfree * 6)) || (rc = dev_alloc(h, &D.s, (size_t)nfree * 6)) || (rc = dev_alloc(h, &D.q2, (size_t)nfree * 6)) || (rc = dev_alloc(h, &D.partials, 2 * kMaxPartials)) ||
      (rc = dev_alloc(h, &D.ticket, 4)) || (rc = dev_alloc(h, &D.sc, 1)) || (rc = dev_alloc(h, &d_stage_p, (size_t)nv * 7)) ||
      (rc = dev_alloc(h, &d_stage_m, (size_t)ne * 7)) || (rc = dev_alloc(h, &d_stage_i, (size_t)ne * 21))) { free_graph(h); return rc; }
  if (edge_type) {
    std::vector<int> ty(edge_type, edge_type + ne);
    if ((rc = dev_upload(h, &d_type, ty)) || (rc = dev_alloc(h, &d_pm, (size_t)ne * 8))) { free_graph(h); return rc; }
  }
  D.Z = d_Z; D.Zinv = d_Zinv; D.info = d_info; D.huber = d_huber; D.edge_type = d_type; D.pm = d_pm;
  CUDA_TRY(cudaMemsetAsync(D.ticket, 0, 4 * sizeof(unsigned int), h->st));
  CUDA_TRY(cudaMemsetAsync(D.sc, 0, sizeof(PgoScalars), h->st));
  CUDA_TRY(cudaMemsetAsync(D.ws, 0, (size_t)std::max(ne, 1) * kEdgeWs * sizeof(double), h->st));
  if (nv) {
    CUDA_TRY(cudaMemcpyAsync(d_stage_p, poses7, (size_t)nv * 7 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    pgo_pack_poses_kernel<<<blocks_for(nv), kPgoThreads, 0, h->st>>>(d_stage_p, nv, D.pose);
    CUDA_TRY(cudaMemcpyAsync(D.pose_bak, D.pose, (size_t)nv * sizeof(Rt), cudaMemcpyDeviceToDevice, h->st));
  }
  if (ne) {
    CUDA_TRY(cudaMemcpyAsync(d_stage_m, meas7, (size_t)ne * 7 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CUDA_TRY(cudaMemcpyAsync(d_stage_i, info21, (size_t)ne * 21 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    pgo_pack_edges_kernel<<<blocks_for(ne), kPgoThreads, 0, h->st>>>(d_stage_m, d_stage_i, ne, d_Z, d_Zinv, d_info, d_type, d_pm, h->floor_plane[0], h->floor_plane[1], h->floor_plane[2],
                                                                     h->floor_plane[3]);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(h->st));
  if (want_chol) {
    analysis.get();
    if ((rc = chol_upload(sym, noff, h->chol, h->st))) { free_graph(h); return rc; }
    h->use_chol = true;
    h->chol_nnz_l = sym.nnz_l_blocks; h->chol_flops = sym.flops; h->chol_fronts = (int)sym.fronts.size();
    h->chol_levels = (int)sym.level_ptr.size() - 1; h->chol_max_front = sym.max_front;
  }
  h->has_graph = true;
  return LVS_OK;
}

int lvs_pgo_set_floor_plane(lvs_pgo_t* h, const double coeffs[4]) {
  if (!h || !coeffs) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (!((coeffs[0] * coeffs[0] + coeffs[1] * coeffs[1]) + coeffs[2] * coeffs[2] > 0.0)) return fail(LVS_ERR_INVALID_ARG, "the plane's normal is zero");
  for (int a = 0; a < 4; a++) h->floor_plane[a] = coeffs[a];
  return LVS_OK;
}

int lvs_pgo_set_poses(lvs_pgo_t* h, const double* poses7) {
  if (!h || !poses7) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (!h->has_graph) return fail(LVS_ERR_EMPTY_GRAPH, "set_graph not called");
  CUDA_TRY(cudaSetDevice(h->device));
  double* stage = nullptr;
  CUDA_TRY(cudaMalloc(&stage, (size_t)std::max(h->D.nv, 1) * 7 * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(stage, poses7, (size_t)h->D.nv * 7 * sizeof(double), cudaMemcpyHostToDevice, h->st);
  if (e == cudaSuccess && h->D.nv) pgo_pack_poses_kernel<<<blocks_for(h->D.nv), kPgoThreads, 0, h->st>>>(stage, h->D.nv, h->D.pose);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  cudaFree(stage);
  if (e != cudaSuccess) return cuda_fail(e, "set_poses", __FILE__, __LINE__);
  return LVS_OK;
}

int lvs_pgo_get_poses(lvs_pgo_t* h, double* poses7) {
  if (!h || !poses7) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  if (!h->has_graph) return fail(LVS_ERR_EMPTY_GRAPH, "set_graph not called");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->D.nv == 0) return LVS_OK;
  double* stage = nullptr;
  CUDA_TRY(cudaMalloc(&stage, (size_t)h->D.nv * 7 * sizeof(double)));
  pgo_unpack_poses_kernel<<<blocks_for(h->D.nv), kPgoThreads, 0, h->st>>>(h->D.pose, h->D.nv, stage);
  cudaError_t e = cudaMemcpyAsync(poses7, stage, (size_t)h->D.nv * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
  cudaFree(stage);
  if (e != cudaSuccess) return cuda_fail(e, "get_poses", __FILE__, __LINE__);
  return LVS_OK;
}

int lvs_pgo_set_solver_options(lvs_pgo_t* h, double pcg_tolerance, int pcg_max_iterations) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  h->pcg_tol_override = pcg_tolerance;
  h->pcg_max_iter_override = pcg_max_iterations;
  return LVS_OK;
}

int lvs_pgo_optimize(lvs_pgo_t* h, int max_iterations, lvs_pgo_stats* stats) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  lvs_pgo_stats local;
  if (!stats) stats = &local;
  memset(stats, 0, sizeof *stats);
  h->trace.clear();
  if (!h->has_graph || h->D.ne < 1) { stats->iterations = -1; stats->status = LVS_ERR_EMPTY_GRAPH; return fail(LVS_ERR_EMPTY_GRAPH, "graph has no edge"); }
  if (h->D.nfree < 1) { stats->iterations = -1; stats->status = LVS_ERR_EMPTY_GRAPH; return fail(LVS_ERR_EMPTY_GRAPH, "every vertex is fixed"); }
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  const bool lm = h->solver == LVS_PGO_LM_CHOL || h->solver == LVS_PGO_LM_PCG;
  double tol; bool carry; int pcg_max;
  solver_tolerance(h, &tol, &carry, &pcg_max);
  double prev_residual = -1.0;
  int launches = 0, lin_launches = 0, pcg_total = 0, trials_total = 0;
  h->solve_launches = 0;
  float lin_ms = 0, solve_ms = 0;
  CUDA_TRY(cudaEventRecord(h->ev[0], h->st));
  // graph->computeActiveErrors(); chi2 = graph->chi2()   (graph_slam.cpp:313-316)
  if ((rc = run_errors(h)) || (rc = fetch_scalars(h))) return rc;
  launches++;
  stats->chi2_before = h->h_sc->chi2_plain;
  double lambda = -1.0, ni = 2.0;
  const double tau = 1e-5, good_lo = 1. / 3., good_hi = 2. / 3.;
  int iters = 0;
  bool failed = false;
  double current_chi = h->h_sc->chi2_robust;
  for (int it = 0; it < max_iterations; it++) {
    lvs_pgo_iter_rec rec;
    memset(&rec, 0, sizeof rec);
    bool terminate = false;
    // computeActiveErrors + activeRobustChi2 + buildSystem
    CUDA_TRY(cudaEventRecord(h->ev[2], h->st));
    if ((rc = run_errors(h)) || (rc = run_linearize(h))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev[3], h->st));
    launches += 3; lin_launches += 2;
    if (lm) {
      if ((rc = fetch_scalars(h))) return rc;
      current_chi = h->h_sc->chi2_robust;
      if (it == 0) { lambda = tau * __longlong_as_double_host(h->h_sc->max_diag_bits); ni = 2; }
    } else {
      CUDA_TRY(cudaStreamSynchronize(h->st));
    }
    { float t = 0; cudaEventElapsedTime(&t, h->ev[2], h->ev[3]); lin_ms += t; }
    if (!lm) {
      CUDA_TRY(cudaEventRecord(h->ev[2], h->st));
      if ((rc = run_pcg(h, 0.0, tol, carry ? prev_residual : -1.0, pcg_max))) return rc;
      CUDA_TRY(cudaEventRecord(h->ev[3], h->st));
      pgo_update_kernel<<<blocks_for(h->D.nfree), kPgoThreads, 0, h->st>>>(h->D);
      if ((rc = run_errors(h)) || (rc = fetch_scalars(h))) return rc;
      launches += 3;
      { float t = 0; cudaEventElapsedTime(&t, h->ev[2], h->ev[3]); solve_ms += t; }
      prev_residual = 0.5 * h->h_sc->dn;
      rec.pcg_iterations = h->h_sc->pcg_iters; pcg_total += h->h_sc->pcg_iters;
      rec.trials = 1; trials_total++;
      rec.chi2 = h->h_sc->chi2_robust;
      if (!h->h_sc->pcg_ok || !std::isfinite(h->h_sc->chi2_robust)) { failed = true; terminate = true; }
    } else {
      double rho = 0;
      int qmax = 0;
      do {
        CUDA_TRY(cudaEventRecord(h->ev[2], h->st));
        if ((rc = run_pcg(h, lambda, tol, carry ? prev_residual : -1.0, pcg_max))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev[3], h->st));
        pgo_update_kernel<<<blocks_for(h->D.nfree), kPgoThreads, 0, h->st>>>(h->D);   // push() + update()
        if ((rc = run_errors(h)) || (rc = fetch_scalars(h))) return rc;
        launches += 3;
        { float t = 0; cudaEventElapsedTime(&t, h->ev[2], h->ev[3]); solve_ms += t; }
        prev_residual = 0.5 * h->h_sc->dn;
        rec.pcg_iterations += h->h_sc->pcg_iters; pcg_total += h->h_sc->pcg_iters;
        double temp_chi = h->h_sc->chi2_robust;
        if (!h->h_sc->pcg_ok) temp_chi = std::numeric_limits<double>::max();
        rho = current_chi - temp_chi;
        const double scale = h->h_sc->scale + 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(temp_chi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, good_hi);
          lambda *= std::max(good_lo, alpha);
          ni = 2;
          current_chi = temp_chi;               // discardTop()
        } else {
          lambda *= ni; ni *= 2;
          pgo_restore_kernel<<<blocks_for(h->D.nfree), kPgoThreads, 0, h->st>>>(h->D);   // pop()
          launches++;
          if (!std::isfinite(lambda)) break;
        }
        qmax++;
      } while (rho < 0 && qmax < 10);
      rec.trials = qmax; trials_total += qmax;
      rec.chi2 = current_chi;
      if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) terminate = true;
    }
    rec.lambda = lambda;
    h->trace.push_back(rec);
    iters++;
    if (terminate) break;
  }
  if ((rc = run_errors(h)) || (rc = fetch_scalars(h))) return rc;
  launches++;
  CUDA_TRY(cudaEventRecord(h->ev[1], h->st));
  CUDA_TRY(cudaStreamSynchronize(h->st));
  float total_ms = 0;
  cudaEventElapsedTime(&total_ms, h->ev[0], h->ev[1]);
  stats->iterations = failed ? 0 : iters;     // SparseOptimizer::optimize: 0 when the algorithm reports Fail
  stats->status = failed ? LVS_ERR_NOT_SPD : LVS_OK;
  stats->chi2_after = h->h_sc->chi2_plain;
  stats->robust_chi2_after = h->h_sc->chi2_robust;
  stats->lambda_final = lambda;
  stats->device_ms = total_ms; stats->linearize_ms = lin_ms; stats->solve_ms = solve_ms;
  stats->lm_trials = trials_total; stats->pcg_iterations = pcg_total; stats->launches = launches + h->solve_launches; stats->linearize_launches = lin_launches;
  return LVS_OK;
}

int lvs_pgo_get_trace(lvs_pgo_t* h, lvs_pgo_iter_rec* recs, int capacity, int* n_out) {
  if (!h || !n_out) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  *n_out = (int)h->trace.size();
  if (recs) for (int i = 0; i < std::min(capacity, *n_out); i++) recs[i] = h->trace[i];
  return LVS_OK;
}

// ---- parity taps
int lvs_pgo_compute_errors(lvs_pgo_t* h, double* err6, double* chi2, double* robust_total) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (!h->has_graph) return fail(LVS_ERR_EMPTY_GRAPH, "set_graph not called");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if (h->D.ne == 0) { if (robust_total) *robust_total = 0; return LVS_OK; }
  if ((rc = run_errors(h)) || (rc = fetch_scalars(h))) return rc;
  if (err6) CUDA_TRY(cudaMemcpy(err6, h->D.err, (size_t)h->D.ne * 6 * sizeof(double), cudaMemcpyDeviceToHost));
  if (chi2) CUDA_TRY(cudaMemcpy(chi2, h->D.chi, (size_t)h->D.ne * sizeof(double), cudaMemcpyDeviceToHost));
  if (robust_total) *robust_total = h->h_sc->chi2_robust;
  return LVS_OK;
}

int lvs_pgo_system_size(lvs_pgo_t* h, int* n_free, int* n_offdiag) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (n_free) *n_free = h->D.nfree;
  if (n_offdiag) *n_offdiag = h->D.noff;
  return LVS_OK;
}

int lvs_pgo_linearize(lvs_pgo_t* h, double* Hd, int32_t* off_ij, double* Ho, double* b) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (!h->has_graph) return fail(LVS_ERR_EMPTY_GRAPH, "set_graph not called");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = run_errors(h)) || (rc = run_linearize(h))) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->st));
  const PgoDev& D = h->D;
  if (Hd) CUDA_TRY(cudaMemcpy(Hd, D.Hd, (size_t)D.nfree * 36 * sizeof(double), cudaMemcpyDeviceToHost));
  if (Ho && D.noff) CUDA_TRY(cudaMemcpy(Ho, D.Ho, (size_t)D.noff * 36 * sizeof(double), cudaMemcpyDeviceToHost));
  if (b) CUDA_TRY(cudaMemcpy(b, D.b, (size_t)D.nfree * 6 * sizeof(double), cudaMemcpyDeviceToHost));
  if (off_ij && D.noff) {
    // reconstruct (row, col) of every off-diagonal block from the SpMV rows
    std::vector<int> rptr(D.nfree + 1);
    CUDA_TRY(cudaMemcpy(rptr.data(), D.rptr, (size_t)(D.nfree + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> rcol(rptr[D.nfree]), rslot(rptr[D.nfree]);
    CUDA_TRY(cudaMemcpy(rcol.data(), D.rcol, rcol.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(rslot.data(), D.rslot, rslot.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int v = 0; v < D.nfree; v++)
      for (int p = rptr[v]; p < rptr[v + 1]; p++)
        if (rslot[p] >= 0 && !(rslot[p] & 1)) { off_ij[2 * (rslot[p] >> 1)] = v; off_ij[2 * (rslot[p] >> 1) + 1] = rcol[p]; }
  }
  return LVS_OK;
}

int lvs_pgo_solve(lvs_pgo_t* h, double lambda, double tolerance, int max_iterations, double* x, int* iterations) {
  if (!h) return fail(LVS_ERR_INVALID_ARG, "handle is NULL");
  if (!h->has_graph || h->D.nfree < 1) return fail(LVS_ERR_EMPTY_GRAPH, "nothing to solve");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = run_pcg(h, lambda, tolerance, -1.0, max_iterations > 0 ? max_iterations : h->D.nfree * 6)) || (rc = fetch_scalars(h))) return rc;
  if (x) CUDA_TRY(cudaMemcpy(x, h->D.x, (size_t)h->D.nfree * 6 * sizeof(double), cudaMemcpyDeviceToHost));
  if (iterations) *iterations = h->h_sc->pcg_iters;
  return LVS_OK;
}

// Symbolic analysis of the direct solver, host only (no device needed): stats = {nnz(L) in 6x6 blocks, fronts, levels,
// largest front dimension, arena bytes, factorisation multiply-adds}; perm_out (may be NULL) receives the elimination order.
int lvs_pgo_chol_analyze(int n_blocks, int n_off, const int32_t* off_ij, long long stats[6], int32_t* perm_out) {
  if (n_blocks < 0 || n_off < 0 || (n_off > 0 && !off_ij) || !stats) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  for (int o = 0; o < n_off; o++)
    if (off_ij[2 * o] < 0 || off_ij[2 * o + 1] >= n_blocks || off_ij[2 * o] >= off_ij[2 * o + 1]) return fail(LVS_ERR_INVALID_ARG, "off block %d is not (row < col)", o);
  CholSymbolic sym;
  chol_analyze(n_blocks, n_off, off_ij, sym);
  stats[0] = sym.nnz_l_blocks; stats[1] = (long long)sym.fronts.size(); stats[2] = (long long)sym.level_ptr.size() - 1;
  stats[3] = sym.max_front; stats[4] = sym.arena * 8; stats[5] = (long long)sym.flops;
  if (perm_out) for (int c = 0; c < n_blocks; c++) perm_out[c] = sym.perm[c];
  return LVS_OK;
}

// Structure of the direct solver attached to this graph (zeros when the solver kind is iterative).
int lvs_pgo_chol_info(lvs_pgo_t* h, long long stats[6]) {
  if (!h || !stats) return fail(LVS_ERR_INVALID_ARG, "NULL argument");
  stats[0] = h->use_chol ? h->chol_nnz_l : 0; stats[1] = h->use_chol ? h->chol_fronts : 0; stats[2] = h->use_chol ? h->chol_levels : 0;
  stats[3] = h->use_chol ? h->chol_max_front : 0; stats[4] = h->use_chol ? h->chol.arena_doubles * 8 : 0; stats[5] = h->use_chol ? (long long)h->chol_flops : 0;
  return LVS_OK;
}

}  // extern "C"
