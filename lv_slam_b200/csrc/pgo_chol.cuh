// Sparse block Cholesky of the pose-graph normal equations — the direct solve behind the "*_var*" solver names.
// Replaces g2o's LinearSolverCholmod / LinearSolverCSparse::solve (3rdtools/g2o-a48ff8c.zip!g2o/g2o/solvers/cholmod/
// linear_solver_cholmod.h:115-154, solvers/csparse/linear_solver_csparse.h:126-307): fill-reducing ordering on the BLOCK pattern
// (g2o's blockorder = true), symbolic analysis once per graph, numeric factorisation + two triangular solves per LM trial.
//
// Method: supernodal multifrontal LL^T on 6x6 blocks.
//   host  (once per set_graph): minimum-degree ordering of the block graph (quotient graph, exact external degrees), elimination
//         tree, postorder, fundamental supernodes, frontal index sets, relative indices child -> parent, level sets;
//   device (per solve): every supernode owns a dense frontal matrix in one arena; fronts of one tree level are independent and are
//         processed together: extend-add of the children's update matrices (fixed child order: deterministic, no atomics), then a
//         dense partial Cholesky of the pivot columns.  The right-hand side travels as one extra row of every front, so the
//         forward substitution falls out of the factorisation; the backward substitution walks the levels top-down.
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace lvs {

struct CholFront {
  int c0, w, r;          // first pivot block column (permuted numbering), pivot block columns, row blocks below the pivots
  int parent;            // parent front or -1
  int level;             // 0 = leaf
  int F;                 // scalar dimension of the front: 6 (w + r) + 1 (the last row carries the right-hand side)
  int rows_off;          // offset of this front's row list R (and of its relative indices) in rows[] / rel[]
  int child_begin, child_end;   // range in child_idx[]
  long long off;         // offset of the front in the arena (doubles); column-major, leading dimension F
};

struct CholSymbolic {
  int n = 0;                               // block rows
  std::vector<int> perm, iperm;            // perm[new] = old, iperm[old] = new
  std::vector<CholFront> fronts;
  std::vector<int> rows, rel;              // R of every front (permuted block indices, ascending); position of each in the parent's front
  std::vector<int> child_idx;
  std::vector<int> level_ptr, level_fronts;   // fronts grouped by level
  std::vector<int> col_front;              // permuted block column -> front
  std::vector<long long> diag_dst, off_dst, rhs_dst;   // scatter targets (arena offsets) of H_vv, H_off and b
  std::vector<unsigned char> off_tr;       // off block has to be transposed into the lower triangle
  std::vector<int> diag_ld, off_ld;        // leading dimension of the target front
  long long arena = 0;                     // doubles
  long long nnz_l_blocks = 0;              // 6x6 blocks of L (diagonal blocks included)
  double flops = 0;                        // multiply-adds of the numeric factorisation (scalar)
  int max_front = 0;                       // largest F
};

// Host-side analysis.  off_ij: n_off pairs (row, col), row < col, of the structurally non-zero upper blocks.
void chol_analyze(int n, int n_off, const int* off_ij, CholSymbolic& S);

struct CholDevice {
  int n = 0, n_off = 0, n_fronts = 0, n_levels = 0, max_front = 0;
  CholFront* fronts = nullptr;
  int *rows = nullptr, *rel = nullptr, *child_idx = nullptr, *level_fronts = nullptr, *perm = nullptr, *col_front = nullptr;
  long long *diag_dst = nullptr, *off_dst = nullptr, *rhs_dst = nullptr;
  int *diag_ld = nullptr, *off_ld = nullptr;
  unsigned char* off_tr = nullptr;
  double* arena = nullptr;
  double* xp = nullptr;                    // solution in permuted order [6 n]
  int* fail_flag = nullptr;                // set when a pivot is not positive (the matrix is not SPD)
  long long* dbg = nullptr;                // LVS_DEBUG_TIMING: phase clocks of the last team front (diagnostics)
  long long arena_doubles = 0;
  std::vector<int> level_ptr;              // host copy
  std::vector<int> level_big;              // host: widest front of each level
  int *small_list = nullptr, *big_list = nullptr;      // fronts of every level split by size (device), with host offsets
  std::vector<int> small_ptr, big_ptr;
  unsigned int* bars = nullptr;            // team barrier counters
  double* back_scratch = nullptr;          // partial sums of the team backward substitution: [coop_grid][2][32]
  double* wscratch = nullptr;              // per team: inverses of the current panel's micro blocks (rank 0 publishes them)
  int coop_grid = 0;                       // CTAs of a cooperative launch (all co-resident)
  cudaStream_t side = nullptr;             // the single-CTA fronts of a level run beside its team launch
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<void*> allocs;
};

int chol_upload(const CholSymbolic& S, int n_off, CholDevice& C, cudaStream_t st);
void chol_free(CholDevice& C);
// (H + lambda I) x = b.  Hd [n][36], Ho [n_off][36] (upper blocks, row-major), b [6 n]; x [6 n] in the original numbering.
// scale_out (device, may be null) receives sum_j x_j (lambda x_j + b_j); *launches is incremented by the kernels issued.
int chol_solve(CholDevice& C, cudaStream_t st, const double* Hd, const double* Ho, const double* b, double lambda, double* x, double* scale_out,
               int* ok_out, int* launches);

}  // namespace lvs
