// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Runs g2o's OWN Levenberg-Marquardt / Gauss-Newton control flow and its PCG linear solver over the restatement's building blocks.  oracle/build_ref.sh unpacks
// core/optimization_algorithm_levenberg.cpp from the reference's 3rdtools/g2o-a48ff8c.zip and oracle/extract_ref_functions.py writes the
// definitions of OptimizationAlgorithmLevenberg::solve, ::computeLambdaInit and ::computeScale (and OptimizationAlgorithmGaussNewton::solve from
// core/optimization_algorithm_gauss_newton.cpp) exactly as they stand into temporary files (G2O_LM_BODIES, G2O_GN_BODIES), compiled here into
// oracle/_ref/liblm_ref.so.  Those three functions decide everything about an LM run: the initial damping
// (tau * max diagonal), the trial loop with push / pop / discardTop, the gain ratio and its scale, the damping schedule, the termination.
// What they call is provided here, on top of oracle/pgo_oracle.cpp (included below, so this library is a second, independent copy of the oracle
// whose optimiser is g2o's code): the members and constructor values of optimization_algorithm_levenberg.{h,cpp:38-52}, a Solver and a
// SparseOptimizer that forward to the restatement's compute_errors / build_system / linear_solve / apply_update, and the outer loop of
// SparseOptimizer::optimize (sparse_optimizer.cpp:366-431).  tests/test_oracle_pgo.py compares whole runs.
#include "pgo_oracle.cpp"

#include <cassert>
#include <iostream>

namespace g2o {
using std::cerr;
using std::endl;

inline double get_monotonic_time() { return 0.0; }
inline bool g2o_isfinite(double x) { return std::isfinite(x); }
struct G2OBatchStatistics {
  double timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0;
  int levenbergIterations = 0, iterationsLinearSolver = 0;
  static G2OBatchStatistics*& current() { static G2OBatchStatistics* p = nullptr; return p; }
  static G2OBatchStatistics* globalStats() { return current(); }      // null except around gref_pcg_solve, which reads the iteration count from it
};
template <typename T>
struct Property {
  T v;
  const T& value() const { return v; }
  void setValue(const T& x) { v = x; }
};

struct OptimizableGraph {
  struct Vertex {
    PGO* g; int k;
    int dimension() const { return 6; }
    double hessian(int i, int j) const { return g->Hd[(size_t)k * 36 + i * 6 + j]; }
  };
};

// SparseOptimizer, as the three functions use it
class SparseOptimizer {
 public:
  PGO* g = nullptr;
  std::vector<std::vector<Iso> > stack;
  std::vector<OptimizableGraph::Vertex> verts;
  std::vector<OptimizableGraph::Vertex*> index;
  void attach(PGO* p) { g = p; verts.clear(); index.clear(); for (int k = 0; k < p->nfree; k++) verts.push_back({p, k}); for (auto& v : verts) index.push_back(&v); }
  void computeActiveErrors() { compute_errors(*g); }
  double activeRobustChi2() const { return robust_chi2(*g); }
  void push() { stack.push_back(g->X); }
  void pop() { g->X = stack.back(); stack.pop_back(); }
  void discardTop() { stack.pop_back(); }
  void update(const double*) { apply_update(*g); }            // the argument is the solver's x(), which apply_update reads
  bool terminate() const { return false; }
  const std::vector<OptimizableGraph::Vertex*>& indexMapping() const { return index; }
};

// Solver (BlockSolver) over the restatement's normal equations: setLambda only records the damping, solve() hands it to linear_solve
class Solver {
 public:
  PGO* g = nullptr;
  SparseOptimizer* opt = nullptr;
  int kind = SOLVER_DENSE;
  double lambda = 0;
  SparseOptimizer* optimizer() const { return opt; }
  bool buildStructure() { return true; }                       // done by the caller (build_structure)
  bool buildSystem() { build_system(*g); return true; }
  bool setLambda(double l, bool) { lambda = l; return true; }
  void restoreDiagonal() {}
  bool solve() { return linear_solve(*g, lambda, kind, -1.0, -1); }
  double* x() { return g->x.data(); }
  const double* b() const { return g->b.data(); }
  size_t vectorSize() const { return (size_t)g->nfree * 6; }
  bool schur() const { return false; }
};

class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
};

// members of OptimizationAlgorithmWithHessian / OptimizationAlgorithmLevenberg (optimization_algorithm_levenberg.h), constructor values of
// optimization_algorithm_levenberg.cpp:38-52
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmLevenberg(Solver& s) : _solver(s) {
    _currentLambda = -1.;
    _tau = 1e-5;
    _goodStepUpperScale = 2. / 3.;
    _goodStepLowerScale = 1. / 3.;
    _userLambdaInit = &_initialLambda; _initialLambda.v = 0.;
    _maxTrialsAfterFailure = &_maxTrials; _maxTrials.v = 10;
    _ni = 2.;
    _levenbergIterations = 0;
  }
  SolverResult solve(int iteration, bool online = false);
  double computeLambdaInit() const;
  double computeScale() const;
  double currentLambda() const { return _currentLambda; }
  int levenbergIteration() const { return _levenbergIterations; }
  SparseOptimizer* _optimizer = nullptr;
  Solver& _solver;

 protected:
  Property<int>* _maxTrialsAfterFailure;
  Property<double>* _userLambdaInit;
  double _currentLambda, _tau, _goodStepLowerScale, _goodStepUpperScale, _ni;
  int _levenbergIterations;
  Property<double> _initialLambda;
  Property<int> _maxTrials;
};

// OptimizationAlgorithmGaussNewton (optimization_algorithm_gauss_newton.{h,cpp}): solve() taken the same way
class OptimizationAlgorithmGaussNewton : public OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmGaussNewton(Solver& s) : _solver(s) {}
  SolverResult solve(int iteration, bool online = false);
  SparseOptimizer* _optimizer = nullptr;
  Solver& _solver;
};

#include G2O_LM_BODIES
#include G2O_GN_BODIES

}  // namespace g2o

// ---------------------------------------------------------------------------------------------------------------------
// g2o's own LinearSolverPCG<MatrixType>::solve, ::multDiag (both overloads) and ::mult (solvers/pcg/linear_solver_pcg.hpp), taken the same
// way (G2O_PCG_BODIES).  Written here: the class declaration and constructor values of linear_solver_pcg.h:40-104, a SparseBlockMatrix that
// hands out the upper blocks of the restatement's damped Hessian column by column, the three one-line block helpers internal::pcg_axy /
// pcg_axpy / pcg_atxpy (in the original they come in an MSVC, a fixed-size and a dynamic-size flavour; the fixed-size one is restated), a
// dynamic vector, and the 6 x 6 inverse of the preconditioner blocks (the restatement's own inv6: Eigen's is not restated anywhere).
#include <Eigen/Core>
namespace Eigen {
template <>
inline Matrix<double, 6, 6> Matrix<double, 6, 6>::inverse() const {
  double m[36], r[36];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) m[i * 6 + j] = (*this)(i, j);
  inv6(m, r);
  Matrix<double, 6, 6> o;
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) o(i, j) = r[i * 6 + j];
  return o;
}
}  // namespace Eigen

namespace g2o {
class VectorXD {      // Eigen::Matrix<double, Dynamic, 1> as the PCG loop uses it
 public:
  std::vector<double> v;
  void setZero(int n) { v.assign((size_t)n, 0.0); }
  int size() const { return (int)v.size(); }
  double dot(const VectorXD& o) const { double s = 0; for (size_t i = 0; i < v.size(); i++) s += v[i] * o.v[i]; return s; }
  VectorXD& operator-=(const VectorXD& o) { for (size_t i = 0; i < v.size(); i++) v[i] -= o.v[i]; return *this; }
  VectorXD operator+(const VectorXD& o) const { VectorXD r(*this); for (size_t i = 0; i < v.size(); i++) r.v[i] = v[i] + o.v[i]; return r; }
  template <int N> Eigen::Matrix<double, N, 1> segment(int off) const { Eigen::Matrix<double, N, 1> m; for (int i = 0; i < N; i++) m(i) = v[(size_t)off + i]; return m; }
  template <int N> struct Seg {
    VectorXD& x; int off;
    Seg& operator=(const Eigen::Matrix<double, N, 1>& m) { for (int i = 0; i < N; i++) x.v[(size_t)off + i] = m(i); return *this; }
    Seg& operator+=(const Eigen::Matrix<double, N, 1>& m) { for (int i = 0; i < N; i++) x.v[(size_t)off + i] += m(i); return *this; }
  };
  template <int N> Seg<N> segment(int off) { return Seg<N>{*this, off}; }
};
inline VectorXD operator*(double a, const VectorXD& x) { VectorXD r(x); for (size_t i = 0; i < r.v.size(); i++) r.v[i] = a * x.v[i]; return r; }
}  // namespace g2o
namespace Eigen {
template <>
class Map<g2o::VectorXD> {
 public:
  Map(double* p, int n) : p_(p), n_(n) {}
  void setZero() { for (int i = 0; i < n_; i++) p_[i] = 0.0; }
  Map& operator+=(const g2o::VectorXD& o) { for (int i = 0; i < n_; i++) p_[i] += o.v[(size_t)i]; return *this; }
  operator g2o::VectorXD() const { g2o::VectorXD r; r.v.assign(p_, p_ + n_); return r; }

 private:
  double* p_;
  int n_;
};
}  // namespace Eigen

namespace g2o {
template <typename MatrixType>
class SparseBlockMatrix {      // the part of core/sparse_block_matrix.h the PCG solver reads: upper blocks per block column, cumulative block ends
 public:
  typedef MatrixType SparseMatrixBlock;
  typedef std::map<int, SparseMatrixBlock*> IntBlockMap;
  std::vector<IntBlockMap> cols_;
  std::vector<int> ends_;
  std::vector<MatrixType> store_;
  const std::vector<IntBlockMap>& blockCols() const { return cols_; }
  const std::vector<int>& rowBlockIndices() const { return ends_; }
  const std::vector<int>& colBlockIndices() const { return ends_; }
  int rows() const { return ends_.empty() ? 0 : ends_.back(); }
  int cols() const { return rows(); }
};
template <typename MatrixType>
class LinearSolverPCG {
 public:
  LinearSolverPCG() { _tolerance = 1e-6; _verbose = false; _absoluteTolerance = true; _residual = -1.0; _maxIter = -1; }      // linear_solver_pcg.h:49-56
  bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b);
  int iterations = -1;

 protected:
  typedef std::vector<MatrixType, Eigen::aligned_allocator<MatrixType> > MatrixVector;
  typedef std::vector<const MatrixType*> MatrixPtrVector;
  double _tolerance, _residual;
  bool _absoluteTolerance, _verbose;
  int _maxIter;
  MatrixPtrVector _diag;
  MatrixVector _J;
  std::vector<std::pair<int, int> > _indices;
  MatrixPtrVector _sparseMat;
  void multDiag(const std::vector<int>& colBlockIndices, MatrixVector& A, const VectorXD& src, VectorXD& dest);
  void multDiag(const std::vector<int>& colBlockIndices, MatrixPtrVector& A, const VectorXD& src, VectorXD& dest);
  void mult(const std::vector<int>& colBlockIndices, const VectorXD& src, VectorXD& dest);
};
namespace internal {      // linear_solver_pcg.hpp:30-71, the fixed-size flavour
template <typename MatrixType>
inline void pcg_axy(const MatrixType& A, const VectorXD& x, int xoff, VectorXD& y, int yoff) {
  y.segment<MatrixType::RowsAtCompileTime>(yoff) = A * x.segment<MatrixType::ColsAtCompileTime>(xoff);
}
template <typename MatrixType>
inline void pcg_axpy(const MatrixType& A, const VectorXD& x, int xoff, VectorXD& y, int yoff) {
  y.segment<MatrixType::RowsAtCompileTime>(yoff) += A * x.segment<MatrixType::ColsAtCompileTime>(xoff);
}
template <typename MatrixType>
inline void pcg_atxpy(const MatrixType& A, const VectorXD& x, int xoff, VectorXD& y, int yoff) {
  y.segment<MatrixType::ColsAtCompileTime>(yoff) += A.transpose() * x.segment<MatrixType::RowsAtCompileTime>(xoff);
}
}  // namespace internal
#include G2O_PCG_BODIES
}  // namespace g2o

extern "C" {

// One LinearSolverPCG::solve of (H + lambda I) x = b on the graph's current linearisation (a fresh solver: no residual carried over)
int gref_pcg_solve(void* h, double lambda, double* x_out, int* iterations) {
  PGO& g = *(PGO*)h;
  build_structure(g);
  compute_errors(g);
  build_system(g);
  typedef Eigen::Matrix<double, 6, 6> M6;
  g2o::SparseBlockMatrix<M6> A;
  const int nb = g.nfree;
  A.cols_.resize((size_t)nb);
  A.store_.reserve((size_t)nb + g.off.size());
  for (int k = 0; k < nb; k++) A.ends_.push_back(6 * (k + 1));
  auto block = [&](const double* src, double damp) { M6 m; for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) m(r, c) = src[r * 6 + c] + (r == c ? damp : 0.0); A.store_.push_back(m); return &A.store_.back(); };
  for (size_t o = 0; o < g.off.size(); o++) A.cols_[(size_t)g.off[o].second][g.off[o].first] = block(&g.Ho[o * 36], 0.0);
  for (int k = 0; k < nb; k++) A.cols_[(size_t)k][k] = block(&g.Hd[(size_t)k * 36], lambda);
  g2o::LinearSolverPCG<M6> pcg;
  std::vector<double> x((size_t)nb * 6, 0.0), b(g.b);
  g2o::G2OBatchStatistics st;
  g2o::G2OBatchStatistics::current() = &st;
  const bool ok = pcg.solve(A, x.data(), b.data());
  g2o::G2OBatchStatistics::current() = nullptr;
  pcg.iterations = st.iterationsLinearSolver;
  std::memcpy(x_out, x.data(), x.size() * sizeof(double));
  if (iterations) *iterations = pcg.iterations;
  return ok ? 1 : 0;
}

// SparseOptimizer::optimize (sparse_optimizer.cpp:366-431) around g2o's own OptimizationAlgorithmLevenberg::solve.  Returns what the
// reference's GraphSLAM::optimize sees: iterations performed, 0 on Fail, -1 on an empty problem.  trace: per iteration (chi2 after the
// iteration, lambda after it, trials); stats: chi2 before, chi2 after, last lambda, total trials, robust chi2 after.
int gref_lm_optimize(void* h, int max_iters, int solver_kind, double* stats, double* trace3, int trace_cap, int* n_trace) {
  PGO& g = *(PGO*)h;
  *n_trace = 0;
  if (g.edges.empty()) return -1;
  build_structure(g);
  if (g.nfree == 0) return -1;                                 // _ivMap.size() == 0
  g.pcg_residual = -1.0;
  compute_errors(g);
  const double chi2_before = plain_chi2(g);
  g2o::SparseOptimizer opt;
  opt.attach(&g);
  g2o::Solver solver;
  solver.g = &g; solver.opt = &opt; solver.kind = solver_kind;
  g2o::OptimizationAlgorithmLevenberg alg(solver);
  alg._optimizer = &opt;
  int cjIterations = 0, total_trials = 0;
  bool ok = true;
  g2o::OptimizationAlgorithm::SolverResult result = g2o::OptimizationAlgorithm::OK;
  for (int i = 0; i < max_iters && !opt.terminate() && ok; i++) {
    result = alg.solve(i, false);
    ok = (result == g2o::OptimizationAlgorithm::OK);
    total_trials += alg.levenbergIteration();
    if (*n_trace < trace_cap) {
      compute_errors(g);
      trace3[*n_trace * 3] = robust_chi2(g); trace3[*n_trace * 3 + 1] = alg.currentLambda(); trace3[*n_trace * 3 + 2] = alg.levenbergIteration();
      (*n_trace)++;
    }
    ++cjIterations;
  }
  compute_errors(g);
  if (stats) { stats[0] = chi2_before; stats[1] = plain_chi2(g); stats[2] = alg.currentLambda(); stats[3] = total_trials; stats[4] = robust_chi2(g); }
  if (result == g2o::OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

// The same outer loop around g2o's own OptimizationAlgorithmGaussNewton::solve; trace3: (chi2 after the iteration, 0, 1)
int gref_gn_optimize(void* h, int max_iters, int solver_kind, double* stats, double* trace3, int trace_cap, int* n_trace) {
  PGO& g = *(PGO*)h;
  *n_trace = 0;
  if (g.edges.empty()) return -1;
  build_structure(g);
  if (g.nfree == 0) return -1;
  g.pcg_residual = -1.0;
  compute_errors(g);
  const double chi2_before = plain_chi2(g);
  g2o::SparseOptimizer opt;
  opt.attach(&g);
  g2o::Solver solver;
  solver.g = &g; solver.opt = &opt; solver.kind = solver_kind;
  g2o::OptimizationAlgorithmGaussNewton alg(solver);
  alg._optimizer = &opt;
  int cjIterations = 0;
  bool ok = true;
  g2o::OptimizationAlgorithm::SolverResult result = g2o::OptimizationAlgorithm::OK;
  for (int i = 0; i < max_iters && !opt.terminate() && ok; i++) {
    result = alg.solve(i, false);
    ok = (result == g2o::OptimizationAlgorithm::OK);
    if (*n_trace < trace_cap) {
      compute_errors(g);
      trace3[*n_trace * 3] = robust_chi2(g); trace3[*n_trace * 3 + 1] = 0; trace3[*n_trace * 3 + 2] = 1;
      (*n_trace)++;
    }
    ++cjIterations;
  }
  compute_errors(g);
  if (stats) { stats[0] = chi2_before; stats[1] = plain_chi2(g); stats[2] = 0; stats[3] = cjIterations; stats[4] = robust_chi2(g); }
  if (result == g2o::OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

}  // extern "C"
