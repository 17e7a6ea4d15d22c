// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Runs g2o's OWN Levenberg-Marquardt control flow over the restatement's building blocks.  oracle/build_ref.sh unpacks
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
  int levenbergIterations = 0;
  static G2OBatchStatistics* globalStats() { return nullptr; }
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

extern "C" {

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
