// Newton / More-Thuente control flow of one align() as a resumable state machine.
//
// The reference runs this loop on the host around its OpenMP derivative pass
// (include/ndt_omp/ndt_omp_impl2.hpp: computeTransformation :88-188, computeStepLengthMT :842-1003,
// updateIntervalMT :718-755, trialValueSelectionMT :759-838, auxilaryFunction_* ndt_omp.h:479-496).
// Here the whole state lives in HBM (AlignState) and is advanced by ONE thread of the last CTA that
// finishes an evaluation kernel, so an align() never returns to the host between evaluations: every
// evaluation launch reads "what to evaluate" from the state and leaves "what to evaluate next" behind.
#pragma once
#include "lvs_math.cuh"
#include "ndt_types.cuh"

namespace lvs {

LVS_HD double dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}

// ndt_omp.h:479-496
LVS_HD double mt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
LVS_HD double mt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

// ndt_omp_impl2.hpp:718-755
LVS_HD bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t,
                               double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

// ndt_omp_impl2.hpp:759-838
LVS_HD double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return (fabs(a_c - a_l) < fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return (fabs(a_c - a_t) >= fabs(a_s - a_t)) ? a_c : a_s;
  }
  if (fabs(g_t) <= fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_n = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    double lim = a_t + 0.66 * (a_u - a_t);
    return (a_t > a_l) ? fmin(lim, a_n) : fmax(lim, a_n);
  }
  double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
  double w = sqrt(z * z - g_t * g_u);
  return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

// Point the next evaluation at the tangent vector x (T = float(SE3::exp(x)), point Jacobians from x).
LVS_HD void state_set_eval_point(AlignState& s, const double* x) {
  Pose P = se3_exp(x);
  pose_to_matrix4f(P, s.T);
  quat_to_mat(P.q, s.Rd);
  for (int i = 0; i < 9; i++) s.Rj[i] = (float)s.Rd[i];
  for (int i = 0; i < 6; i++) s.x_t[i] = x[i];
}

// pcl::Registration::align prologue + the first lines of computeTransformation (:88-129): the state
// after this call asks for the derivative pass at the initial guess.
LVS_HD void align_state_init(AlignState& s, const float* guess16, int trace_on) {
  const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  bool differs = false;
  for (int i = 0; i < 16; i++) if (guess16[i] != I4[i]) differs = true;
  // output cloud = guess * input when guess != Identity, else the input itself; either way T * input
  for (int i = 0; i < 16; i++) { s.T[i] = differs ? guess16[i] : I4[i]; s.final_T[i] = s.T[i]; }
  matrix4f_to_se3_log(guess16, s.p);                 // Sophus::SE3(R, t).log()  (:119-120)
  Pose P = se3_exp(s.p);                             // the point Jacobian uses SE3::exp(p), not the guess itself
  quat_to_mat(P.q, s.Rd);
  for (int i = 0; i < 9; i++) s.Rj[i] = (float)s.Rd[i];
  for (int i = 0; i < 6; i++) { s.x_t[i] = s.p[i]; s.g[i] = 0; s.dir[i] = 0; }
  for (int i = 0; i < 36; i++) s.H[i] = 0;
  s.score = 0;
  s.nr_iterations = 0; s.converged = 0; s.n_eval = 0; s.n_hess = 0;
  s.delta_norm = 0; s.phi_0 = 0; s.d_phi_0 = 0; s.a_l = s.f_l = s.g_l = s.a_u = s.f_u = s.g_u = 0;
  s.a_t = 0; s.phi_t = s.d_phi_t = s.psi_t = s.d_psi_t = 0;
  s.open_interval = 1; s.interval_converged = 0; s.step_iterations = 0;
  s.trans_probability = 0;
  s.n_trace = 0; s.trace_on = trace_on;
  s.eval_kind = EVAL_DERIV_H;
  s.phase = PH_INIT;
}

struct StepOut { bool finished; };

#ifdef __CUDACC__
// The pseudo-inverse fallback is cold (H numerically singular): kept out of line so that its 36-element work arrays do not inflate the
// stack frame and the register pressure of the state machine around it.
static __device__ __noinline__ void svd6_solve_cold(const double* A, const double* b, double* x) { svd6_solve(A, b, x); }

// lu6_solve (lvs_math.cuh) carried out by one warp on an augmented 6x7 matrix in shared memory: the same compare-and-swap pivoting
// (through a row permutation instead of physical swaps), the same multiplier and the same element updates, each done by its own
// lane - bit-identical results, ~1 us instead of a 36-double register array spilling on a single thread.  M[i][6] = right-hand side;
// perm: 6 ints of shared memory; x: 6 doubles of shared memory.  Returns false (on every lane) where lu6_solve would.
__device__ __forceinline__ bool lu6_solve_warp(double (*M)[7], int* perm, double* x, int lane) {
  double amax = 0.0;
  for (int e = lane; e < 36; e += 32) amax = fmax(amax, fabs(M[e / 6][e % 6]));
  for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (!(amax > 0.0) || !(amax < 1.0e300)) return false;
  const double tiny = amax * 1e-10;
  if (lane < 6) perm[lane] = lane;
  __syncwarp();
  for (int k = 0; k < 6; k++) {
    {
      // pivot = the largest |entry| of column k among the remaining rows, the first one on a tie - the row lu6_solve's
      // compare-and-swap chain brings to position k (the order it leaves the OTHER rows in differs, which no result depends on:
      // every row's arithmetic is independent of its position and later pivots are chosen by value)
      const int pi = (lane >= k && lane < 6) ? perm[lane] : 0;
      double v = (lane >= k && lane < 6) ? fabs(M[pi][k]) : -1.0;
      int at = lane;
#pragma unroll
      for (int o = 4; o; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o, 8);
        const int oa = __shfl_down_sync(0xffffffffu, at, o, 8);
        if (ov > v || (ov == v && oa < at)) { v = ov; at = oa; }      // the lower index wins ties
      }
      at = __shfl_sync(0xffffffffu, at, 0);
      if (lane == 0 && at != k) { const int t = perm[k]; perm[k] = perm[at]; perm[at] = t; }
    }
    __syncwarp();
    const int pk = perm[k];
    const double piv = M[pk][k];
    if (!(fabs(piv) > tiny)) return false;
    const double inv = 1.0 / piv;
    const int w = 6 - k;                                   // columns k+1 .. 6 of the rows below
    if (lane < (5 - k) * w) {
      const int pi = perm[k + 1 + lane / w], j = k + 1 + lane % w;
      const double f = M[pi][k] * inv;
      M[pi][j] -= f * M[pk][j];
    }
    __syncwarp();
  }
  if (lane == 0) {
    for (int k = 5; k >= 0; k--) {
      const int pk = perm[k];
      double t = M[pk][6];
      for (int j = k + 1; j < 6; j++) t -= M[pk][j] * x[j];
      x[k] = t / M[pk][k];
    }
  }
  __syncwarp();
  return true;
}
#endif

LVS_HD bool mt_keep_searching(const AlignState& s) {   // loop condition at :920
  const double nu = 0.9;
  return !s.interval_converged && s.step_iterations < 10 && !(s.psi_t <= 0 && s.d_phi_t <= -nu * s.d_phi_0);
}

// One pass through the state machine after an evaluation has deposited (score, g, H) in the state.
// Returns true when the align is finished.  n_src = number of source points (trans_probability divisor).
// Device callers may hand in two results computed by other warps while this thread was busy elsewhere (eval_finish):
//   newton_pre   H^-1 (-g) of this evaluation by elimination (null: the elimination failed -> pseudo-inverse fallback here)
//   compose_pre  log(exp(dir * a_t) * exp(p)) for the dir, a_t and p the state held on entry, compose_a_t = that a_t (null: not computed)
LVS_HD bool align_state_advance(AlignState& s, const AlignConsts& c, int n_src, TraceRec* trace, const double* newton_pre = nullptr,
                                const double* compose_pre = nullptr, double compose_a_t = 0.0) {
  const double mu = 1.e-4;
  bool need_newton = false;   // start a new outer iteration (solve + line-search setup)
  bool step_done = false;     // computeStepLengthMT returned s.a_t

  switch (s.phase) {
    case PH_INIT:
      s.n_eval++;
      need_newton = true;
      break;
    case PH_MT_FIRST:
      s.n_eval++;
      s.phi_t = -s.score;
      s.d_phi_t = -dot6(s.g, s.dir);
      s.psi_t = mt_psi(s.a_t, s.phi_t, s.phi_0, s.d_phi_0, mu);
      s.d_psi_t = mt_dpsi(s.d_phi_t, s.d_phi_0, mu);
      break;
    case PH_MT_TRIAL:
      s.n_eval++;
      s.phi_t = -s.score;
      s.d_phi_t = -dot6(s.g, s.dir);
      s.psi_t = mt_psi(s.a_t, s.phi_t, s.phi_0, s.d_phi_0, mu);
      s.d_psi_t = mt_dpsi(s.d_phi_t, s.d_phi_0, mu);
      if (s.open_interval && (s.psi_t <= 0 && s.d_psi_t >= 0)) {
        s.open_interval = 0;
        s.f_l = s.f_l + s.phi_0 - mu * s.d_phi_0 * s.a_l; s.g_l = s.g_l + mu * s.d_phi_0;
        s.f_u = s.f_u + s.phi_0 - mu * s.d_phi_0 * s.a_u; s.g_u = s.g_u + mu * s.d_phi_0;
      }
      if (s.open_interval) s.interval_converged = mt_update_interval(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.psi_t, s.d_psi_t);
      else s.interval_converged = mt_update_interval(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.phi_t, s.d_phi_t);
      s.step_iterations++;
      break;
    case PH_HESS27:
      s.n_hess++;
      step_done = true;
      break;
    default:
      s.eval_kind = EVAL_NONE;
      return true;
  }

  if (s.phase == PH_MT_FIRST || s.phase == PH_MT_TRIAL) {
    if (mt_keep_searching(s)) {
      // next trial (:922-946)
      if (s.open_interval) s.a_t = mt_trial_value(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.psi_t, s.d_psi_t);
      else s.a_t = mt_trial_value(s.a_l, s.f_l, s.g_l, s.a_u, s.f_u, s.g_u, s.a_t, s.phi_t, s.d_phi_t);
      s.a_t = fmin(s.a_t, c.step_size);
      s.a_t = fmax(s.a_t, c.trans_eps / 2);
      double x[6];
      for (int i = 0; i < 6; i++) x[i] = s.p[i] + s.dir[i] * s.a_t;
      state_set_eval_point(s, x);
      for (int i = 0; i < 16; i++) s.final_T[i] = s.T[i];
      s.eval_kind = EVAL_DERIV_NOH;
      s.phase = PH_MT_TRIAL;
      return false;
    }
    if (s.step_iterations) {      // :999-1000 — Hessian at the accepted point, all-double radius-search pass
      s.eval_kind = EVAL_HESS27;
      s.phase = PH_HESS27;
      return false;
    }
    step_done = true;
  }

  for (int guard = 0; guard < 4096; guard++) {
    if (step_done) {
      // back in computeTransformation (:159-183)
      double delta[6], pn[6];
      if (compose_pre && s.a_t == compose_a_t) {       // dir and p cannot have changed since entry when a_t has not
        for (int i = 0; i < 6; i++) pn[i] = compose_pre[i];
        compose_pre = nullptr;                           // valid for the first composition of this call only
      } else {
        for (int i = 0; i < 6; i++) delta[i] = s.dir[i] * s.a_t;
        se3_log(se3_mul(se3_exp(delta), se3_exp(s.p)), pn);
      }
      if (trace && s.trace_on && s.n_trace < kMaxTrace) {
        TraceRec& t = trace[s.n_trace];
        for (int i = 0; i < 6; i++) { t.p_before[i] = s.p[i]; t.dir[i] = s.dir[i]; t.p_after[i] = pn[i]; }
        t.step = s.a_t; t.score = s.score; t.trials = s.step_iterations; t.hessian_recomputed = s.step_iterations ? 1 : 0;
      }
      if (s.trace_on) s.n_trace++;
      for (int i = 0; i < 6; i++) s.p[i] = pn[i];
      // pclomp_ground tests the step length in the first iteration too (ndt_ground_impl.hpp:173)
      if (s.nr_iterations > c.max_iter || ((s.nr_iterations || c.variant == LVS_NDT_GROUND) && (fabs(s.a_t) < c.trans_eps))) s.converged = 1;
      s.nr_iterations++;
      if (s.converged) {
        s.trans_probability = s.score / (double)n_src;
        s.eval_kind = EVAL_NONE; s.phase = PH_DONE;
        return true;
      }
      need_newton = true;
      step_done = false;
    }
    if (need_newton) {
      need_newton = false;
      // :138-152 — descent direction through the 6x6 SVD, negative gradient for maximisation
      double ng[6], dp[6];
      for (int i = 0; i < 6; i++) ng[i] = -s.g[i];
      // JacobiSVD::solve == H^-1 (-g) when H is numerically full rank (the normal case, cond ~1e4): elimination with partial
      // pivoting is ~50x cheaper on a single GPU thread; the one-sided Jacobi SVD keeps the pseudo-inverse semantics otherwise.
      // newton_pre (device): the caller's warp has already run the elimination on this evaluation's (H, -g); null = it failed
#ifdef __CUDA_ARCH__
      if (newton_pre) { for (int i = 0; i < 6; i++) dp[i] = newton_pre[i]; }
      else svd6_solve_cold(s.H, ng, dp);
#else
      if (!lu6_solve(s.H, ng, dp)) svd6_solve(s.H, ng, dp);
#endif
      double nrm = sqrt(dot6(dp, dp));
      if (nrm == 0 || nrm != nrm) {
        s.trans_probability = s.score / (double)n_src;
        s.converged = (nrm == nrm) ? 1 : 0;
        s.eval_kind = EVAL_NONE; s.phase = PH_DONE;
        return true;
      }
      for (int i = 0; i < 6; i++) s.dir[i] = dp[i] / nrm;
      s.delta_norm = nrm;
      // computeStepLengthMT prologue (:846-900)
      s.phi_0 = -s.score;
      s.d_phi_0 = -dot6(s.g, s.dir);
      s.step_iterations = 0;
      if (s.d_phi_0 >= 0) {
        if (s.d_phi_0 == 0) { s.a_t = 0; step_done = true; continue; }   // returns 0 without touching anything
        s.d_phi_0 *= -1;
        for (int i = 0; i < 6; i++) s.dir[i] *= -1;
      }
      s.a_l = 0; s.a_u = 0;
      s.f_l = mt_psi(s.a_l, s.phi_0, s.phi_0, s.d_phi_0, mu); s.g_l = mt_dpsi(s.d_phi_0, s.d_phi_0, mu);
      s.f_u = mt_psi(s.a_u, s.phi_0, s.phi_0, s.d_phi_0, mu); s.g_u = mt_dpsi(s.d_phi_0, s.d_phi_0, mu);
      s.interval_converged = ((c.step_size - c.trans_eps / 2) > 0) ? 1 : 0;   // sic (:891)
      s.open_interval = 1;
      s.a_t = nrm;
      s.a_t = fmin(s.a_t, c.step_size);
      s.a_t = fmax(s.a_t, c.trans_eps / 2);
      double x[6];
      for (int i = 0; i < 6; i++) x[i] = s.p[i] + s.dir[i] * s.a_t;
      state_set_eval_point(s, x);
      for (int i = 0; i < 16; i++) s.final_T[i] = s.T[i];
      s.eval_kind = EVAL_DERIV_H;
      // lean_final_evaluation: with the More-Thuente loop dead (interval_converged set above) the step length is final, so the test of
      // :175-179 can be made now; when it will end the align, only the score of the coming pass is ever read (:187) - skip its Hessian
      if (c.lean_final && s.interval_converged &&
          (s.nr_iterations > c.max_iter || ((s.nr_iterations || c.variant == LVS_NDT_GROUND) && (fabs(s.a_t) < c.trans_eps))))
        s.eval_kind = EVAL_DERIV_NOH;
      s.phase = PH_MT_FIRST;
      return false;
    }
  }
  s.eval_kind = EVAL_NONE; s.phase = PH_DONE;
  return true;
}

}  // namespace lvs
