// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or call
// anything under oracle/.
//
// PARITY: the reference (BurryChen/lv_slam) ships no test, fixture or golden vector for the pose-graph path and g2o as a whole cannot be
// built in this image (it needs Eigen).  The functions of g2o that this file restates are nevertheless taken from the
// reference's own 3rdtools/g2o-a48ff8c.zip at build time and compiled against interface stand-ins (oracle/build_ref.sh, oracle/ref_stubs/,
// oracle/g2o_ref_harness.cpp, oracle/lm_ref_harness.cpp, oracle/prior_ref_api.cpp, oracle/dquat_ref_api.cpp) as the CHECKERS of this file:
//   compute_dq_dR, computeEdgeSE3Gradient, the MQT mappings (EdgeSE3 error, both Jacobians, oplus), RobustKernelHuber::robustify   bit for bit
//   OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale, OptimizationAlgorithmGaussNewton::solve over this file's blocks  bit for bit
//   BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm (1e-13: another association of the products), ::linearizeOplus (numeric)
//   the reference's own edge_se3_prior{xy,xyz,quat,vec}.hpp and edge_se3_plane.hpp (with g2o's plane3d.h), CSparse
// (tests/test_oracle_pgo.py).  LinearSolverPCG::solve / multDiag / mult: same
// iteration counts and solutions.  Not pinned that way: BlockSolver's block bookkeeping, Eigen's rounding.  The older
// pins remain: (i) g2o's own property test restated in tests/ (analytic vs numeric EdgeSE3 Jacobian, test_slam3d_jacobian.cpp:109-140),
// (ii) closed-form small graphs, (iii) CSparse cross-checked against a dense Cholesky.
//
// CPU restatement of lv_slam::GraphSLAM::optimize (src/global_graph/graph_slam.cpp:298-331) over g2o a48ff8c
// (vendored as 3rdtools/g2o-a48ff8c.zip; paths below are inside the zip, g2o/g2o/...):
//   VertexSE3::oplusImpl                    types/slam3d/vertex_se3.h:105-114
//   EdgeSE3::computeError / linearizeOplus  types/slam3d/edge_se3.cpp:78-105
//   toVectorMQT / fromVectorMQT / normalize types/slam3d/isometry3d_mappings.cpp:38-44,77-99,117-122
//   computeEdgeSE3Gradient, skew/skewT      types/slam3d/isometry3d_gradients.h:43-84,193-263
//   compute_dq_dR (+ 4 generated cases)     types/slam3d/dquat2mat.cpp:35-83, dquat2mat_maxima_generated.cpp:27-237
//   BaseBinaryEdge::constructQuadraticForm  core/base_binary_edge.hpp:63-129
//   RobustKernelHuber::robustify            core/robust_kernel_impl.cpp:65-78 ; robustInformation core/base_edge.h:94-100
//   BlockSolver::buildSystem/setLambda      core/block_solver.hpp:463-566
//   OptimizationAlgorithmLevenberg::solve   core/optimization_algorithm_levenberg.cpp:58-175
//   OptimizationAlgorithmGaussNewton::solve core/optimization_algorithm_gauss_newton.cpp:50-92
//   SparseOptimizer::optimize / activeRobustChi2 / update   core/sparse_optimizer.cpp:102-116,366-446
//   LinearSolverCSparse::solve              solvers/csparse/linear_solver_csparse.h (cs_schol/cs_chol/cs_*solve of EXTERNAL/csparse)
//   LinearSolverPCG::solve                  solvers/pcg/linear_solver_pcg.hpp:80-158 (block-Jacobi PCG, tolerance 1e-6)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace {

struct Iso { double R[9]; double t[3]; };   // row-major rotation, translation

Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = (a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j]) + a.R[i * 3 + 2] * b.R[6 + j];
    r.t[i] = ((a.R[i * 3] * b.t[0] + a.R[i * 3 + 1] * b.t[1]) + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
  }
  return r;
}
Iso iso_inv(const Iso& a) {   // Eigen Isometry inverse: R^T, -R^T t
  Iso r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; i++) r.t[i] = -((r.R[i * 3] * a.t[0] + r.R[i * 3 + 1] * a.t[1]) + r.R[i * 3 + 2] * a.t[2]);
  return r;
}

struct Q { double w, x, y, z; };
Q quat_from_R(const double* m) {   // Eigen::Quaterniond(Matrix3d)
  Q q;
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0); q.w = 0.5 * t; t = 0.5 / t;
    q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * t; t = 0.5 / t;
    q.w = (m[k * 3 + j] - m[j * 3 + k]) * t;
    v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
void quat_to_R(const Q& q, double* r) {   // Eigen toRotationMatrix
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r[0] = 1 - (tyy + tzz); r[1] = txy - twz; r[2] = txz + twy;
  r[3] = txy + twz; r[4] = 1 - (txx + tzz); r[5] = tyz - twx;
  r[6] = txz - twy; r[7] = tyz + twx; r[8] = 1 - (txx + tyy);
}
void quat_normalize(Q& q) { double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); q.w /= n; q.x /= n; q.y /= n; q.z /= n; }

Iso iso_from_qt7(const double* v) {   // x y z qx qy qz qw, quaternion normalised (EdgeSE3::read / fromVectorQT)
  Iso r;
  Q q = {v[6], v[3], v[4], v[5]};
  quat_normalize(q);
  quat_to_R(q, r.R);
  r.t[0] = v[0]; r.t[1] = v[1]; r.t[2] = v[2];
  return r;
}
void iso_to_qt7(const Iso& a, double* v) {   // toVectorQT
  Q q = quat_from_R(a.R);
  quat_normalize(q);
  v[0] = a.t[0]; v[1] = a.t[1]; v[2] = a.t[2]; v[3] = q.x; v[4] = q.y; v[5] = q.z; v[6] = q.w;
}

// toVectorMQT (isometry3d_mappings.cpp:94-99)
void to_vector_mqt(const Iso& d, double* e) {
  Q q = quat_from_R(d.R);
  quat_normalize(q);
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  e[0] = d.t[0]; e[1] = d.t[1]; e[2] = d.t[2]; e[3] = q.x; e[4] = q.y; e[5] = q.z;
}
// fromVectorMQT (isometry3d_mappings.cpp:84-91,117-122)
Iso from_vector_mqt(const double* v) {
  Iso r;
  double w = 1 - ((v[3] * v[3] + v[4] * v[4]) + v[5] * v[5]);
  if (w < 0) { for (int i = 0; i < 9; i++) r.R[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  else { w = std::sqrt(w); Q q = {w, v[3], v[4], v[5]}; quat_to_R(q, r.R); }
  r.t[0] = v[0]; r.t[1] = v[1]; r.t[2] = v[2];
  return r;
}

// compute_dq_dR (dquat2mat.cpp:35-83): D is 3x9 with respect to the COLUMN-major vec(R) = (r00 r10 r20 r01 r11 r21 r02 r12 r22).
void compute_dq_dR(double D[3][9], const double* R /*row-major*/) {
  const double r00 = R[0], r10 = R[3], r20 = R[6], r01 = R[1], r11 = R[4], r21 = R[7], r02 = R[2], r12 = R[5], r22 = R[8];
  std::memset(D, 0, sizeof(double) * 27);
  double tr = r00 + r11 + r22, S, qw;
  int which;
  if (tr > 0) { S = std::sqrt(tr + 1.0) * 2; qw = 0.25 * S; which = 0; }
  else if ((r00 > r11) & (r00 > r22)) { S = std::sqrt(1.0 + r00 - r11 - r22) * 2; qw = (r21 - r12) / S; which = 1; }
  else if (r11 > r22) { S = std::sqrt(1.0 + r11 - r00 - r22) * 2; qw = (r02 - r20) / S; which = 2; }
  else { S = std::sqrt(1.0 + r22 - r00 - r11) * 2; qw = (r10 - r01) / S; which = 3; }
  S *= .25;
  if (which == 0) {
    double a1 = 1 / std::pow(S, 3), a2 = -0.03125 * (r21 - r12) * a1, a3 = 1 / S, a4 = 0.25 * a3, a5 = -0.25 * a3;
    double a6 = 0.03125 * (r20 - r02) * a1, a7 = -0.03125 * (r10 - r01) * a1;
    D[0][0] = a2; D[0][4] = a2; D[0][5] = a4; D[0][7] = a5; D[0][8] = a2;
    D[1][0] = a6; D[1][2] = a5; D[1][4] = a6; D[1][6] = a4; D[1][8] = a6;
    D[2][0] = a7; D[2][1] = a4; D[2][3] = a5; D[2][4] = a7; D[2][8] = a7;
  } else if (which == 1) {
    double a1 = 1 / S, a2 = -0.125 * a1, a3 = 1 / std::pow(S, 3), a4 = r10 + r01, a5 = 0.25 * a1, a6 = 0.03125 * a3 * a4, a7 = r20 + r02;
    double a8 = 0.03125 * a3 * a7;
    D[0][0] = 0.125 * a1; D[0][4] = a2; D[0][8] = a2;
    D[1][0] = -0.03125 * a3 * a4; D[1][1] = a5; D[1][3] = a5; D[1][4] = a6; D[1][8] = a6;
    D[2][0] = -0.03125 * a3 * a7; D[2][2] = a5; D[2][4] = a8; D[2][6] = a5; D[2][8] = a8;
  } else if (which == 2) {
    double a1 = 1 / std::pow(S, 3), a2 = r10 + r01, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4, a6 = -0.125 * a4, a7 = r21 + r12;
    double a8 = 0.03125 * a1 * a7;
    D[0][0] = a3; D[0][1] = a5; D[0][3] = a5; D[0][4] = -0.03125 * a1 * a2; D[0][8] = a3;
    D[1][0] = a6; D[1][4] = 0.125 * a4; D[1][8] = a6;
    D[2][0] = a8; D[2][4] = -0.03125 * a1 * a7; D[2][5] = a5; D[2][7] = a5; D[2][8] = a8;
  } else {
    double a1 = 1 / std::pow(S, 3), a2 = r20 + r02, a3 = 0.03125 * a1 * a2, a4 = 1 / S, a5 = 0.25 * a4, a6 = r21 + r12, a7 = 0.03125 * a1 * a6;
    double a8 = -0.125 * a4;
    D[0][0] = a3; D[0][2] = a5; D[0][4] = a3; D[0][6] = a5; D[0][8] = -0.03125 * a1 * a2;
    D[1][0] = a7; D[1][4] = a7; D[1][5] = a5; D[1][7] = a5; D[1][8] = -0.03125 * a1 * a6;
    D[2][0] = a8; D[2][4] = a8; D[2][8] = 0.125 * a4;
  }
  if (qw <= 0) for (int i = 0; i < 3; i++) for (int j = 0; j < 9; j++) D[i][j] *= -1;
}

void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
}

// computeEdgeSE3Gradient (isometry3d_gradients.h:193-263).  Ji, Jj row-major 6x6.
void edge_gradient(const Iso& Z, const Iso& Xi, const Iso& Xj, double* Ji, double* Jj, Iso* E_out) {
  const Iso A = iso_inv(Z);
  const Iso B = iso_mul(iso_inv(Xi), Xj);
  const Iso E = iso_mul(A, B);
  if (E_out) *E_out = E;
  double D[3][9];
  compute_dq_dR(D, E.R);
  std::memset(Ji, 0, 36 * sizeof(double));
  std::memset(Jj, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Ji[i * 6 + j] = -A.R[i * 3 + j]; Jj[i * 6 + j] = E.R[i * 3 + j]; }
  {   // dte/dqi = Ra * skewT(tb)
    const double x = 2 * B.t[0], y = 2 * B.t[1], z = 2 * B.t[2];
    const double S[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    double M[9];
    mat3_mul(A.R, S, M);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ji[i * 6 + 3 + j] = M[i * 3 + j];
  }
  auto fill_rot_block = [&](const double* Rl, const double* Sx, const double* Sy, const double* Sz, double* J) {
    double Mx[9], My[9], Mz[9];
    mat3_mul(Rl, Sx, Mx); mat3_mul(Rl, Sy, My); mat3_mul(Rl, Sz, Mz);
    const double* Ms[3] = {Mx, My, Mz};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 9; k++) s += D[r][k] * Ms[c][(k % 3) * 3 + (k / 3)];   // column-major vec of M_c
        J[(3 + r) * 6 + 3 + c] = s;
      }
  };
  {   // dre/dqi: skewT(Sx,Sy,Sz,Rb)
    const double* R = B.R;
    const double r11 = 2 * R[0], r12 = 2 * R[1], r13 = 2 * R[2], r21 = 2 * R[3], r22 = 2 * R[4], r23 = 2 * R[5], r31 = 2 * R[6], r32 = 2 * R[7], r33 = 2 * R[8];
    const double Sx[9] = {0, 0, 0, r31, r32, r33, -r21, -r22, -r23};
    const double Sy[9] = {-r31, -r32, -r33, 0, 0, 0, r11, r12, r13};
    const double Sz[9] = {r21, r22, r23, -r11, -r12, -r13, 0, 0, 0};
    fill_rot_block(A.R, Sx, Sy, Sz, Ji);
  }
  {   // dre/dqj: skew(Sx,Sy,Sz,Identity)
    const double Sx[9] = {0, 0, 0, 0, 0, -2, 0, 2, 0};
    const double Sy[9] = {0, 0, 2, 0, 0, 0, -2, 0, 0};
    const double Sz[9] = {0, -2, 0, 2, 0, 0, 0, 0, 0};
    fill_rot_block(E.R, Sx, Sy, Sz, Jj);
  }
}

// type 0: EdgeSE3.  Types 1-4: the reference's unary priors on a VertexSE3 (include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp; i == j), whose
// measurement is pm[] and whose D x D information sits in the top-left corner of info (D = 2 for xy, 3 otherwise; the rest is zero, so the
// 6-vector / 6 x 6 code paths of the binary edge are reused with zero padding).
// Type 5: EdgeSE3Plane (include/g2o/edge_se3_plane.hpp) against a FIXED VertexPlane - how the nodelet uses it: one floor plane node, fixed at
// creation (global_graph_nodelet.cpp:601-611) - which makes it a unary constraint on the pose; pm = measured plane (4), the vertex's plane (4).
enum { EDGE_SE3 = 0, PRIOR_XY = 1, PRIOR_XYZ = 2, PRIOR_QUAT = 3, PRIOR_VEC = 4, PRIOR_PLANE = 5 };
struct Edge { int i, j; Iso Z, Zinv; double info[36]; double huber; int type = EDGE_SE3; double pm[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };   // huber <= 0: no kernel

// g2o Plane3D (g2o!types/slam3d_addons/plane3d.h): coefficients scaled to a unit normal; rotation(n) = AngleAxis(azimuth, Z) * AngleAxis(-elevation, Y)
void plane_normalize(double* c) { const double n = std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]); const double k = 1. / n; for (int a = 0; a < 4; a++) c[a] = c[a] * k; }
double plane_azimuth(const double* v) { return std::atan2(v[1], v[0]); }
double plane_elevation(const double* v) { return std::atan2(v[2], std::sqrt(v[0] * v[0] + v[1] * v[1])); }
void plane_rotation(const double* n, double* R) {
  const double az = plane_azimuth(n), el = plane_elevation(n);
  // Eigen: AngleAxis * AngleAxis is a quaternion product, then toRotationMatrix
  const Q q1 = {std::cos(0.5 * az), 0.0, 0.0, std::sin(0.5 * az)}, q2 = {std::cos(-0.5 * el), 0.0, std::sin(-0.5 * el), 0.0};
  Q q;
  q.w = q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z;
  q.x = q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y;
  q.y = q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z;
  q.z = q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x;
  quat_to_R(q, R);
}

// setMeasurement of the prior edges: PriorQuat keeps w >= 0 (edge_se3_priorquat.hpp:52-57), PriorVec normalises direction and measurement
// (edge_se3_priorvec.hpp:50-53); meas6 = xy | xyz | qx qy qz qw | direction(3) measurement(3)
void prior_set_measurement(int type, const double* m, double* pm, const double* floor_plane = nullptr) {
  for (int a = 0; a < 8; a++) pm[a] = 0.0;
  if (type == PRIOR_PLANE) {                                   // Plane3D(coeffs) of the measurement and of the vertex estimate
    for (int a = 0; a < 4; a++) { pm[a] = m[a]; pm[4 + a] = floor_plane ? floor_plane[a] : (a == 2 ? 1.0 : 0.0); }
    plane_normalize(pm); plane_normalize(pm + 4);
  }
  if (type == PRIOR_XY) { pm[0] = m[0]; pm[1] = m[1]; }
  else if (type == PRIOR_XYZ) { pm[0] = m[0]; pm[1] = m[1]; pm[2] = m[2]; }
  else if (type == PRIOR_QUAT) { const double sg = m[3] < 0.0 ? -1.0 : 1.0; for (int a = 0; a < 4; a++) pm[a] = sg * m[a]; }
  else if (type == PRIOR_VEC) {
    for (int h = 0; h < 2; h++) {
      const double* v = m + 3 * h;
      const double n = std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
      for (int a = 0; a < 3; a++) pm[3 * h + a] = v[a] / n;
    }
  }
}

// computeError of the prior edges, zero-padded to 6 (edge_se3_priorxy.hpp:41-46, priorxyz:41-46, priorquat:41-50, priorvec:41-50)
void prior_error(int type, const double* pm, const Iso& X, double* e) {
  for (int a = 0; a < 6; a++) e[a] = 0.0;
  if (type == PRIOR_XY) { e[0] = X.t[0] - pm[0]; e[1] = X.t[1] - pm[1]; }
  else if (type == PRIOR_XYZ) { for (int a = 0; a < 3; a++) e[a] = X.t[a] - pm[a]; }
  else if (type == PRIOR_QUAT) {
    Q q = quat_from_R(X.R);                                  // Eigen::Quaterniond(linear()), not normalised
    const double dot = ((pm[0] * q.x + pm[1] * q.y) + pm[2] * q.z) + pm[3] * q.w;
    if (dot < 0.0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    e[0] = q.x - pm[0]; e[1] = q.y - pm[1]; e[2] = q.z - pm[2];
  } else if (type == PRIOR_VEC) {
    // linear().inverse() * direction: Eigen's 3 x 3 inverse (cofactors over the determinant), not the transpose
    const double* m = X.R;
    double cof[9];
    cof[0] = m[4] * m[8] - m[5] * m[7]; cof[1] = m[5] * m[6] - m[3] * m[8]; cof[2] = m[3] * m[7] - m[4] * m[6];
    cof[3] = m[2] * m[7] - m[1] * m[8]; cof[4] = m[0] * m[8] - m[2] * m[6]; cof[5] = m[1] * m[6] - m[0] * m[7];
    cof[6] = m[1] * m[5] - m[2] * m[4]; cof[7] = m[2] * m[3] - m[0] * m[5]; cof[8] = m[0] * m[4] - m[1] * m[3];
    const double det = (m[0] * cof[0] + m[1] * cof[1]) + m[2] * cof[2];
    const double id = 1.0 / det;
    // inverse[r][c] = cof[c][r] / det
    for (int r = 0; r < 3; r++) {
      const double v = ((cof[0 * 3 + r] * id) * pm[0] + (cof[1 * 3 + r] * id) * pm[1]) + (cof[2 * 3 + r] * id) * pm[2];
      e[r] = v - pm[3 + r];
    }
  } else if (type == PRIOR_PLANE) {
    // local_plane = X^-1 * plane (operator*(Isometry3d, Plane3D)); error = local_plane.ominus(measurement) (edge_se3_plane.hpp:40-47)
    const Iso w2n = iso_inv(X);
    double lp[4];
    for (int r = 0; r < 3; r++) lp[r] = (w2n.R[r * 3] * pm[4] + w2n.R[r * 3 + 1] * pm[5]) + w2n.R[r * 3 + 2] * pm[6];
    lp[3] = pm[7] - ((w2n.t[0] * lp[0] + w2n.t[1] * lp[1]) + w2n.t[2] * lp[2]);
    plane_normalize(lp);
    double R[9], n[3];
    plane_rotation(lp, R);                                     // ominus: rotation(normal()).transpose() * measurement.normal()
    for (int r = 0; r < 3; r++) n[r] = (R[r] * pm[0] + R[3 + r] * pm[1]) + R[6 + r] * pm[2];
    e[0] = plane_azimuth(n); e[1] = plane_elevation(n); e[2] = (-lp[3]) - (-pm[3]);
  }
}

// BaseUnaryEdge::linearizeOplus (g2o!core/base_unary_edge.hpp: central differences, delta = 1e-9, through VertexSE3::oplus): J row-major
// 6 x 6, rows past the edge's dimension zero.  (The 12 oplus calls count towards the vertex's 1000-call re-orthogonalisation in g2o; like
// the update itself that cadence is not modelled here.)
void prior_jacobian(int type, const double* pm, const Iso& X, double* J) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int a = 0; a < 36; a++) J[a] = 0.0;
  for (int d = 0; d < 6; d++) {
    double add[6] = {0, 0, 0, 0, 0, 0}, e1[6], e2[6];
    add[d] = delta;
    prior_error(type, pm, iso_mul(X, from_vector_mqt(add)), e1);
    add[d] = -delta;
    prior_error(type, pm, iso_mul(X, from_vector_mqt(add)), e2);
    for (int r = 0; r < 6; r++) J[r * 6 + d] = scalar * (e1[r] - e2[r]);
  }
}

// ---------------------------------------------------------------- CSparse (vendored by the reference, loaded when built)
struct cs { int nzmax, m, n; int* p; int* i; double* x; int nz; };
struct css { int* pinv; int* q; int* parent; int* cp; int* leftmost; int m2; double lnz, unz; };
struct csn { cs* L; cs* U; int* pinv; double* B; };
struct CSparseApi {
  void* lib = nullptr;
  css* (*schol)(int, const cs*) = nullptr;
  csn* (*chol)(const cs*, const css*) = nullptr;
  int (*ipvec)(const int*, const double*, double*, int) = nullptr;
  int (*pvec)(const int*, const double*, double*, int) = nullptr;
  int (*lsolve)(const cs*, double*) = nullptr;
  int (*ltsolve)(const cs*, double*) = nullptr;
  css* (*sfree)(css*) = nullptr;
  csn* (*nfree)(csn*) = nullptr;
  bool ok() const { return lib != nullptr; }
};
CSparseApi g_cs;

bool load_csparse(const char* path) {
  if (g_cs.ok()) return true;
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  CSparseApi a;
  a.lib = h;
  a.schol = (css * (*)(int, const cs*)) dlsym(h, "cs_schol");
  a.chol = (csn * (*)(const cs*, const css*)) dlsym(h, "cs_chol");
  a.ipvec = (int (*)(const int*, const double*, double*, int))dlsym(h, "cs_ipvec");
  a.pvec = (int (*)(const int*, const double*, double*, int))dlsym(h, "cs_pvec");
  a.lsolve = (int (*)(const cs*, double*))dlsym(h, "cs_lsolve");
  a.ltsolve = (int (*)(const cs*, double*))dlsym(h, "cs_ltsolve");
  a.sfree = (css * (*)(css*)) dlsym(h, "cs_sfree");
  a.nfree = (csn * (*)(csn*)) dlsym(h, "cs_nfree");
  if (!a.schol || !a.chol || !a.ipvec || !a.pvec || !a.lsolve || !a.ltsolve || !a.sfree || !a.nfree) { dlclose(h); return false; }
  g_cs = a;
  return true;
}

enum SolverKind { SOLVER_CSPARSE = 0, SOLVER_PCG = 1, SOLVER_DENSE = 2 };
enum Algorithm { ALG_LM = 0, ALG_GN = 1 };

struct IterRec { double chi2, lambda; int trials, pcg_iters; };

struct PGO {
  double floor_plane[4] = {0, 0, 1, 0};   // the fixed VertexPlane the EdgeSE3Plane rows refer to
  std::vector<Iso> X;
  std::vector<uint8_t> fixed;
  std::vector<Edge> edges;
  // structure (BlockSolver::buildStructure): free vertices in ascending id, upper-triangular unique blocks
  std::vector<int> hidx;                 // vertex -> block index or -1
  int nfree = 0;
  std::vector<std::pair<int, int>> off;  // unique (bi < bj) blocks, sorted by (column bj, row bi) like g2o's per-column maps
  std::map<std::pair<int, int>, int> off_index;
  std::vector<int> edge_off;             // edge -> off-diagonal slot or -1
  std::vector<uint8_t> edge_transposed;
  // values
  std::vector<double> Hd, Ho, b, x;      // [nfree][36], [noff][36] (row-major block (bi,bj)), [6 nfree], [6 nfree]
  std::vector<double> err, Ji, Jj;       // per edge
  std::vector<IterRec> trace;
  css* symbolic = nullptr;               // reused across solves like LinearSolverCSparse does
  double pcg_residual = -1.0;
  int last_pcg_iters = 0;
  long n_linearize = 0, n_solve = 0;
  ~PGO() { if (symbolic && g_cs.ok()) g_cs.sfree(symbolic); }
};

void build_structure(PGO& g) {
  const int nv = (int)g.X.size();
  g.hidx.assign(nv, -1);
  g.nfree = 0;
  for (int v = 0; v < nv; v++) if (!g.fixed[v]) g.hidx[v] = g.nfree++;
  g.off.clear(); g.off_index.clear();
  std::vector<std::pair<std::pair<int, int>, int>> tmp;   // ((col, row), dummy)
  for (const Edge& e : g.edges) {
    int a = g.hidx[e.i], c = g.hidx[e.j];
    if (a < 0 || c < 0 || a == c) continue;
    int bi = std::min(a, c), bj = std::max(a, c);
    tmp.push_back({{bj, bi}, 0});
  }
  std::sort(tmp.begin(), tmp.end());
  tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
  for (auto& t : tmp) { g.off_index[{t.first.second, t.first.first}] = (int)g.off.size(); g.off.push_back({t.first.second, t.first.first}); }
  g.edge_off.assign(g.edges.size(), -1);
  g.edge_transposed.assign(g.edges.size(), 0);
  for (size_t k = 0; k < g.edges.size(); k++) {
    int a = g.hidx[g.edges[k].i], c = g.hidx[g.edges[k].j];
    if (a < 0 || c < 0 || a == c) continue;
    g.edge_off[k] = g.off_index[{std::min(a, c), std::max(a, c)}];
    g.edge_transposed[k] = a > c;
  }
  g.Hd.assign((size_t)g.nfree * 36, 0.0);
  g.Ho.assign(g.off.size() * 36, 0.0);
  g.b.assign((size_t)g.nfree * 6, 0.0);
  g.x.assign((size_t)g.nfree * 6, 0.0);
  g.err.assign(g.edges.size() * 6, 0.0);
  g.Ji.assign(g.edges.size() * 36, 0.0);
  g.Jj.assign(g.edges.size() * 36, 0.0);
  if (g.symbolic && g_cs.ok()) { g_cs.sfree(g.symbolic); }
  g.symbolic = nullptr;
}

double edge_chi2(const Edge& e, const double* er) {   // _error.dot(information() * _error)
  double s = 0;
  for (int r = 0; r < 6; r++) {
    double t = 0;
    for (int c = 0; c < 6; c++) t += e.info[r * 6 + c] * er[c];
    s += er[r] * t;
  }
  return s;
}

void huber(double e, double delta, double rho[3]) {   // robust_kernel_impl.cpp:65-78
  double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else { double sq = std::sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}

void compute_errors(PGO& g) {   // SparseOptimizer::computeActiveErrors
  for (size_t k = 0; k < g.edges.size(); k++) {
    const Edge& e = g.edges[k];
    if (e.type != EDGE_SE3) { prior_error(e.type, e.pm, g.X[e.i], &g.err[k * 6]); continue; }
    Iso d = iso_mul(iso_mul(e.Zinv, iso_inv(g.X[e.i])), g.X[e.j]);
    to_vector_mqt(d, &g.err[k * 6]);
  }
}
double robust_chi2(const PGO& g) {   // activeRobustChi2
  double chi = 0;
  for (size_t k = 0; k < g.edges.size(); k++) {
    double c = edge_chi2(g.edges[k], &g.err[k * 6]);
    if (g.edges[k].huber > 0) { double rho[3]; huber(c, g.edges[k].huber, rho); chi += rho[0]; }
    else chi += c;
  }
  return chi;
}
double plain_chi2(const PGO& g) {   // OptimizableGraph::chi2
  double chi = 0;
  for (size_t k = 0; k < g.edges.size(); k++) chi += edge_chi2(g.edges[k], &g.err[k * 6]);
  return chi;
}

// C += A^T * W * B (6x6 row-major)
void atwb_add(const double* A, const double* W, const double* B, double* C) {
  double WB[36];
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < 6; k++) s += W[r * 6 + k] * B[k * 6 + c]; WB[r * 6 + c] = s; }
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < 6; k++) s += A[k * 6 + r] * WB[k * 6 + c]; C[r * 6 + c] += s; }
}

void build_system(PGO& g) {   // BlockSolver::buildSystem + constructQuadraticForm
  g.n_linearize++;
  std::fill(g.Hd.begin(), g.Hd.end(), 0.0);
  std::fill(g.Ho.begin(), g.Ho.end(), 0.0);
  std::fill(g.b.begin(), g.b.end(), 0.0);
  for (size_t k = 0; k < g.edges.size(); k++) {
    const Edge& e = g.edges[k];
    double* A = &g.Ji[k * 36];
    double* B = &g.Jj[k * 36];
    const bool unary = e.type != EDGE_SE3;
    if (unary) { prior_jacobian(e.type, e.pm, g.X[e.i], A); for (int q = 0; q < 36; q++) B[q] = 0.0; }
    else edge_gradient(e.Z, g.X[e.i], g.X[e.j], A, B, nullptr);
    const int a = g.hidx[e.i], c = unary ? -1 : g.hidx[e.j];
    if (a < 0 && c < 0) continue;
    const double* er = &g.err[k * 6];
    double w = 1.0;
    if (e.huber > 0) { double rho[3]; huber(edge_chi2(e, er), e.huber, rho); w = rho[1]; }
    double W[36], omega_r[6];
    for (int r = 0; r < 36; r++) W[r] = w * e.info[r];
    for (int r = 0; r < 6; r++) { double s = 0; for (int q = 0; q < 6; q++) s += e.info[r * 6 + q] * er[q]; omega_r[r] = -s * w; }
    if (a >= 0) {
      for (int r = 0; r < 6; r++) { double s = 0; for (int q = 0; q < 6; q++) s += A[q * 6 + r] * omega_r[q]; g.b[a * 6 + r] += s; }
      atwb_add(A, W, A, &g.Hd[(size_t)a * 36]);
      if (c >= 0 && a != c) {
        if (g.edge_transposed[k]) atwb_add(B, W, A, &g.Ho[(size_t)g.edge_off[k] * 36]);
        else atwb_add(A, W, B, &g.Ho[(size_t)g.edge_off[k] * 36]);
      }
    }
    if (c >= 0) {
      for (int r = 0; r < 6; r++) { double s = 0; for (int q = 0; q < 6; q++) s += B[q * 6 + r] * omega_r[q]; g.b[c * 6 + r] += s; }
      atwb_add(B, W, B, &g.Hd[(size_t)c * 36]);
    }
  }
}

// ---------------------------------------------------------------- linear solvers on (H + lambda I) x = b
bool solve_dense(PGO& g, double lambda) {
  const int n = g.nfree * 6;
  std::vector<double> M((size_t)n * n, 0.0);
  for (int v = 0; v < g.nfree; v++)
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) M[(size_t)(v * 6 + r) * n + v * 6 + c] = g.Hd[(size_t)v * 36 + r * 6 + c] + (r == c ? lambda : 0.0);
  for (size_t o = 0; o < g.off.size(); o++)
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
      double v = g.Ho[o * 36 + r * 6 + c];
      M[(size_t)(g.off[o].first * 6 + r) * n + g.off[o].second * 6 + c] = v;
      M[(size_t)(g.off[o].second * 6 + c) * n + g.off[o].first * 6 + r] = v;
    }
  // Cholesky M = L L^T (lower, in place)
  for (int j = 0; j < n; j++) {
    double d = M[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= M[(size_t)j * n + k] * M[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    M[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = M[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= M[(size_t)i * n + k] * M[(size_t)j * n + k];
      M[(size_t)i * n + j] = s / d;
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) { double s = g.b[i]; for (int k = 0; k < i; k++) s -= M[(size_t)i * n + k] * y[k]; y[i] = s / M[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= M[(size_t)k * n + i] * g.x[k]; g.x[i] = s / M[(size_t)i * n + i]; }
  return true;
}

bool solve_csparse(PGO& g, double lambda) {   // LinearSolverCSparse::solve: upper-triangular CCS, symbolic once, numeric each time
  if (!g_cs.ok()) return false;
  const int n = g.nfree * 6;
  // per block column: off-diagonal blocks (rows above) then the diagonal block; g.off is sorted by (column, row)
  std::vector<int> colstart(g.nfree + 1, 0);
  for (auto& o : g.off) colstart[o.second + 1]++;
  for (int c = 0; c < g.nfree; c++) colstart[c + 1] += colstart[c];
  std::vector<int> Ap(n + 1), Ai;
  std::vector<double> Ax;
  size_t nnz = g.off.size() * 36 + (size_t)g.nfree * 21;
  Ai.reserve(nnz); Ax.reserve(nnz);
  for (int bc = 0; bc < g.nfree; bc++)
    for (int c = 0; c < 6; c++) {
      Ap[bc * 6 + c] = (int)Ai.size();
      for (int o = colstart[bc]; o < colstart[bc + 1]; o++)
        for (int r = 0; r < 6; r++) { Ai.push_back(g.off[o].first * 6 + r); Ax.push_back(g.Ho[(size_t)o * 36 + r * 6 + c]); }
      for (int r = 0; r <= c; r++) { Ai.push_back(bc * 6 + r); Ax.push_back(g.Hd[(size_t)bc * 36 + r * 6 + c] + (r == c ? lambda : 0.0)); }
    }
  Ap[n] = (int)Ai.size();
  cs A;
  A.nzmax = (int)Ai.size(); A.m = n; A.n = n; A.p = Ap.data(); A.i = Ai.data(); A.x = Ax.data(); A.nz = -1;
  if (!g.symbolic) g.symbolic = g_cs.schol(1, &A);   // AMD ordering of A + A^T
  if (!g.symbolic) return false;
  csn* N = g_cs.chol(&A, g.symbolic);
  if (!N) return false;                               // not positive definite
  std::vector<double> w(n);
  g_cs.ipvec(g.symbolic->pinv, g.b.data(), w.data(), n);
  g_cs.lsolve(N->L, w.data());
  g_cs.ltsolve(N->L, w.data());
  g_cs.pvec(g.symbolic->pinv, w.data(), g.x.data(), n);
  g_cs.nfree(N);
  return true;
}

bool inv6(const double* M, double* R) {   // Eigen general 6x6 inverse is PartialPivLU based; Gauss-Jordan with partial pivoting here
  double a[6][12];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { a[i][j] = M[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int k = 0; k < 6; k++) {
    int p = k;
    for (int i = k + 1; i < 6; i++) if (std::fabs(a[i][k]) > std::fabs(a[p][k])) p = i;
    if (a[p][k] == 0.0) return false;
    if (p != k) for (int j = 0; j < 12; j++) std::swap(a[p][j], a[k][j]);
    double inv = 1.0 / a[k][k];
    for (int j = 0; j < 12; j++) a[k][j] *= inv;
    for (int i = 0; i < 6; i++) if (i != k) { double f = a[i][k]; if (f != 0.0) for (int j = 0; j < 12; j++) a[i][j] -= f * a[k][j]; }
  }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) R[i * 6 + j] = a[i][6 + j];
  return true;
}

// LinearSolverPCG::solve (linear_solver_pcg.hpp:80-158).  tol < 0 selects g2o's defaults (1e-6, absolute-tolerance carry-over).
bool solve_pcg(PGO& g, double lambda, double tol, int max_iter) {
  const int n = g.nfree * 6;
  std::vector<double> J((size_t)g.nfree * 36), D((size_t)g.nfree * 36);
  for (int v = 0; v < g.nfree; v++) {
    for (int k = 0; k < 36; k++) D[(size_t)v * 36 + k] = g.Hd[(size_t)v * 36 + k] + ((k % 7) == 0 ? lambda : 0.0);
    if (!inv6(&D[(size_t)v * 36], &J[(size_t)v * 36])) return false;
  }
  auto mult_diag = [&](const std::vector<double>& M, const std::vector<double>& s, std::vector<double>& d) {
    for (int v = 0; v < g.nfree; v++)
      for (int r = 0; r < 6; r++) { double t = 0; for (int c = 0; c < 6; c++) t += M[(size_t)v * 36 + r * 6 + c] * s[v * 6 + c]; d[v * 6 + r] = t; }
  };
  auto mult = [&](const std::vector<double>& s, std::vector<double>& d) {
    mult_diag(D, s, d);
    for (size_t o = 0; o < g.off.size(); o++) {
      const double* a = &g.Ho[o * 36];
      const int ro = g.off[o].first * 6, co = g.off[o].second * 6;
      for (int r = 0; r < 6; r++) { double t = 0; for (int c = 0; c < 6; c++) t += a[r * 6 + c] * s[co + c]; d[ro + r] += t; }
      for (int c = 0; c < 6; c++) { double t = 0; for (int r = 0; r < 6; r++) t += a[r * 6 + c] * s[ro + r]; d[co + c] += t; }
    }
  };
  std::vector<double> r(g.b), d(n, 0.0), q(n, 0.0), s(n, 0.0);
  std::fill(g.x.begin(), g.x.end(), 0.0);
  mult_diag(J, r, d);
  double dn = 0;
  for (int i = 0; i < n; i++) dn += r[i] * d[i];
  const bool g2o_defaults = tol < 0;
  double d0 = (g2o_defaults ? 1e-6 : tol) * dn;
  if (g2o_defaults && g.pcg_residual > 0.0 && g.pcg_residual > d0) d0 = g.pcg_residual;
  const int maxit = max_iter < 0 ? n : max_iter;
  int it;
  for (it = 0; it < maxit; ++it) {
    if (dn <= d0) break;
    mult(d, q);
    double dq = 0;
    for (int i = 0; i < n; i++) dq += d[i] * q[i];
    const double a = dn / dq;
    for (int i = 0; i < n; i++) { g.x[i] += a * d[i]; r[i] -= a * q[i]; }
    mult_diag(J, r, s);
    const double dold = dn;
    dn = 0;
    for (int i = 0; i < n; i++) dn += r[i] * s[i];
    const double ba = dn / dold;
    for (int i = 0; i < n; i++) d[i] = s[i] + ba * d[i];
  }
  g.pcg_residual = 0.5 * dn;
  g.last_pcg_iters = it;
  return true;
}

bool linear_solve(PGO& g, double lambda, int solver, double pcg_tol, int pcg_max_iter) {
  g.n_solve++;
  g.last_pcg_iters = 0;
  if (solver == SOLVER_CSPARSE) return solve_csparse(g, lambda);
  if (solver == SOLVER_PCG) return solve_pcg(g, lambda, pcg_tol, pcg_max_iter);
  return solve_dense(g, lambda);
}

void apply_update(PGO& g) {   // SparseOptimizer::update -> VertexSE3::oplusImpl (the 1000-call re-orthogonalisation is never reached here)
  for (size_t v = 0; v < g.X.size(); v++) {
    if (g.hidx[v] < 0) continue;
    g.X[v] = iso_mul(g.X[v], from_vector_mqt(&g.x[(size_t)g.hidx[v] * 6]));
  }
}

// SparseOptimizer::optimize driving OptimizationAlgorithmLevenberg / GaussNewton.  Returns the reference's return value:
// number of iterations performed, 0 when the algorithm reports Fail, -1 on an empty problem.
int optimize(PGO& g, int max_iters, int algorithm, int solver, double pcg_tol, int pcg_max_iter, double* stats) {
  g.trace.clear();
  g.pcg_residual = -1.0;
  if (g.edges.empty()) return -1;
  build_structure(g);
  if (g.nfree == 0) return -1;
  compute_errors(g);
  const double chi2_before = plain_chi2(g);
  double lambda = -1, ni = 2;
  const double tau = 1e-5, good_lo = 1. / 3., good_hi = 2. / 3.;
  int iters = 0, total_trials = 0;
  bool fail = false;
  for (int it = 0; it < max_iters; it++) {
    IterRec rec{0, 0, 0, 0};
    bool terminate = false;
    if (algorithm == ALG_GN) {
      compute_errors(g);
      build_system(g);
      bool ok = linear_solve(g, 0.0, solver, pcg_tol, pcg_max_iter);
      rec.pcg_iters = g.last_pcg_iters;
      if (!ok) { fail = true; iters++; g.trace.push_back(rec); break; }
      apply_update(g);
      compute_errors(g);
      rec.chi2 = robust_chi2(g);
      rec.trials = 1;
    } else {
      compute_errors(g);
      double current_chi = robust_chi2(g), temp_chi = current_chi;
      build_system(g);
      if (it == 0) {   // computeLambdaInit
        double mx = 0;
        for (int v = 0; v < g.nfree; v++) for (int j = 0; j < 6; j++) mx = std::max(std::fabs(g.Hd[(size_t)v * 36 + j * 7]), mx);
        lambda = tau * mx; ni = 2;
      }
      double rho = 0;
      int qmax = 0;
      do {
        std::vector<Iso> backup = g.X;                     // push()
        bool ok2 = linear_solve(g, lambda, solver, pcg_tol, pcg_max_iter);
        rec.pcg_iters += g.last_pcg_iters;
        apply_update(g);
        compute_errors(g);
        temp_chi = robust_chi2(g);
        if (!ok2) temp_chi = std::numeric_limits<double>::max();
        rho = current_chi - temp_chi;
        double scale = 0;
        for (int j = 0; j < g.nfree * 6; j++) scale += g.x[j] * (lambda * g.x[j] + g.b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(temp_chi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, good_hi);
          double sf = std::max(good_lo, alpha);
          lambda *= sf; ni = 2; current_chi = temp_chi;   // discardTop()
        } else {
          lambda *= ni; ni *= 2;
          g.X = backup;                                    // pop()
          if (!std::isfinite(lambda)) break;
        }
        qmax++;
      } while (rho < 0 && qmax < 10);
      rec.trials = qmax; rec.chi2 = current_chi;
      total_trials += qmax;
      if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) terminate = true;
    }
    rec.lambda = lambda;
    g.trace.push_back(rec);
    iters++;
    if (terminate) break;
  }
  compute_errors(g);
  if (stats) { stats[0] = chi2_before; stats[1] = plain_chi2(g); stats[2] = lambda; stats[3] = total_trials; stats[4] = robust_chi2(g); }
  return fail ? 0 : iters;
}

}  // namespace

extern "C" {

// tap: compute_dq_dR of a rotation matrix (row-major in, [3][9] row-major out), for the comparison with g2o's own generated code
void opgo_compute_dq_dR(const double* R9, double* D27) {
  double D[3][9];
  compute_dq_dR(D, R9);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 9; c++) D27[r * 9 + c] = D[r][c];
}

int opgo_load_csparse(const char* path) { return load_csparse(path) ? 1 : 0; }
int opgo_have_csparse() { return g_cs.ok() ? 1 : 0; }

void* opgo_create() { return new PGO(); }
void opgo_destroy(void* h) { delete (PGO*)h; }

// poses7 / meas7: x y z qx qy qz qw.  info21: upper triangle, row-major, as in g2o files.  huber_delta <= 0: no kernel.
// edge_type (may be NULL: all EdgeSE3): 0 EdgeSE3, 1-4 the unary priors, whose meas7 row carries xy | xyz | qx qy qz qw | direction +
// measurement and whose info21 row carries the D x D information in the top-left corner of the 6 x 6 upper triangle.
void opgo_set_graph_typed(void* h, int nv, const double* poses7, const uint8_t* fixed, int ne, const int32_t* ij, const double* meas7,
                          const double* info21, const double* huber_delta, const int32_t* edge_type) {
  PGO& g = *(PGO*)h;
  g.X.resize(nv); g.fixed.assign(nv, 0);
  for (int v = 0; v < nv; v++) { g.X[v] = iso_from_qt7(poses7 + 7 * v); if (fixed) g.fixed[v] = fixed[v]; }
  g.edges.resize(ne);
  for (int k = 0; k < ne; k++) {
    Edge& e = g.edges[k];
    e.i = ij[2 * k]; e.j = ij[2 * k + 1];
    e.type = edge_type ? edge_type[k] : EDGE_SE3;
    if (e.type != EDGE_SE3) {
      e.j = e.i;
      prior_set_measurement(e.type, meas7 + 7 * k, e.pm, g.floor_plane);
      const double ident[7] = {0, 0, 0, 0, 0, 0, 1};
      e.Z = iso_from_qt7(ident);
    } else e.Z = iso_from_qt7(meas7 + 7 * k);
    e.Zinv = iso_inv(e.Z);
    const double* u = info21 + 21 * k;
    int p = 0;
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { e.info[r * 6 + c] = u[p]; e.info[c * 6 + r] = u[p]; p++; }
    e.huber = huber_delta ? huber_delta[k] : 0.0;
  }
  build_structure(g);
}

void opgo_set_graph(void* h, int nv, const double* poses7, const uint8_t* fixed, int ne, const int32_t* ij, const double* meas7,
                    const double* info21, const double* huber_delta) {
  opgo_set_graph_typed(h, nv, poses7, fixed, ne, ij, meas7, info21, huber_delta, nullptr);
}
void opgo_set_floor_plane(void* h, const double* coeffs4) { for (int a = 0; a < 4; a++) ((PGO*)h)->floor_plane[a] = coeffs4[a]; }
// meas: the measurement, followed for PRIOR_PLANE (type 5) by the fixed plane's 4 coefficients at meas[4..7]
void opgo_prior_error(int type, const double* meas, const double* x7, double* e6) {
  double pm[8];
  prior_set_measurement(type, meas, pm, meas + 4);
  prior_error(type, pm, iso_from_qt7(x7), e6);
}
void opgo_prior_jacobian(int type, const double* meas, const double* x7, double* J36) {
  double pm[8];
  prior_set_measurement(type, meas, pm, meas + 4);
  prior_jacobian(type, pm, iso_from_qt7(x7), J36);
}

void opgo_get_poses(void* h, double* poses7) { PGO& g = *(PGO*)h; for (size_t v = 0; v < g.X.size(); v++) iso_to_qt7(g.X[v], poses7 + 7 * v); }
void opgo_get_poses_matrix(void* h, double* Rt12) { PGO& g = *(PGO*)h; for (size_t v = 0; v < g.X.size(); v++) { std::memcpy(Rt12 + 12 * v, g.X[v].R, 72); std::memcpy(Rt12 + 12 * v + 9, g.X[v].t, 24); } }

// errors [ne][6], per-edge chi2 [ne]; returns the robustified total (activeRobustChi2)
double opgo_compute_errors(void* h, double* err6, double* chi2) {
  PGO& g = *(PGO*)h;
  compute_errors(g);
  if (err6) std::memcpy(err6, g.err.data(), g.err.size() * 8);
  if (chi2) for (size_t k = 0; k < g.edges.size(); k++) chi2[k] = edge_chi2(g.edges[k], &g.err[k * 6]);
  return robust_chi2(g);
}

// Linearisation at the current estimate: per-edge Jacobians (row-major 6x6) and the assembled system.
void opgo_linearize(void* h, double* Ji, double* Jj) {
  PGO& g = *(PGO*)h;
  compute_errors(g);
  build_system(g);
  if (Ji) std::memcpy(Ji, g.Ji.data(), g.Ji.size() * 8);
  if (Jj) std::memcpy(Jj, g.Jj.data(), g.Jj.size() * 8);
}
int opgo_num_free(void* h) { return ((PGO*)h)->nfree; }
int opgo_num_offdiag(void* h) { return (int)((PGO*)h)->off.size(); }
void opgo_get_system(void* h, double* Hd, int32_t* off_ij, double* Ho, double* b) {
  PGO& g = *(PGO*)h;
  if (Hd) std::memcpy(Hd, g.Hd.data(), g.Hd.size() * 8);
  if (Ho) std::memcpy(Ho, g.Ho.data(), g.Ho.size() * 8);
  if (b) std::memcpy(b, g.b.data(), g.b.size() * 8);
  if (off_ij) for (size_t o = 0; o < g.off.size(); o++) { off_ij[2 * o] = g.off[o].first; off_ij[2 * o + 1] = g.off[o].second; }
}
// one linear solve of (H + lambda I) x = b on the last linearisation
int opgo_solve(void* h, double lambda, int solver, double pcg_tol, int pcg_max_iter, double* x) {
  PGO& g = *(PGO*)h;
  bool ok = linear_solve(g, lambda, solver, pcg_tol, pcg_max_iter);
  if (x) std::memcpy(x, g.x.data(), g.x.size() * 8);
  return ok ? 1 : 0;
}
int opgo_last_pcg_iters(void* h) { return ((PGO*)h)->last_pcg_iters; }
// nnz(L) (scalar entries) of CSparse's symbolic factorisation under its own AMD ordering, after a SOLVER_CSPARSE solve; -1 if none
double opgo_csparse_lnz(void* h) { PGO& g = *(PGO*)h; return g.symbolic ? g.symbolic->lnz : -1.0; }

// stats5: chi2 before (plain), chi2 after (plain), final lambda, LM trials, robust chi2 after
int opgo_optimize(void* h, int max_iters, int algorithm, int solver, double pcg_tol, int pcg_max_iter, double* stats5) {
  return optimize(*(PGO*)h, max_iters, algorithm, solver, pcg_tol, pcg_max_iter, stats5);
}
int opgo_trace_len(void* h) { return (int)((PGO*)h)->trace.size(); }
void opgo_get_trace(void* h, double* out /*[n][4]: chi2 lambda trials pcg_iters*/) {
  PGO& g = *(PGO*)h;
  for (size_t k = 0; k < g.trace.size(); k++) { out[4 * k] = g.trace[k].chi2; out[4 * k + 1] = g.trace[k].lambda; out[4 * k + 2] = g.trace[k].trials; out[4 * k + 3] = g.trace[k].pcg_iters; }
}

// stand-alone taps for the unit tests
void opgo_edge_error(const double* z7, const double* xi7, const double* xj7, double* e6) {
  Iso Z = iso_from_qt7(z7), Xi = iso_from_qt7(xi7), Xj = iso_from_qt7(xj7);
  to_vector_mqt(iso_mul(iso_mul(iso_inv(Z), iso_inv(Xi)), Xj), e6);
}
void opgo_edge_jacobians(const double* z7, const double* xi7, const double* xj7, double* Ji, double* Jj) {
  edge_gradient(iso_from_qt7(z7), iso_from_qt7(xi7), iso_from_qt7(xj7), Ji, Jj, nullptr);
}
void opgo_huber(double e, double delta, double* rho3) { huber(e, delta, rho3); }      // tap: the Huber kernel of the squared error
void opgo_oplus_matrix(const double* x7, const double* delta6, double* R9, double* t3) {      // the same, as rotation matrix + translation
  Iso r = iso_mul(iso_from_qt7(x7), from_vector_mqt(delta6));
  std::memcpy(R9, r.R, 72); std::memcpy(t3, r.t, 24);
}
void opgo_oplus(const double* x7, const double* delta6, double* out7) {
  Iso r = iso_mul(iso_from_qt7(x7), from_vector_mqt(delta6));
  iso_to_qt7(r, out7);
}

}  // extern "C"
