// SE(3) edge arithmetic of the pose-graph path, fp64, host + device.  Follows g2o a48ff8c (vendored by the reference as
// 3rdtools/g2o-a48ff8c.zip; paths inside the zip, g2o/g2o/...):
//   EdgeSE3::computeError                  types/slam3d/edge_se3.cpp:78-83
//   toVectorMQT / fromVectorMQT / normalize types/slam3d/isometry3d_mappings.cpp:38-44,77-99,117-122
//   computeEdgeSE3Gradient, skew / skewT   types/slam3d/isometry3d_gradients.h:43-84,193-263
//   compute_dq_dR + generated cases        types/slam3d/dquat2mat.cpp:35-83, dquat2mat_maxima_generated.cpp:27-237
//   VertexSE3::oplusImpl                   types/slam3d/vertex_se3.h:105-114
//   RobustKernelHuber::robustify           core/robust_kernel_impl.cpp:65-78
#pragma once
#include "lvs_math.cuh"

namespace lvs {

struct Rt {            // Eigen::Isometry3d as row-major rotation + translation (96 B)
  double R[9];
  double t[3];
};

LVS_HD Rt rt_mul(const Rt& a, const Rt& b) {
  Rt r;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = (a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j]) + a.R[i * 3 + 2] * b.R[6 + j];
    r.t[i] = ((a.R[i * 3] * b.t[0] + a.R[i * 3 + 1] * b.t[1]) + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
  }
  return r;
}

LVS_HD Rt rt_inv(const Rt& a) {
  Rt r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
#pragma unroll
  for (int i = 0; i < 3; i++) r.t[i] = -((r.R[i * 3] * a.t[0] + r.R[i * 3 + 1] * a.t[1]) + r.R[i * 3 + 2] * a.t[2]);
  return r;
}

LVS_HD Rt rt_from_qt7(const double* v) {   // x y z qx qy qz qw; the quaternion is normalised like EdgeSE3::read does
  Rt r;
  Q4 q = {v[6], v[3], v[4], v[5]};
  quat_normalize(q);
  quat_to_mat(q, r.R);
  r.t[0] = v[0]; r.t[1] = v[1]; r.t[2] = v[2];
  return r;
}

LVS_HD void rt_to_qt7(const Rt& a, double* v) {   // toVectorQT
  Q4 q = quat_from_mat(a.R);
  quat_normalize(q);
  v[0] = a.t[0]; v[1] = a.t[1]; v[2] = a.t[2]; v[3] = q.x; v[4] = q.y; v[5] = q.z; v[6] = q.w;
}

LVS_HD void rt_to_vector_mqt(const Rt& d, double* e) {
  Q4 q = quat_from_mat(d.R);
  quat_normalize(q);
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  e[0] = d.t[0]; e[1] = d.t[1]; e[2] = d.t[2]; e[3] = q.x; e[4] = q.y; e[5] = q.z;
}

LVS_HD Rt rt_from_vector_mqt(const double* v) {
  Rt r;
  double w = 1 - ((v[3] * v[3] + v[4] * v[4]) + v[5] * v[5]);
  if (w < 0) {
#pragma unroll
    for (int i = 0; i < 9; i++) r.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    Q4 q = {sqrt(w), v[3], v[4], v[5]};
    quat_to_mat(q, r.R);
  }
  r.t[0] = v[0]; r.t[1] = v[1]; r.t[2] = v[2];
  return r;
}

// e = toVectorMQT(Z^-1 * Xi^-1 * Xj)
LVS_HD void edge_error(const Rt& Zinv, const Rt& Xi, const Rt& Xj, double* e) { rt_to_vector_mqt(rt_mul(rt_mul(Zinv, rt_inv(Xi)), Xj), e); }

// ---- unary priors on a VertexSE3 (the reference's own edge types, include/g2o/edge_se3_prior{xy,xyz,quat,vec}.hpp; edge kinds
// LVS_PGO_EDGE_PRIOR_* of include/lvslam_b200.h).  Errors are zero-padded to 6 so that the 6-vector / 6 x 6 code of the binary edge is reused.
// setMeasurement: PriorQuat keeps w >= 0 (edge_se3_priorquat.hpp:52-57), PriorVec normalises direction and measurement
// (edge_se3_priorvec.hpp:50-53).  m = xy | xyz | qx qy qz qw | direction(3) measurement(3).
// Kind 5: EdgeSE3Plane (include/g2o/edge_se3_plane.hpp) against a FIXED VertexPlane - the floor constraint as the nodelet builds it: one plane node,
// fixed at creation (global_graph_nodelet.cpp:601-611) - i.e. a unary constraint on the pose.  pm = measured plane (4), the vertex's plane (4), both
// as g2o's Plane3D keeps them (g2o types/slam3d_addons/plane3d.h: scaled to a unit normal).
LVS_HD void plane_normalize(double* c) { const double n = sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]); const double k = 1. / n; for (int a = 0; a < 4; a++) c[a] = c[a] * k; }
LVS_HD double plane_azimuth(const double* v) { return atan2(v[1], v[0]); }
LVS_HD double plane_elevation(const double* v) { return atan2(v[2], sqrt(v[0] * v[0] + v[1] * v[1])); }
LVS_HD void plane_rotation(const double* n, double* R) {      // Plane3D::rotation: AngleAxis(azimuth, Z) * AngleAxis(-elevation, Y), a quaternion product in Eigen
  const double az = plane_azimuth(n), el = plane_elevation(n);
  const Q4 q1 = {cos(0.5 * az), 0.0, 0.0, sin(0.5 * az)}, q2 = {cos(-0.5 * el), 0.0, sin(-0.5 * el), 0.0};
  Q4 q;
  q.w = q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z;
  q.x = q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y;
  q.y = q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z;
  q.z = q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x;
  quat_to_mat(q, R);
}

LVS_HD void prior_set_measurement(int type, const double* m, double* pm, const double* floor_plane) {
  for (int a = 0; a < 8; a++) pm[a] = 0.0;
  if (type == 5) {
    for (int a = 0; a < 4; a++) { pm[a] = m[a]; pm[4 + a] = floor_plane[a]; }
    plane_normalize(pm); plane_normalize(pm + 4);
  }
  if (type == 1) { pm[0] = m[0]; pm[1] = m[1]; }
  else if (type == 2) { pm[0] = m[0]; pm[1] = m[1]; pm[2] = m[2]; }
  else if (type == 3) { const double sg = m[3] < 0.0 ? -1.0 : 1.0; for (int a = 0; a < 4; a++) pm[a] = sg * m[a]; }
  else if (type == 4) {
    for (int h = 0; h < 2; h++) {
      const double* v = m + 3 * h;
      const double n = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
      for (int a = 0; a < 3; a++) pm[3 * h + a] = v[a] / n;
    }
  }
}

// computeError (edge_se3_priorxy.hpp:41-46, priorxyz:41-46, priorquat:41-50, priorvec:41-50)
LVS_HD void prior_error(int type, const double* pm, const Rt& X, double* e) {
#pragma unroll
  for (int a = 0; a < 6; a++) e[a] = 0.0;
  if (type == 1) { e[0] = X.t[0] - pm[0]; e[1] = X.t[1] - pm[1]; }
  else if (type == 2) { e[0] = X.t[0] - pm[0]; e[1] = X.t[1] - pm[1]; e[2] = X.t[2] - pm[2]; }
  else if (type == 3) {
    Q4 q = quat_from_mat(X.R);                                 // Eigen::Quaterniond(linear()), not normalised
    const double dot = ((pm[0] * q.x + pm[1] * q.y) + pm[2] * q.z) + pm[3] * q.w;
    const double sg = dot < 0.0 ? -1.0 : 1.0;
    e[0] = sg * q.x - pm[0]; e[1] = sg * q.y - pm[1]; e[2] = sg * q.z - pm[2];
  } else if (type == 4) {
    // linear().inverse() * direction: Eigen's 3 x 3 inverse (cofactors over the determinant), not the transpose
    const double* m = X.R;
    double cof[9];
    cof[0] = m[4] * m[8] - m[5] * m[7]; cof[1] = m[5] * m[6] - m[3] * m[8]; cof[2] = m[3] * m[7] - m[4] * m[6];
    cof[3] = m[2] * m[7] - m[1] * m[8]; cof[4] = m[0] * m[8] - m[2] * m[6]; cof[5] = m[1] * m[6] - m[0] * m[7];
    cof[6] = m[1] * m[5] - m[2] * m[4]; cof[7] = m[2] * m[3] - m[0] * m[5]; cof[8] = m[0] * m[4] - m[1] * m[3];
    const double det = (m[0] * cof[0] + m[1] * cof[1]) + m[2] * cof[2];
    const double id = 1.0 / det;
#pragma unroll
    for (int r = 0; r < 3; r++) e[r] = (((cof[r] * id) * pm[0] + (cof[3 + r] * id) * pm[1]) + (cof[6 + r] * id) * pm[2]) - pm[3 + r];
  } else if (type == 5) {
    // local_plane = X^-1 * plane (operator*(Isometry3d, Plane3D)); error = local_plane.ominus(measurement) (edge_se3_plane.hpp:40-47)
    const Rt w2n = rt_inv(X);
    double lp[4];
#pragma unroll
    for (int r = 0; r < 3; r++) lp[r] = (w2n.R[r * 3] * pm[4] + w2n.R[r * 3 + 1] * pm[5]) + w2n.R[r * 3 + 2] * pm[6];
    lp[3] = pm[7] - ((w2n.t[0] * lp[0] + w2n.t[1] * lp[1]) + w2n.t[2] * lp[2]);
    plane_normalize(lp);
    double R[9], n[3];
    plane_rotation(lp, R);
#pragma unroll
    for (int r = 0; r < 3; r++) n[r] = (R[r] * pm[0] + R[3 + r] * pm[1]) + R[6 + r] * pm[2];
    e[0] = plane_azimuth(n); e[1] = plane_elevation(n); e[2] = (-lp[3]) - (-pm[3]);
  }
}

// BaseUnaryEdge::linearizeOplus (g2o core/base_unary_edge.hpp): central differences with delta = 1e-9 through VertexSE3::oplus.
// J row-major 6 x 6, rows past the edge's dimension zero.
LVS_HD void prior_jacobian(int type, const double* pm, const Rt& X, double* J) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  for (int d = 0; d < 6; d++) {
    double add[6] = {0, 0, 0, 0, 0, 0}, e1[6], e2[6];
    add[d] = delta;
    prior_error(type, pm, rt_mul(X, rt_from_vector_mqt(add)), e1);
    add[d] = -delta;
    prior_error(type, pm, rt_mul(X, rt_from_vector_mqt(add)), e2);
    for (int r = 0; r < 6; r++) J[r * 6 + d] = scalar * (e1[r] - e2[r]);
  }
}

LVS_HD double chi2_of(const double* info /*6x6*/, const double* e) {
  double s = 0;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double t = 0;
#pragma unroll
    for (int c = 0; c < 6; c++) t += info[r * 6 + c] * e[c];
    s += e[r] * t;
  }
  return s;
}

// rho[0] = rho(e), rho[1] = rho'(e)
LVS_HD void huber_rho(double e, double delta, double* rho0, double* rho1) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { *rho0 = e; *rho1 = 1.0; }
  else { const double sq = sqrt(e); *rho0 = 2 * sq * delta - dsqr; *rho1 = delta / sq; }
}

// dq/dR, 3x9 with respect to the COLUMN-major vec(R); R row-major here.
LVS_HD void compute_dq_dR(double* D /*[3][9]*/, const double* R) {
  const double r00 = R[0], r10 = R[3], r20 = R[6], r01 = R[1], r11 = R[4], r21 = R[7], r02 = R[2], r12 = R[5], r22 = R[8];
#pragma unroll
  for (int i = 0; i < 27; i++) D[i] = 0.0;
  const double tr = r00 + r11 + r22;
  double S, qw;
  int which;
  if (tr > 0) { S = sqrt(tr + 1.0) * 2; qw = 0.25 * S; which = 0; }
  else if ((r00 > r11) & (r00 > r22)) { S = sqrt(1.0 + r00 - r11 - r22) * 2; qw = (r21 - r12) / S; which = 1; }
  else if (r11 > r22) { S = sqrt(1.0 + r11 - r00 - r22) * 2; qw = (r02 - r20) / S; which = 2; }
  else { S = sqrt(1.0 + r22 - r00 - r11) * 2; qw = (r10 - r01) / S; which = 3; }
  S *= .25;
  const double i1 = 1 / S, i3 = 1 / (S * S * S);
#define DQ(r, c) D[(r) * 9 + (c)]
  if (which == 0) {
    const double a2 = -0.03125 * (r21 - r12) * i3, a4 = 0.25 * i1, a5 = -0.25 * i1, a6 = 0.03125 * (r20 - r02) * i3, a7 = -0.03125 * (r10 - r01) * i3;
    DQ(0, 0) = a2; DQ(0, 4) = a2; DQ(0, 5) = a4; DQ(0, 7) = a5; DQ(0, 8) = a2;
    DQ(1, 0) = a6; DQ(1, 2) = a5; DQ(1, 4) = a6; DQ(1, 6) = a4; DQ(1, 8) = a6;
    DQ(2, 0) = a7; DQ(2, 1) = a4; DQ(2, 3) = a5; DQ(2, 4) = a7; DQ(2, 8) = a7;
  } else if (which == 1) {
    const double a2 = -0.125 * i1, a4 = r10 + r01, a5 = 0.25 * i1, a6 = 0.03125 * i3 * a4, a7 = r20 + r02, a8 = 0.03125 * i3 * a7;
    DQ(0, 0) = 0.125 * i1; DQ(0, 4) = a2; DQ(0, 8) = a2;
    DQ(1, 0) = -0.03125 * i3 * a4; DQ(1, 1) = a5; DQ(1, 3) = a5; DQ(1, 4) = a6; DQ(1, 8) = a6;
    DQ(2, 0) = -0.03125 * i3 * a7; DQ(2, 2) = a5; DQ(2, 4) = a8; DQ(2, 6) = a5; DQ(2, 8) = a8;
  } else if (which == 2) {
    const double a2 = r10 + r01, a3 = 0.03125 * i3 * a2, a5 = 0.25 * i1, a6 = -0.125 * i1, a7 = r21 + r12, a8 = 0.03125 * i3 * a7;
    DQ(0, 0) = a3; DQ(0, 1) = a5; DQ(0, 3) = a5; DQ(0, 4) = -0.03125 * i3 * a2; DQ(0, 8) = a3;
    DQ(1, 0) = a6; DQ(1, 4) = 0.125 * i1; DQ(1, 8) = a6;
    DQ(2, 0) = a8; DQ(2, 4) = -0.03125 * i3 * a7; DQ(2, 5) = a5; DQ(2, 7) = a5; DQ(2, 8) = a8;
  } else {
    const double a2 = r20 + r02, a3 = 0.03125 * i3 * a2, a5 = 0.25 * i1, a6 = r21 + r12, a7 = 0.03125 * i3 * a6, a8 = -0.125 * i1;
    DQ(0, 0) = a3; DQ(0, 2) = a5; DQ(0, 4) = a3; DQ(0, 6) = a5; DQ(0, 8) = -0.03125 * i3 * a2;
    DQ(1, 0) = a7; DQ(1, 4) = a7; DQ(1, 5) = a5; DQ(1, 7) = a5; DQ(1, 8) = -0.03125 * i3 * a6;
    DQ(2, 0) = a8; DQ(2, 4) = a8; DQ(2, 8) = 0.125 * i1;
  }
#undef DQ
  if (qw <= 0) {
#pragma unroll
    for (int i = 0; i < 27; i++) D[i] = -D[i];
  }
}

// rot block of a Jacobian: J[3+r][3+c] = sum_k D[r][k] * vec_colmajor(Rl * S_c)[k]
LVS_HD void rot_block(const double* D, const double* Rl, const double* Sx, const double* Sy, const double* Sz, double* J) {
  double M[3][9];
  mat3_mul(Rl, Sx, M[0]); mat3_mul(Rl, Sy, M[1]); mat3_mul(Rl, Sz, M[2]);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 9; k++) s += D[r * 9 + k] * M[c][(k % 3) * 3 + (k / 3)];
      J[(3 + r) * 6 + 3 + c] = s;
    }
}

// computeEdgeSE3Gradient: Ji, Jj row-major 6x6.
LVS_HD void edge_gradient(const Rt& Z, const Rt& Xi, const Rt& Xj, double* Ji, double* Jj) {
  const Rt A = rt_inv(Z);
  const Rt B = rt_mul(rt_inv(Xi), Xj);
  const Rt E = rt_mul(A, B);
  double D[27];
  compute_dq_dR(D, E.R);
#pragma unroll
  for (int i = 0; i < 36; i++) { Ji[i] = 0.0; Jj[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { Ji[i * 6 + j] = -A.R[i * 3 + j]; Jj[i * 6 + j] = E.R[i * 3 + j]; }
  {
    const double x = 2 * B.t[0], y = 2 * B.t[1], z = 2 * B.t[2];
    const double S[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    double M[9];
    mat3_mul(A.R, S, M);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Ji[i * 6 + 3 + j] = M[i * 3 + j];
  }
  {
    const double* R = B.R;
    const double r11 = 2 * R[0], r12 = 2 * R[1], r13 = 2 * R[2], r21 = 2 * R[3], r22 = 2 * R[4], r23 = 2 * R[5], r31 = 2 * R[6], r32 = 2 * R[7], r33 = 2 * R[8];
    const double Sx[9] = {0, 0, 0, r31, r32, r33, -r21, -r22, -r23};
    const double Sy[9] = {-r31, -r32, -r33, 0, 0, 0, r11, r12, r13};
    const double Sz[9] = {r21, r22, r23, -r11, -r12, -r13, 0, 0, 0};
    rot_block(D, A.R, Sx, Sy, Sz, Ji);
  }
  {
    const double Sx[9] = {0, 0, 0, 0, 0, -2, 0, 2, 0};
    const double Sy[9] = {0, 0, 2, 0, 0, 0, -2, 0, 0};
    const double Sz[9] = {0, -2, 0, 2, 0, 0, 0, 0, 0};
    rot_block(D, E.R, Sx, Sy, Sz, Jj);
  }
}

// C = A^T * W * B, 6x6 row-major
LVS_HD void atwb(const double* A, const double* W, const double* B, double* C) {
  double WB[36];
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += W[r * 6 + k] * B[k * 6 + c];
      WB[r * 6 + c] = s;
    }
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += A[k * 6 + r] * WB[k * 6 + c];
      C[r * 6 + c] = s;
    }
}

// 6x6 inverse by Gauss-Jordan with partial pivoting (block-Jacobi preconditioner).  Static indices only.
LVS_HD bool inv6(const double* M, double* R) {
  double a[6][12];
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) { a[i][j] = M[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; k++) {
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      if (fabs(a[i][k]) > fabs(a[k][k])) {
#pragma unroll
        for (int j = 0; j < 12; j++) { double t = a[k][j]; a[k][j] = a[i][j]; a[i][j] = t; }
      }
    }
    if (a[k][k] == 0.0) ok = false;
    const double inv = 1.0 / a[k][k];
#pragma unroll
    for (int j = 0; j < 12; j++) a[k][j] *= inv;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (i != k) {
        const double f = a[i][k];
#pragma unroll
        for (int j = 0; j < 12; j++) a[i][j] -= f * a[k][j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) R[i * 6 + j] = a[i][6 + j];
  return ok;
}

}  // namespace lvs
