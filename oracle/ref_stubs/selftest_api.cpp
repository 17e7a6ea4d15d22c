// ORACLE — TEST INFRASTRUCTURE ONLY.  Entry points that exercise the Eigen stand-in (eigen_min.h) on its own, so that tests/test_oracle_ndt.py can hold its
// semantics against numpy: the checkers compiled from the reference rest on this header doing what the Eigen operations of the same name do.
#include <Eigen/Core>

extern "C" {

// out: [0..8] A*B, [9..17] A^T*B, [18..26] A.inverse(), [27..29] A*v, [30] v.dot(w), [31..33] v.cross(w), [34] v.norm(), [35..43] outer v w^T
void est_matrix3(const double* a9, const double* b9, const double* v3, const double* w3, double* out) {
  Eigen::Matrix3d A, B;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A(r, c) = a9[r * 3 + c]; B(r, c) = b9[r * 3 + c]; }
  const Eigen::Vector3d v(v3[0], v3[1], v3[2]), w(w3[0], w3[1], w3[2]);
  const Eigen::Matrix3d AB = A * B, AtB = A.transpose() * B, Ai = A.inverse(), O = v * w.transpose();
  const Eigen::Vector3d Av = A * v, x = v.cross(w);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { out[r * 3 + c] = AB(r, c); out[9 + r * 3 + c] = AtB(r, c); out[18 + r * 3 + c] = Ai(r, c); out[35 + r * 3 + c] = O(r, c); }
  for (int i = 0; i < 3; i++) { out[27 + i] = Av(i); out[31 + i] = x(i); }
  out[30] = v.dot(w); out[34] = v.norm();
}

// quaternion (w x y z in) : out [0..8] toRotationMatrix, [9..12] Quaterniond(that matrix) as w x y z, [13..15] q._transformVector(v),
// [16..19] (q * p) as w x y z, [20..28] (AngleAxis(a, Z) * AngleAxis(b, Y)).toRotationMatrix()
void est_quaternion(const double* q4, const double* p4, const double* v3, double a, double b, double* out) {
  Eigen::Quaterniond q(q4[0], q4[1], q4[2], q4[3]), p(p4[0], p4[1], p4[2], p4[3]);
  const Eigen::Matrix3d R = q.toRotationMatrix();
  const Eigen::Quaterniond back(R), qp = q * p;
  const Eigen::Vector3d tv = q._transformVector(Eigen::Vector3d(v3[0], v3[1], v3[2]));
  const Eigen::Matrix3d AA = (Eigen::AngleAxisd(a, Eigen::Vector3d::UnitZ()) * Eigen::AngleAxisd(b, Eigen::Vector3d::UnitY())).toRotationMatrix();
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { out[r * 3 + c] = R(r, c); out[20 + r * 3 + c] = AA(r, c); }
  out[9] = back.w(); out[10] = back.x(); out[11] = back.y(); out[12] = back.z();
  for (int i = 0; i < 3; i++) out[13 + i] = tv(i);
  out[16] = qp.w(); out[17] = qp.x(); out[18] = qp.y(); out[19] = qp.z();
}

// isometries as row-major 4 x 4 : out [0..15] (A * B).matrix(), [16..31] A.inverse().matrix(); block views: [32..34] translation(), [35..43] linear()
void est_isometry(const double* a16, const double* b16, double* out) {
  Eigen::Isometry3d A, B;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { A.matrix()(r, c) = a16[r * 4 + c]; B.matrix()(r, c) = b16[r * 4 + c]; }
  const Eigen::Isometry3d AB = A * B, Ai = A.inverse();
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { out[r * 4 + c] = AB.matrix()(r, c); out[16 + r * 4 + c] = Ai.matrix()(r, c); }
  const Eigen::Vector3d t = A.translation();
  const Eigen::Matrix3d L = A.linear();
  for (int i = 0; i < 3; i++) { out[32 + i] = t(i); for (int c = 0; c < 3; c++) out[35 + i * 3 + c] = L(i, c); }
}

}  // extern "C"
