// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry point over g2o's own compute_dq_dR (types/slam3d/dquat2mat.cpp with its Maxima-generated cases, unpacked from the reference's
// 3rdtools/g2o-a48ff8c.zip by oracle/build_ref.sh and compiled as they are): the derivative of the quaternion's vector part with respect to
// the column-major rotation matrix, the table behind EdgeSE3::linearizeOplus.  tests/test_oracle_pgo.py holds oracle/pgo_oracle.cpp to it.
#include "dquat2mat.h"

extern "C" void gref_compute_dq_dR(const double* R9_rowmajor, double* D27_rowmajor /* [3][9] */) {
  Eigen::Matrix<double, 3, 9, Eigen::ColMajor> D;
  D.setZero();
  const double* R = R9_rowmajor;
  // arguments r11 r21 r31 r12 ... : the matrix column by column
  g2o::internal::compute_dq_dR(D, R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 9; c++) D27_rowmajor[r * 9 + c] = D(r, c);
}
