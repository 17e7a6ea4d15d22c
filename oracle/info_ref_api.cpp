// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points over the reference's OWN src/global_graph/information_matrix_calculator.cpp, compiled as it is (oracle/build_ref.sh) against
// stand-ins for the ROS / PCL / Eigen headers it includes (oracle/ref_stubs/): edge information matrices and the fitness score behind them.
#include <global_graph/information_matrix_calculator.hpp>
#include <cstring>

typedef pcl::PointCloud<pcl::PointXYZI> Cloud;
static Cloud::Ptr load(const float* xyz, size_t n, size_t stride) {
  Cloud::Ptr c(new Cloud());
  c->points.resize(n);
  for (size_t i = 0; i < n; i++) { c->points[i].x = xyz[i * stride]; c->points[i].y = xyz[i * stride + 1]; c->points[i].z = xyz[i * stride + 2]; c->points[i].intensity = 0.0f; }
  return c;
}
static Eigen::Isometry3d iso(const double* T16_rowmajor) {
  Eigen::Isometry3d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T.matrix()(r, c) = T16_rowmajor[r * 4 + c];
  return T;
}

extern "C" {

double iref_fitness_score(const float* xyz1, size_t n1, size_t s1, const float* xyz2, size_t n2, size_t s2, const double* relpose16, double max_range) {
  return lv_slam::InformationMatrixCalculator::calc_fitness_score(load(xyz1, n1, s1), load(xyz2, n2, s2), iso(relpose16), max_range);
}

// InformationMatrixCalculator(nh) with the given overrides of its parameters (name / value pairs), then calc_information_matrix
void iref_information_matrix(const float* xyz1, size_t n1, size_t s1, const float* xyz2, size_t n2, size_t s2, const double* relpose16, int n_params,
                             const char* const* names, const double* values, double* inf36) {
  ros::NodeHandle nh;
  for (int i = 0; i < n_params; i++) nh.values[names[i]] = values[i];
  lv_slam::InformationMatrixCalculator calc(nh);
  const Eigen::MatrixXd inf = calc.calc_information_matrix(load(xyz1, n1, s1), load(xyz2, n2, s2), iso(relpose16));
  for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) inf36[r * 6 + c] = inf(r, c);
}

}  // extern "C"
