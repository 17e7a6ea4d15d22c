// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in: the point type information_matrix_calculator.hpp names.
#pragma once
#include <Eigen/Core>
namespace pcl {
struct PointXYZI { float x, y, z, intensity; };
}
