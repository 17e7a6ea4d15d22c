// lv_slam::GraphSLAM as far as optimize() needs it (include/global_graph/graph_slam.hpp:40-149) plus the solver_type_ member
// INTEGRATION.md asks the maintainer to add.
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include "../g2o/core/sparse_optimizer.h"
namespace lv_slam {
class GraphSLAM {
 public:
  int optimize(int num_iterations);
  std::unique_ptr<g2o::HyperGraph> graph;
  std::string solver_type_;
};
}  // namespace lv_slam
