// Reference-side binding for the pose-graph path: a replacement body for lv_slam::GraphSLAM::optimize
// (/root/reference/src/global_graph/graph_slam.cpp:298-331).  The g2o graph stays the container the rest of global_graph
// works with (KeyFrame holds g2o::VertexSE3*, the nodelet reads ->estimate()); only the solve moves: pack -> lvs_pgo_* -> setEstimate.
// Needs g2o + Eigen, which are NOT in the build image of this repository; compiled on the lv_slam side.
#include <global_graph/graph_slam.hpp>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>
#include <g2o/types/slam3d/edge_se3.h>
#include <g2o/types/slam3d/vertex_se3.h>
#include <algorithm>
#include <cstdint>
#include <iostream>
#include <map>
#include <string>
#include <vector>
#include "lvslam_b200.h"

namespace lv_slam {

static void to_qt7(const Eigen::Isometry3d& T, double* v) {
  Eigen::Quaterniond q(T.linear());
  q.normalize();
  v[0] = T.translation().x(); v[1] = T.translation().y(); v[2] = T.translation().z();
  v[3] = q.x(); v[4] = q.y(); v[5] = q.z(); v[6] = q.w();
}

static int solver_kind(const std::string& s) {   // the names GraphSLAM's constructor hands to g2o's factory
  const bool gn = s.compare(0, 2, "gn") == 0, pcg = s.find("pcg") != std::string::npos;
  return gn ? (pcg ? LVS_PGO_GN_PCG : LVS_PGO_GN_CHOL) : (pcg ? LVS_PGO_LM_PCG : LVS_PGO_LM_CHOL);
}

int GraphSLAM::optimize(int num_iterations) {
  g2o::SparseOptimizer* graph = dynamic_cast<g2o::SparseOptimizer*>(this->graph.get());
  if (graph->edges().size() < 1) return -1;
  // vertices in ascending id, edges in ascending internal id: the order g2o's initializeOptimization() establishes
  std::vector<g2o::VertexSE3*> vs;
  std::map<int, int> index;
  for (auto& kv : graph->vertices()) if (auto* v = dynamic_cast<g2o::VertexSE3*>(kv.second)) vs.push_back(v);
  std::sort(vs.begin(), vs.end(), [](g2o::VertexSE3* a, g2o::VertexSE3* b) { return a->id() < b->id(); });
  for (size_t i = 0; i < vs.size(); i++) index[vs[i]->id()] = (int)i;
  std::vector<g2o::EdgeSE3*> es;
  for (auto* e : graph->edges()) if (auto* s = dynamic_cast<g2o::EdgeSE3*>(e)) es.push_back(s);
  std::sort(es.begin(), es.end(), [](g2o::EdgeSE3* a, g2o::EdgeSE3* b) { return a->internalId() < b->internalId(); });
  std::vector<double> poses(7 * vs.size()), meas(7 * es.size()), info(21 * es.size()), huber(es.size(), 0.0);
  std::vector<uint8_t> fixed(vs.size());
  std::vector<int32_t> ij(2 * es.size());
  for (size_t i = 0; i < vs.size(); i++) { to_qt7(vs[i]->estimate(), &poses[7 * i]); fixed[i] = vs[i]->fixed(); }
  for (size_t k = 0; k < es.size(); k++) {
    ij[2 * k] = index[es[k]->vertices()[0]->id()];
    ij[2 * k + 1] = index[es[k]->vertices()[1]->id()];
    to_qt7(es[k]->measurement(), &meas[7 * k]);
    int p = 0;
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) info[21 * k + p++] = es[k]->information()(r, c);
    if (auto* hk = dynamic_cast<g2o::RobustKernelHuber*>(es[k]->robustKernel())) huber[k] = hk->delta();
  }
  lvs_pgo_t* h = nullptr;
  if (lvs_pgo_create(solver_kind(solver_type_), 0, nullptr, &h) != LVS_OK) { std::cerr << "lvslam_b200: " << lvs_last_error() << std::endl; return 0; }
  lvs_pgo_stats st{};
  int rc = lvs_pgo_set_graph(h, (int)vs.size(), poses.data(), fixed.data(), (int)es.size(), ij.data(), meas.data(), info.data(), huber.data());
  if (rc == LVS_OK) rc = lvs_pgo_optimize(h, num_iterations, &st);
  if (rc == LVS_OK) rc = lvs_pgo_get_poses(h, poses.data());
  lvs_pgo_destroy(h);
  if (rc != LVS_OK) { std::cerr << "lvslam_b200: " << lvs_last_error() << std::endl; return 0; }
  for (size_t i = 0; i < vs.size(); i++) {
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.linear() = Eigen::Quaterniond(poses[7 * i + 6], poses[7 * i + 3], poses[7 * i + 4], poses[7 * i + 5]).toRotationMatrix();
    T.translation() = Eigen::Vector3d(poses[7 * i], poses[7 * i + 1], poses[7 * i + 2]);
    vs[i]->setEstimate(T);
  }
  std::cout << "iterations: " << st.iterations << " / " << num_iterations << std::endl;
  std::cout << "chi2: (before)" << st.chi2_before << " -> (after)" << st.chi2_after << std::endl;
  std::cout << "time: " << st.device_ms * 1e-3 << "[sec]" << std::endl;
  return st.iterations;
}

}  // namespace lv_slam
