// Reference-side binding for the pose-graph path: a replacement body for lv_slam::GraphSLAM::optimize
// (/root/reference/src/global_graph/graph_slam.cpp:298-331).  The g2o graph stays the container the rest of global_graph
// works with (KeyFrame holds g2o::VertexSE3*, the nodelet reads ->estimate()); only the solve moves: pack -> lvs_pgo_* -> setEstimate.
// Needs g2o + Eigen, which are NOT in the build image of this repository; compiled on the lv_slam side.
#include <global_graph/graph_slam.hpp>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>
#include <g2o/types/slam3d/edge_se3.h>
#include <g2o/types/slam3d/vertex_se3.h>
#include <g2o/edge_se3_priorxy.hpp>       // the reference's own unary priors (include/g2o/ of lv_slam): GPS / IMU constraints
#include <g2o/edge_se3_priorxyz.hpp>
#include <g2o/edge_se3_priorquat.hpp>
#include <g2o/edge_se3_priorvec.hpp>
#include <g2o/edge_se3_plane.hpp>        // floor constraint: EdgeSE3Plane against the nodelet's one fixed VertexPlane
#include <g2o/types/slam3d_addons/vertex_plane.h>
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

// the D x D information of a prior edge into the top-left corner of a 6 x 6 upper triangle (row-major, 21 numbers)
template <typename E>
static void pack_info(const E* e, int D, double* u21) {
  std::fill(u21, u21 + 21, 0.0);
  for (int r = 0; r < D; r++) for (int c = r; c < D; c++) u21[r * 6 - r * (r - 1) / 2 + (c - r)] = e->information()(r, c);
}
template <typename E>
static double huber_of(const E* e) {
  auto* hk = dynamic_cast<g2o::RobustKernelHuber*>(e->robustKernel());
  return hk ? hk->delta() : 0.0;
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
  // EdgeSE3, the unary priors on a VertexSE3 (add_se3_prior_{xy,xyz,quat,vec}_edge, graph_slam.cpp:194-240) and EdgeSE3Plane against a FIXED plane
  // vertex (the floor node, global_graph_nodelet.cpp:601-611); edges to a free plane vertex and plane-plane edges are left out like any other type
  g2o::VertexPlane* floor = nullptr;
  struct Item { g2o::HyperGraph::Edge* e; long long id; int type; };
  std::vector<Item> es;
  for (auto* e : graph->edges()) {
    if (auto* s = dynamic_cast<g2o::EdgeSE3*>(e)) es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_SE3});
    else if (auto* s = dynamic_cast<g2o::EdgeSE3PriorXY*>(e)) es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_PRIOR_XY});
    else if (auto* s = dynamic_cast<g2o::EdgeSE3PriorXYZ*>(e)) es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_PRIOR_XYZ});
    else if (auto* s = dynamic_cast<g2o::EdgeSE3PriorQuat*>(e)) es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_PRIOR_QUAT});
    else if (auto* s = dynamic_cast<g2o::EdgeSE3PriorVec*>(e)) es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_PRIOR_VEC});
    else if (auto* s = dynamic_cast<g2o::EdgeSE3Plane*>(e)) {
      auto* pl = dynamic_cast<g2o::VertexPlane*>(s->vertices()[1]);
      if (pl && pl->fixed() && (!floor || floor == pl)) { floor = pl; es.push_back({e, (long long)s->internalId(), LVS_PGO_EDGE_SE3_PLANE}); }
    }
  }
  std::sort(es.begin(), es.end(), [](const Item& a, const Item& b) { return a.id < b.id; });
  std::vector<double> poses(7 * vs.size()), meas(7 * es.size(), 0.0), info(21 * es.size()), huber(es.size(), 0.0);
  std::vector<uint8_t> fixed(vs.size());
  std::vector<int32_t> ij(2 * es.size()), type(es.size());
  for (size_t i = 0; i < vs.size(); i++) { to_qt7(vs[i]->estimate(), &poses[7 * i]); fixed[i] = vs[i]->fixed(); }
  for (size_t k = 0; k < es.size(); k++) {
    type[k] = es[k].type;
    ij[2 * k] = index[es[k].e->vertices()[0]->id()];
    ij[2 * k + 1] = es[k].type == LVS_PGO_EDGE_SE3 ? index[es[k].e->vertices()[1]->id()] : ij[2 * k];
    double* m = &meas[7 * k];
    switch (es[k].type) {
      case LVS_PGO_EDGE_SE3: {
        auto* e = static_cast<g2o::EdgeSE3*>(es[k].e);
        to_qt7(e->measurement(), m);
        int p = 0;
        for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) info[21 * k + p++] = e->information()(r, c);
        huber[k] = huber_of(e);
        break;
      }
      case LVS_PGO_EDGE_PRIOR_XY: {
        auto* e = static_cast<g2o::EdgeSE3PriorXY*>(es[k].e);
        m[0] = e->measurement()(0, 0); m[1] = e->measurement()(1, 0);
        pack_info(e, 2, &info[21 * k]); huber[k] = huber_of(e);
        break;
      }
      case LVS_PGO_EDGE_PRIOR_XYZ: {
        auto* e = static_cast<g2o::EdgeSE3PriorXYZ*>(es[k].e);
        for (int a = 0; a < 3; a++) m[a] = e->measurement()(a, 0);
        pack_info(e, 3, &info[21 * k]); huber[k] = huber_of(e);
        break;
      }
      case LVS_PGO_EDGE_PRIOR_QUAT: {
        auto* e = static_cast<g2o::EdgeSE3PriorQuat*>(es[k].e);
        m[0] = e->measurement().x(); m[1] = e->measurement().y(); m[2] = e->measurement().z(); m[3] = e->measurement().w();
        pack_info(e, 3, &info[21 * k]); huber[k] = huber_of(e);
        break;
      }
      case LVS_PGO_EDGE_SE3_PLANE: {
        auto* e = static_cast<g2o::EdgeSE3Plane*>(es[k].e);
        const Eigen::Vector4d pc = e->measurement().toVector();
        for (int a = 0; a < 4; a++) m[a] = pc(a, 0);
        pack_info(e, 3, &info[21 * k]); huber[k] = huber_of(e);
        break;
      }
      default: {
        auto* e = static_cast<g2o::EdgeSE3PriorVec*>(es[k].e);
        for (int a = 0; a < 6; a++) m[a] = e->measurement()(a, 0);
        pack_info(e, 3, &info[21 * k]); huber[k] = huber_of(e);
        break;
      }
    }
  }
  lvs_pgo_t* h = nullptr;
  if (lvs_pgo_create(solver_kind(solver_type_), 0, nullptr, &h) != LVS_OK) { std::cerr << "lvslam_b200: " << lvs_last_error() << std::endl; return 0; }
  lvs_pgo_stats st{};
  int rc = LVS_OK;
  if (floor) {
    const Eigen::Vector4d fc = floor->estimate().toVector();
    const double c4[4] = {fc(0, 0), fc(1, 0), fc(2, 0), fc(3, 0)};
    rc = lvs_pgo_set_floor_plane(h, c4);
  }
  if (rc == LVS_OK) rc = lvs_pgo_set_graph_typed(h, (int)vs.size(), poses.data(), fixed.data(), (int)es.size(), ij.data(), meas.data(), info.data(), huber.data(), type.data());
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
