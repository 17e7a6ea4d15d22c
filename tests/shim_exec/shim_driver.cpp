// Executes the reference-side bindings of shim/ on the GPU (tests/test_shim_exec_gpu.py): the calls a lv_slam nodelet makes -
// setInputTarget / setInputSource / align / getFinalTransformation / hasConverged / getFitnessScore on the registration classes
// (scan_matching_odom_nodelet.cpp:109-126,197,220-226; loop_detector.hpp:155-184,219-262) and GraphSLAM::optimize on a g2o graph
// (global_graph_nodelet.cpp:670-764) - against the interface stand-ins of tests/shim_stubs.  Compiled three times: as is (pclomp), with
// -DLVS_SHIM_PCA (pclpca) and with -DLVS_SHIM_GROUND (pclomp_ground, registration only).  Input: raw little-endian arrays written by the test; output: one "key v v v ..." line per result.
#include <ndt_b200.h>
#include <global_graph/graph_slam.hpp>
#include <global_graph/information_matrix_calculator.hpp>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/types/slam3d/edge_se3.h>
#include <g2o/edge_se3_priorxy.hpp>
#include <g2o/edge_se3_priorxyz.hpp>
#include <g2o/edge_se3_priorquat.hpp>
#include <g2o/edge_se3_priorvec.hpp>
#include <g2o/edge_se3_plane.hpp>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#if defined(LVS_SHIM_PCA)
namespace ns = pclpca;
#define NDT_CLASS NormalDistributionsTransform
#elif defined(LVS_SHIM_GROUND)
namespace ns = pclomp_ground;
#define NDT_CLASS NormalDistributionsTransformGround
#else
namespace ns = pclomp;
#define NDT_CLASS NormalDistributionsTransform
#endif
pcl::PointCloud<pcl::PointXYZI>::Ptr lvs_prefilter(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr&, bool, double, double, float);   // shim/aux_b200.cpp

template <typename T>
static std::vector<T> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "cannot open " << path << std::endl; std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}

static pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_from(const std::vector<float>& xyz) {
  pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
  c->points.resize(xyz.size() / 3);
  for (size_t i = 0; i < c->points.size(); i++) { auto& p = c->points[i]; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.intensity = (float)(i % 7); }
  c->width = (uint32_t)c->points.size(); c->height = 1;
  return c;
}

#if !defined(LVS_SHIM_PCA) && !defined(LVS_SHIM_GROUND)
static Eigen::Isometry3d iso_from7(const double* v) {
  Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
  T.linear() = Eigen::Quaterniond(v[6], v[3], v[4], v[5]).toRotationMatrix();
  T.translation() = Eigen::Vector3d(v[0], v[1], v[2]);
  return T;
}
#endif

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  // ---------------- registration, through the pcl::Registration base pointer like registrations.cpp:78-98 hands it out
  auto tgt = cloud_from(slurp<float>(dir + "/tgt.f32")), src = cloud_from(slurp<float>(dir + "/src.f32"));
  const std::vector<float> g = slurp<float>(dir + "/guess.f32");
  Eigen::Matrix4f guess;
  for (int i = 0; i < 16; i++) guess.data()[i] = g[i];                       // column-major
  boost::shared_ptr<ns::NDT_CLASS<pcl::PointXYZI, pcl::PointXYZI> > ndt(new ns::NDT_CLASS<pcl::PointXYZI, pcl::PointXYZI>());
#ifdef LVS_SHIM_GROUND
  ndt->setResolution(10.0f);                                                   // ground_s2k, scan_matching_odom_nodelet.cpp:121-126
#else
  ndt->setResolution(1.0f);
#endif
  ndt->setNumThreads(4);
#if defined(LVS_SHIM_PCA) || defined(LVS_SHIM_GROUND)
  ndt->setNeighborhoodSearchMethod(ns::DIRECT1);                               // scan_matching_odom_nodelet.cpp:109-119
#else
  ndt->setNeighborhoodSearchMethod(ns::DIRECT7);
#endif
  pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>::Ptr reg = ndt;
  reg->setTransformationEpsilon(0.01);
#ifdef LVS_SHIM_GROUND
  reg->setMaximumIterations(64);
#else
  reg->setMaximumIterations(30);
#endif
  reg->setInputTarget(tgt);
  reg->setInputSource(src);
  pcl::PointCloud<pcl::PointXYZI> aligned;
  reg->align(aligned, guess);
  const Eigen::Matrix4f F = reg->getFinalTransformation();
  std::printf("final");
  for (int i = 0; i < 16; i++) std::printf(" %.9g", F.data()[i]);
  std::printf("\nconverged %d\niterations %d\ntrans_probability %.17g\n", (int)reg->hasConverged(), ndt->getFinalNumIteration(), ndt->getTransformationProbability());
  std::printf("fitness %.17g\nfitness_capped %.17g\n", ns::lvs_fitness_score<pcl::PointXYZI>(reg, std::numeric_limits<double>::max()), ns::lvs_fitness_score<pcl::PointXYZI>(reg, 0.25));
  std::printf("aligned_n %zu\naligned_first %.9g %.9g %.9g\naligned_last %.9g %.9g %.9g\n", aligned.points.size(), aligned.points[0].x, aligned.points[0].y, aligned.points[0].z,
              aligned.points.back().x, aligned.points.back().y, aligned.points.back().z);
  // second align from the first result, as the odometry nodelet does for its first pair (scan_matching_odom_nodelet.cpp:222-226)
  reg->align(aligned, F);
  std::printf("second_iterations %d\n", ndt->getFinalNumIteration());
#ifdef LVS_SHIM_PCA
  {
    const auto& cells = ndt->getTargetCells();
    long long wsum = 0;
    int usable = 0;
    for (size_t i = 0; i < cells.keys.size(); i++) { wsum += cells.weight[i]; usable += cells.nr_points[i] >= 6; }
    std::printf("cells %zu usable %d weight_sum %lld\n", cells.keys.size(), usable, wsum);
  }
#elif defined(LVS_SHIM_GROUND)
#else
  // ---------------- the stages either side: InformationMatrixCalculator::calc_fitness_score and the prefilter (shim/aux_b200.cpp)
  {
    Eigen::Isometry3d rel = Eigen::Isometry3d::Identity();
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) rel.linear()(r, c) = F(r, c); rel.translation().v[r] = F(r, 3); }
    std::printf("info_fitness %.17g\n", lv_slam::InformationMatrixCalculator::calc_fitness_score(tgt, src, rel));
    auto pf = lvs_prefilter(src, true, 0.5, 100.0, 0.1f);
    double sx = 0, si = 0;
    for (auto& p : pf->points) { sx += p.x; si += p.intensity; }
    std::printf("prefilter %zu %.9g %.9g\n", pf->points.size(), sx, si);
  }
  // ---------------- pose graph: a g2o graph filled the way GraphSLAM::add_se3_node / add_se3_edge / add_robust_kernel do, then optimize()
  {
    const std::vector<double> poses = slurp<double>(dir + "/poses7.f64"), meas = slurp<double>(dir + "/meas7.f64"), info21 = slurp<double>(dir + "/info21.f64");
    const std::vector<int32_t> ij = slurp<int32_t>(dir + "/ij.i32"), etype = slurp<int32_t>(dir + "/etype.i32");
    const std::vector<double> huber = slurp<double>(dir + "/huber.f64");
    lv_slam::GraphSLAM gs;
    gs.solver_type_ = "lm_var_cholmod";
    g2o::SparseOptimizer* graph = new g2o::SparseOptimizer();
    gs.graph.reset(graph);
    std::vector<g2o::VertexSE3*> vs;
    for (size_t i = 0; i < poses.size() / 7; i++) {
      auto* v = new g2o::VertexSE3();
      v->setId((int)graph->vertices().size());
      v->setEstimate(iso_from7(&poses[7 * i]));
      graph->addVertex(v);
      vs.push_back(v);
    }
    // edges in file order; the unary priors the way GraphSLAM::add_se3_prior_{xy,xyz,quat,vec}_edge build them (graph_slam.cpp:194-240)
    g2o::VertexPlane* floor_node = nullptr;
    auto kernel = [&](size_t k) -> g2o::RobustKernel* {
      if (!(huber[k] > 0)) return nullptr;
      auto* hk = new g2o::RobustKernelHuber();
      hk->setDelta(huber[k]);
      return hk;
    };
    auto info_dd = [&](size_t k, auto& I, int D) {
      for (int r = 0; r < D; r++) for (int c = r; c < D; c++) { const double v = info21[21 * k + r * 6 - r * (r - 1) / 2 + (c - r)]; I(r, c) = v; I(c, r) = v; }
    };
    for (size_t k = 0; k < ij.size() / 2; k++) {
      const double* m = &meas[7 * k];
      if (etype[k] == 0) {
        auto* e = new g2o::EdgeSE3();
        e->setMeasurement(iso_from7(m));
        Eigen::Matrix<double, 6, 6> I;
        info_dd(k, I, 6);
        e->setInformation(I);
        e->vertices().push_back(vs[ij[2 * k]]);
        e->vertices().push_back(vs[ij[2 * k + 1]]);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      } else if (etype[k] == 1) {
        auto* e = new g2o::EdgeSE3PriorXY();
        Eigen::Matrix<double, 2, 1> z; z(0, 0) = m[0]; z(1, 0) = m[1];
        Eigen::Matrix<double, 2, 2> I; info_dd(k, I, 2);
        e->setMeasurement(z); e->setInformation(I); e->vertices().push_back(vs[ij[2 * k]]);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      } else if (etype[k] == 2) {
        auto* e = new g2o::EdgeSE3PriorXYZ();
        Eigen::Matrix<double, 3, 3> I; info_dd(k, I, 3);
        e->setMeasurement(Eigen::Vector3d(m[0], m[1], m[2])); e->setInformation(I); e->vertices().push_back(vs[ij[2 * k]]);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      } else if (etype[k] == 3) {
        auto* e = new g2o::EdgeSE3PriorQuat();
        Eigen::Matrix<double, 3, 3> I; info_dd(k, I, 3);
        e->setMeasurement(Eigen::Quaterniond(m[3], m[0], m[1], m[2])); e->setInformation(I); e->vertices().push_back(vs[ij[2 * k]]);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      } else if (etype[k] == 5) {
        // the floor constraint: one plane node, fixed at creation (global_graph_nodelet.cpp:601-611); the plane itself comes in floor.f64
        if (!floor_node) {
          const std::vector<double> fc = slurp<double>(dir + "/floor.f64");
          Eigen::Vector4d c4; for (int a = 0; a < 4; a++) c4(a, 0) = fc[a];
          floor_node = new g2o::VertexPlane();
          floor_node->setId((int)graph->vertices().size());
          floor_node->setEstimate(g2o::Plane3D(c4));
          floor_node->setFixed(true);
          graph->addVertex(floor_node);
        }
        auto* e = new g2o::EdgeSE3Plane();
        Eigen::Vector4d z; for (int a = 0; a < 4; a++) z(a, 0) = m[a];
        Eigen::Matrix<double, 3, 3> I; info_dd(k, I, 3);
        e->setMeasurement(g2o::Plane3D(z)); e->setInformation(I);
        e->vertices().push_back(vs[ij[2 * k]]); e->vertices().push_back(floor_node);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      } else {
        auto* e = new g2o::EdgeSE3PriorVec();
        Eigen::Matrix<double, 6, 1> z; for (int a = 0; a < 6; a++) z(a, 0) = m[a];
        Eigen::Matrix<double, 3, 3> I; info_dd(k, I, 3);
        e->setMeasurement(z); e->setInformation(I); e->vertices().push_back(vs[ij[2 * k]]);
        if (auto* hk = kernel(k)) e->setRobustKernel(hk);
        graph->addEdge(e);
      }
    }
    const int it = gs.optimize(100);
    std::printf("pgo_iterations %d\n", it);
    for (size_t i = 0; i < vs.size(); i++) {
      const Eigen::Isometry3d& T = vs[i]->estimate();
      Eigen::Quaterniond q(T.linear());
      std::printf("pose %zu %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", i, T.translation().x(), T.translation().y(), T.translation().z(), q.x(), q.y(), q.z(), q.w());
    }
    lv_slam::GraphSLAM empty;
    empty.graph.reset(new g2o::SparseOptimizer());
    std::printf("pgo_empty %d\n", empty.optimize(10));
  }
#endif
  return 0;
}
