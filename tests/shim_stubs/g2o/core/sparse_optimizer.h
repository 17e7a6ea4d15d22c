#pragma once
#include <map>
#include <set>
#include <vector>
#include "../../eigen_stub.h"
namespace g2o {
class RobustKernel { public: virtual ~RobustKernel() {} double delta() const { return d_; } void setDelta(double d) { d_ = d; } private: double d_ = 1.0; };
class HyperGraph {
 public:
  class Vertex { public: virtual ~Vertex() {} int id() const { return id_; } private: int id_ = 0; };
  class Edge { public: virtual ~Edge() {} std::vector<Vertex*>& vertices() { return v_; } long long internalId() const { return 0; } private: std::vector<Vertex*> v_; };
  typedef std::map<int, Vertex*> VertexIDMap;
  typedef std::set<Edge*> EdgeSet;
  virtual ~HyperGraph() {}
  VertexIDMap& vertices() { return vs_; }
  EdgeSet& edges() { return es_; }
 private:
  VertexIDMap vs_;
  EdgeSet es_;
};
class SparseOptimizer : public HyperGraph {};
}  // namespace g2o
