#pragma once
#include <map>
#include <set>
#include <vector>
#include "../../eigen_stub.h"
namespace g2o {
class RobustKernel { public: virtual ~RobustKernel() {} double delta() const { return d_; } void setDelta(double d) { d_ = d; } private: double d_ = 1.0; };
class HyperGraph {
 public:
  class Vertex { public: virtual ~Vertex() {} int id() const { return id_; } void setId(int i) { id_ = i; } private: int id_ = 0; };
  class Edge {
   public:
    Edge() : iid_(next_id()++) {}
    virtual ~Edge() {}
    std::vector<Vertex*>& vertices() { return v_; }
    long long internalId() const { return iid_; }       // g2o numbers edges in creation order
   private:
    static long long& next_id() { static long long n = 0; return n; }
    std::vector<Vertex*> v_;
    long long iid_;
  };
  typedef std::map<int, Vertex*> VertexIDMap;
  typedef std::set<Edge*> EdgeSet;
  virtual ~HyperGraph() { for (auto& kv : vs_) delete kv.second; for (auto* e : es_) delete e; }
  bool addVertex(Vertex* v) { return vs_.emplace(v->id(), v).second; }
  bool addEdge(Edge* e) { return es_.insert(e).second; }
  VertexIDMap& vertices() { return vs_; }
  EdgeSet& edges() { return es_; }
 private:
  VertexIDMap vs_;
  EdgeSet es_;
};
class SparseOptimizer : public HyperGraph {};
}  // namespace g2o
