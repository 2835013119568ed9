// Compile-only check (tests/test_cpp_shim.py, CPU): every helper of include/predict_b200.hxx that
// shim_check.cxx does not call is instantiated here, for a host graph class and for DeviceGraph.
#include <cstdint>
#include <tuple>
#include <vector>
#include "predict_b200.hxx"

struct MiniGraph {
  using key_type = uint32_t;
  std::vector<uint64_t> off;
  std::vector<uint32_t> keys;
  size_t span() const { return off.size() - 1; }
  bool hasVertex(size_t u) const { return u < span(); }
  size_t degree(size_t u) const { return u < span() ? off[u + 1] - off[u] : 0; }
  template <class F> void forEachEdgeKey(size_t u, F fn) const {
    for (uint64_t i = off[u]; i < off[u + 1]; ++i) fn(keys[i]);
  }
};

size_t instantiate(const MiniGraph& x, const nlp_b200::DeviceGraph& dg) {
  size_t words = 0;
  std::vector<std::tuple<uint32_t, uint32_t>> a = nlp_b200::generateEdgeDeletionsB200(x, 12345u, 100, &words, true);
  auto b = nlp_b200::generateEdgeDeletionsB200(dg, 1u, 10);
  nlp_b200::setHeldBackEdges(a);
  nlp_b200::setFetchEdges(false);
  const nlp_b200::LinkEvaluation ev = nlp_b200::evaluateLastPrediction();
  // the batch loop on the GPU: base graph, rollback, apply; several GPUs
  nlp_b200::checkpointGraph(dg);
  nlp_b200::rollbackGraph(dg);
  nlp_b200::applyBatchUpdateB200(dg, b);
  nlp_b200::joinCommunicatorFromEnv();
  // graph ingest on the GPU (main.cxx:243-245)
  nlp_b200::DeviceGraph fromFile(nlp_b200::DeviceGraph::FromMtx{}, "graph.mtx", false, true);
  words += fromFile.span() + fromFile.size();
  return a.size() + b.size() + words + ev.common;
}
