// TEST INFRASTRUCTURE ONLY -- C-callable wrapper around the UNMODIFIED reference batch generator
// (inc/batch.hxx: generateEdgeDeletions + tidyBatchUpdateU, as runBatches calls them at
// main.cxx:165-168), compiled from the sources where they lie (-I/root/reference, see
// oracle/Makefile) into oracle/_ref/libnlpref_batch.so.  Nothing from the reference is copied.
// Used by tests/test_batch_oracle.py to pin oracle/batch_oracle.c.
#include <cstdint>
#include <cstdlib>
#include <random>
#include <tuple>
#include <vector>
#include "inc/main.hxx"

using namespace std;

extern "C" {

// offsets[span+1], keys[M] -> reference DiGraph (the class main.cxx uses), then the reference's
// own sampling with std::default_random_engine(seed).  Returns the number of directed deletions
// written to out_u / out_v (capacity `cap`), or -1 if the capacity is too small.
int64_t nlpref_edge_deletions(const uint64_t* offsets, const uint32_t* keys, uint32_t span, uint32_t seed,
                              uint64_t batch_size, uint32_t* out_u, uint32_t* out_v, uint64_t cap) {
  using K = uint32_t;
  DiGraph<K, None, None> x;
  for (uint32_t u = 1; u < span; ++u) x.addVertex(u);
  for (uint32_t u = 0; u < span; ++u)
    for (uint64_t e = offsets[u]; e < offsets[u + 1]; ++e) x.addEdge(u, keys[e]);
  updateOmpU(x);
  default_random_engine rnd(seed);
  auto deletions = generateEdgeDeletions(rnd, x, size_t(batch_size), 1, x.span() - 1, true);
  vector<tuple<K, K>> insertions;
  tidyBatchUpdateU(deletions, insertions, x);
  if (deletions.size() > cap) return -1;
  size_t i = 0;
  for (const auto& [u, v] : deletions) { out_u[i] = u; out_v[i] = v; ++i; }
  return (int64_t)deletions.size();
}

// The reference's applyBatchUpdateOmpU (inc/batch.hxx:239-247, as runBatches calls it at
// main.cxx:169) on a DiGraph built from the CSR: removes the directed pairs (del_u[i], del_v[i]),
// then dumps the graph back as CSR.  out_keys needs room for the input's M entries; returns the
// number of entries left (out_offsets[span]).
int64_t nlpref_apply_deletions(const uint64_t* offsets, const uint32_t* keys, uint32_t span,
                               const uint32_t* del_u, const uint32_t* del_v, uint64_t ndel,
                               uint64_t* out_offsets, uint32_t* out_keys) {
  using K = uint32_t;
  DiGraph<K, None, None> y;
  for (uint32_t u = 1; u < span; ++u) y.addVertex(u);
  for (uint32_t u = 0; u < span; ++u)
    for (uint64_t e = offsets[u]; e < offsets[u + 1]; ++e) y.addEdge(u, keys[e]);
  updateOmpU(y);
  vector<tuple<K, K>> deletions, insertions;
  deletions.reserve(ndel);
  for (uint64_t i = 0; i < ndel; ++i) deletions.push_back(make_tuple(del_u[i], del_v[i]));
  applyBatchUpdateOmpU(y, deletions, insertions);
  uint64_t at = 0;
  for (uint32_t u = 0; u < span; ++u) {
    out_offsets[u] = at;
    if (y.hasVertex(u)) y.forEachEdgeKey(u, [&](K v) { out_keys[at++] = v; });
  }
  out_offsets[span] = at;
  return (int64_t)at;
}

// The reference's own ingest, as main.cxx:243-245 runs it: readMtxOmpW on the file at `path`, then
// symmetrizeOmp unless `symmetric`, then removeSelfLoopsOmpU; the graph is dumped as CSR.  First call
// with out_keys == NULL to learn the sizes (returns the entry count, *out_span = span).
int64_t nlpref_read_mtx(const char* path, int symmetric, int drop_self_loops, uint32_t* out_span,
                        uint64_t* out_offsets, uint32_t* out_keys) {
  using K = uint32_t;
  DiGraph<K, None, float> x;
  auto fl = [](auto u) { return true; };
  readMtxOmpW(x, path, false);
  if (!symmetric) x = symmetrizeOmp(x);
  if (drop_self_loops) removeSelfLoopsOmpU(x, fl);
  const uint32_t span = (uint32_t)x.span();
  *out_span = span;
  uint64_t at = 0;
  for (uint32_t u = 0; u < span; ++u) {
    if (out_offsets) out_offsets[u] = at;
    if (x.hasVertex(u)) x.forEachEdgeKey(u, [&](K v) { if (out_keys) out_keys[at] = v; ++at; });
  }
  if (out_offsets) out_offsets[span] = at;
  return (int64_t)at;
}

}  // extern "C"
