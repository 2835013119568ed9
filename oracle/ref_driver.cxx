// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-callable wrapper around the UNMODIFIED reference templates, compiled from the
// sources where they lie (`-I/root/reference`, see oracle/Makefile) into
// oracle/_ref/libnlpref.so.  Nothing from the reference is copied: this file only
// #includes `inc/main.hxx` and dispatches a runtime (measure, D, omp) triple to the
// matching template instantiation of
//   predictLinks<Measure>[Omp]<MINDEGREE1>(x, {repeat, maxEdges, minScore})
// (reference inc/predict.hxx:502-831), on a `DiGraphCsr` (inc/Graph.hxx:383-639)
// filled from the caller's CSR arrays.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <tuple>
#include <algorithm>
#include "inc/main.hxx"

using namespace std;

namespace {

using K = uint32_t;
using W = float;
using RefGraph = DiGraphCsr<K, None, None, size_t>;
using RefResult = PredictLinkResult<K, W>;
using RefOptions = PredictLinkOptions<W>;

// Measures in main.cxx:212-220 order.
enum Measure { CN = 0, JC, SI, SC, HP, HD, LHN, AA, RA };

template <int D>
RefResult run_d(const RefGraph& x, int measure, bool omp, const RefOptions& o) {
  switch (measure) {
    case CN:  return omp ? predictLinksCommonNeighborsOmp<D>(x, o)         : predictLinksCommonNeighbors<D>(x, o);
    case JC:  return omp ? predictLinksJaccardCoefficientOmp<D>(x, o)      : predictLinksJaccardCoefficient<D>(x, o);
    case SI:  return omp ? predictLinksSorensenIndexOmp<D>(x, o)           : predictLinksSorensenIndex<D>(x, o);
    case SC:  return omp ? predictLinksSaltonCosineSimilarityOmp<D>(x, o)  : predictLinksSaltonCosineSimilarity<D>(x, o);
    case HP:  return omp ? predictLinksHubPromotedOmp<D>(x, o)             : predictLinksHubPromoted<D>(x, o);
    case HD:  return omp ? predictLinksHubDepressedOmp<D>(x, o)            : predictLinksHubDepressed<D>(x, o);
    case LHN: return omp ? predictLinksLeichtHolmeNermanScoreOmp<D>(x, o)  : predictLinksLeichtHolmeNermanScore<D>(x, o);
    case AA:  return omp ? predictLinksAdamicAdarCoefficientOmp<D>(x, o)   : predictLinksAdamicAdarCoefficient<D>(x, o);
    case RA:  return omp ? predictLinksResourceAllocationScoreOmp<D>(x, o) : predictLinksResourceAllocationScore<D>(x, o);
  }
  return RefResult();
}

bool run(const RefGraph& x, int measure, int D, bool omp, const RefOptions& o, RefResult& out) {
  switch (D) {  // the eleven thresholds main.cxx:67-80 sweeps
    case 0:    out = run_d<0>(x, measure, omp, o);    return true;
    case 2:    out = run_d<2>(x, measure, omp, o);    return true;
    case 4:    out = run_d<4>(x, measure, omp, o);    return true;
    case 8:    out = run_d<8>(x, measure, omp, o);    return true;
    case 16:   out = run_d<16>(x, measure, omp, o);   return true;
    case 32:   out = run_d<32>(x, measure, omp, o);   return true;
    case 64:   out = run_d<64>(x, measure, omp, o);   return true;
    case 128:  out = run_d<128>(x, measure, omp, o);  return true;
    case 256:  out = run_d<256>(x, measure, omp, o);  return true;
    case 512:  out = run_d<512>(x, measure, omp, o);  return true;
    case 1024: out = run_d<1024>(x, measure, omp, o); return true;
  }
  return false;
}

struct RefHandle {
  RefGraph  x;
  RefResult last;
  RefHandle(size_t n, size_t m) : x(n, m) {}
};

}  // namespace


extern "C" {

// Build a reference DiGraphCsr from CSR arrays (offsets[S+1], keys[M]).
void* nlpref_graph_create(const uint64_t* offsets, const uint32_t* keys, uint32_t span) {
  size_t S = span, M = offsets[S];
  RefHandle* h = new RefHandle(S, M);
  for (size_t u = 0; u <= S; ++u) h->x.offsets[u] = offsets[u];
  for (size_t u = 0; u <  S; ++u) h->x.degrees[u] = K(offsets[u+1] - offsets[u]);
  if (M) memcpy(h->x.edgeKeys.data(), keys, M * sizeof(K));
  return h;
}

void nlpref_graph_destroy(void* hp) { delete static_cast<RefHandle*>(hp); }

// Run one reference prediction; returns the number of predicted edges (kept inside the
// handle until fetched), or -1 when (measure, D) has no instantiation here.
// NOTE: the reference's OpenMP merge is undefined when #candidates < max_edges
// (inc/predict.hxx:424,442,452-453); callers must use omp=0 in that regime.
int64_t nlpref_predict(void* hp, int measure, int D, int omp, int threads, uint64_t max_edges,
                       float min_score, int repeat, float* time_ms, float* scoring_ms) {
  RefHandle* h = static_cast<RefHandle*>(hp);
  if (threads > 0) omp_set_num_threads(threads);
  RefOptions o(repeat, size_t(max_edges), min_score);
  if (measure < 0 || measure > 8) return -1;
  if (!run(h->x, measure, D, omp != 0, o, h->last)) return -1;
  if (time_ms)    *time_ms    = h->last.time;
  if (scoring_ms) *scoring_ms = h->last.scoringTime;
  return int64_t(h->last.edges.size());
}

// Copy the last result out as SoA (never memcpy a std::tuple).
void nlpref_fetch(void* hp, uint32_t* u, uint32_t* v, float* score) {
  RefHandle* h = static_cast<RefHandle*>(hp);
  size_t n = h->last.edges.size();
  for (size_t i = 0; i < n; ++i) {
    u[i]     = get<0>(h->last.edges[i]);
    v[i]     = get<1>(h->last.edges[i]);
    score[i] = get<2>(h->last.edges[i]);
  }
}

int nlpref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
