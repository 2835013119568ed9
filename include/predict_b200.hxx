// predict_b200.hxx -- host-side C++ mirror of the reference's link-prediction API
// (puzzlef/neighborhood-link-prediction-openmp, inc/predict.hxx) on top of the C ABI in
// nlp_b200.h.  Same names, same template parameter order, same option / result structs:
//
//   PredictLinkOptions<W>            inc/predict.hxx:33-55
//   PredictLinkResult<K, W>          inc/predict.hxx:65-102
//   predictLinks<Measure>[Omp]<MINDEGREE1=4, MAXFACTOR2=0, FORCEHEAP=false, G, W=float>(x, o={})
//                                    inc/predict.hxx:502-831 (18 entry points)
//
// `G` only needs what the reference templates use: key_type, span(), hasVertex(u), degree(u),
// forEachEdgeKey(u, f) -- so the reference's DiGraph and DiGraphCsr (inc/Graph.hxx:23-372,
// 383-639) work unchanged, and so does nlp_b200::DeviceGraph below (a graph that is already
// resident on the GPU; use it to run many predictions on one upload, as main.cxx:212-220 does).
//
// Everything lives in namespace nlp_b200 so the header can sit next to the reference's own
// predict.hxx in one translation unit (tests do that).  For a drop-in build define
// NLP_B200_DROP_IN before including: the 20 names are then also visible unqualified.
//
// Differences from the reference, all deliberate:
//   * ties are ordered (score desc, u asc, v asc) instead of by heap accident;
//   * when fewer than maxEdges candidates exist the shorter list is returned (the sequential
//     reference's behaviour, inc/predict.hxx:250-255; the OpenMP merge is undefined there);
//   * errors of the GPU path (no device, out of memory) throw std::runtime_error -- there is no
//     CPU fallback;
//   * W must be float (main.cxx:20); FORCEHEAP is accepted and ignored.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <array>
#include <vector>

#include "nlp_b200.h"

namespace nlp_b200 {

template <class W>
struct PredictLinkOptions {
  int    repeat;     // number of times to repeat the scoring phase [1]
  size_t maxEdges;   // maximum number of edges to predict [-1]
  W      minScore;   // minimum score above which to consider a link [0]
  PredictLinkOptions(int repeat = 1, size_t maxEdges = size_t(-1), W minScore = W()) :
  repeat(repeat), maxEdges(maxEdges), minScore(minScore) {}
};

template <class K, class W>
struct PredictLinkResult {
  std::vector<std::tuple<K, K, W>> edges;   // predicted links (u < v), score descending
  float time;                               // total time, milliseconds
  float scoringTime;                        // scoring phase, milliseconds
  nlp_result stats;                         // counters of the GPU path (not in the reference)
  PredictLinkResult() : edges(), time(), scoringTime(), stats() {}
  PredictLinkResult(std::vector<std::tuple<K, K, W>>&& edges, float time = 0, float scoringTime = 0) :
  edges(std::move(edges)), time(time), scoringTime(scoringTime), stats() {}
};

namespace detail {

inline void check(nlp_handle* h, int rc, const char* what) {
  if (rc == NLP_OK) return;
  throw std::runtime_error(std::string("nlp_b200: ") + what + ": " + nlp_last_error(h));
}

// One predictor handle per process (the reference call is blocking and serial as well).
struct Session {
  nlp_handle* h = nullptr;
  std::mutex  mu;
  const void* resident = nullptr;   // DeviceGraph currently bound to the handle
  std::vector<uint64_t> offsets;    // staging for pack()
  std::vector<uint32_t> keys;
  std::array<uint64_t, 4> fingerprint{};   // {hash, second independent hash, span, entries}
  bool     have_fingerprint = false;
  bool     fetch_edges = true;      // setFetchEdges(false): results stay on the GPU (evaluateLastPrediction)
  static Session& get() { static Session s; return s; }
  nlp_handle* handle() {
    if (!h) {
      const char* d = std::getenv("NLP_B200_DEVICE");
      const int rc = nlp_create(&h, d ? std::atoi(d) : 0);
      if (rc != NLP_OK) throw std::runtime_error(std::string("nlp_b200: nlp_create: ") + nlp_last_error(nullptr));
      // NLP_B200_REUSE=1: share the sorted wedge records of a threshold between the measures
      // (main.cxx:212-220 runs all nine at every threshold on one graph); see nlp_set_reuse
      const char* r = std::getenv("NLP_B200_REUSE");
      if (r && *r == '1') nlp_set_reuse(h, 1);
    }
    return h;
  }
  ~Session() { if (h) nlp_destroy(h); }
};

inline uint64_t mix(uint64_t a, uint64_t x) {
  a ^= x + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2);
  return a * 0xBF58476D1CE4E5B9ull;
}

inline uint64_t mix2(uint64_t a, uint64_t x) {
  a = (a ^ x) * 0x100000001B3ull;
  return a ^ (a >> 29);
}

// G -> CSR (64-bit offsets, 32-bit keys), with a content fingerprint: two independent 64-bit
// hashes plus span and entry count (a graph is only taken for the resident one when all four agree).
// Rows must be sorted ascending, as update() leaves them (_bitset.hxx:20); the library checks it.
template <class G>
inline std::array<uint64_t, 4> pack(const G& x, std::vector<uint64_t>& off, std::vector<uint32_t>& keys) {
  const size_t S = x.span();
  off.assign(S + 1, 0);
  for (size_t u = 0; u < S; ++u)
    off[u + 1] = off[u] + (x.hasVertex(u) ? (uint64_t)x.degree(u) : 0);
  keys.resize(off[S]);
  std::vector<uint64_t> rowhash(S, 0), rowhash2(S, 0);
  #pragma omp parallel for schedule(dynamic, 2048)
  for (size_t u = 0; u < S; ++u) {
    if (!x.hasVertex(u)) continue;
    uint64_t i = off[u], hsh = off[u + 1], hs2 = 0xCBF29CE484222325ull ^ u;
    x.forEachEdgeKey(u, [&](auto v) { keys[i++] = (uint32_t)v; hsh = mix(hsh, (uint64_t)v); hs2 = mix2(hs2, (uint64_t)v + 1); });
    rowhash[u] = hsh; rowhash2[u] = hs2;
  }
  uint64_t f = mix(S, off[S]), f2 = 0x9AE16A3B2F90404Full;
  for (size_t u = 0; u < S; ++u) { f = mix(f, rowhash[u]); f2 = mix2(f2, rowhash2[u]); }
  return {f, f2, (uint64_t)S, off[S]};
}

}  // namespace detail


// A graph that already lives on the GPU: upload once, predict many times.
// Satisfies nothing of the host read API on purpose -- it is only accepted by the entry points.
class DeviceGraph {
 public:
  using key_type = uint32_t;
  template <class G>
  explicit DeviceGraph(const G& x) {
    detail::Session& s = detail::Session::get();
    std::lock_guard<std::mutex> lock(s.mu);
    detail::pack(x, s.offsets, s.keys);
    upload(s, s.offsets.data(), s.keys.data(), (uint32_t)(s.offsets.size() - 1));
  }
  // Host CSR arrays directly (offsets[span+1], keys[offsets[span]], rows sorted ascending).
  DeviceGraph(const uint64_t* offsets, const uint32_t* keys, uint32_t span) {
    detail::Session& s = detail::Session::get();
    std::lock_guard<std::mutex> lock(s.mu);
    upload(s, offsets, keys, span);
  }
  ~DeviceGraph() {
    detail::Session& s = detail::Session::get();
    std::lock_guard<std::mutex> lock(s.mu);
    if (s.resident == this) s.resident = nullptr;
  }
  // The Matrix Market file at `path`, ingested ON THE GPU the way main.cxx:243-245 does it on the
  // host: readMtxOmpW, symmetrizeOmp unless `symmetric` (main.cxx's second argument), then
  // removeSelfLoopsOmpU -- entry for entry the reference's graph (nlp_ingest_mtx).
  struct FromMtx {};
  DeviceGraph(FromMtx, const std::string& path, bool symmetric = false, bool dropSelfLoops = true) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("nlp_b200: cannot open " + path);
    const std::streamsize bytes = f.tellg();
    std::vector<char> text((size_t)bytes);
    f.seekg(0);
    if (bytes && !f.read(text.data(), bytes)) throw std::runtime_error("nlp_b200: cannot read " + path);
    detail::Session& s = detail::Session::get();
    std::lock_guard<std::mutex> lock(s.mu);
    nlp_handle* h = s.handle();
    uint32_t span = 0; uint64_t entries = 0;
    const uint32_t flags = (symmetric ? 0u : NLP_INGEST_SYMMETRIZE) | (dropSelfLoops ? NLP_INGEST_DROP_SELF_LOOPS : 0u);
    detail::check(h, nlp_ingest_mtx(h, text.data(), (uint64_t)text.size(), flags, &span, &entries), "nlp_ingest_mtx");
    s.resident = this; s.have_fingerprint = false; span_ = span; entries_ = entries;
  }
  DeviceGraph(const DeviceGraph&) = delete;
  DeviceGraph& operator=(const DeviceGraph&) = delete;
  size_t span() const { return span_; }
  size_t size() const { return entries_; }         // directed entries (known for ingested graphs)
 private:
  void upload(detail::Session& s, const uint64_t* off, const uint32_t* keys, uint32_t span) {
    nlp_handle* h = s.handle();
    detail::check(h, nlp_set_graph(h, off, keys, span), "nlp_set_graph");
    s.resident = this; s.have_fingerprint = false; span_ = span;
  }
  size_t span_ = 0, entries_ = 0;
  template <class K, class W, class G> friend PredictLinkResult<K, W> predictLinksB200(const G&, int, unsigned, unsigned, const PredictLinkOptions<W>&);
};


// The one function all 18 entry points forward to.
template <class K, class W, class G>
inline PredictLinkResult<K, W> predictLinksB200(const G& x, int measure, unsigned minDegree1, unsigned maxFactor2,
                                                const PredictLinkOptions<W>& o) {
  static_assert(std::is_same<W, float>::value, "nlp_b200: the GPU path scores in float (main.cxx:20: TYPE = float)");
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  nlp_handle* h = s.handle();
  if constexpr (std::is_same<G, DeviceGraph>::value) {
    if (s.resident != &x) throw std::runtime_error("nlp_b200: this DeviceGraph is no longer resident (another graph was uploaded since)");
  } else {
    // content-addressed: the same graph handed over again (main.cxx runs 99 predictions per
    // batch on one graph) is packed and compared, but not uploaded again
    const auto f = detail::pack(x, s.offsets, s.keys);
    if (!(s.have_fingerprint && s.fingerprint == f && s.resident == nullptr)) {
      detail::check(h, nlp_set_graph(h, s.offsets.data(), s.keys.data(), (uint32_t)(s.offsets.size() - 1)), "nlp_set_graph");
      s.fingerprint = f; s.have_fingerprint = true; s.resident = nullptr;
    }
  }
  nlp_options opt;
  opt.measure = measure; opt.min_degree1 = minDegree1; opt.max_factor2 = maxFactor2;
  opt.repeat = o.repeat; opt.max_edges = o.maxEdges == size_t(-1) ? NLP_UNBOUNDED : (uint64_t)o.maxEdges;
  opt.min_score = o.minScore;
  PredictLinkResult<K, W> a;
  detail::check(h, nlp_predict(h, &opt, &a.stats), "nlp_predict");
  a.time = a.stats.time_ms;
  a.scoringTime = a.stats.scoring_ms;
  if (!s.fetch_edges) return a;      // the edges stay in GPU memory (evaluateLastPrediction)
  const size_t n = (size_t)a.stats.count;
  std::vector<uint32_t> u(n), v(n);
  std::vector<float> sc(n);
  detail::check(h, nlp_fetch(h, u.data(), v.data(), sc.data(), n), "nlp_fetch");
  a.edges.resize(n);   // SoA -> tuples (never memcpy a std::tuple: member order is unspecified)
  for (size_t i = 0; i < n; ++i) a.edges[i] = std::make_tuple((K)u[i], (K)v[i], (W)sc[i]);
  return a;
}


// ---- evaluation on the GPU (main.cxx:48-57, 94-133, 199-206) ------------------------------------
// main.cxx turns every prediction into both directions, sorts, uniques and intersects it with the
// sorted directed list of removed edges on the host.  These three calls do the same where the
// prediction already lies:
//
//   setHeldBackEdges(deletions0);        // once per batch: main.cxx:206-207's sorted directed list
//   setFetchEdges(false);                // optional: leave PredictLinkResult::edges empty (no D2H copy)
//   auto p1 = fn<deg>(y, {repeat, deletions0.size()/2});
//   auto ev = evaluateLastPrediction();  // ev.precision, ev.recall as glog prints them (main.cxx:201-202)
struct LinkEvaluation {
  size_t predicted;   // |insertions1|: directed predicted edges (main.cxx:51-54)
  size_t truth;       // |insertions0|
  size_t common;      // |common1| (main.cxx:55)
  double precision;   // main.cxx:201
  double recall;      // main.cxx:202
  float  time;        // device time, milliseconds
};

// `edges` = sorted directed (u, v[, w]) tuples, e.g. main.cxx's deletions0.
template <class Tuple>
inline void setHeldBackEdges(const std::vector<Tuple>& edges) {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  nlp_handle* h = s.handle();
  std::vector<uint32_t> u(edges.size()), v(edges.size());
  for (size_t i = 0; i < edges.size(); ++i) { u[i] = (uint32_t)std::get<0>(edges[i]); v[i] = (uint32_t)std::get<1>(edges[i]); }
  detail::check(h, nlp_set_truth(h, u.data(), v.data(), (uint64_t)edges.size()), "nlp_set_truth");
}

inline void setFetchEdges(bool on) {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  s.fetch_edges = on;
}

// ---- batch generation on the GPU (inc/batch.hxx:99-112, 200-208 as called at main.cxx:165-168) ----
// The same removed edges as
//   default_random_engine rnd(seed);
//   auto deletions = generateEdgeDeletions(rnd, x, batchSize, 1, x.span()-1, true);
//   tidyBatchUpdateU(deletions, insertions, x);
// draw for draw (nlp_generate_deletions).  `words` receives the number of engine outputs the batch
// consumed: `rnd.discard(*words)` leaves a host engine where the reference's would be.
// With `holdBack` the list also becomes the ground truth of evaluateLastPrediction() without
// leaving the GPU (it is main.cxx's sorted `deletions0`).
template <class G>
inline auto generateEdgeDeletionsB200(const G& x, uint32_t seed, size_t batchSize, size_t* words = nullptr, bool holdBack = false) {
  using K = typename G::key_type;
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  nlp_handle* h = s.handle();
  if constexpr (std::is_same<G, DeviceGraph>::value) {
    if (s.resident != &x) throw std::runtime_error("nlp_b200: this DeviceGraph is no longer resident (another graph was uploaded since)");
  } else {
    const auto f = detail::pack(x, s.offsets, s.keys);
    if (!(s.have_fingerprint && s.fingerprint == f && s.resident == nullptr)) {
      detail::check(h, nlp_set_graph(h, s.offsets.data(), s.keys.data(), (uint32_t)(s.offsets.size() - 1)), "nlp_set_graph");
      s.fingerprint = f; s.have_fingerprint = true; s.resident = nullptr;
    }
  }
  uint64_t n = 0, w = 0;
  detail::check(h, nlp_generate_deletions(h, seed, (uint64_t)batchSize, &n, &w), "nlp_generate_deletions");
  if (words) *words = (size_t)w;
  if (holdBack) {
    const uint32_t *du = nullptr, *dv = nullptr;
    detail::check(h, nlp_deletions_device(h, &du, &dv, &n), "nlp_deletions_device");
    detail::check(h, nlp_set_truth(h, du, dv, n), "nlp_set_truth");
  }
  std::vector<uint32_t> u((size_t)n), v((size_t)n);
  detail::check(h, nlp_fetch_deletions(h, u.data(), v.data(), n), "nlp_fetch_deletions");
  std::vector<std::tuple<K, K>> a((size_t)n);
  for (size_t i = 0; i < (size_t)n; ++i) a[i] = std::make_tuple((K)u[i], (K)v[i]);
  return a;
}

// ---- applying a batch on the GPU (inc/batch.hxx:239-247 as called at main.cxx:164-169) ------------
// runBatches copies the loaded graph for every batch (y = duplicate(x)), applies the tidied
// deletions to the copy (applyBatchUpdateOmpU) and predicts on it.  With the graph resident:
//
//   nlp_b200::DeviceGraph x(g);                        // once: upload (validated on the device)
//   nlp_b200::checkpointGraph(x);                      // x is the base of every batch
//   for (every batch) {
//     nlp_b200::rollbackGraph(x);                      // y = duplicate(x): no copy, no upload
//     auto deletions = nlp_b200::generateEdgeDeletionsB200(x, seed, batchSize, nullptr, true);
//     nlp_b200::applyBatchUpdateB200(x, deletions);    // the resident graph is y now
//     ... predictLinks*<D>(x, {repeat, deletions.size() / 2}) ...
//   }
inline void checkpointGraph(const DeviceGraph& x) {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  if (s.resident != &x) throw std::runtime_error("nlp_b200: this DeviceGraph is no longer resident");
  detail::check(s.handle(), nlp_graph_checkpoint(s.handle()), "nlp_graph_checkpoint");
}

inline void rollbackGraph(const DeviceGraph& x) {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  if (s.resident != &x) throw std::runtime_error("nlp_b200: this DeviceGraph is no longer resident");
  detail::check(s.handle(), nlp_graph_rollback(s.handle()), "nlp_graph_rollback");
}

// `deletions` = unique directed (u, v) pairs, as tidyBatchUpdateU leaves them (both directions of
// every removed edge).  The resident graph loses one stored copy per pair.
template <class Tuple>
inline void applyBatchUpdateB200(const DeviceGraph& x, const std::vector<Tuple>& deletions) {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  if (s.resident != &x) throw std::runtime_error("nlp_b200: this DeviceGraph is no longer resident");
  nlp_handle* h = s.handle();
  std::vector<uint32_t> u(deletions.size()), v(deletions.size());
  for (size_t i = 0; i < deletions.size(); ++i) { u[i] = (uint32_t)std::get<0>(deletions[i]); v[i] = (uint32_t)std::get<1>(deletions[i]); }
  detail::check(h, nlp_apply_deletions(h, u.data(), v.data(), (uint64_t)deletions.size()), "nlp_apply_deletions");
}

// ---- several GPUs -----------------------------------------------------------
// One process per GPU (mpirun / torchrun / a shell loop), the same program on every rank.  Every
// rank uploads the graph to its own GPU; after joinCommunicatorFromEnv() every predictLinks* call is
// a collective: the sources are partitioned by wedge work, the library merges the ranks' candidates
// over its NCCL communicator, and EVERY rank returns the full result.
//   RANK, WORLD_SIZE          as the launcher sets them
//   LOCAL_RANK                -> the CUDA device (unless NLP_B200_DEVICE is set)
//   NLP_B200_COMM_FILE        a path all ranks can reach: rank 0 writes the 128-byte id there
inline void joinCommunicatorFromEnv() {
  const char* wr = std::getenv("WORLD_SIZE"); const char* rk = std::getenv("RANK");
  const int world = wr ? std::atoi(wr) : 1, rank = rk ? std::atoi(rk) : 0;
  if (world <= 1) return;
  if (!std::getenv("NLP_B200_DEVICE")) { if (const char* lr = std::getenv("LOCAL_RANK")) setenv("NLP_B200_DEVICE", lr, 1); }
  const char* path = std::getenv("NLP_B200_COMM_FILE");
  if (!path) throw std::runtime_error("nlp_b200: set NLP_B200_COMM_FILE to a path every rank can read");
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  nlp_handle* h = s.handle();
  unsigned char id[NLP_COMM_ID_BYTES];
  const std::string tmp = std::string(path) + ".tmp";
  if (rank == 0) {
    if (nlp_comm_unique_id(id) != NLP_OK) throw std::runtime_error(std::string("nlp_b200: nlp_comm_unique_id: ") + nlp_last_error(nullptr));
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id, 1, sizeof id, f) != sizeof id) throw std::runtime_error("nlp_b200: cannot write NLP_B200_COMM_FILE");
    std::fclose(f);
    std::rename(tmp.c_str(), path);
  } else {
    for (int tries = 0;; ++tries) {
      FILE* f = std::fopen(path, "rb");
      if (f) {
        const size_t got = std::fread(id, 1, sizeof id, f);
        std::fclose(f);
        if (got == sizeof id) break;
      }
      if (tries > 6000) throw std::runtime_error("nlp_b200: no communicator id in NLP_B200_COMM_FILE after 10 minutes");
      struct timespec ts = {0, 100000000L};
      nanosleep(&ts, nullptr);
    }
  }
  detail::check(h, nlp_comm_init(h, id, rank, world), "nlp_comm_init");
}

inline LinkEvaluation evaluateLastPrediction() {
  detail::Session& s = detail::Session::get();
  std::lock_guard<std::mutex> lock(s.mu);
  nlp_handle* h = s.handle();
  nlp_evaluation e;
  detail::check(h, nlp_evaluate(h, &e), "nlp_evaluate");
  return LinkEvaluation{(size_t)e.predicted, (size_t)e.truth, (size_t)e.common, e.precision, e.recall, e.ms};
}


// The 18 entry points (inc/predict.hxx:502-831).  The sequential and the OpenMP names run the
// same GPU path; measure order as main.cxx:212-220.
#define NLP_B200_ENTRY(NAME, MEASURE)                                                              \
  template <int MINDEGREE1 = 4, int MAXFACTOR2 = 0, bool FORCEHEAP = false, class G, class W = float> \
  inline auto NAME(const G& x, const PredictLinkOptions<W>& o = {}) {                              \
    using K = typename G::key_type;                                                                \
    return predictLinksB200<K, W>(x, MEASURE, (unsigned)MINDEGREE1, (unsigned)MAXFACTOR2, o);      \
  }                                                                                                \
  template <int MINDEGREE1 = 4, int MAXFACTOR2 = 0, bool FORCEHEAP = false, class G, class W = float> \
  inline auto NAME##Omp(const G& x, const PredictLinkOptions<W>& o = {}) {                         \
    using K = typename G::key_type;                                                                \
    return predictLinksB200<K, W>(x, MEASURE, (unsigned)MINDEGREE1, (unsigned)MAXFACTOR2, o);      \
  }

NLP_B200_ENTRY(predictLinksCommonNeighbors,        NLP_COMMON_NEIGHBORS)      // inc/predict.hxx:503, 520
NLP_B200_ENTRY(predictLinksJaccardCoefficient,     NLP_JACCARD_COEFFICIENT)   // inc/predict.hxx:541, 558
NLP_B200_ENTRY(predictLinksSorensenIndex,          NLP_SORENSEN_INDEX)        // inc/predict.hxx:579, 596
NLP_B200_ENTRY(predictLinksSaltonCosineSimilarity, NLP_SALTON_COSINE)         // inc/predict.hxx:617, 634
NLP_B200_ENTRY(predictLinksHubPromoted,            NLP_HUB_PROMOTED)          // inc/predict.hxx:655, 672
NLP_B200_ENTRY(predictLinksHubDepressed,           NLP_HUB_DEPRESSED)         // inc/predict.hxx:693, 710
NLP_B200_ENTRY(predictLinksLeichtHolmeNermanScore, NLP_LEICHT_HOLME_NERMAN)   // inc/predict.hxx:731, 748
NLP_B200_ENTRY(predictLinksAdamicAdarCoefficient,  NLP_ADAMIC_ADAR)           // inc/predict.hxx:769, 787
NLP_B200_ENTRY(predictLinksResourceAllocationScore, NLP_RESOURCE_ALLOCATION)  // inc/predict.hxx:809, 827
#undef NLP_B200_ENTRY

}  // namespace nlp_b200


#ifdef NLP_B200_DROP_IN
using nlp_b200::PredictLinkOptions;
using nlp_b200::PredictLinkResult;
using nlp_b200::predictLinksCommonNeighbors;         using nlp_b200::predictLinksCommonNeighborsOmp;
using nlp_b200::predictLinksJaccardCoefficient;      using nlp_b200::predictLinksJaccardCoefficientOmp;
using nlp_b200::predictLinksSorensenIndex;           using nlp_b200::predictLinksSorensenIndexOmp;
using nlp_b200::predictLinksSaltonCosineSimilarity;  using nlp_b200::predictLinksSaltonCosineSimilarityOmp;
using nlp_b200::predictLinksHubPromoted;             using nlp_b200::predictLinksHubPromotedOmp;
using nlp_b200::predictLinksHubDepressed;            using nlp_b200::predictLinksHubDepressedOmp;
using nlp_b200::predictLinksLeichtHolmeNermanScore;  using nlp_b200::predictLinksLeichtHolmeNermanScoreOmp;
using nlp_b200::predictLinksAdamicAdarCoefficient;   using nlp_b200::predictLinksAdamicAdarCoefficientOmp;
using nlp_b200::predictLinksResourceAllocationScore; using nlp_b200::predictLinksResourceAllocationScoreOmp;
#endif
