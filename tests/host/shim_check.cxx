// Exercises include/predict_b200.hxx (the C++ mirror of inc/predict.hxx:502-831) on a graph
// class that offers only the four accessors the reference templates use.  Reads a CSR from a
// binary file, runs every entry point for the thresholds main.cxx:67-80 sweeps, and writes the
// results to a binary file that tests/test_cpp_shim.py compares with the oracle.
//
//   shim_check <graph.bin> <out.bin> <maxEdges|-1>
//   graph.bin : u64 span, u64 M, u64 offsets[span+1], u32 keys[M]
//   out.bin   : per case  u32 measure, u32 omp, u32 D, u64 n, then n x (u32 u, u32 v, f32 score)
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <tuple>
#include <vector>
#include "predict_b200.hxx"

using namespace std;

struct MiniGraph {
  using key_type = uint32_t;
  vector<uint64_t> off;
  vector<uint32_t> keys;
  size_t span() const { return off.size() - 1; }
  bool hasVertex(size_t u) const { return u < span(); }
  size_t degree(size_t u) const { return u < span() ? off[u + 1] - off[u] : 0; }
  template <class F> void forEachEdgeKey(size_t u, F fn) const {
    for (uint64_t i = off[u]; i < off[u + 1]; ++i) fn(keys[i]);
  }
};

static FILE* out;

template <class R>
static void emit(uint32_t measure, uint32_t omp, uint32_t D, const R& r) {
  uint64_t n = r.edges.size();
  fwrite(&measure, 4, 1, out); fwrite(&omp, 4, 1, out); fwrite(&D, 4, 1, out); fwrite(&n, 8, 1, out);
  for (const auto& [u, v, s] : r.edges) { uint32_t a = u, b = v; float c = s; fwrite(&a, 4, 1, out); fwrite(&b, 4, 1, out); fwrite(&c, 4, 1, out); }
  if (!(r.time >= r.scoringTime && r.scoringTime >= 0)) throw runtime_error("time < scoringTime");
}

template <int D, class G>
static void run_d(const G& x, const nlp_b200::PredictLinkOptions<float>& o) {
  using namespace nlp_b200;
  emit(0, 0, D, predictLinksCommonNeighbors<D>(x, o));          emit(0, 1, D, predictLinksCommonNeighborsOmp<D>(x, o));
  emit(1, 0, D, predictLinksJaccardCoefficient<D>(x, o));       emit(1, 1, D, predictLinksJaccardCoefficientOmp<D>(x, o));
  emit(2, 0, D, predictLinksSorensenIndex<D>(x, o));            emit(2, 1, D, predictLinksSorensenIndexOmp<D>(x, o));
  emit(3, 0, D, predictLinksSaltonCosineSimilarity<D>(x, o));   emit(3, 1, D, predictLinksSaltonCosineSimilarityOmp<D>(x, o));
  emit(4, 0, D, predictLinksHubPromoted<D>(x, o));              emit(4, 1, D, predictLinksHubPromotedOmp<D>(x, o));
  emit(5, 0, D, predictLinksHubDepressed<D>(x, o));             emit(5, 1, D, predictLinksHubDepressedOmp<D>(x, o));
  emit(6, 0, D, predictLinksLeichtHolmeNermanScore<D>(x, o));   emit(6, 1, D, predictLinksLeichtHolmeNermanScoreOmp<D>(x, o));
  emit(7, 0, D, predictLinksAdamicAdarCoefficient<D>(x, o));    emit(7, 1, D, predictLinksAdamicAdarCoefficientOmp<D>(x, o));
  emit(8, 0, D, predictLinksResourceAllocationScore<D>(x, o));  emit(8, 1, D, predictLinksResourceAllocationScoreOmp<D>(x, o));
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: shim_check graph.bin out.bin maxEdges\n"); return 2; }
  MiniGraph x;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  uint64_t S = 0, M = 0;
  if (fread(&S, 8, 1, f) != 1 || fread(&M, 8, 1, f) != 1) return 2;
  x.off.resize(S + 1); x.keys.resize(M);
  if (fread(x.off.data(), 8, S + 1, f) != S + 1) return 2;
  if (M && fread(x.keys.data(), 4, M, f) != M) return 2;
  fclose(f);
  const long long k = atoll(argv[3]);
  try {
    out = fopen(argv[2], "wb");
    nlp_b200::PredictLinkOptions<float> o(1, k < 0 ? size_t(-1) : (size_t)k);
    run_d<0>(x, o); run_d<2>(x, o); run_d<16>(x, o); run_d<1024>(x, o);
    // the default template arguments and the default options (MINDEGREE1 = 4, all candidates)
    emit(1, 1, 4, nlp_b200::predictLinksJaccardCoefficientOmp(x));
    // a graph that stays on the GPU across predictions
    nlp_b200::DeviceGraph dg(x);
    emit(1, 1, 8, nlp_b200::predictLinksJaccardCoefficientOmp<8>(dg, o));
    auto last = nlp_b200::predictLinksAdamicAdarCoefficientOmp<8>(dg, {1, 100});
    emit(7, 1, 8, last);
    // evaluation on the device (main.cxx:48-57, 201-202): hold back the first half of that very
    // prediction, predict again without fetching the edges, expect precision 1/2 and recall 1
    {
      vector<tuple<uint32_t, uint32_t, float>> truth;
      const size_t half = last.edges.size() / 2;
      for (size_t i = 0; i < half; ++i) {
        const auto& [u, v, w] = last.edges[i];
        truth.push_back({u, v, 1.0f}); truth.push_back({v, u, 1.0f});
      }
      sort(truth.begin(), truth.end());
      nlp_b200::setHeldBackEdges(truth);
      nlp_b200::setFetchEdges(false);
      auto again = nlp_b200::predictLinksAdamicAdarCoefficientOmp<8>(dg, {1, 100});
      nlp_b200::setFetchEdges(true);
      const auto ev = nlp_b200::evaluateLastPrediction();
      if (!again.edges.empty() || again.stats.count != last.edges.size()) throw runtime_error("setFetchEdges(false) still fetched");
      if (ev.predicted != 2 * last.edges.size() || ev.truth != truth.size() || ev.common != truth.size())
        throw runtime_error("evaluateLastPrediction: wrong counts");
      if (half && (ev.recall != 1.0 || ev.precision != double(truth.size()) / double(2 * last.edges.size())))
        throw runtime_error("evaluateLastPrediction: wrong precision / recall");
    }
    // graph ingest on the GPU (main.cxx:243-245): the edges u < v of x as a "general" Matrix Market
    // file, symmetrized and stripped of self-loops on the device, must predict like the same graph
    // built on the host
    {
      const string path = string(argv[2]) + ".mtx";
      MiniGraph y;
      vector<vector<uint32_t>> rows(S);
      size_t lines = 0;
      for (uint64_t u = 0; u < S; ++u)
        for (uint64_t e = x.off[u]; e < x.off[u + 1]; ++e) {
          const uint32_t v = x.keys[e];
          if (u == 0 || u >= v || (e > x.off[u] && x.keys[e - 1] == v)) continue;
          rows[u].push_back(v); rows[v].push_back((uint32_t)u); ++lines;
        }
      FILE* m = fopen(path.c_str(), "w");
      if (!m) throw runtime_error("cannot write " + path);
      fprintf(m, "%%%%MatrixMarket matrix coordinate real general\n%% written by shim_check\n%llu %llu %zu\n",
              (unsigned long long)(S - 1), (unsigned long long)(S - 1), lines);
      for (uint64_t u = 0; u < S; ++u)
        for (uint32_t v : rows[u]) if (u < v) fprintf(m, "%llu %u 1.5\n", (unsigned long long)u, v);
      fclose(m);
      y.off.assign(S + 1, 0);
      for (uint64_t u = 0; u < S; ++u) { sort(rows[u].begin(), rows[u].end()); y.off[u + 1] = y.off[u] + rows[u].size(); }
      for (uint64_t u = 0; u < S; ++u) y.keys.insert(y.keys.end(), rows[u].begin(), rows[u].end());
      const auto want = nlp_b200::predictLinksJaccardCoefficientOmp<8>(y, {1, 500});
      nlp_b200::DeviceGraph fromFile(nlp_b200::DeviceGraph::FromMtx{}, path, /*symmetric=*/false, /*dropSelfLoops=*/true);
      remove(path.c_str());
      if (fromFile.span() != S || fromFile.size() != y.keys.size()) throw runtime_error("DeviceGraph(FromMtx): wrong span or size");
      const auto got = nlp_b200::predictLinksJaccardCoefficientOmp<8>(fromFile, {1, 500});
      if (got.edges != want.edges) throw runtime_error("DeviceGraph(FromMtx): prediction differs from the host-built graph");
    }
    fclose(out);
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_check: %s\n", e.what());
    return 3;
  }
  printf("shim_check ok\n");
  return 0;
}
