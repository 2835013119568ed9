// C ABI (include/nlp_b200.h) over the hand-written sm_100a kernels: host-side orchestration of
// one prediction = frontier -> wedge enumeration / counting / scoring -> global top-K.
// Replaces predictLinksWithIntersectionOmp (reference inc/predict.hxx:409-467).
// There is no CPU fallback anywhere in this file: every step is a kernel launch.
#include "../../include/nlp_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "frontier.cuh"
#include "wedge.cuh"
#include "select.cuh"
#include "pairs.cuh"
#include "bucket.cuh"
#include "evaluate.cuh"
#include "batch.cuh"
#include "ingest.cuh"
#include "mtx_header.hpp"

#include <dlfcn.h>
#include <nccl.h>      // types only: libnccl.so.2 is dlopen'ed by nlp_comm_init

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace nlp;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void*  p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct nlp_handle {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_frontier = nullptr, ev_scored = nullptr, ev_done = nullptr;
  cudaEvent_t ev_phase[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool phases_valid = false;
  float phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // pruned-buffer mode: per-phase time summed over passes
  // graph
  const uint64_t* d_off = nullptr;
  const uint32_t* d_keys = nullptr;
  DevBuf own_off, own_keys, spare_off, spare_keys;   // spare: destination of nlp_apply_deletions when the graph is the handle's own
  DevBuf del_bits;
  // nlp_graph_checkpoint: the base graph of a batch loop (main.cxx:164: y = duplicate(x) per batch)
  DevBuf base_off, base_keys;                // owned copy-free: the handle's own CSR moves here
  const uint64_t* base_d_off = nullptr; const uint32_t* base_d_keys = nullptr;
  uint32_t base_S = 0; int base_sym = 0; uint32_t base_maxmult = 1; unsigned long long base_id = 0; bool has_base = false;
  uint32_t S = 0;
  uint64_t M = 0;
  uint32_t maxdeg = 0;
  uint32_t maxmult = 0;                      // largest multiplicity of an entry in a row; 0 = not measured yet
  bool has_graph = false;
  DevBuf deg, work, work64, elig, maxdeg_dev;
  DevBuf chunk_base, chunk_src, chunk_cnt;   // long rows (deg > LONG_ROW) cut into CHUNK-entry pieces
  uint64_t nchunks = 0;
  DevBuf ecount, ekeys;                      // LHub: compacted eligible first-hop lists
  DevBuf scan_tiles, scan_total;
  // pair path (LHub on symmetric graphs)
  DevBuf it_u, it_cnt, it_dw, it_ptr, it_off, sym_flag;
  int sym_state = 0;                         // 0 unknown, 1 symmetric rows, 2 not symmetric
  unsigned long long graph_id = 0;           // content fingerprint of the resident graph (0 = not taken)
  std::map<unsigned long long, int> known_sym;   // fingerprint -> sym_state of graphs seen before
  // item / record counts of the pair path per (D, rank, world): a pure function of the resident
  // graph, so only the first prediction at a threshold pays the two host round trips for them
  std::map<uint64_t, std::pair<uint64_t, uint64_t>> pair_sizes;
  int path_mode = NLP_PATH_AUTO;
  int coop_mode = 1;                         // 0: per-warp wedge streaming in k_hash / k_dense (count measures)
  int range_half = 1;                        // k_range: half-word counters for sources with deg < 2^15 (NLP_B200_RANGE_HALF=0: off)
  uint32_t range_div = 4;                    // weight of the per-row window cost in the k_range / k_dense rule (frontier.cuh)
  uint32_t bucket_cap = BK_CAP_COUNT;        // records per bucket of the count measures (NLP_B200_BUCKET_CAP = 4096 | 8192)
  int range_mode = 1;
  int flt_range_mode = 1;                    // 0: float measures keep the single-warp dense-table path for hub-heavy sources (NLP_B200_RANGE_FLT=0)                        // 0: hub-heavy count sources use k_dense (HBM tables) instead of k_range
  DevBuf list[NBINS], defer[NBINS];
  DevBuf gtable;
  uint32_t gtable_n = 0;
  // control
  DevBuf ctr, thr;
  Counters* h_ctr = nullptr;          // pinned
  unsigned long long* h_hist = nullptr;  // pinned, 12*256
  SelectState* h_sel = nullptr;       // pinned
  // candidates (ping-pong SoA)
  DevBuf cu[2], cv[2], cs[2];
  uint64_t cand_cap = 0;
  // dense spill tables
  DevBuf tables, touched, range_cursors, range_touched;
  DevBuf fence_slot, fence_tab;              // k_range: fence tables of the long rows (per graph, built at first use)
  bool fence_valid = false;
  uint32_t fence_ncell = 0;
  uint64_t fence_rows = 0;
  int range_fence = 1;                       // NLP_B200_RANGE_FENCE=0: no fence tables
  int range_quarter = 1;                     // NLP_B200_RANGE_QUARTER=0: no byte counters
  DevBuf flt_cnt, flt_off, flt_items, flt_ids, flt_defer, flt_tlist;   // k_range_flt: (source, window range) items
  FltItems flt_it{nullptr, nullptr, nullptr};
  uint64_t flt_n = 0;
  // select / sort scratch
  DevBuf counts, totals, hist, sel, cursor2, sel11;
  Select11* h_sel11 = nullptr;         // pinned (header + digit histograms)
  bool sel11_l0 = false;               // the scoring kernel already accumulated the first select histogram
  DevBuf oc_counts, oc_off;                  // ordered compaction (pair path top-K)
  // pair path: records sorted by (u, v) at (pair_pu, pair_pv), aligned scores in cs[pair_score_buf]
  bool pair_pending = false;
  const uint32_t* pair_pu = nullptr;
  const uint32_t* pair_pv = nullptr;
  int pair_score_buf = 0;
  bool pair_from_cache = false;
  uint64_t pair_n = 0, pair_kept = 0;
  // nlp_set_reuse: sorted wedge records kept per (D, rank, world) so that further measures at the
  // same threshold only reduce + score + select (the nine measures share their counts)
  struct PairCache { DevBuf u, v, w; uint64_t P = 0, first_hop = 0, elig = 0, wedges = 0, stamp = 0; };
  std::map<uint64_t, PairCache> pair_cache;
  int reuse = 0;
  uint64_t cache_bytes = 0, cache_stamp = 0;
  // bucket path (bucket.cuh): the per-(D, bucket size) plan -- items grouped by source, bucket and
  // slot offsets -- is a pure function of the resident graph, built once and kept until the graph
  // changes; `arena` is one allocation that all its arrays are carved from
  struct BucketPlan {
    DevBuf arena;
    BucketPlanDev dev;
    PairItems big_items{nullptr, nullptr, nullptr, nullptr};
    const unsigned long long* b_off = nullptr; const unsigned long long* bg_first = nullptr;
    const unsigned long long* bg_roff = nullptr; const uint32_t* bg_item = nullptr;
    uint64_t E = 0, P = 0, Es = 0, Ps = 0, ns = 0, nbig = 0, Eb = 0, Pb = 0;
    uint64_t first_hop = 0, elig = 0, wedges = 0, stamp = 0;
    std::vector<unsigned long long> h_bg_first;   // host copies: the big sources are few
    std::vector<uint32_t> h_bg_item;
    uint64_t part_key = 0, part_kb0 = 0, part_kb1 = 0, part_slo = 0, part_shi = 0;   // this rank's share (cached per rank/world)
    bool usable = false;                     // false: too large for the scratch budget (source path instead)
  };
  std::map<uint64_t, BucketPlan> plans;
  // nlp_set_reuse on the bucket path: the distinct pairs of a threshold with their counts AFTER the
  // exclusion (12 B per record slot of this rank), shared by the seven count measures
  struct CountStore { DevBuf arena; const uint32_t* u = nullptr; const uint32_t* v = nullptr; const uint32_t* c = nullptr;
                      uint64_t n = 0, candidates = 0, stamp = 0; };
  std::map<uint64_t, CountStore> count_store;
  uint64_t plan_bytes = 0;
  DevBuf plan_tmp;                           // scratch of a plan build (kept: no allocation churn)
  DevBuf al_u, al_v, al_s, al_c;             // record-aligned output of the bucket path (pair, score bits, count)
  const uint32_t* pair_ps = nullptr;         // aligned score bits when the records live outside the candidate buffers
  // asynchronous fetch: result -> staging (device copy on the compute stream) -> caller memory
  // (copy stream), double buffered, so the next prediction overlaps the transfer
  cudaStream_t copy_stream = nullptr;
  cudaStream_t stream2 = nullptr;            // big-source detour of the bucket path, concurrent with k_bucket
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<DevBuf> arena_pool;            // plan arenas of earlier graphs, reused (no cudaMalloc / cudaFree per batch)
  DevBuf stg_u[2], stg_v[2], stg_s[2];
  cudaEvent_t ev_stg_ready[2] = {nullptr, nullptr}, ev_stg_done[2] = {nullptr, nullptr};
  bool stg_busy[2] = {false, false};
  int stg_next = 0;
  // nlp_generate_deletions: per-slot tables of the random stream, orbit, result (sorted unique directed pairs)
  DevBuf bt_d, bt_uat, bt_hit, bt_jump[2], bt_starts, bt_pos, del_u, del_v;
  uint64_t del_n = 0;
  bool has_deletions = false;
  // held-back edges for nlp_evaluate: packed (u << 32 | v) keys, ascending
  DevBuf truth_key, truth_tmp, eval_ctr;
  uint64_t truth_n = 0;
  bool has_truth = false;
  cudaEvent_t ev_eval0 = nullptr, ev_eval1 = nullptr;
  // result
  int res_buf = 0;
  uint64_t res_count = 0;
  bool has_result = false;
  // multi-GPU: NCCL communicator (nlp_comm_init); with it nlp_predict merges across the ranks itself
  void* comm = nullptr;
  bool part_ordered = false;                 // this prediction's partition gave every rank an ascending source range
  DevBuf gcnt, gsend, grecv;
  uint64_t gathered_bytes = 0;               // payload bytes received by all-gathers so far
  // config
  int rank = 0, world = 1;
  uint64_t scratch_limit = 0;
  uint64_t budget_base = 0;                  // see scratch_budget()
  uint64_t launches = 0;
  std::string err;
};

namespace {

#define NLP_CUDA(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      (h)->err = std::string(#expr) + " (nlp_b200.cu:" + std::to_string(__LINE__) + "): " + cudaGetErrorString(e__); \
      cudaGetLastError();   /* a failed cudaMalloc must not resurface at the next launch check */ \
      return NLP_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

int fail(nlp_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

int ensure(nlp_handle* h, DevBuf& b, size_t bytes, bool zero = false) {
  if (bytes <= b.cap && b.p) return NLP_OK;
  if (b.p) { NLP_CUDA(h, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
  if (bytes == 0) bytes = 16;
  NLP_CUDA(h, cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  if (zero) NLP_CUDA(h, cudaMemsetAsync(b.p, 0, bytes, h->stream));
  return NLP_OK;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.cap = 0;
}

#define NLP_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != NLP_OK) return rc__; \
  } while (0)

// NLP_B200_DEBUG_SYNC=1 in the environment synchronises after every launch, so a faulting
// kernel is reported at its own launch site.
static const bool g_debug_sync = [] { const char* e = getenv("NLP_B200_DEBUG_SYNC"); return e && *e == '1'; }();

#define NLP_LAUNCHED(h)                                   \
  do {                                                    \
    (h)->launches++;                                      \
    NLP_CUDA(h, cudaGetLastError());                      \
    if (g_debug_sync) NLP_CUDA(h, cudaStreamSynchronize((h)->stream)); \
  } while (0)

inline unsigned grid_for(uint64_t items, unsigned per_block, unsigned cap_blocks) {
  uint64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap_blocks) g = cap_blocks;
  return (unsigned)g;
}

int read_counters(nlp_handle* h) {
  NLP_CUDA(h, cudaMemcpyAsync(h->h_ctr, h->ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLP_OK;
}

DevGraph dev_graph(const nlp_handle* h) {
  DevGraph g;
  g.off = h->d_off; g.keys = h->d_keys; g.deg = (const uint32_t*)h->deg.p; g.S = h->S;
  return g;
}

// ---- graph -----------------------------------------------------------------------------------
// out[i] = sum of in[0..i); returns the grand total in *total (host).  in/out may alias for u64.
template <class TIn>
int exclusive_scan(nlp_handle* h, const TIn* in, uint64_t n, unsigned long long* out, uint64_t* total) {
  if (total) *total = 0;
  if (!n) return NLP_OK;
  const uint32_t ntiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
  NLP_TRY(ensure(h, h->scan_tiles, (size_t)ntiles * 8));
  NLP_TRY(ensure(h, h->scan_total, 16));
  k_scan_tiles<TIn><<<ntiles, SCAN_THREADS, 0, h->stream>>>(in, n, (unsigned long long*)h->scan_tiles.p);
  NLP_LAUNCHED(h);
  k_scan_spine<<<1, SCAN_THREADS, 0, h->stream>>>((unsigned long long*)h->scan_tiles.p, ntiles,
                                                  (unsigned long long*)h->scan_total.p);
  NLP_LAUNCHED(h);
  k_scan_apply<TIn><<<ntiles, SCAN_THREADS, 0, h->stream>>>(in, n, (const unsigned long long*)h->scan_tiles.p, out);
  NLP_LAUNCHED(h);
  if (!total) return NLP_OK;                 // the caller knows the total already: no host round trip
  unsigned long long t = 0;
  NLP_CUDA(h, cudaMemcpyAsync(&t, h->scan_total.p, 8, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  *total = t;
  return NLP_OK;
}

int measure_budget(nlp_handle* h);

void clear_pair_cache(nlp_handle* h) {
  for (auto& kv : h->pair_cache) { release(kv.second.u); release(kv.second.v); release(kv.second.w); }
  h->pair_cache.clear();
  h->cache_bytes = 0;
}

// The arenas of dropped plans are kept for the plans of the next graph: a batch of the harness
// rebuilds its plans on a graph of nearly the same size, and cudaFree / cudaMalloc of hundreds of
// megabytes costs milliseconds (sometimes hundreds) each.
void retire_arena(nlp_handle* h, DevBuf& a) {
  if (!a.p) return;
  if (h->arena_pool.size() >= 8) {                   // bounded: drop the smallest
    size_t k = 0;
    for (size_t i = 1; i < h->arena_pool.size(); ++i) if (h->arena_pool[i].cap < h->arena_pool[k].cap) k = i;
    if (h->arena_pool[k].cap < a.cap) std::swap(h->arena_pool[k], a);
    release(a);
    return;
  }
  h->arena_pool.push_back(a);
  a.p = nullptr; a.cap = 0;
}

bool take_arena(nlp_handle* h, DevBuf& a, size_t bytes) {
  size_t best = h->arena_pool.size();
  for (size_t i = 0; i < h->arena_pool.size(); ++i)
    if (h->arena_pool[i].cap >= bytes && (best == h->arena_pool.size() || h->arena_pool[i].cap < h->arena_pool[best].cap)) best = i;
  if (best < h->arena_pool.size()) {
    a = h->arena_pool[best];
    h->arena_pool.erase(h->arena_pool.begin() + best);
    return true;
  }
  const size_t want = bytes + bytes / 16;            // a little headroom for the next, slightly different graph
  if (cudaMalloc(&a.p, want) != cudaSuccess) {
    cudaGetLastError();
    for (auto& b : h->arena_pool) release(b);        // give the pooled memory back and try once more
    h->arena_pool.clear();
    if (cudaMalloc(&a.p, bytes) != cudaSuccess) { cudaGetLastError(); a.p = nullptr; a.cap = 0; return false; }
    a.cap = bytes;
    return true;
  }
  a.cap = want;
  return true;
}

void clear_count_store(nlp_handle* h) {
  for (auto& kv : h->count_store) { h->plan_bytes -= std::min<uint64_t>(h->plan_bytes, kv.second.arena.cap); retire_arena(h, kv.second.arena); }
  h->count_store.clear();
}

void clear_plans(nlp_handle* h) {
  clear_count_store(h);
  for (auto& kv : h->plans) retire_arena(h, kv.second.arena);
  h->plans.clear();
  h->plan_bytes = 0;
}

// trusted: the CSR was produced by this library from a validated graph (nlp_apply_deletions): no
// validation pass; `known_sym` (1 symmetric, 2 not, 0 unknown) and the old multiplicity bound carry over.
int finish_graph(nlp_handle* h, bool trusted = false, int known_sym = 0) {
  const uint32_t S = h->S;
  const uint32_t old_maxmult = h->maxmult;
  NLP_TRY(ensure(h, h->deg, (size_t)S * 4));
  NLP_TRY(ensure(h, h->work, (size_t)S * 4));
  NLP_TRY(ensure(h, h->work64, (size_t)S * 8));
  NLP_TRY(ensure(h, h->chunk_base, (size_t)S * 8));
  NLP_TRY(ensure(h, h->elig, ((size_t)S + 31) / 32 * 4));
  NLP_TRY(ensure(h, h->maxdeg_dev, 32));
  for (int b = 0; b < NBINS; ++b) {
    NLP_TRY(ensure(h, h->list[b], (size_t)S * 4));
    if (b >= 2) NLP_TRY(ensure(h, h->defer[b], (size_t)S * 4));
  }
  NLP_CUDA(h, cudaMemsetAsync(h->maxdeg_dev.p, 0, 32, h->stream));
  uint64_t m = 0;
  NLP_CUDA(h, cudaMemcpyAsync(&m, h->d_off + S, 8, cudaMemcpyDeviceToHost, h->stream));
  h->nchunks = 0;
  if (S) {
    k_degrees<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(h->d_off, S, (uint32_t*)h->deg.p,
                                                                        (unsigned long long*)h->chunk_base.p,
                                                                        (uint32_t*)h->maxdeg_dev.p);
    NLP_LAUNCHED(h);
    // chunk table of the long rows: chunk_base[u] = first chunk of row u, chunk_src[c] = its row
    NLP_TRY(exclusive_scan<unsigned long long>(h, (const unsigned long long*)h->chunk_base.p, S,
                                               (unsigned long long*)h->chunk_base.p, &h->nchunks));
    NLP_TRY(ensure(h, h->chunk_src, (size_t)h->nchunks * 4));
    NLP_TRY(ensure(h, h->chunk_cnt, (size_t)h->nchunks * 4));
    if (h->nchunks) {
      k_chunk_fill<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(
          (const uint32_t*)h->deg.p, (const unsigned long long*)h->chunk_base.p, S, (uint32_t*)h->chunk_src.p);
      NLP_LAUNCHED(h);
    }
  }
  // validate the entries (keys below span, rows sorted) and measure the largest entry
  // multiplicity in the same pass; the offsets were checked by k_degrees -- a CSR whose offsets are
  // not monotone must not be walked at all
  uint32_t info[4] = {0, 0, 0, 0};                  // {max degree, validation flags, max multiplicity, -}
  NLP_CUDA(h, cudaMemcpyAsync(info, h->maxdeg_dev.p, 16, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (info[1] & 1u) return fail(h, NLP_ERR_ARG, "graph: offsets are not non-decreasing");
  h->M = m;
  unsigned long long fp = 0;
  if (m && S && !trusted) {
    DevGraph g = dev_graph(h);
    k_validate_entries<<<grid_for(m, 256, h->num_sms * 16), 256, 0, h->stream>>>(g, m, (unsigned int*)h->maxdeg_dev.p + 1,
                                                                                   (unsigned int*)h->maxdeg_dev.p + 2,
                                                                                   (unsigned long long*)h->maxdeg_dev.p + 2);
    NLP_LAUNCHED(h);
    NLP_CUDA(h, cudaMemcpyAsync(info, h->maxdeg_dev.p, 16, cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(&fp, (const char*)h->maxdeg_dev.p + 16, 8, cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (info[1] & 2u) return fail(h, NLP_ERR_ARG, "graph: a key is not below span");
    if (info[1] & 4u) return fail(h, NLP_ERR_ARG, "graph: a row is not sorted ascending");
  }
  const uint32_t md = info[0];
  h->maxdeg = md;
  h->gtable_n = 0;
  h->maxmult = trusted ? (old_maxmult > 1u ? old_maxmult : 1u) : (info[2] > 1u ? info[2] : 1u);
  // what is already known about this very graph (same span, entry count and content fingerprint:
  // the base graph of a batch loop is bound again for every batch)
  h->graph_id = trusted ? 0ull : (fp ^ ((unsigned long long)S << 40) ^ (m * 0x9E3779B97F4A7C15ull)) | 1ull;
  h->sym_state = known_sym;
  if (!trusted) {
    auto known = h->known_sym.find(h->graph_id);
    if (known != h->known_sym.end()) h->sym_state = known->second;
  }
  h->pair_sizes.clear();
  clear_pair_cache(h);
  clear_plans(h);
  h->fence_valid = false;
  NLP_TRY(measure_budget(h));
  h->has_graph = true;
  h->has_result = false;
  return NLP_OK;
}

// Adamic-Adar term table, gtable[d] = 1.0 / log((double)d) computed with the host libm so it
// is the very double the reference's lambda produces (inc/predict.hxx:788).
int ensure_gtable(nlp_handle* h) {
  const uint32_t n = h->maxdeg + 1;
  if (h->gtable_n == n && h->gtable.p) return NLP_OK;
  std::vector<double> t(n);
  for (uint32_t d = 0; d < n; ++d) t[d] = 1.0 / std::log((double)d);
  NLP_TRY(ensure(h, h->gtable, (size_t)n * 8));
  NLP_CUDA(h, cudaMemcpyAsync(h->gtable.p, t.data(), (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->gtable_n = n;
  return NLP_OK;
}

// ---- top-K ------------------------------------------------------------------------------------
// Sort the first n entries of candidate buffer `buf` by the canonical key; returns the buffer
// that holds the sorted entries.
int radix_pass(nlp_handle* h, int& buf, uint64_t n, uint32_t nblocks, int pass, bool has3) {
  const int word = pass / 4, shift = (pass % 4) * 8;
  const int o = buf ^ 1;
  k_tilehist<<<nblocks, SORT_THREADS, 0, h->stream>>>((const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p,
                                                      (const uint32_t*)h->cs[buf].p, n, word, shift,
                                                      (uint32_t*)h->counts.p, nblocks);
  NLP_LAUNCHED(h);
  k_rowscan<<<256, 256, 0, h->stream>>>((uint32_t*)h->counts.p, nblocks, (uint32_t*)h->totals.p);
  NLP_LAUNCHED(h);
  if (has3) {
    k_scatter<true><<<nblocks, SORT_THREADS, scatter_smem_bytes(true), h->stream>>>(
        (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p,
        (uint32_t*)h->cu[o].p, (uint32_t*)h->cv[o].p, (uint32_t*)h->cs[o].p, n, word, shift,
        (const uint32_t*)h->counts.p, nblocks, (const uint32_t*)h->totals.p);
  } else {
    k_scatter<false><<<nblocks, SORT_THREADS, scatter_smem_bytes(false), h->stream>>>(
        (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p,
        (uint32_t*)h->cu[o].p, (uint32_t*)h->cv[o].p, (uint32_t*)h->cs[o].p, n, word, shift,
        (const uint32_t*)h->counts.p, nblocks, (const uint32_t*)h->totals.p);
  }
  NLP_LAUNCHED(h);
  buf = o;
  return NLP_OK;
}

int radix_sort(nlp_handle* h, int buf, uint64_t n, int* out_buf, int first_pass = 0) {
  *out_buf = buf;
  if (n < 2) return NLP_OK;
  NLP_CUDA(h, cudaMemsetAsync(h->hist.p, 0, 12 * 256 * 8, h->stream));
  k_prehist<<<grid_for(n, 256 * 8, h->num_sms * 8), 256, 0, h->stream>>>(
      (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p, n,
      (unsigned long long*)h->hist.p);
  NLP_LAUNCHED(h);
  NLP_CUDA(h, cudaMemcpyAsync(h->h_hist, h->hist.p, 12 * 256 * 8, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  const uint32_t nblocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  NLP_TRY(ensure(h, h->counts, (size_t)nblocks * 256 * 4));
  for (int pass = first_pass; pass < 12; ++pass) {
    bool constant = false;
    for (int d = 0; d < 256; ++d)
      if (h->h_hist[pass * 256 + d] == n) { constant = true; break; }
    if (constant) continue;                      // every key has the same digit here
    NLP_TRY(radix_pass(h, buf, n, nblocks, pass, true));
  }
  *out_buf = buf;
  return NLP_OK;
}

// Stable sort of (u, v[, payload]) records by (u, v) only; ids are < span, so the digits above
// its bit width are skipped without looking at the data.
int radix_sort_pairs(nlp_handle* h, int buf, uint64_t n, bool has3, int* out_buf) {
  *out_buf = buf;
  if (n < 2) return NLP_OK;
  const uint32_t nblocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  NLP_TRY(ensure(h, h->counts, (size_t)nblocks * 256 * 4));
  const uint32_t top = h->S ? h->S - 1 : 0;
  for (int pass = 0; pass < 8; ++pass) {
    if ((top >> ((pass % 4) * 8)) == 0) continue;
    NLP_TRY(radix_pass(h, buf, n, nblocks, pass, has3));
  }
  *out_buf = buf;
  return NLP_OK;
}

// Best min(n, K) of the n entries in buffer `buf`, canonical order.
int top_k(nlp_handle* h, int buf, uint64_t n, uint64_t K, int* out_buf, uint64_t* out_n) {
  if (K < n) {
    // MSD radix select: narrow to (items before the K-th's bucket) + (that bucket)
    NLP_CUDA(h, cudaMemsetAsync(h->sel.p, 0, sizeof(SelectState), h->stream));
    const uint64_t slack = std::max<uint64_t>(K / 8, 65536);
    uint64_t above = 0, bucket = n;
    uint32_t bits = 0;
    bool done = false;
    while (!done) {
      for (int lvl = 0; lvl < 4; ++lvl) {        // four levels queued back to back, one read of the state
        k_select_hist<<<grid_for(n, 256 * 8, h->num_sms * 8), 256, 0, h->stream>>>(
            (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p, n,
            (SelectState*)h->sel.p);
        NLP_LAUNCHED(h);
        k_select_step<<<1, 1, 0, h->stream>>>((SelectState*)h->sel.p, (unsigned long long)K, (unsigned long long)slack, 96u);
        NLP_LAUNCHED(h);
      }
      NLP_CUDA(h, cudaMemcpyAsync(h->h_sel, h->sel.p, offsetof(SelectState, hist), cudaMemcpyDeviceToHost, h->stream));
      NLP_CUDA(h, cudaStreamSynchronize(h->stream));
      above = h->h_sel->above; bucket = h->h_sel->bucket; bits = h->h_sel->bits; done = h->h_sel->done != 0;
    }
    if (bits > 0 && above + bucket < n) {
      const int o = buf ^ 1;
      NLP_CUDA(h, cudaMemsetAsync(h->cursor2.p, 0, 8, h->stream));
      k_select_compact<<<grid_for(n, 256 * 8, h->num_sms * 8), 256, 0, h->stream>>>(
          (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p, n,
          (const SelectState*)h->sel.p, (uint32_t*)h->cu[o].p, (uint32_t*)h->cv[o].p, (uint32_t*)h->cs[o].p,
          (unsigned long long*)h->cursor2.p);
      NLP_LAUNCHED(h);
      buf = o;
      n = above + bucket;
    }
  }
  NLP_TRY(radix_sort(h, buf, n, out_buf));
  *out_n = std::min(n, K);
  return NLP_OK;
}

// Grow the six candidate arrays together or not at all: the new arrays are allocated first, the old
// ones are freed only once all six exist, so a failed allocation leaves the handle as it was
// (cand_cap still describes real buffers).  *ok = false reports an out-of-memory condition to
// callers that can fall back to another path; with ok == nullptr it is an error.
int ensure_candidates(nlp_handle* h, uint64_t cap, bool* ok = nullptr) {
  if (ok) *ok = true;
  if (cap < 1024) cap = 1024;
  if (cap <= h->cand_cap) return NLP_OK;
  // padded to whole sort tiles: k_scatter reads full tiles with TMA bulk copies
  const uint64_t padded = (cap + SORT_TILE - 1) / SORT_TILE * SORT_TILE + SORT_TILE;
  void* fresh[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 6; ++i) {
    const cudaError_t e = cudaMalloc(&fresh[i], padded * 4);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int k = 0; k < i; ++k) cudaFree(fresh[k]);
      if (ok) { *ok = false; return NLP_OK; }
      h->err = std::string("candidate buffers (") + std::to_string(padded * 24) + " bytes): " + cudaGetErrorString(e);
      return NLP_ERR_CUDA;
    }
  }
  DevBuf* bufs[6] = {&h->cu[0], &h->cv[0], &h->cs[0], &h->cu[1], &h->cv[1], &h->cs[1]};
  for (int i = 0; i < 6; ++i) {
    if (bufs[i]->p) cudaFree(bufs[i]->p);
    bufs[i]->p = fresh[i]; bufs[i]->cap = padded * 4;
  }
  h->cand_cap = cap;
  h->has_result = false;
  return NLP_OK;
}

// ---- wedge kernels ---------------------------------------------------------------------------
struct HashCfg { int threads; int log2_slots; };

inline HashCfg hash_cfg(int bin, bool flt) {
  HashCfg c;
  c.log2_slots = bin == 2 ? 10 : bin == 3 ? 12 : 14;
  c.threads = flt ? 32 : (bin == 2 ? 64 : bin == 3 ? 128 : 512);
  return c;
}

template <bool FLT, bool ADMIT>
int launch_hash(nlp_handle* h, const Params& p, int bin, const uint32_t* list, uint32_t n, uint32_t* deferred) {
  if (!n) return NLP_OK;
  const HashCfg c = hash_cfg(bin, FLT);
  // slots {key,value} + u16 list of the claimed slots (at most bin_limit of them)
  const size_t smem = (((size_t)8) << c.log2_slots) + (size_t)bin_limit(bin) * 2;
  auto kern = k_hash<FLT, ADMIT>;
  NLP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  int occ = 1;
  NLP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, c.threads, smem));
  if (occ < 1) occ = 1;
  const unsigned grid = (unsigned)std::min<uint64_t>(n, (uint64_t)h->num_sms * occ);
  kern<<<grid, c.threads, smem, h->stream>>>(p, list, n, bin, deferred, c.log2_slots);
  NLP_LAUNCHED(h);
  return NLP_OK;
}

template <bool FLT, bool ADMIT>
int launch_dense(nlp_handle* h, const Params& p, const uint32_t* list, uint32_t n, uint32_t* deferred,
                 unsigned slots, uint64_t touched_cap) {
  if (!n) return NLP_OK;
  const unsigned grid = (unsigned)std::min<uint64_t>(n, slots);
  k_dense<FLT, ADMIT><<<grid, FLT ? 32 : 512, 0, h->stream>>>(p, list, n, 5, deferred, (uint32_t*)h->tables.p,
                                                              (uint32_t*)h->touched.p, touched_cap);
  NLP_LAUNCHED(h);
  return NLP_OK;
}

// Count measures: hub-heavy sources on windowed shared-memory counters (k_range).
constexpr uint32_t RANGE_COUNTERS = 52 * 1024;      // 208 KB of u32 counters per block (+13 KB static)

// Fence tables of the rows that are long against the number of window cells (wedge.cuh, range_batch):
// fence[c] = first position of the row with key >= c * RANGE_COUNTERS.  A row gets one when it holds
// at least two entries per cell, so all tables together take at most 2 bytes per entry of the graph.
int ensure_fences(nlp_handle* h) {
  if (h->fence_valid) return NLP_OK;
  const uint32_t S = h->S;
  const uint32_t ncell = (uint32_t)(((uint64_t)S + RANGE_COUNTERS - 1) / RANGE_COUNTERS);
  h->fence_ncell = ncell;
  h->fence_rows = 0;
  h->fence_valid = true;
  uint32_t min_deg = std::max<uint32_t>(2u * ncell, 64u);
  if (!h->range_fence || !S || h->maxdeg < min_deg) return NLP_OK;
  const DevGraph g = dev_graph(h);
  NLP_TRY(ensure(h, h->fence_slot, (size_t)S * 4));
  DevBuf tmp;
  NLP_TRY(ensure(h, tmp, (size_t)S * 8));
  uint64_t rows = 0;
  for (int attempt = 0; attempt < 8; ++attempt, min_deg *= 2u) {
    k_fence_mark<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(g, min_deg, (uint32_t*)h->fence_slot.p);
    NLP_LAUNCHED(h);
    const int rc = exclusive_scan<uint32_t>(h, (const uint32_t*)h->fence_slot.p, S, (unsigned long long*)tmp.p, &rows);
    if (rc != NLP_OK) { release(tmp); return rc; }
    if (rows * (uint64_t)(ncell + 1u) * 4u <= h->budget_base / 8) break;
    rows = 0;
  }
  if (rows) {
    const int rc = ensure(h, h->fence_tab, (size_t)rows * (ncell + 1u) * 4u);
    if (rc != NLP_OK) { release(tmp); return rc; }
    k_fence_fill<<<grid_for((uint64_t)S * 32, 256, h->num_sms * 16), 256, 0, h->stream>>>(g, min_deg, RANGE_COUNTERS, ncell,
                                                                                            (const unsigned long long*)tmp.p,
                                                                                            (uint32_t*)h->fence_slot.p, (uint32_t*)h->fence_tab.p);
    NLP_LAUNCHED(h);
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));      // tmp is freed below
  }
  release(tmp);
  h->fence_rows = rows;
  return NLP_OK;
}

template <bool FLT, bool ADMIT>
int launch_range(nlp_handle* h, const Params& p, const uint32_t* list, uint32_t n, uint32_t* deferred) {
  if (!n) return NLP_OK;
  const uint64_t stride = ((uint64_t)h->maxdeg + CHUNK + 31) / 32 * 32;
  if constexpr (FLT) {
    // one warp per source: per-warp row records (16 B) + deg(w) of every first-hop row
    const size_t smem = (size_t)RFLT_WARPS * RFLT_WIN * 4;
    NLP_CUDA(h, cudaFuncSetAttribute(k_range_flt<ADMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)n + RFLT_WARPS - 1) / RFLT_WARPS, (uint64_t)h->num_sms);
    NLP_TRY(ensure(h, h->range_cursors, (size_t)h->num_sms * RFLT_WARPS * stride * 16));
    NLP_TRY(ensure(h, h->range_touched, (size_t)h->num_sms * RFLT_WARPS * stride * 8));
    NLP_TRY(ensure(h, h->flt_tlist, (size_t)h->num_sms * RFLT_WARPS * RFLT_WIN * 4));
    k_range_flt<ADMIT><<<grid, RFLT_WARPS * 32, smem, h->stream>>>(p, h->flt_it, list, n, 6, deferred, (uint4*)h->range_cursors.p,
                                                                   (double*)h->range_touched.p, stride, (uint32_t*)h->flt_tlist.p);
    NLP_LAUNCHED(h);
    return NLP_OK;
  } else {
  NLP_TRY(ensure_fences(h));
  RangeFences fx;
  fx.slot_of = h->fence_rows ? (const uint32_t*)h->fence_slot.p : nullptr;
  fx.fence = (const uint32_t*)h->fence_tab.p;
  fx.ncell = h->fence_ncell;
  const size_t smem = (size_t)RANGE_COUNTERS * 4;
  NLP_CUDA(h, cudaFuncSetAttribute(k_range<ADMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)std::min<uint64_t>(n, (uint64_t)h->num_sms);
  // per block: row cursor + row end for every first-hop entry of its current source
  NLP_TRY(ensure(h, h->range_cursors, (size_t)h->num_sms * 2 * stride * 8));
  // per block: the vertices of its current window that have a count (at most 4 * RANGE_COUNTERS)
  NLP_TRY(ensure(h, h->range_touched, (size_t)h->num_sms * 4 * RANGE_COUNTERS * 4));
  k_range<ADMIT><<<grid, RANGE_THREADS, smem, h->stream>>>(p, fx, list, n, 6, deferred, RANGE_COUNTERS,
                                                           (unsigned long long*)h->range_cursors.p, stride,
                                                           (uint32_t*)h->range_touched.p);
  NLP_LAUNCHED(h);
  return NLP_OK;
  }
}

template <bool FLT>
int launch_tiny(nlp_handle* h, const Params& p, int bin, const uint32_t* list, uint32_t n) {
  if (!n) return NLP_OK;
  if (bin == 0) {
    k_tiny<8, FLT><<<grid_for(n, 8 * 4, h->num_sms * 16), 256, 0, h->stream>>>(p, list, n);
  } else {
    k_tiny<32, FLT><<<grid_for(n, 8, h->num_sms * 16), 256, 0, h->stream>>>(p, list, n);
  }
  NLP_LAUNCHED(h);
  return NLP_OK;
}

// Bytes of GPU scratch a prediction may use: what was free when the graph was set plus what the
// handle already held for scratch then.  Measured once per graph -- cudaMemGetInfo costs
// milliseconds when the process holds many allocations, far too much for every prediction.
int scratch_budget(nlp_handle* h, uint64_t* out) {
  uint64_t budget = h->budget_base;
  if (h->scratch_limit && h->scratch_limit < budget) budget = h->scratch_limit;
  *out = budget;
  return NLP_OK;
}

int measure_budget(nlp_handle* h) {
  size_t free_b = 0, total_b = 0;
  NLP_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
  // what is free now, plus what the handle already holds of exactly the buffers the plans below
  // are checked against (candidate arrays, aligned output, plan arenas).  Scratch that a path keeps
  // for itself (eligible lists, spill tables, item descriptors) is NOT headroom: it stays allocated.
  uint64_t budget = (uint64_t)free_b;
  for (int b = 0; b < 2; ++b) budget += h->cu[b].cap + h->cv[b].cap + h->cs[b].cap;
  budget += h->al_u.cap + h->al_v.cap + h->al_s.cap + h->al_c.cap + h->plan_tmp.cap;
  for (const auto& a : h->arena_pool) budget += a.cap;
  h->budget_base = budget / 10 * 8;
  return NLP_OK;
}

// Rows symmetric (entry multiplicities mirrored)?  Checked once per graph on the device; the
// w-centric LHub paths are only admissible then.
int check_symmetry(nlp_handle* h) {
  if (h->sym_state != 0) return NLP_OK;
  const DevGraph g = dev_graph(h);
  NLP_TRY(ensure(h, h->sym_flag, 32));
  NLP_CUDA(h, cudaMemsetAsync(h->sym_flag.p, 0, 32, h->stream));
  if (h->M) {
    k_symmetry<<<grid_for(h->M, 256, h->num_sms * 16), 256, 0, h->stream>>>(g, h->M, (unsigned int*)h->sym_flag.p,
                                                                              (unsigned long long*)h->sym_flag.p + 1);
    NLP_LAUNCHED(h);
  }
  unsigned long long sf[3] = {0, 0, 0};              // {asym flag, entries u < w, entries u > w}
  NLP_CUDA(h, cudaMemcpyAsync(sf, h->sym_flag.p, 24, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  const bool f = (unsigned int)sf[0] != 0u || sf[1] != sf[2];
  h->sym_state = f ? 2 : 1;
  if (h->graph_id) {
    if (h->known_sym.size() >= 64) h->known_sym.clear();
    h->known_sym[h->graph_id] = h->sym_state;
    if (h->has_base && h->graph_id == h->base_id) h->base_sym = h->sym_state;
  }
  // the check is graph preparation, not part of the prediction: restart the clock
  NLP_CUDA(h, cudaEventRecord(h->ev_start, h->stream));
  return NLP_OK;
}

// ---- bucket path (bucket.cuh) ----------------------------------------------------------------
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += (n ? n : 1) * sizeof(T);
    return p;
  }
};

// cudaMalloc that reports failure instead of raising an ABI error: the caller falls back to
// another path (the sticky-free allocation error is cleared).
bool try_ensure(nlp_handle* h, DevBuf& b, size_t bytes) {
  (void)h;
  if (bytes <= b.cap && b.p) return true;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  if (bytes == 0) bytes = 16;
  if (cudaMalloc(&b.p, bytes) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; b.cap = 0; return false; }
  b.cap = bytes;
  return true;
}

// Build the plan of threshold D for buckets of 2 * half records.  plan.usable stays false when it
// does not fit the scratch budget (the caller runs another path).
int build_plan(nlp_handle* h, uint32_t D, uint32_t half, nlp_handle::BucketPlan& plan) {
  const uint32_t S = h->S;
  const DevGraph g = dev_graph(h);
  Counters* hc = h->h_ctr;
  plan.usable = false;
  uint64_t budget = 0;
  NLP_TRY(scratch_budget(h, &budget));
  NLP_CUDA(h, cudaMemsetAsync(h->ctr.p, 0, sizeof(Counters), h->stream));
  k_plan_rows<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(g, D, (uint32_t*)h->work.p, (Counters*)h->ctr.p);
  NLP_LAUNCHED(h);
  uint64_t E = 0;
  NLP_TRY(exclusive_scan<uint32_t>(h, (const uint32_t*)h->work.p, S, (unsigned long long*)h->work64.p, &E));
  NLP_TRY(read_counters(h));
  plan.first_hop = hc->first_hop; plan.elig = hc->eligible_first_hop; plan.wedges = hc->wedges;
  plan.E = E;
  if (E == 0) { plan.usable = true; return NLP_OK; }            // no wedge with v > u at all
  if (E >= 0xfffffff0ull || E * 120 > budget / 2) return NLP_OK;
  NLP_TRY(ensure_candidates(h, E));
  // scratch of the build, one allocation
  uint32_t *it_cnt, *it_dw, *g_cnt, *head, *f_item, *f_cnt, *f_src;
  unsigned long long *it_ptr, *hs, *rc, *si, *sr, *ks, *src_roff;
  auto carve_tmp = [&](Carver& c) {
    it_cnt = c.take<uint32_t>(E); it_dw = c.take<uint32_t>(E); it_ptr = c.take<unsigned long long>(E);
    g_cnt = c.take<uint32_t>(E); head = c.take<uint32_t>(E); hs = c.take<unsigned long long>(E);
    rc = c.take<unsigned long long>(E); f_item = c.take<uint32_t>(E); f_cnt = c.take<uint32_t>(E);
    f_src = c.take<uint32_t>(E); si = c.take<unsigned long long>(E); sr = c.take<unsigned long long>(E);
    ks = c.take<unsigned long long>(E); src_roff = c.take<unsigned long long>(E + 1);
  };
  { Carver c(nullptr); carve_tmp(c); if (!try_ensure(h, h->plan_tmp, c.off + 256)) return NLP_OK; }
  { Carver c(h->plan_tmp.p); carve_tmp(c); }
  const unsigned gS = grid_for(S, 256, h->num_sms * 8), gE = grid_for(E, 256, h->num_sms * 16);
  k_plan_items<<<grid_for((uint64_t)S, 8 * 32, h->num_sms * 16), 256, 0, h->stream>>>(g, (const uint32_t*)h->work.p, (const unsigned long long*)h->work64.p,
                                          (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, it_cnt, it_dw, it_ptr);
  NLP_LAUNCHED(h);
  int buf = 0;
  {   // stable sort of (u, item index) by u: only the digits below bits(S - 1)
    const uint32_t nblocks = (uint32_t)((E + SORT_TILE - 1) / SORT_TILE);
    NLP_TRY(ensure(h, h->counts, (size_t)nblocks * 256 * 4));
    const uint32_t top = S ? S - 1 : 0;
    for (int pass = 4; pass < 8; ++pass) {
      if ((top >> ((pass % 4) * 8)) == 0) continue;
      NLP_TRY(radix_pass(h, buf, E, nblocks, pass, false));
    }
  }
  const uint32_t* su = (const uint32_t*)h->cu[buf].p;
  const uint32_t* sidx = (const uint32_t*)h->cv[buf].p;
  k_plan_heads<<<gE, 256, 0, h->stream>>>(su, sidx, it_cnt, E, g_cnt, head);
  NLP_LAUNCHED(h);
  uint64_t nsrc = 0, P = 0;
  NLP_TRY(exclusive_scan<uint32_t>(h, head, E, hs, &nsrc));
  NLP_TRY(exclusive_scan<uint32_t>(h, g_cnt, E, rc, &P));
  plan.P = P;
  if (P >= 0xfffffff0ull) return NLP_OK;
  k_plan_sources<<<gE, 256, 0, h->stream>>>(head, hs, rc, E, nsrc, P, src_roff);
  NLP_LAUNCHED(h);
  k_plan_class<<<gE, 256, 0, h->stream>>>(head, hs, src_roff, g_cnt, E, half, f_item, f_cnt, f_src);
  NLP_LAUNCHED(h);
  uint64_t Es = 0, Ps = 0, ns = 0;
  NLP_TRY(exclusive_scan<uint32_t>(h, f_item, E, si, &Es));
  NLP_TRY(exclusive_scan<uint32_t>(h, f_cnt, E, sr, &Ps));
  NLP_TRY(exclusive_scan<uint32_t>(h, f_src, nsrc, ks, &ns));
  const uint64_t nbig = nsrc - ns, Eb = E - Es, Pb = P - Ps;
  plan.Es = Es; plan.Ps = Ps; plan.ns = ns; plan.nbig = nbig; plan.Eb = Eb; plan.Pb = Pb;
  // the plan itself, one allocation
  PlanScatterOut o;
  uint32_t* bk_first = nullptr;
  auto carve_plan = [&](Carver& c) {
    o.sm_u = c.take<uint32_t>(ns); o.sm_item = c.take<uint32_t>(ns + 1);
    o.sm_soff = c.take<unsigned long long>(ns + 1); o.sm_roff = c.take<unsigned long long>(ns);
    o.s_cnt = c.take<uint32_t>(Es); o.s_dw = c.take<uint32_t>(Es); o.s_ptr = c.take<unsigned long long>(Es);
    o.s_src = c.take<uint32_t>(Es); o.s_loff = c.take<unsigned long long>(Es);
    o.b_u = c.take<uint32_t>(Eb); o.b_cnt = c.take<uint32_t>(Eb); o.b_dw = c.take<uint32_t>(Eb);
    o.b_ptr = c.take<unsigned long long>(Eb); o.b_off = c.take<unsigned long long>(Eb);
    o.bg_first = c.take<unsigned long long>(nbig + 1); o.bg_roff = c.take<unsigned long long>(nbig);
    o.bg_item = c.take<uint32_t>(nbig + 1);
    bk_first = c.take<uint32_t>((Ps + half - 1) / half + 2);
  };
  size_t bytes = 0;
  { Carver c(nullptr); carve_plan(c); bytes = c.off + 256; }
  // keep the plans of a sweep (main.cxx: 11 thresholds) within a quarter of the budget, oldest first out
  while (h->plan_bytes + bytes > budget / 4 && !h->plans.empty()) {
    auto lru = h->plans.end();
    for (auto i = h->plans.begin(); i != h->plans.end(); ++i)
      if (&i->second != &plan && (lru == h->plans.end() || i->second.stamp < lru->second.stamp)) lru = i;
    if (lru == h->plans.end()) break;
    h->plan_bytes -= lru->second.arena.cap;
    retire_arena(h, lru->second.arena);
    h->plans.erase(lru);
  }
  if (h->plan_bytes + bytes > budget / 4) return NLP_OK;
  if (!take_arena(h, plan.arena, bytes)) return NLP_OK;
  h->plan_bytes += plan.arena.cap;
  { Carver c(plan.arena.p); carve_plan(c); }
  k_plan_scatter<<<gE, 256, 0, h->stream>>>(su, sidx, it_dw, it_ptr, g_cnt, head, hs, rc, f_item, si, sr, ks,
                                            E, Es, Ps, ns, nbig, P, o);
  NLP_LAUNCHED(h);
  {
    const uint64_t nb = (Ps + half - 1) / half;
    k_plan_buckets<<<grid_for(nb + 1, 256, h->num_sms * 8), 256, 0, h->stream>>>(o.sm_soff, (uint32_t)ns, half, nb, bk_first);
    NLP_LAUNCHED(h);
  }
  plan.dev.bk_first = bk_first;
  plan.dev.sm_u = o.sm_u; plan.dev.sm_item = o.sm_item; plan.dev.sm_soff = o.sm_soff; plan.dev.sm_roff = o.sm_roff;
  plan.dev.s_cnt = o.s_cnt; plan.dev.s_dw = o.s_dw; plan.dev.s_ptr = o.s_ptr; plan.dev.s_src = o.s_src; plan.dev.s_loff = o.s_loff;
  plan.dev.ns = (uint32_t)ns; plan.dev.half = half;
  plan.big_items.u = o.b_u; plan.big_items.cnt = o.b_cnt; plan.big_items.dw = o.b_dw; plan.big_items.ptr = o.b_ptr;
  plan.b_off = o.b_off; plan.bg_first = o.bg_first; plan.bg_roff = o.bg_roff; plan.bg_item = o.bg_item;
  plan.h_bg_first.assign(nbig + 1, 0); plan.h_bg_item.assign(nbig + 1, 0);
  NLP_CUDA(h, cudaMemcpyAsync(plan.h_bg_first.data(), o.bg_first, (nbig + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(plan.h_bg_item.data(), o.bg_item, (nbig + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  plan.usable = true;
  return NLP_OK;
}

// Big sources of a plan (more records than half a bucket): wedge records to the candidate buffers,
// global stable radix sort by (u, v), run reduce + exclusion + scoring (pairs.cuh), results into
// their slots of the aligned output.  Launches on h->stream.
template <bool FLT>
int big_detour(nlp_handle* h, const nlp_handle::BucketPlan& plan, const Params& p, uint64_t kb0, uint64_t kb1,
               uint32_t* al_u, uint32_t* al_v, uint32_t* al_s, uint32_t* al_c, bool capture) {
  const uint64_t ib0 = plan.h_bg_item[kb0], ib1 = plan.h_bg_item[kb1];
  const uint64_t r0 = plan.h_bg_first[kb0], r1 = plan.h_bg_first[kb1];
  const uint64_t Eb = ib1 - ib0, Pb = r1 - r0;
  if (!Pb) return NLP_OK;
  PairItems bi = plan.big_items;
  bi.u += ib0; bi.cnt += ib0; bi.dw += ib0; bi.ptr += ib0;
  k_pair_emit<FLT><<<grid_for(Eb, 256, h->num_sms * 16), 256, 0, h->stream>>>(
      p.g.keys, Eb, bi, plan.b_off + ib0, (unsigned long long)r0, (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, (uint32_t*)h->cs[0].p);
  NLP_LAUNCHED(h);
  int sb = 0;
  NLP_TRY(radix_sort_pairs(h, 0, Pb, FLT, &sb));
  // reuse store (count measures): the counts after the exclusion come along, in the idle payload array
  uint32_t* pc = (capture && !FLT) ? (uint32_t*)h->cs[sb].p : nullptr;
  k_pair_reduce<FLT><<<grid_for(Pb, 256, h->num_sms * 16), 256, 0, h->stream>>>(
      p, (const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p, (const uint32_t*)h->cs[sb].p, Pb, (uint32_t*)h->cs[sb ^ 1].p, pc);
  NLP_LAUNCHED(h);
  k_big_place<<<grid_for(Pb, 256, h->num_sms * 16), 256, 0, h->stream>>>(
      (const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p, (const uint32_t*)h->cs[sb ^ 1].p, pc, Pb, r0,
      plan.bg_first, plan.bg_roff, (uint32_t)kb0, (uint32_t)kb1, al_u, al_v, al_s, al_c);
  NLP_LAUNCHED(h);
  return NLP_OK;
}

// LHub bucket path.  *used stays false when the graph's rows are not symmetric or the plan / the
// aligned output do not fit the scratch budget; the caller then runs the source-centric kernels.
template <bool FLT>
int bucket_pass(nlp_handle* h, const nlp_options* opt, nlp_result* res, int* out_buf, uint64_t* out_fill, bool* used) {
  *used = false;
  NLP_TRY(check_symmetry(h));
  if (h->sym_state != 1) return NLP_OK;
  const uint32_t cap_rec = FLT ? BK_CAP_FLT : h->bucket_cap;
  const uint32_t half = cap_rec / 2u;
  const uint64_t key = ((uint64_t)opt->min_degree1 << 16) | half;
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    nlp_handle::BucketPlan& np = h->plans[key];
    const int rc = build_plan(h, opt->min_degree1, half, np);
    if (rc != NLP_OK) { h->plan_bytes -= std::min<uint64_t>(h->plan_bytes, np.arena.cap); retire_arena(h, np.arena); h->plans.erase(key); return rc; }
    it = h->plans.find(key);
  }
  nlp_handle::BucketPlan& plan = it->second;
  plan.stamp = ++h->cache_stamp;
  if (!plan.usable) return NLP_OK;
  const uint64_t P = plan.P;
  uint64_t budget = 0;
  NLP_TRY(scratch_budget(h, &budget));
  if ((P + 2 * (uint64_t)SORT_TILE) * 40 + h->plan_bytes > budget) return NLP_OK;
  const uint64_t padded = (P + OC_TILE) / OC_TILE * OC_TILE + 1024;
  if (!try_ensure(h, h->al_u, padded * 4) || !try_ensure(h, h->al_v, padded * 4) || !try_ensure(h, h->al_s, padded * 4) ||
      !try_ensure(h, h->al_c, padded * 4)) return NLP_OK;
  {
    bool ok = true;
    NLP_TRY(ensure_candidates(h, P, &ok));
    if (!ok) return NLP_OK;
  }
  Counters* hc = h->h_ctr;
  NLP_CUDA(h, cudaMemsetAsync(h->ctr.p, 0, sizeof(Counters), h->stream));
  NLP_CUDA(h, cudaMemsetAsync(h->thr.p, 0, sizeof(Threshold), h->stream));
  Params p;
  memset(&p, 0, sizeof p);
  p.g = dev_graph(h); p.D = opt->min_degree1; p.F2 = opt->max_factor2; p.measure = opt->measure; p.min_score = opt->min_score;
  p.gtable = (const double*)h->gtable.p;
  p.ctr = (Counters*)h->ctr.p; p.thr = (const Threshold*)h->thr.p;
  p.cap = h->cand_cap;
  h->phases_valid = false;
  uint32_t* al_u = (uint32_t*)h->al_u.p; uint32_t* al_v = (uint32_t*)h->al_v.p; uint32_t* al_s = (uint32_t*)h->al_s.p;
  uint32_t* al_c = (uint32_t*)h->al_c.p;
  const uint64_t rank = (uint64_t)h->rank, world = (uint64_t)h->world;
  // this rank's share: a contiguous range of buckets (equal parts of the prefix sum of the wedge
  // records per source) = an ascending range of sources, the big sources inside it, and one
  // contiguous range of slots
  const uint64_t nb = (plan.Ps + half - 1) / half;
  const uint64_t b0 = nb * rank / world, b1 = nb * (rank + 1) / world;
  uint64_t kb0 = 0, kb1 = plan.nbig, slo = 0, shi = P;
  if (world > 1) {
    const uint64_t pkey = (rank << 32) | world;
    if (plan.part_key != pkey) {
      NLP_TRY(ensure(h, h->gcnt, (size_t)(world + 8) * 8));
      unsigned long long out[4] = {0, 0, 0, 0};
      k_plan_partition<<<1, 1, 0, h->stream>>>(plan.dev, plan.big_items.u, plan.bg_item, plan.bg_roff, (uint32_t)plan.nbig, P,
                                               b0, b1, rank == 0, rank + 1 == world, (unsigned long long*)h->gcnt.p);
      NLP_LAUNCHED(h);
      NLP_CUDA(h, cudaMemcpyAsync(out, h->gcnt.p, 32, cudaMemcpyDeviceToHost, h->stream));
      NLP_CUDA(h, cudaStreamSynchronize(h->stream));
      plan.part_key = pkey; plan.part_kb0 = out[0]; plan.part_kb1 = out[1]; plan.part_slo = out[2]; plan.part_shi = out[3];
    }
    kb0 = plan.part_kb0; kb1 = plan.part_kb1; slo = plan.part_slo; shi = plan.part_shi;
  }
  // reuse store (nlp_set_reuse, count measures): the pairs of this threshold with their counts are
  // resident -> score them and go straight to the select
  const uint64_t skey = ((uint64_t)opt->min_degree1 << 24) | (rank << 12) | world;
  const bool capture = !FLT && h->reuse != 0;
  if (capture) {
    auto hit = h->count_store.find(skey);
    if (hit != h->count_store.end() && hit->second.n == shi - slo) {
      nlp_handle::CountStore& cs = hit->second;
      cs.stamp = ++h->cache_stamp;
      NLP_CUDA(h, cudaEventRecord(h->ev_frontier, h->stream));
      for (int i = 0; i < 3; ++i) NLP_CUDA(h, cudaEventRecord(h->ev_phase[i], h->stream));
      NLP_CUDA(h, cudaMemsetAsync(h->sel11.p, 0, sizeof(Select11), h->stream));
      if (cs.n) {
        k_score<false><<<grid_for(cs.n, 256, h->num_sms * 16), 256, 0, h->stream>>>(
            p, cs.u, cs.v, const_cast<uint32_t*>(cs.c), al_s, 0, cs.n, (Select11*)h->sel11.p, false, false);
        NLP_LAUNCHED(h);
      }
      h->sel11_l0 = true;
      for (int i = 3; i < 7; ++i) NLP_CUDA(h, cudaEventRecord(h->ev_phase[i], h->stream));
      h->phases_valid = true;
      NLP_TRY(read_counters(h));
      res->first_hop = plan.first_hop; res->eligible_first_hop = plan.elig; res->wedges = plan.wedges;
      res->candidates = hc->candidates; res->kept = hc->kept; res->emitted = hc->kept;
      res->passes = 1; res->path = NLP_PATH_PAIR; res->pair_records = P;
      res->bin_sources[0] = plan.ns; res->bin_sources[1] = plan.nbig; res->bin_sources[7] = 1;   // [7]: served from the reuse store
      h->pair_pending = true; h->pair_from_cache = true; h->pair_n = cs.n; h->pair_kept = hc->kept;
      h->pair_pu = cs.u; h->pair_pv = cs.v; h->pair_ps = al_s; h->pair_score_buf = 0;
      h->part_ordered = true;
      *out_buf = 0; *out_fill = hc->kept; *used = true;
      return NLP_OK;
    }
  }
  NLP_CUDA(h, cudaEventRecord(h->ev_frontier, h->stream));
  NLP_CUDA(h, cudaEventRecord(h->ev_phase[0], h->stream));
  // big sources: global sort of their records (pairs.cuh), then into their slots -- issued on a
  // second stream BEFORE k_bucket, so the two run side by side (disjoint buffers; k_score waits for both)
  const bool detour = kb1 > kb0;
  if (detour) {
    NLP_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    NLP_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    std::swap(h->stream, h->stream2);              // the helpers launch on h->stream
    int rc = big_detour<FLT>(h, plan, p, kb0, kb1, al_u, al_v, al_s, al_c, capture);
    if (rc == NLP_OK && cudaEventRecord(h->ev_join, h->stream) != cudaSuccess) rc = fail(h, NLP_ERR_CUDA, "cudaEventRecord(ev_join)");
    std::swap(h->stream, h->stream2);
    NLP_TRY(rc);
  }
  // small sources: one block per bucket
  if (b1 > b0) {
    const uint32_t top = h->S ? h->S - 1 : 0;
    int key_passes = 0;
    while (key_passes < 4 && (top >> (8 * key_passes)) != 0) ++key_passes;
    auto kern = cap_rec == 4096u ? k_bucket<FLT, 4096u> : k_bucket<FLT, 8192u>;
    const uint32_t smem = bucket_smem_bytes(FLT, cap_rec);
    NLP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (uint64_t s0 = b0; s0 < b1; s0 += 0x7fffffffull) {          // grid.x limit
      const unsigned grid = (unsigned)std::min<uint64_t>(b1 - s0, 0x7fffffffull);
      kern<<<grid, BK_THREADS, smem, h->stream>>>(p, plan.dev, s0, key_passes, al_u, al_v, al_c);
      NLP_LAUNCHED(h);
    }
  }
  NLP_CUDA(h, cudaEventRecord(h->ev_phase[1], h->stream));
  if (detour) NLP_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  NLP_CUDA(h, cudaEventRecord(h->ev_phase[2], h->stream));
  // exclusion + scoring of the small sources' pairs, one thread per slot; also the first
  // histogram of the top-K select (Select11)
  NLP_CUDA(h, cudaMemsetAsync(h->sel11.p, 0, sizeof(Select11), h->stream));
  if (shi > slo) {
    k_score<FLT><<<grid_for(shi - slo, 256, h->num_sms * 16), 256, 0, h->stream>>>(p, al_u, al_v, al_c, al_s, slo, shi, (Select11*)h->sel11.p, true, capture);
    NLP_LAUNCHED(h);
  }
  h->sel11_l0 = true;
  for (int i = 3; i < 7; ++i) NLP_CUDA(h, cudaEventRecord(h->ev_phase[i], h->stream));
  h->phases_valid = true;
  NLP_TRY(read_counters(h));
  if (hc->overflow) return fail(h, NLP_ERR_CAPACITY, "internal: bucket overflow (inconsistent plan)");
  if (capture && shi > slo) {
    // keep (u, v, count after the exclusion) of this rank's slots for the other count measures at
    // this threshold; least recently used thresholds go first, the arenas are pooled
    const uint64_t n = shi - slo, bytes = n * 12 + 1024;
    const uint64_t limit = budget / 4;
    while (h->plan_bytes + bytes > limit && !h->count_store.empty()) {
      auto lru = h->count_store.begin();
      for (auto i = h->count_store.begin(); i != h->count_store.end(); ++i) if (i->second.stamp < lru->second.stamp) lru = i;
      h->plan_bytes -= std::min<uint64_t>(h->plan_bytes, lru->second.arena.cap);
      retire_arena(h, lru->second.arena);
      h->count_store.erase(lru);
    }
    nlp_handle::CountStore cs;
    if (h->plan_bytes + bytes <= limit && take_arena(h, cs.arena, bytes)) {
      uint32_t* a = (uint32_t*)cs.arena.p;
      NLP_CUDA(h, cudaMemcpyAsync(a, al_u + slo, n * 4, cudaMemcpyDeviceToDevice, h->stream));
      NLP_CUDA(h, cudaMemcpyAsync(a + n, al_v + slo, n * 4, cudaMemcpyDeviceToDevice, h->stream));
      NLP_CUDA(h, cudaMemcpyAsync(a + 2 * n, al_c + slo, n * 4, cudaMemcpyDeviceToDevice, h->stream));
      cs.u = a; cs.v = a + n; cs.c = a + 2 * n; cs.n = n; cs.candidates = hc->candidates; cs.stamp = ++h->cache_stamp;
      h->plan_bytes += cs.arena.cap;
      auto old = h->count_store.find(skey);
      if (old != h->count_store.end()) { h->plan_bytes -= std::min<uint64_t>(h->plan_bytes, old->second.arena.cap); retire_arena(h, old->second.arena); h->count_store.erase(old); }
      h->count_store[skey] = cs;
    }
  }
  res->first_hop = plan.first_hop; res->eligible_first_hop = plan.elig; res->wedges = plan.wedges;
  res->candidates = hc->candidates; res->kept = hc->kept; res->emitted = hc->kept;
  res->passes = 1; res->path = NLP_PATH_PAIR; res->pair_records = P;
  res->bin_sources[0] = plan.ns; res->bin_sources[1] = plan.nbig;
  h->pair_pending = true; h->pair_from_cache = true; h->pair_n = shi - slo; h->pair_kept = hc->kept;
  h->pair_pu = al_u + slo; h->pair_pv = al_v + slo; h->pair_ps = al_s + slo; h->pair_score_buf = 0;
  h->part_ordered = true;
  *out_buf = 0; *out_fill = hc->kept; *used = true;
  return NLP_OK;
}

// LHub pair path (pairs.cuh).  *used stays false when the graph's rows are not symmetric or the
// wedge records do not fit the scratch budget; the caller then runs the source-centric kernels.
template <bool FLT>
int pair_pass(nlp_handle* h, const nlp_options* opt, nlp_result* res, int* out_buf, uint64_t* out_fill, bool* used) {
  *used = false;
  const uint32_t S = h->S;
  const DevGraph g = dev_graph(h);
  NLP_TRY(check_symmetry(h));
  if (h->sym_state != 1) return NLP_OK;
  Counters* hc = h->h_ctr;
  NLP_CUDA(h, cudaMemsetAsync(h->ctr.p, 0, sizeof(Counters), h->stream));
  NLP_CUDA(h, cudaMemsetAsync(h->thr.p, 0, sizeof(Threshold), h->stream));
  const uint64_t size_key = ((uint64_t)opt->min_degree1 << 32) | ((uint64_t)h->rank << 16) | (uint64_t)h->world;
  Params p;
  memset(&p, 0, sizeof p);
  p.g = g; p.D = opt->min_degree1; p.F2 = opt->max_factor2; p.measure = opt->measure; p.min_score = opt->min_score;
  p.gtable = (const double*)h->gtable.p;
  p.ctr = (Counters*)h->ctr.p; p.thr = (const Threshold*)h->thr.p;
  h->phases_valid = false;

  // ---- sorted records of this threshold already resident (nlp_set_reuse)? -----------------------
  if (h->reuse) {
    auto hit = h->pair_cache.find(size_key);
    if (hit != h->pair_cache.end()) {
      nlp_handle::PairCache& c = hit->second;
      c.stamp = ++h->cache_stamp;
      const uint64_t P = c.P;
      NLP_TRY(ensure_candidates(h, P));
      for (int i = 0; i < 3; ++i) NLP_CUDA(h, cudaEventRecord(i == 0 ? h->ev_frontier : h->ev_phase[i - 1], h->stream));
      NLP_CUDA(h, cudaEventRecord(h->ev_phase[2], h->stream));
      if (P) {
        p.cap = h->cand_cap;
        k_pair_reduce<FLT><<<grid_for(P, 256, h->num_sms * 16), 256, 0, h->stream>>>(
            p, (const uint32_t*)c.u.p, (const uint32_t*)c.v.p, (const uint32_t*)c.w.p, P, (uint32_t*)h->cs[0].p);
        NLP_LAUNCHED(h);
      }
      for (int i = 3; i < 7; ++i) NLP_CUDA(h, cudaEventRecord(h->ev_phase[i], h->stream));
      h->phases_valid = true;
      NLP_TRY(read_counters(h));
      res->first_hop = c.first_hop; res->eligible_first_hop = c.elig; res->wedges = c.wedges;
      res->candidates = hc->candidates; res->kept = hc->kept; res->emitted = hc->kept;
      res->passes = 1; res->path = NLP_PATH_PAIR_SORT; res->pair_records = P;
      h->pair_pending = true; h->pair_from_cache = true; h->pair_n = P; h->pair_kept = hc->kept; h->pair_ps = nullptr;
      h->part_ordered = false;
      h->pair_pu = (const uint32_t*)c.u.p; h->pair_pv = (const uint32_t*)c.v.p; h->pair_score_buf = 0;
      *out_buf = 1; *out_fill = hc->kept; *used = true;
      return NLP_OK;
    }
  }

  k_pair_rows<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(g, opt->min_degree1, h->rank, h->world,
                                                                        (uint32_t*)h->work.p, (Counters*)h->ctr.p);
  NLP_LAUNCHED(h);
  uint64_t E = 0, P = 0;
  const auto known = h->pair_sizes.find(size_key);
  const bool cached = known != h->pair_sizes.end();
  if (cached) { E = known->second.first; P = known->second.second; }
  NLP_TRY(exclusive_scan<uint32_t>(h, (const uint32_t*)h->work.p, S, (unsigned long long*)h->work64.p, cached ? nullptr : &E));
  PairItems it{nullptr, nullptr, nullptr, nullptr};
  uint64_t budget = 0;
  NLP_TRY(scratch_budget(h, &budget));
  if (E) {
    if (E * 28 > budget / 2) return NLP_OK;
    // out of memory here is not an error: the source-centric kernels take over (*used stays false)
    if (!try_ensure(h, h->it_u, E * 4) || !try_ensure(h, h->it_cnt, E * 4) || !try_ensure(h, h->it_dw, E * 4) ||
        !try_ensure(h, h->it_ptr, E * 8) || !try_ensure(h, h->it_off, E * 8)) return NLP_OK;
    it.u = (uint32_t*)h->it_u.p; it.cnt = (uint32_t*)h->it_cnt.p; it.dw = (uint32_t*)h->it_dw.p;
    it.ptr = (unsigned long long*)h->it_ptr.p;
    k_pair_items<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(
        g, (const uint32_t*)h->work.p, (const unsigned long long*)h->work64.p, h->rank, h->world, it, (Counters*)h->ctr.p);
    NLP_LAUNCHED(h);
    NLP_TRY(exclusive_scan<uint32_t>(h, (const uint32_t*)h->it_cnt.p, E, (unsigned long long*)h->it_off.p, cached ? nullptr : &P));
  }
  if (!cached) h->pair_sizes[size_key] = std::make_pair(E, P);
  NLP_CUDA(h, cudaEventRecord(h->ev_frontier, h->stream));
  NLP_CUDA(h, cudaEventRecord(h->ev_phase[0], h->stream));
  NLP_TRY(scratch_budget(h, &budget));
  // items (28 B each), the six candidate arrays, the sort's tile counts, the reuse store: one budget
  const uint64_t pair_need = E * 28 + (P + 2 * (uint64_t)SORT_TILE) * 24 + P / 4 + (h->reuse ? h->cache_bytes + P * 12 : 0);
  if (P >= 0xfffffff0ull || pair_need > budget) return NLP_OK;
  {
    bool ok = true;
    NLP_TRY(ensure_candidates(h, P, &ok));
    if (!ok) return NLP_OK;
  }
  p.cap = h->cand_cap;
  // with reuse on, the records always carry deg(w) so that they serve the float measures too
  const bool payload = FLT || h->reuse;
  int cur = 0, sb = 0;
  if (P) {
    if (payload)
      k_pair_emit<true><<<grid_for(E, 256, h->num_sms * 16), 256, 0, h->stream>>>(
          g.keys, E, it, (const unsigned long long*)h->it_off.p, 0ull, (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, (uint32_t*)h->cs[0].p);
    else
      k_pair_emit<false><<<grid_for(E, 256, h->num_sms * 16), 256, 0, h->stream>>>(
          g.keys, E, it, (const unsigned long long*)h->it_off.p, 0ull, (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, (uint32_t*)h->cs[0].p);
    NLP_LAUNCHED(h);
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[1], h->stream));
    NLP_TRY(radix_sort_pairs(h, 0, P, payload, &sb));
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[2], h->stream));
    cur = sb ^ 1;
    k_pair_reduce<FLT><<<grid_for(P, 256, h->num_sms * 16), 256, 0, h->stream>>>(
        p, (const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p, (const uint32_t*)h->cs[sb].p, P, (uint32_t*)h->cs[cur].p);
    NLP_LAUNCHED(h);
    for (int i = 3; i < 7; ++i) NLP_CUDA(h, cudaEventRecord(h->ev_phase[i], h->stream));
    h->phases_valid = true;
  }
  NLP_TRY(read_counters(h));
  if (hc->overflow) return fail(h, NLP_ERR_CAPACITY, "internal: candidate buffer overflow (pair path)");
  res->first_hop = hc->first_hop;
  res->eligible_first_hop = hc->eligible_first_hop;
  res->wedges = hc->wedges;
  res->candidates = hc->candidates;
  res->kept = hc->kept;
  res->emitted = hc->kept;
  res->passes = 1;
  res->path = NLP_PATH_PAIR_SORT;
  res->pair_records = P;
  h->pair_pending = true; h->pair_from_cache = false; h->pair_n = P; h->pair_kept = hc->kept; h->pair_ps = nullptr;
  h->part_ordered = false;
  h->pair_pu = (const uint32_t*)h->cu[sb].p; h->pair_pv = (const uint32_t*)h->cv[sb].p; h->pair_score_buf = cur;
  if (h->reuse && P) {
    // keep a copy of the sorted records (after the reduce kernel is queued: same stream, so the
    // copy simply follows it); evict least-recently-used thresholds beyond a quarter of the budget
    const uint64_t bytes = P * 12;
    const uint64_t limit = budget / 4;
    while (h->cache_bytes + bytes > limit && !h->pair_cache.empty()) {
      auto lru = h->pair_cache.begin();
      for (auto i = h->pair_cache.begin(); i != h->pair_cache.end(); ++i) if (i->second.stamp < lru->second.stamp) lru = i;
      h->cache_bytes -= lru->second.P * 12;
      release(lru->second.u); release(lru->second.v); release(lru->second.w);
      h->pair_cache.erase(lru);
    }
    if (h->cache_bytes + bytes <= limit) {
      nlp_handle::PairCache& c = h->pair_cache[size_key];
      NLP_TRY(ensure(h, c.u, P * 4)); NLP_TRY(ensure(h, c.v, P * 4)); NLP_TRY(ensure(h, c.w, P * 4));
      NLP_CUDA(h, cudaMemcpyAsync(c.u.p, h->cu[sb].p, P * 4, cudaMemcpyDeviceToDevice, h->stream));
      NLP_CUDA(h, cudaMemcpyAsync(c.v.p, h->cv[sb].p, P * 4, cudaMemcpyDeviceToDevice, h->stream));
      NLP_CUDA(h, cudaMemcpyAsync(c.w.p, h->cs[sb].p, P * 4, cudaMemcpyDeviceToDevice, h->stream));
      c.P = P; c.first_hop = hc->first_hop; c.elig = hc->eligible_first_hop; c.wedges = hc->wedges; c.stamp = ++h->cache_stamp;
      h->cache_bytes += bytes;
    }
  }
  *out_buf = cur;
  *out_fill = hc->kept;
  *used = true;
  return NLP_OK;
}

// Sources of k_range with deg(u) below the returned limit may count in half words: a count is at
// most deg(u) x (largest multiplicity of an entry in a row) and must stay below 2^15.  The
// multiplicity is measured by the validation pass when the graph is set (k_validate_entries).
int half_word_limit(nlp_handle* h, bool range_on, uint32_t* limit) {
  *limit = 0;
  if (!range_on || !h->range_half) return NLP_OK;
  *limit = 32768u / (h->maxmult > 1u ? h->maxmult : 1u);   // the multiplicity comes from the validation pass of nlp_set_graph
  return NLP_OK;
}

template <bool FLT>
int scoring_pass(nlp_handle* h, const nlp_options* opt, nlp_result* res, int* out_buf, uint64_t* out_fill) {
  const uint32_t S = h->S;
  const uint64_t K = opt->max_edges;
  const bool lhub = opt->min_degree1 != 0;
  Counters* hc = h->h_ctr;

  if (lhub && h->path_mode != NLP_PATH_SOURCE) {
    bool used = false;
    if (h->path_mode == NLP_PATH_PAIR_SORT) NLP_TRY(pair_pass<FLT>(h, opt, res, out_buf, out_fill, &used));
    else NLP_TRY(bucket_pass<FLT>(h, opt, res, out_buf, out_fill, &used));
    if (used) return NLP_OK;
  }
  res->path = NLP_PATH_SOURCE;
  h->pair_pending = false;
  h->part_ordered = false;
  NLP_CUDA(h, cudaMemsetAsync(h->ctr.p, 0, sizeof(Counters), h->stream));
  NLP_CUDA(h, cudaMemsetAsync(h->thr.p, 0, sizeof(Threshold), h->stream));
  const DevGraph g = dev_graph(h);
  // (a) frontier
  WorkOut wo;
  wo.work64 = (unsigned long long*)h->work64.p;
  wo.ecount = nullptr; wo.ekeys = nullptr;
  wo.chunk_cnt = (uint32_t*)h->chunk_cnt.p;
  const uint64_t nrb = ((uint64_t)S + 31) / 32;
  const unsigned gshort = grid_for(nrb, 8, h->num_sms * 16);
  const unsigned glong = (unsigned)std::min<uint64_t>(h->nchunks, (uint64_t)h->num_sms * 16);
  if (lhub) {
    NLP_TRY(ensure(h, h->ecount, (size_t)S * 4));
    NLP_TRY(ensure(h, h->ekeys, (size_t)h->M * 4));
    wo.ecount = (uint32_t*)h->ecount.p; wo.ekeys = (uint32_t*)h->ekeys.p;
    k_elig<<<grid_for(nrb, 256, h->num_sms * 8), 256, 0, h->stream>>>(
        (const uint32_t*)h->deg.p, S, opt->min_degree1, (uint32_t*)h->elig.p);
    NLP_LAUNCHED(h);
    k_work_short<true><<<gshort, 256, 0, h->stream>>>(g, (const uint32_t*)h->elig.p, h->rank, h->world, wo, (Counters*)h->ctr.p);
    NLP_LAUNCHED(h);
    if (glong) {
      k_work_long<true><<<glong, 256, 0, h->stream>>>(g, (const uint32_t*)h->elig.p, h->rank, h->world,
                                                       (const uint32_t*)h->chunk_src.p, (const unsigned long long*)h->chunk_base.p,
                                                       (uint32_t)h->nchunks, wo, (Counters*)h->ctr.p);
      NLP_LAUNCHED(h);
    }
  } else {
    k_work_short<false><<<gshort, 256, 0, h->stream>>>(g, nullptr, h->rank, h->world, wo, (Counters*)h->ctr.p);
    NLP_LAUNCHED(h);
    if (glong) {
      k_work_long<false><<<glong, 256, 0, h->stream>>>(g, nullptr, h->rank, h->world,
                                                        (const uint32_t*)h->chunk_src.p, (const unsigned long long*)h->chunk_base.p,
                                                        (uint32_t)h->nchunks, wo, (Counters*)h->ctr.p);
      NLP_LAUNCHED(h);
    }
  }
  BinLists bl;
  for (int b = 0; b < NBINS; ++b) bl.list[b] = (uint32_t*)h->list[b].p;
  // count measures may send hub-heavy sources to the windowed shared-memory counters (k_range);
  // the float measures need the ordered single-warp accumulation of k_dense
  // the float measures walk windows too, one WARP per source (k_range_flt: rows one after the other
  // keep the reference's accumulation order) -- on graphs without repeated entries in a row, and
  // while the per-warp row records fit
  uint32_t range_c = (!FLT && h->maxdeg < (1u << 22) && h->range_mode != 0) ? RANGE_COUNTERS : 0u;
  uint32_t range_fixed = 256, range_div = h->range_div;
  if (FLT && h->range_mode != 0 && h->flt_range_mode != 0 && h->maxmult <= 1u && h->maxdeg < (1u << 22)) {
    const uint64_t stride = ((uint64_t)h->maxdeg + CHUNK + 31) / 32 * 32;
    // (against the memory that was free when the graph was set, not against nlp_set_scratch_limit:
    // that limit is about the candidate buffer and the spill tables)
    if ((uint64_t)h->num_sms * RFLT_WARPS * stride * 24 <= h->budget_base / 4) { range_c = RFLT_WIN; range_fixed = 32; range_div = 4; }
  }
  uint32_t half_deg = 0;
  NLP_TRY(half_word_limit(h, range_c != 0u && !FLT, &half_deg));
  const uint32_t quarter_deg = (half_deg && h->range_quarter) ? 128u / (h->maxmult > 1u ? h->maxmult : 1u) : 0u;
  k_bin<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(g, (const unsigned long long*)h->work64.p, h->rank, h->world,
                                                                  range_c, range_fixed, range_div, half_deg, quarter_deg, (uint32_t*)h->work.p, bl, (Counters*)h->ctr.p);
  NLP_LAUNCHED(h);
  NLP_CUDA(h, cudaEventRecord(h->ev_frontier, h->stream));
  NLP_TRY(read_counters(h));

  uint64_t nb[NBINS], total_need = 0;
  for (int b = 0; b < NBINS; ++b) { nb[b] = hc->bin_count[b]; total_need += hc->bin_bound[b]; }
  res->first_hop = hc->first_hop;
  res->eligible_first_hop = hc->eligible_first_hop;
  res->wedges = hc->wedges;
  res->frontier_sources = hc->frontier;
  for (int b = 0; b < 8; ++b) res->bin_sources[b] = b < NBINS ? nb[b] : 0;

  // scratch plan: dense spill tables first, candidate buffer with what is left
  uint64_t budget = 0;
  NLP_TRY(scratch_budget(h, &budget));
  unsigned dense_slots = 0;
  uint64_t touched_cap = 0;
  if (nb[5]) {
    touched_cap = std::min<uint64_t>(hc->max_bound, S);
    const uint64_t per_slot = (uint64_t)S * 4 + touched_cap * 4;
    uint64_t want = FLT ? (uint64_t)h->num_sms * 16 : (uint64_t)h->num_sms * 2;
    want = std::min<uint64_t>(want, nb[5]);
    const uint64_t afford = std::max<uint64_t>(1, (budget / 3) / per_slot);
    dense_slots = (unsigned)std::min<uint64_t>(want, afford);
    NLP_TRY(ensure(h, h->tables, (uint64_t)dense_slots * S * 4, true));   // zeroed when (re)allocated
    NLP_TRY(ensure(h, h->touched, (uint64_t)dense_slots * touched_cap * 4));
    budget -= std::min<uint64_t>(budget, (uint64_t)dense_slots * per_slot);
  }
  const uint64_t cap_limit = std::min<uint64_t>(budget / 24, 0xfffffff0ull);
  uint64_t want_cap = total_need;
  if (K != NLP_UNBOUNDED) {
    // room for every candidate when that is cheap; otherwise 16 K, at least 2^27 entries (3 GB)
    // and up to a quarter of the scratch budget: a source reserves its BOUND while it runs, so
    // the buffer also limits how many hub-heavy sources can be in flight at once
    const uint64_t k16 = K > (1ull << 58) ? (1ull << 62) : 16 * K;
    want_cap = std::min<uint64_t>(total_need, std::max<uint64_t>(std::max<uint64_t>(k16, 1ull << 27), cap_limit / 4) + S + (1ull << 20));
  }
  const uint64_t cap = std::min<uint64_t>(want_cap, cap_limit);
  const bool admit = cap < total_need;
  if (admit) {
    const uint64_t kk = std::min<uint64_t>(K, total_need);
    if (K == NLP_UNBOUNDED || cap < kk + S + 4096)
      return fail(h, NLP_ERR_CAPACITY, "candidate buffer too small for this request (raise nlp_set_scratch_limit or bound max_edges)");
  }
  NLP_TRY(ensure_candidates(h, cap));

  // float measures on warp windows: the bin-6 sources are cut into (source, window range) items,
  // and the bin's work list becomes the list of item ids (also the unit of deferral)
  const uint32_t* list6 = (const uint32_t*)h->list[6].p;
  uint32_t* defer6 = (uint32_t*)h->defer[6].p;
  uint64_t n6 = nb[6];
  if (FLT && nb[6]) {
    NLP_TRY(ensure(h, h->flt_cnt, nb[6] * 4));
    NLP_TRY(ensure(h, h->flt_off, nb[6] * 8));
    k_flt_item_counts<<<grid_for(nb[6], 256, h->num_sms * 8), 256, 0, h->stream>>>(g, list6, (uint32_t)nb[6], (uint32_t*)h->flt_cnt.p);
    NLP_LAUNCHED(h);
    uint64_t nitems = 0;
    NLP_TRY(exclusive_scan<uint32_t>(h, (const uint32_t*)h->flt_cnt.p, nb[6], (unsigned long long*)h->flt_off.p, &nitems));
    if (nitems >= 0xfffffff0ull) return fail(h, NLP_ERR_CAPACITY, "too many window items");
    NLP_TRY(ensure(h, h->flt_items, nitems * 12 + 64));
    NLP_TRY(ensure(h, h->flt_ids, nitems * 4));
    NLP_TRY(ensure(h, h->flt_defer, nitems * 4));
    h->flt_it.u = (uint32_t*)h->flt_items.p; h->flt_it.w0 = h->flt_it.u + nitems; h->flt_it.w1 = h->flt_it.w0 + nitems;
    k_flt_item_fill<<<grid_for(nb[6], 256, h->num_sms * 8), 256, 0, h->stream>>>(g, list6, (uint32_t)nb[6],
                                                                               (const unsigned long long*)h->flt_off.p, h->flt_it, (uint32_t*)h->flt_ids.p);
    NLP_LAUNCHED(h);
    h->flt_n = nitems;
    list6 = (const uint32_t*)h->flt_ids.p; defer6 = (uint32_t*)h->flt_defer.p; n6 = nitems;
  }

  int cur = 0;
  Params p;
  p.g = g; p.D = opt->min_degree1; p.F2 = opt->max_factor2; p.measure = opt->measure; p.min_score = opt->min_score;
  p.coop = (h->maxdeg < (1u << 22) && h->coop_mode != 0) ? 1u : 0u;
  p.range_half = half_deg;
  p.range_quarter = quarter_deg;
  p.elig = lhub ? (const uint32_t*)h->elig.p : nullptr;
  p.ekeys = lhub ? (const uint32_t*)h->ekeys.p : nullptr;
  p.ecount = lhub ? (const uint32_t*)h->ecount.p : nullptr;
  p.chunk_base = (const unsigned long long*)h->chunk_base.p;
  p.chunk_cnt = (const uint32_t*)h->chunk_cnt.p;
  p.gtable = (const double*)h->gtable.p;
  p.work = (const uint32_t*)h->work.p;
  p.cap = cap; p.ctr = (Counters*)h->ctr.p; p.thr = (const Threshold*)h->thr.p;
  p.soft_cap = cap;
  const uint64_t pass_quota = std::max<uint64_t>(K > (1ull << 58) ? cap : 16 * K, 1ull << 22);   // writes per pass before a cut
  auto bind = [&](int b) { p.cu = (uint32_t*)h->cu[b].p; p.cv = (uint32_t*)h->cv[b].p; p.cs = (float*)h->cs[b].p; };
  bind(cur);

  res->passes = 1;
  h->phases_valid = false;
  for (int i = 0; i < 8; ++i) h->phase_acc[i] = 0.f;
  if (!admit) {
    // everything fits: one pass, no admission control, no host round trip until the end
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[0], h->stream));
    NLP_TRY((launch_range<FLT, false>(h, p, list6, (uint32_t)n6, nullptr)));
    NLP_TRY((launch_dense<FLT, false>(h, p, (const uint32_t*)h->list[5].p, (uint32_t)nb[5], nullptr, dense_slots, touched_cap)));
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[1], h->stream));
    for (int b = 4; b >= 2; --b) {
      NLP_TRY((launch_hash<FLT, false>(h, p, b, (const uint32_t*)h->list[b].p, (uint32_t)nb[b], nullptr)));
      NLP_CUDA(h, cudaEventRecord(h->ev_phase[6 - b], h->stream));
    }
    NLP_TRY((launch_tiny<FLT>(h, p, 1, (const uint32_t*)h->list[1].p, (uint32_t)nb[1])));
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[5], h->stream));
    NLP_TRY((launch_tiny<FLT>(h, p, 0, (const uint32_t*)h->list[0].p, (uint32_t)nb[0])));
    NLP_CUDA(h, cudaEventRecord(h->ev_phase[6], h->stream));
    NLP_TRY(read_counters(h));
    h->phases_valid = true;
  } else {
    // the buffer cannot hold every candidate: admit sources while there is room, then keep the
    // best K (which fixes the pruning threshold) and continue with the deferred sources
    uint32_t* lists[NBINS]; uint32_t* defers[NBINS];
    for (int b = 0; b < NBINS; ++b) { lists[b] = (uint32_t*)h->list[b].p; defers[b] = (uint32_t*)h->defer[b].p; }
    lists[6] = const_cast<uint32_t*>(list6); defers[6] = defer6;
    uint64_t remaining[NBINS];
    for (int b = 0; b < NBINS; ++b) remaining[b] = nb[b];
    remaining[6] = n6;
    uint64_t tiny_pos[2] = {0, 0};
    uint64_t fill = 0;
    for (;; res->passes++) {
      // reset queues / deferred counts, seed the reservation with the current fill
      NLP_CUDA(h, cudaMemsetAsync((char*)h->ctr.p + offsetof(Counters, deferred), 0, 16 * 8, h->stream));
      NLP_CUDA(h, cudaMemcpyAsync((char*)h->ctr.p + offsetof(Counters, reserved), &fill, 8, cudaMemcpyHostToDevice, h->stream));
      p.soft_cap = std::min<uint64_t>(cap, fill + pass_quota);
      NLP_CUDA(h, cudaEventRecord(h->ev_phase[0], h->stream));
      NLP_TRY((launch_range<FLT, true>(h, p, lists[6], (uint32_t)remaining[6], defers[6])));
      NLP_TRY((launch_dense<FLT, true>(h, p, lists[5], (uint32_t)remaining[5], defers[5], dense_slots, touched_cap)));
      NLP_CUDA(h, cudaEventRecord(h->ev_phase[1], h->stream));
      for (int b = 4; b >= 2; --b) {
        NLP_TRY((launch_hash<FLT, true>(h, p, b, lists[b], (uint32_t)remaining[b], defers[b])));
        NLP_CUDA(h, cudaEventRecord(h->ev_phase[6 - b], h->stream));
      }
      NLP_TRY(read_counters(h));
      fill = hc->cursor;
      uint64_t fill_ub = fill;
      for (int b = 1; b >= 0; --b) {
        const uint64_t G = b ? 32 : 8;
        const uint64_t left = nb[b] - tiny_pos[b];
        const uint64_t take = std::min<uint64_t>(left, (cap - fill_ub) / G);
        if (take) {
          NLP_TRY((launch_tiny<FLT>(h, p, b, (const uint32_t*)h->list[b].p + tiny_pos[b], (uint32_t)take)));
          tiny_pos[b] += take; fill_ub += take * G;
        }
        NLP_CUDA(h, cudaEventRecord(h->ev_phase[b ? 5 : 6], h->stream));
      }
      NLP_TRY(read_counters(h));
      for (int i = 1; i <= 6; ++i) {   // per-phase device time, summed over the passes
        float ms = 0.f;
        NLP_CUDA(h, cudaEventElapsedTime(&ms, h->ev_phase[i - 1], h->ev_phase[i]));
        h->phase_acc[i] += ms;
      }
      fill = hc->cursor;
      bool done = tiny_pos[0] == nb[0] && tiny_pos[1] == nb[1];
      for (int b = 2; b < NBINS; ++b) { remaining[b] = hc->deferred[b]; done = done && remaining[b] == 0; std::swap(lists[b], defers[b]); }
      if (done) break;
      if (fill > K) {
        int ob; uint64_t on;
        NLP_TRY(top_k(h, cur, fill, K, &ob, &on));
        cur = ob; bind(cur); fill = on;
        // threshold = the K-th best so far
        Threshold t;
        uint32_t tu, tv, ts;
        NLP_CUDA(h, cudaMemcpyAsync(&tu, (uint32_t*)h->cu[cur].p + (on - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        NLP_CUDA(h, cudaMemcpyAsync(&tv, (uint32_t*)h->cv[cur].p + (on - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        NLP_CUDA(h, cudaMemcpyAsync(&ts, (uint32_t*)h->cs[cur].p + (on - 1), 4, cudaMemcpyDeviceToHost, h->stream));
        NLP_CUDA(h, cudaStreamSynchronize(h->stream));
        t.active = 1; t.key = desc_key(ts); t.u = tu; t.v = tv;
        NLP_CUDA(h, cudaMemcpyAsync(h->thr.p, &t, sizeof t, cudaMemcpyHostToDevice, h->stream));
        NLP_CUDA(h, cudaMemcpyAsync((char*)h->ctr.p + offsetof(Counters, cursor), &fill, 8, cudaMemcpyHostToDevice, h->stream));
        NLP_CUDA(h, cudaStreamSynchronize(h->stream));
      }
      if (res->passes > 100000) return fail(h, NLP_ERR_CAPACITY, "candidate buffer passes did not converge");
    }
  }
  if (hc->overflow) return fail(h, NLP_ERR_CAPACITY, "internal: candidate buffer overflow");
  res->candidates = hc->candidates;
  res->kept = hc->kept;
  res->emitted = hc->cursor;
  *out_buf = cur;
  *out_fill = hc->cursor;
  return NLP_OK;
}

// ---- multi-GPU: NCCL communicator behind the C ABI -----------------------
// libnccl is dlopen'ed on first use, so the library loads (and every single-GPU entry works)
// on a box without NCCL; <nccl.h> is only needed for the types.
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return nullptr; }
  auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) api.err = std::string("dlsym ") + n; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.GetErrorString) {
    dlclose(api.lib); api.lib = nullptr; return nullptr;
  }
  return &api;
}

#define NLP_NCCL(h, expr)                                                                      \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      (h)->err = std::string(#expr) + " (nlp_b200.cu:" + std::to_string(__LINE__) + "): " + nccl_api()->GetErrorString(r__); \
      return NLP_ERR_COMM;                                                                     \
    }                                                                                          \
  } while (0)

inline bool distributed(const nlp_handle* h) { return h->comm != nullptr; }   // also with one rank: same code path

// MSD radix select with the digit histograms summed over the ranks: every rank ends with the
// same prefix / above / bucket, i.e. the GLOBAL cutoff ("all-reduce of a
// score histogram to agree the cutoff").  Levels are queued four at a time; `done` is read back.
int select_narrow(nlp_handle* h, const uint32_t* cu, const uint32_t* cv, const uint32_t* cs, uint64_t n,
                  uint64_t K, uint32_t max_bits, bool dist) {
  NLP_CUDA(h, cudaMemsetAsync(h->sel.p, 0, sizeof(SelectState), h->stream));
  const uint64_t slack = std::max<uint64_t>(K / 8, 65536);
  NcclApi* api = dist ? nccl_api() : nullptr;
  bool done = false;
  while (!done) {
    for (int lvl = 0; lvl < 4; ++lvl) {
      k_select_hist<<<grid_for(n, 256 * 8, h->num_sms * 8), 256, 0, h->stream>>>(cu, cv, cs, n, (SelectState*)h->sel.p);
      NLP_LAUNCHED(h);
      if (api) {
        unsigned long long* hist = (unsigned long long*)((char*)h->sel.p + offsetof(SelectState, hist));
        NLP_NCCL(h, api->AllReduce(hist, hist, 256, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->stream));
      }
      k_select_step<<<1, 1, 0, h->stream>>>((SelectState*)h->sel.p, (unsigned long long)K, (unsigned long long)slack, max_bits);
      NLP_LAUNCHED(h);
    }
    NLP_CUDA(h, cudaMemcpyAsync(h->h_sel, h->sel.p, offsetof(SelectState, hist), cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));
    done = h->h_sel->done != 0;
  }
  return NLP_OK;
}

// Sum of one count per rank (one tiny all-gather + read-back).
int global_count(nlp_handle* h, uint64_t mine, uint64_t* sum) {
  NcclApi* api = nccl_api();
  const int W = h->world;
  NLP_TRY(ensure(h, h->gcnt, (size_t)(W + 1) * 8));
  unsigned long long* d_cnt = (unsigned long long*)h->gcnt.p;
  unsigned long long m = mine;
  NLP_CUDA(h, cudaMemcpyAsync(d_cnt + W, &m, 8, cudaMemcpyHostToDevice, h->stream));
  NLP_NCCL(h, api->AllGather(d_cnt + W, d_cnt, 1, ncclUint64, (ncclComm_t)h->comm, h->stream));
  std::vector<unsigned long long> cnt(W);
  NLP_CUDA(h, cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)W * 8, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  *sum = 0;
  for (int r = 0; r < W; ++r) *sum += cnt[r];
  return NLP_OK;
}

// All-gather the m local survivors in candidate buffer `buf` (counts first, then ONE padded
// payload all-gather); afterwards buffer 0 holds the candidates of all ranks in rank order.
int gather_candidates(nlp_handle* h, int buf, uint64_t m, uint64_t* total, const std::vector<unsigned long long>* known = nullptr) {
  NcclApi* api = nccl_api();
  const int W = h->world;
  if (W > 16) return fail(h, NLP_ERR_ARG, "multi-GPU merge: more than 16 ranks");
  std::vector<unsigned long long> cnt(W);
  if (known) {
    cnt = *known;                                  // the caller already exchanged what the counts follow from
  } else {
    NLP_TRY(ensure(h, h->gcnt, (size_t)(W + 1) * 8));
    unsigned long long* d_cnt = (unsigned long long*)h->gcnt.p;
    unsigned long long mine = m;
    NLP_CUDA(h, cudaMemcpyAsync(d_cnt + W, &mine, 8, cudaMemcpyHostToDevice, h->stream));
    NLP_NCCL(h, api->AllGather(d_cnt + W, d_cnt, 1, ncclUint64, (ncclComm_t)h->comm, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)W * 8, cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  uint64_t width = 1, sum = 0;
  GatherOffsets go;
  go.world = W;
  for (int r = 0; r < W; ++r) { go.off[r] = sum; width = std::max<uint64_t>(width, cnt[r]); sum += cnt[r]; }
  go.off[W] = sum;
  if (sum >= 0xfffffff0ull) return fail(h, NLP_ERR_CAPACITY, "multi-GPU merge: too many candidates");
  NLP_TRY(ensure(h, h->gsend, (size_t)width * 12));
  NLP_TRY(ensure(h, h->grecv, (size_t)width * 12 * W));
  uint32_t* send = (uint32_t*)h->gsend.p;
  if (m) {
    NLP_CUDA(h, cudaMemcpyAsync(send, h->cu[buf].p, m * 4, cudaMemcpyDeviceToDevice, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(send + width, h->cv[buf].p, m * 4, cudaMemcpyDeviceToDevice, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(send + 2 * width, h->cs[buf].p, m * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  NLP_NCCL(h, api->AllGather(send, h->grecv.p, (size_t)width * 3, ncclUint32, (ncclComm_t)h->comm, h->stream));
  NLP_TRY(ensure_candidates(h, sum));           // may reallocate the candidate buffers: the local survivors are in `send`
  if (sum) {
    k_gather_unpack<<<grid_for(sum, 256, h->num_sms * 16), 256, 0, h->stream>>>(
        (const uint32_t*)h->grecv.p, (unsigned long long)width, go, (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, (uint32_t*)h->cs[0].p);
    NLP_LAUNCHED(h);
  }
  *total = sum;
  h->gathered_bytes += (uint64_t)width * 12 * W;
  return NLP_OK;
}

// Stable sort of buffer `buf` by the score digits that are not constant (known from digit
// histograms the caller already has: `constant` bit d set = digit d of desc_key(score) is the same
// for every entry).
int radix_sort_score(nlp_handle* h, int buf, uint64_t n, unsigned constant, int* out_buf) {
  *out_buf = buf;
  if (n < 2) return NLP_OK;
  const uint32_t nblocks = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
  NLP_TRY(ensure(h, h->counts, (size_t)nblocks * 256 * 4));
  for (int d = 0; d < 4; ++d) {
    if ((constant >> d) & 1u) continue;
    NLP_TRY(radix_pass(h, buf, n, nblocks, 8 + d, true));
  }
  *out_buf = buf;
  return NLP_OK;
}

// Top-K of the w-centric LHub paths, one GPU or several.  The kept pairs lie in ascending (u, v)
// order in record-aligned arrays (pair_pu / pair_pv / score bits): exact radix select on the score
// (Select11; with a communicator the histograms are all-reduced, so every rank finds the GLOBAL
// cutoff), in-order compaction of exactly the survivors (better than the K-th score, plus the first
// `need` of its tie class in (u, v) order), [one all-gather,] stable sort by score.  Replaces the
// per-thread heaps and the serial T-way merge of inc/predict.hxx:313-336, 431-460.
int ordered_top_k(nlp_handle* h, uint64_t K, int* out_buf, uint64_t* out_n) {
  const uint32_t* pu = h->pair_pu;
  const uint32_t* pv = h->pair_pv;
  const uint64_t n = h->pair_n;
  const int sbuf = h->pair_score_buf, ob = sbuf ^ 1;
  const bool external = h->pair_ps != nullptr;       // bucket path: records and scores in the aligned arrays
  const uint32_t* sbits = external ? h->pair_ps : (const uint32_t*)h->cs[sbuf].p;
  const bool dist = distributed(h);
  NcclApi* api = dist ? nccl_api() : nullptr;
  Select11* st = (Select11*)h->sel11.p;
  const bool l0 = h->sel11_l0;
  h->sel11_l0 = false;
  *out_buf = external ? 0 : ob; *out_n = 0;
  if (!dist && (!n || !h->pair_kept)) return NLP_OK;
  const unsigned gsel = grid_for(n, 256 * 8, h->num_sms * 8);
  if (!l0) NLP_CUDA(h, cudaMemsetAsync(st, 0, sizeof(Select11), h->stream));
  for (int lvl = 0; lvl < 3; ++lvl) {
    if (n && !(lvl == 0 && l0)) {
      k_sel11_hist<<<gsel, 256, 0, h->stream>>>(sbits, n, st);
      NLP_LAUNCHED(h);
    }
    if (api) NLP_NCCL(h, api->AllReduce(st->hist, st->hist, 2048, ncclUint64, ncclSum, (ncclComm_t)h->comm, h->stream));
    k_sel11_step<<<1, 256, 0, h->stream>>>(st, (unsigned long long)K);
    NLP_LAUNCHED(h);
  }
  const uint32_t ntiles = (uint32_t)((n + OC_TILE - 1) / OC_TILE);
  if (n) {
    NLP_TRY(ensure(h, h->oc_counts, (size_t)ntiles * 8));
    NLP_TRY(ensure(h, h->oc_off, (size_t)ntiles * 8));
    k_ordered_count2<<<ntiles, OC_THREADS, 0, h->stream>>>(sbits, n, st, (unsigned long long*)h->oc_counts.p);
    NLP_LAUNCHED(h);
  }
  NLP_CUDA(h, cudaMemcpyAsync(h->h_sel11, st, offsetof(Select11, hist), cudaMemcpyDeviceToHost, h->stream));
  uint64_t packed = 0;
  if (n) NLP_TRY(exclusive_scan<unsigned long long>(h, (const unsigned long long*)h->oc_counts.p, ntiles, (unsigned long long*)h->oc_off.p, &packed));
  else NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  const uint64_t better = packed >> 32, tie = packed & 0xffffffffull;
  const uint64_t need = h->h_sel11->need;
  uint64_t need_r = std::min(need, tie);
  uint32_t key_or = h->h_sel11->key_or, key_nor = h->h_sel11->key_nor;   // which bits of the survivors' keys vary
  std::vector<unsigned long long> counts;
  if (dist) {
    // ONE small exchange: (better, tie, key bits) of every rank.  Ranks with ascending source ranges
    // take the tie class in rank order; everybody can then compute every rank's survivor count.
    const int W = h->world;
    NLP_TRY(ensure(h, h->gcnt, (size_t)(4 * W + 4) * 8));
    unsigned long long* d_cnt = (unsigned long long*)h->gcnt.p;
    unsigned long long mine[4] = {better, tie, key_or, key_nor};
    NLP_CUDA(h, cudaMemcpyAsync(d_cnt + 4 * W, mine, 32, cudaMemcpyHostToDevice, h->stream));
    NLP_NCCL(h, api->AllGather(d_cnt + 4 * W, d_cnt, 4, ncclUint64, (ncclComm_t)h->comm, h->stream));
    std::vector<unsigned long long> all(4 * W);
    NLP_CUDA(h, cudaMemcpyAsync(all.data(), d_cnt, (size_t)W * 32, cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));
    counts.resize(W);
    uint64_t tie_before = 0;
    key_or = 0; key_nor = 0;
    for (int r = 0; r < W; ++r) {
      const uint64_t b = all[4 * r], t = all[4 * r + 1];
      uint64_t nr = std::min<uint64_t>(need, t);
      if (h->part_ordered) nr = need > tie_before ? std::min<uint64_t>(need - tie_before, t) : 0;
      if (r == h->rank) need_r = nr;
      counts[r] = b + nr;
      tie_before += t;
      key_or |= (uint32_t)all[4 * r + 2]; key_nor |= (uint32_t)all[4 * r + 3];
    }
  }
  const uint64_t m = better + need_r;
  // Survivors go to buffer `ob`.  When the records live in candidate buffer `ob` themselves (no
  // cache), its u/v arrays are still being read: write to (cu[sbuf], cv[sbuf], cs[ob]) instead and
  // swap the two score arrays afterwards, which makes that triple buffer `sbuf`.  External records
  // (bucket path): survivors simply go to buffer 0.
  int res = ob;
  if (n) {
    uint32_t *ou = (uint32_t*)h->cu[ob].p, *ov = (uint32_t*)h->cv[ob].p, *os = (uint32_t*)h->cs[ob].p;
    if (external) { res = 0; ou = (uint32_t*)h->cu[0].p; ov = (uint32_t*)h->cv[0].p; os = (uint32_t*)h->cs[0].p; }
    else if (!h->pair_from_cache) { ou = (uint32_t*)h->cu[sbuf].p; ov = (uint32_t*)h->cv[sbuf].p; res = sbuf; }
    k_ordered_write2<<<ntiles, OC_THREADS, 0, h->stream>>>(pu, pv, sbits, n, st, (unsigned long long)need_r,
                                                           (const unsigned long long*)h->oc_off.p, ou, ov, os);
    NLP_LAUNCHED(h);
    if (!external && !h->pair_from_cache) std::swap(h->cs[0], h->cs[1]);
  } else if (external) res = 0;
  unsigned constant = 0;                             // digits of the score key in which no bit varies
  const uint32_t varying = key_or & key_nor;
  for (int d = 0; d < 4; ++d)
    if (((varying >> (8 * d)) & 255u) == 0u) constant |= 1u << d;
  uint64_t total = m;
  if (dist) {
    NLP_TRY(gather_candidates(h, res, m, &total, &counts));
    res = 0;
    if (!h->part_ordered) return top_k(h, 0, total, K, out_buf, out_n);
  }
  NLP_TRY(radix_sort_score(h, res, total, constant, out_buf));
  *out_n = std::min(total, K);
  return NLP_OK;
}

// The merge of a multi-GPU prediction on the source-centric kernels (replaces the serial T-way heap
// merge of inc/predict.hxx:431-460): `fill` unordered candidates in buffer `buf`; global cutoff by
// all-reduced select histograms on the 96-bit key, local compaction of the survivors, one
// all-gather, final on-device select + sort.  Every rank ends with the same result.
int dist_top_k(nlp_handle* h, int buf, uint64_t fill, uint64_t K, int* out_buf, uint64_t* out_n) {
  uint64_t m = fill, total = 0, fill_all = 0;
  int lb = buf;
  NLP_TRY(global_count(h, fill, &fill_all));
  if (K < fill_all) NLP_TRY(select_narrow(h, (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p, fill, K, 96u, true));
  if (K < fill_all && fill) {
    const int o = buf ^ 1;
    NLP_CUDA(h, cudaMemsetAsync(h->cursor2.p, 0, 8, h->stream));
    k_select_compact<<<grid_for(fill, 256 * 8, h->num_sms * 8), 256, 0, h->stream>>>(
        (const uint32_t*)h->cu[buf].p, (const uint32_t*)h->cv[buf].p, (const uint32_t*)h->cs[buf].p, fill,
        (const SelectState*)h->sel.p, (uint32_t*)h->cu[o].p, (uint32_t*)h->cv[o].p, (uint32_t*)h->cs[o].p,
        (unsigned long long*)h->cursor2.p);
    NLP_LAUNCHED(h);
    unsigned long long got = 0;
    NLP_CUDA(h, cudaMemcpyAsync(&got, h->cursor2.p, 8, cudaMemcpyDeviceToHost, h->stream));
    NLP_CUDA(h, cudaStreamSynchronize(h->stream));
    lb = o; m = got;
  }
  NLP_TRY(gather_candidates(h, lb, m, &total));
  return top_k(h, 0, total, K, out_buf, out_n);
}

}  // namespace


extern "C" {

int nlp_create(nlp_handle** out, int device) {
  if (!out) { g_create_error = "nlp_create: null output pointer"; return NLP_ERR_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("nlp_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
    return NLP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "nlp_create: bad device index"; return NLP_ERR_ARG; }
  nlp_handle* h = new nlp_handle();
  h->device = device;
  if (const char* e = getenv("NLP_B200_RANGE")) h->range_mode = atoi(e);   // experiment knobs
  if (const char* e = getenv("NLP_B200_RANGE_HALF")) h->range_half = atoi(e);
  if (const char* e = getenv("NLP_B200_RANGE_QUARTER")) h->range_quarter = atoi(e);
  if (const char* e = getenv("NLP_B200_RANGE_FENCE")) h->range_fence = atoi(e);
  if (const char* e = getenv("NLP_B200_RANGE_DIV")) h->range_div = (uint32_t)std::max(1, atoi(e));
  if (const char* e = getenv("NLP_B200_RANGE_FLT")) h->flt_range_mode = atoi(e);
  if (const char* e = getenv("NLP_B200_COOP")) h->coop_mode = atoi(e);
  if (const char* e = getenv("NLP_B200_BUCKET_CAP")) h->bucket_cap = atoi(e) == 8192 ? 8192u : 4096u;
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string("nlp_create: ") + what + ": " + cudaGetErrorString(err);
    delete h;
    return (int)NLP_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (prop.major < 10) {
    g_create_error = "nlp_create: kernels are built for sm_100a (B200) only";
    delete h;
    return NLP_ERR_CUDA;
  }
  h->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_stg_ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_stg_done[i], cudaEventDisableTiming);
  }
  cudaEventCreate(&h->ev_start); cudaEventCreate(&h->ev_frontier); cudaEventCreate(&h->ev_scored); cudaEventCreate(&h->ev_done);
  for (int i = 0; i < 7; ++i) cudaEventCreate(&h->ev_phase[i]);
  cudaEventCreate(&h->ev_eval0); cudaEventCreate(&h->ev_eval1);
  if ((e = cudaMallocHost((void**)&h->h_ctr, sizeof(Counters))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost((void**)&h->h_hist, 12 * 256 * 8)) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost((void**)&h->h_sel, sizeof(SelectState))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost((void**)&h->h_sel11, sizeof(Select11))) != cudaSuccess) return bail("cudaMallocHost", e);
  {
    // the big-source detour runs next to k_bucket, whose grid fills the GPU: a higher priority lets
    // the detour's many small kernels be scheduled as blocks of k_bucket retire instead of after it
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (const char* pe = getenv("NLP_B200_DETOUR_PRIORITY")) if (atoi(pe) == 0) hi = lo;
    if ((e = cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  int rc = NLP_OK;
  if ((rc = ensure(h, h->ctr, sizeof(Counters), true)) || (rc = ensure(h, h->thr, sizeof(Threshold), true)) ||
      (rc = ensure(h, h->totals, 256 * 4)) || (rc = ensure(h, h->hist, 12 * 256 * 8)) ||
      (rc = ensure(h, h->sel, sizeof(SelectState), true)) || (rc = ensure(h, h->cursor2, 16, true)) ||
      (rc = ensure(h, h->sel11, sizeof(Select11), true))) {
    g_create_error = h->err;
    delete h;
    return rc;
  }
  if ((e = cudaFuncSetAttribute(k_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_smem_bytes(true))) != cudaSuccess ||
      (e = cudaFuncSetAttribute(k_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_smem_bytes(false))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(k_scatter)", e);
  *out = h;
  return NLP_OK;
}

int nlp_destroy(nlp_handle* h) {
  if (!h) return NLP_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm) { nccl_api()->CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; }
  release(h->gcnt); release(h->gsend); release(h->grecv);
  clear_pair_cache(h);
  clear_plans(h);
  release(h->plan_tmp); release(h->al_u); release(h->al_v); release(h->al_s); release(h->al_c);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (int i = 0; i < 2; ++i) {
    release(h->stg_u[i]); release(h->stg_v[i]); release(h->stg_s[i]);
    if (h->ev_stg_ready[i]) cudaEventDestroy(h->ev_stg_ready[i]);
    if (h->ev_stg_done[i]) cudaEventDestroy(h->ev_stg_done[i]);
  }
  release(h->own_off); release(h->own_keys); release(h->spare_off); release(h->spare_keys); release(h->base_off); release(h->base_keys); release(h->del_bits); release(h->deg); release(h->work); release(h->work64); release(h->elig); release(h->maxdeg_dev);
  release(h->chunk_base); release(h->chunk_src); release(h->chunk_cnt); release(h->ecount); release(h->ekeys);
  release(h->scan_tiles); release(h->scan_total);
  release(h->it_u); release(h->it_cnt); release(h->it_dw); release(h->it_ptr); release(h->it_off); release(h->sym_flag);
  for (int b = 0; b < NBINS; ++b) { release(h->list[b]); release(h->defer[b]); }
  release(h->gtable); release(h->ctr); release(h->thr);
  for (int b = 0; b < 2; ++b) { release(h->cu[b]); release(h->cv[b]); release(h->cs[b]); }
  release(h->tables); release(h->touched); release(h->range_cursors); release(h->range_touched); release(h->fence_slot); release(h->fence_tab); release(h->flt_cnt); release(h->flt_off); release(h->flt_items); release(h->flt_ids); release(h->flt_defer); release(h->flt_tlist); release(h->counts); release(h->totals); release(h->hist);
  release(h->sel); release(h->cursor2); release(h->oc_counts); release(h->oc_off);
  release(h->truth_key); release(h->truth_tmp); release(h->eval_ctr);
  release(h->bt_d); release(h->bt_uat); release(h->bt_hit); release(h->bt_jump[0]); release(h->bt_jump[1]);
  release(h->bt_starts); release(h->bt_pos); release(h->del_u); release(h->del_v);
  if (h->ev_eval0) cudaEventDestroy(h->ev_eval0);
  if (h->ev_eval1) cudaEventDestroy(h->ev_eval1);
  if (h->h_ctr) cudaFreeHost(h->h_ctr);
  if (h->h_hist) cudaFreeHost(h->h_hist);
  if (h->h_sel) cudaFreeHost(h->h_sel);
  if (h->h_sel11) cudaFreeHost(h->h_sel11);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  release(h->sel11);
  for (auto& b : h->arena_pool) release(b);
  h->arena_pool.clear();
  for (int i = 0; i < 7; ++i) cudaEventDestroy(h->ev_phase[i]);
  cudaEventDestroy(h->ev_start); cudaEventDestroy(h->ev_frontier); cudaEventDestroy(h->ev_scored); cudaEventDestroy(h->ev_done);
  cudaStreamDestroy(h->stream);
  delete h;
  return NLP_OK;
}

int nlp_set_graph(nlp_handle* h, const uint64_t* offsets, const uint32_t* keys, uint32_t span) {
  if (!h) return NLP_ERR_ARG;
  if (!offsets) return fail(h, NLP_ERR_ARG, "nlp_set_graph: null offsets");
  if (offsets[0] != 0) return fail(h, NLP_ERR_ARG, "nlp_set_graph: offsets[0] must be 0");
  const uint64_t M = offsets[span];
  if (M && !keys) return fail(h, NLP_ERR_ARG, "nlp_set_graph: null keys");
  NLP_CUDA(h, cudaSetDevice(h->device));
  h->has_graph = false;
  NLP_TRY(ensure(h, h->own_off, ((size_t)span + 1) * 8));
  NLP_TRY(ensure(h, h->own_keys, (size_t)M * 4));
  NLP_CUDA(h, cudaMemcpyAsync(h->own_off.p, offsets, ((size_t)span + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  if (M) NLP_CUDA(h, cudaMemcpyAsync(h->own_keys.p, keys, (size_t)M * 4, cudaMemcpyHostToDevice, h->stream));
  h->d_off = (const uint64_t*)h->own_off.p;
  h->d_keys = (const uint32_t*)h->own_keys.p;
  h->S = span;
  return finish_graph(h);
}

int nlp_set_graph_device(nlp_handle* h, const uint64_t* d_offsets, const uint32_t* d_keys, uint32_t span) {
  if (!h) return NLP_ERR_ARG;
  if (!d_offsets) return fail(h, NLP_ERR_ARG, "nlp_set_graph_device: null offsets");
  NLP_CUDA(h, cudaSetDevice(h->device));
  h->has_graph = false;
  release(h->own_off); release(h->own_keys);
  h->d_off = d_offsets;
  h->d_keys = d_keys;
  h->S = span;
  const int rc = finish_graph(h);
  if (rc == NLP_OK && h->M && !d_keys) { h->has_graph = false; return fail(h, NLP_ERR_ARG, "nlp_set_graph_device: null keys"); }
  return rc;
}

int nlp_set_partition(nlp_handle* h, int rank, int world) {
  if (!h) return NLP_ERR_ARG;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, NLP_ERR_ARG, "nlp_set_partition: need 0 <= rank < world");
  if (h->comm && (rank != h->rank || world != h->world)) return fail(h, NLP_ERR_ARG, "nlp_set_partition: a communicator is active (nlp_comm_destroy first)");
  if (rank != h->rank || world != h->world) clear_count_store(h);
  h->rank = rank; h->world = world;
  return NLP_OK;
}

int nlp_comm_unique_id(void* id) {
  if (!id) return NLP_ERR_ARG;
  NcclApi* api = nccl_api();
  if (!api) { g_create_error = "nlp_comm_unique_id: NCCL not available"; return NLP_ERR_COMM; }
  ncclUniqueId uid;
  if (api->GetUniqueId(&uid) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return NLP_ERR_COMM; }
  static_assert(sizeof(uid) == NLP_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &uid, sizeof uid);
  return NLP_OK;
}

int nlp_comm_init(nlp_handle* h, const void* id, int rank, int world) {
  if (!h) return NLP_ERR_ARG;
  if (!id) return fail(h, NLP_ERR_ARG, "nlp_comm_init: null id");
  if (world < 1 || rank < 0 || rank >= world) return fail(h, NLP_ERR_ARG, "nlp_comm_init: need 0 <= rank < world");
  NcclApi* api = nccl_api();
  if (!api) return fail(h, NLP_ERR_COMM, "nlp_comm_init: libnccl.so.2 could not be loaded");
  NLP_CUDA(h, cudaSetDevice(h->device));
  if (h->comm) { api->CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; }
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  ncclComm_t c = nullptr;
  NLP_NCCL(h, api->CommInitRank(&c, world, uid, rank));
  h->comm = c;
  h->rank = rank; h->world = world;
  return NLP_OK;
}

int nlp_comm_destroy(nlp_handle* h) {
  if (!h) return NLP_ERR_ARG;
  if (h->comm) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    nccl_api()->CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
  }
  h->rank = 0; h->world = 1;
  return NLP_OK;
}

uint64_t nlp_comm_bytes(const nlp_handle* h) { return h ? h->gathered_bytes : 0; }

int nlp_set_reuse(nlp_handle* h, int on) {
  if (!h) return NLP_ERR_ARG;
  NLP_CUDA(h, cudaSetDevice(h->device));
  if (!h->pair_cache.empty()) { NLP_CUDA(h, cudaStreamSynchronize(h->stream)); clear_pair_cache(h); }
  clear_count_store(h);        // arenas go back to the pool: no cudaFree / cudaMalloc in steady state (stream-ordered reuse)
  h->reuse = on ? 1 : 0;
  return NLP_OK;
}

int nlp_set_path(nlp_handle* h, int path) {
  if (!h) return NLP_ERR_ARG;
  if (path != NLP_PATH_AUTO && path != NLP_PATH_SOURCE && path != NLP_PATH_PAIR && path != NLP_PATH_PAIR_SORT) return fail(h, NLP_ERR_ARG, "nlp_set_path: unknown path");
  h->path_mode = path;
  return NLP_OK;
}

int nlp_set_scratch_limit(nlp_handle* h, uint64_t bytes) {
  if (!h) return NLP_ERR_ARG;
  h->scratch_limit = bytes;
  return NLP_OK;
}

int nlp_predict(nlp_handle* h, const nlp_options* opt, nlp_result* res) {
  if (!h) return NLP_ERR_ARG;
  if (!opt || !res) return fail(h, NLP_ERR_ARG, "nlp_predict: null argument");
  if (opt->measure < 0 || opt->measure >= NLP_NUM_MEASURES) return fail(h, NLP_ERR_ARG, "nlp_predict: unknown measure");
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_predict: no graph set");
  NLP_CUDA(h, cudaSetDevice(h->device));
  memset(res, 0, sizeof *res);
  h->has_result = false;
  const bool flt = opt->measure == NLP_ADAMIC_ADAR || opt->measure == NLP_RESOURCE_ALLOCATION;
  if (opt->measure == NLP_ADAMIC_ADAR) NLP_TRY(ensure_gtable(h));
  NLP_TRY(ensure_candidates(h, 1024));
  const int reps = opt->repeat > 0 ? opt->repeat : 1;
  float scoring_sum = 0.f, frontier_sum = 0.f;
  int buf = 0;
  uint64_t fill = 0;
  if (opt->max_edges == 0 || h->S == 0) {           // inc/predict.hxx:429: nothing to do
    h->res_buf = 0; h->res_count = 0; h->has_result = true;
    return NLP_OK;
  }
  for (int r = 0; r < reps; ++r) {
    NLP_CUDA(h, cudaEventRecord(h->ev_start, h->stream));
    if (flt) NLP_TRY(scoring_pass<true>(h, opt, res, &buf, &fill));
    else     NLP_TRY(scoring_pass<false>(h, opt, res, &buf, &fill));
    NLP_CUDA(h, cudaEventRecord(h->ev_scored, h->stream));
    NLP_CUDA(h, cudaEventSynchronize(h->ev_scored));
    float ms = 0.f, fms = 0.f;
    NLP_CUDA(h, cudaEventElapsedTime(&ms, h->ev_start, h->ev_scored));
    NLP_CUDA(h, cudaEventElapsedTime(&fms, h->ev_start, h->ev_frontier));
    scoring_sum += ms; frontier_sum += fms;
  }
  int ob = buf;
  uint64_t on = 0;
  if (h->pair_pending)     NLP_TRY(ordered_top_k(h, opt->max_edges, &ob, &on));
  else if (distributed(h)) NLP_TRY(dist_top_k(h, buf, fill, opt->max_edges, &ob, &on));
  else                     NLP_TRY(top_k(h, buf, fill, opt->max_edges, &ob, &on));
  NLP_CUDA(h, cudaEventRecord(h->ev_done, h->stream));
  NLP_CUDA(h, cudaEventSynchronize(h->ev_done));
  float sel = 0.f;
  NLP_CUDA(h, cudaEventElapsedTime(&sel, h->ev_scored, h->ev_done));
  h->res_buf = ob; h->res_count = on; h->has_result = true;
  res->count = on;
  res->scoring_ms = scoring_sum / reps;
  res->frontier_ms = frontier_sum / reps;
  res->select_ms = sel;
  res->time_ms = res->scoring_ms + sel;
  NLP_CUDA(h, cudaEventElapsedTime(&res->phase_ms[0], h->ev_start, h->ev_frontier));
  if (h->phases_valid)
    for (int i = 1; i <= 6; ++i) NLP_CUDA(h, cudaEventElapsedTime(&res->phase_ms[i], h->ev_phase[i - 1], h->ev_phase[i]));
  else if (res->path == NLP_PATH_SOURCE)
    for (int i = 1; i <= 6; ++i) res->phase_ms[i] = h->phase_acc[i];
  res->phase_ms[7] = sel;
  return NLP_OK;
}

int nlp_fetch(nlp_handle* h, uint32_t* u, uint32_t* v, float* score, uint64_t capacity) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_result) return fail(h, NLP_ERR_NO_RESULT, "nlp_fetch: no result");
  const uint64_t n = std::min<uint64_t>(capacity, h->res_count);
  if (!n) return NLP_OK;
  if (!u || !v || !score) return fail(h, NLP_ERR_ARG, "nlp_fetch: null output");
  NLP_CUDA(h, cudaSetDevice(h->device));
  const int b = h->res_buf;
  NLP_CUDA(h, cudaMemcpyAsync(u, h->cu[b].p, n * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(v, h->cv[b].p, n * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(score, h->cs[b].p, n * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLP_OK;
}

int nlp_fetch_async(nlp_handle* h, uint32_t* u, uint32_t* v, float* score, uint64_t capacity) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_result) return fail(h, NLP_ERR_NO_RESULT, "nlp_fetch_async: no result");
  const uint64_t n = std::min<uint64_t>(capacity, h->res_count);
  if (!n) return NLP_OK;
  if (!u || !v || !score) return fail(h, NLP_ERR_ARG, "nlp_fetch_async: null output");
  NLP_CUDA(h, cudaSetDevice(h->device));
  const int s = h->stg_next;
  h->stg_next ^= 1;
  if (h->stg_busy[s]) NLP_CUDA(h, cudaEventSynchronize(h->ev_stg_done[s]));   // slot still feeding an older transfer
  NLP_TRY(ensure(h, h->stg_u[s], n * 4));
  NLP_TRY(ensure(h, h->stg_v[s], n * 4));
  NLP_TRY(ensure(h, h->stg_s[s], n * 4));
  const int b = h->res_buf;
  NLP_CUDA(h, cudaMemcpyAsync(h->stg_u[s].p, h->cu[b].p, n * 4, cudaMemcpyDeviceToDevice, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(h->stg_v[s].p, h->cv[b].p, n * 4, cudaMemcpyDeviceToDevice, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(h->stg_s[s].p, h->cs[b].p, n * 4, cudaMemcpyDeviceToDevice, h->stream));
  NLP_CUDA(h, cudaEventRecord(h->ev_stg_ready[s], h->stream));
  NLP_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_stg_ready[s], 0));
  NLP_CUDA(h, cudaMemcpyAsync(u, h->stg_u[s].p, n * 4, cudaMemcpyDefault, h->copy_stream));
  NLP_CUDA(h, cudaMemcpyAsync(v, h->stg_v[s].p, n * 4, cudaMemcpyDefault, h->copy_stream));
  NLP_CUDA(h, cudaMemcpyAsync(score, h->stg_s[s].p, n * 4, cudaMemcpyDefault, h->copy_stream));
  NLP_CUDA(h, cudaEventRecord(h->ev_stg_done[s], h->copy_stream));
  h->stg_busy[s] = true;
  return NLP_OK;
}

int nlp_fetch_wait(nlp_handle* h) {
  if (!h) return NLP_ERR_ARG;
  NLP_CUDA(h, cudaSetDevice(h->device));
  NLP_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  h->stg_busy[0] = h->stg_busy[1] = false;
  return NLP_OK;
}

int nlp_result_device(nlp_handle* h, const uint32_t** d_u, const uint32_t** d_v, const float** d_score, uint64_t* count) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_result) return fail(h, NLP_ERR_NO_RESULT, "nlp_result_device: no result");
  const int b = h->res_buf;
  if (d_u) *d_u = (const uint32_t*)h->cu[b].p;
  if (d_v) *d_v = (const uint32_t*)h->cv[b].p;
  if (d_score) *d_score = (const float*)h->cs[b].p;
  if (count) *count = h->res_count;
  return NLP_OK;
}

int nlp_merge(nlp_handle* h, const uint32_t* d_u, const uint32_t* d_v, const float* d_score, uint64_t n,
              uint64_t max_edges, float* select_ms) {
  if (!h) return NLP_ERR_ARG;
  if (n && (!d_u || !d_v || !d_score)) return fail(h, NLP_ERR_ARG, "nlp_merge: null input");
  if (n >= 0xfffffff0ull) return fail(h, NLP_ERR_CAPACITY, "nlp_merge: too many candidates");
  NLP_CUDA(h, cudaSetDevice(h->device));
  // The inputs must not point into this handle's own result buffers (nlp_result_device):
  // growing the candidate buffer frees them.  Copy the local result out first (nlp_fetch).
  const int b = 0;
  h->has_result = false;
  NLP_TRY(ensure_candidates(h, n));
  NLP_CUDA(h, cudaEventRecord(h->ev_start, h->stream));
  if (n) {
    NLP_CUDA(h, cudaMemcpyAsync(h->cu[b].p, d_u, n * 4, cudaMemcpyDeviceToDevice, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(h->cv[b].p, d_v, n * 4, cudaMemcpyDeviceToDevice, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(h->cs[b].p, d_score, n * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  int ob = b;
  uint64_t on = 0;
  NLP_TRY(top_k(h, b, n, max_edges, &ob, &on));
  NLP_CUDA(h, cudaEventRecord(h->ev_done, h->stream));
  NLP_CUDA(h, cudaEventSynchronize(h->ev_done));
  if (select_ms) NLP_CUDA(h, cudaEventElapsedTime(select_ms, h->ev_start, h->ev_done));
  h->res_buf = ob; h->res_count = on; h->has_result = true;
  return NLP_OK;
}

int nlp_generate_deletions(nlp_handle* h, uint32_t seed, uint64_t batch_size, uint64_t* count, uint64_t* words) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_generate_deletions: no graph set");
  NLP_CUDA(h, cudaSetDevice(h->device));
  h->has_deletions = false;
  h->del_n = 0;
  if (count) *count = 0;
  if (words) *words = 0;
  const uint64_t B = batch_size;
  const uint32_t S = h->S;
  if (B == 0 || S <= 1) { h->has_deletions = true; return NLP_OK; }
  const uint64_t P = 6 * B + 16;                    // a deletion uses at most 4 failed tries + 2 = 6 slots
  if (P >= 0xfffffff0ull) return fail(h, NLP_ERR_CAPACITY, "nlp_generate_deletions: batch too large");
  NLP_TRY(ensure(h, h->bt_d, P * 8));
  NLP_TRY(ensure(h, h->bt_uat, P * 4));
  NLP_TRY(ensure(h, h->bt_hit, P * 4));
  NLP_TRY(ensure(h, h->bt_jump[0], P * 4));
  NLP_TRY(ensure(h, h->bt_jump[1], P * 4));
  NLP_TRY(ensure(h, h->bt_starts, B * 4));
  NLP_TRY(ensure(h, h->bt_pos, 2 * B * 8));
  h->has_result = false;                            // the candidate buffers are reused for the pairs
  NLP_TRY(ensure_candidates(h, 2 * B));
  const DevGraph g = dev_graph(h);
  uint32_t seed0 = seed % 2147483647u;              // linear_congruential_engine::seed
  if (seed0 == 0) seed0 = 1;
  const unsigned gp = grid_for(P, 256, h->num_sms * 16), gb = grid_for(B, 256, h->num_sms * 16);
  k_batch_slots<<<gp, 256, 0, h->stream>>>(g.deg, S, seed0, P, (double*)h->bt_d.p, (uint32_t*)h->bt_uat.p);
  NLP_LAUNCHED(h);
  k_batch_next<<<gp, 256, 0, h->stream>>>((const uint32_t*)h->bt_uat.p, P, (uint32_t*)h->bt_hit.p, (uint32_t*)h->bt_jump[0].p);
  NLP_LAUNCHED(h);
  NLP_CUDA(h, cudaMemsetAsync(h->bt_starts.p, 0, 4, h->stream));       // the first deletion starts at slot 0
  int cur = 0;
  for (uint64_t known = 1; known < B;) {            // orbit of slot 0 by pointer doubling: jump = next^known
    const uint64_t take = std::min<uint64_t>(known, B - known);
    k_batch_extend<<<grid_for(take, 256, h->num_sms * 16), 256, 0, h->stream>>>((uint32_t*)h->bt_starts.p, known, take,
                                                                                  (const uint32_t*)h->bt_jump[cur].p);
    NLP_LAUNCHED(h);
    known += take;
    if (known < B) {
      k_batch_square<<<gp, 256, 0, h->stream>>>((const uint32_t*)h->bt_jump[cur].p, (uint32_t*)h->bt_jump[cur ^ 1].p, P);
      NLP_LAUNCHED(h);
      cur ^= 1;
    }
  }
  k_batch_emit<<<gb, 256, 0, h->stream>>>(g, (const uint32_t*)h->bt_starts.p, B, (const uint32_t*)h->bt_hit.p,
                                          (const uint32_t*)h->bt_uat.p, (const double*)h->bt_d.p,
                                          (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p);
  NLP_LAUNCHED(h);
  int sb = 0;
  NLP_TRY(radix_sort_pairs(h, 0, 2 * B, false, &sb));                  // sortEdgesByIdU, inc/batch.hxx:171-178
  uint32_t* head = (uint32_t*)h->cs[sb].p;                             // the score array is idle here
  k_batch_heads<<<grid_for(2 * B, 256, h->num_sms * 16), 256, 0, h->stream>>>((const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p,
                                                                               2 * B, head);
  NLP_LAUNCHED(h);
  uint64_t m = 0;
  NLP_TRY(exclusive_scan<uint32_t>(h, head, 2 * B, (unsigned long long*)h->bt_pos.p, &m));
  NLP_TRY(ensure(h, h->del_u, m * 4));
  NLP_TRY(ensure(h, h->del_v, m * 4));
  k_batch_compact<<<grid_for(2 * B, 256, h->num_sms * 16), 256, 0, h->stream>>>(
      (const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p, 2 * B, head, (const unsigned long long*)h->bt_pos.p,
      (uint32_t*)h->del_u.p, (uint32_t*)h->del_v.p);
  NLP_LAUNCHED(h);
  // engine words the batch consumed: two per slot, up to where the deletion after the last would start
  uint32_t last = 0, last_hit = 0;
  NLP_CUDA(h, cudaMemcpyAsync(&last, (const uint32_t*)h->bt_starts.p + (B - 1), 4, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(&last_hit, (const uint32_t*)h->bt_hit.p + last, 4, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  h->del_n = m;
  h->has_deletions = true;
  if (count) *count = m;
  if (words) *words = 2ull * (last_hit != BATCH_NONE ? (uint64_t)last_hit + 2ull : (uint64_t)last + 5ull);
  return NLP_OK;
}

int nlp_apply_deletions(nlp_handle* h, const uint32_t* del_u, const uint32_t* del_v, uint64_t n) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_apply_deletions: no graph set");
  if (n && (!del_u || !del_v)) return fail(h, NLP_ERR_ARG, "nlp_apply_deletions: null input");
  if (n >= 0xfffffff0ull) return fail(h, NLP_ERR_CAPACITY, "nlp_apply_deletions: batch too large");
  NLP_CUDA(h, cudaSetDevice(h->device));
  const uint32_t S = h->S;
  const uint64_t M = h->M;
  const DevGraph g = dev_graph(h);
  // Rows symmetric before?  Then a batch that holds both directions of every pair (as
  // tidyBatchUpdateU leaves it) keeps them symmetric, which k_del_mirror verifies for the price of
  // two bisections per request instead of a full symmetry pass over the new graph.
  NLP_TRY(check_symmetry(h));
  const int base_sym = h->sym_state;
  h->has_result = false;                               // the candidate buffers stage the request list
  NLP_TRY(ensure_candidates(h, n));
  uint32_t* du = (uint32_t*)h->cu[0].p; uint32_t* dv = (uint32_t*)h->cv[0].p;
  if (n) {
    NLP_CUDA(h, cudaMemcpyAsync(du, del_u, n * 4, cudaMemcpyDefault, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(dv, del_v, n * 4, cudaMemcpyDefault, h->stream));
  }
  const uint64_t nwords = (M + 31) / 32 + 2;
  NLP_TRY(ensure(h, h->del_bits, nwords * 4));
  NLP_CUDA(h, cudaMemsetAsync(h->del_bits.p, 0, nwords * 4, h->stream));
  uint32_t* marks = (uint32_t*)h->work.p;              // [S]: marks per row, then the new degrees
  NLP_CUDA(h, cudaMemsetAsync(marks, 0, (size_t)S * 4, h->stream));
  if (n) {
    k_del_mark<<<grid_for(n, 256, h->num_sms * 16), 256, 0, h->stream>>>(g, du, dv, n, (uint32_t*)h->del_bits.p, marks);
    NLP_LAUNCHED(h);
  }
  unsigned int asym = 0;
  if (n && base_sym == 1) {
    NLP_TRY(ensure(h, h->sym_flag, 32));
    NLP_CUDA(h, cudaMemsetAsync(h->sym_flag.p, 0, 4, h->stream));
    k_del_mirror<<<grid_for(n, 256, h->num_sms * 16), 256, 0, h->stream>>>(g, du, dv, n, (const uint32_t*)h->del_bits.p,
                                                                            (unsigned int*)h->sym_flag.p);
    NLP_LAUNCHED(h);
    NLP_CUDA(h, cudaMemcpyAsync(&asym, h->sym_flag.p, 4, cudaMemcpyDeviceToHost, h->stream));   // read by the synchronize below
  }
  // the new CSR goes to buffers this handle owns; a borrowed graph (nlp_set_graph_device) is left untouched
  const bool cur_is_own = h->d_off == (const uint64_t*)h->own_off.p && h->own_off.p != nullptr;
  DevBuf& n_off = cur_is_own ? h->spare_off : h->own_off;
  DevBuf& n_keys = cur_is_own ? h->spare_keys : h->own_keys;
  NLP_TRY(ensure(h, n_off, ((size_t)S + 1) * 8));
  NLP_TRY(ensure(h, n_keys, (size_t)M * 4));
  uint64_t M2 = 0;
  if (S) {
    k_del_newdeg<<<grid_for(S, 256, h->num_sms * 8), 256, 0, h->stream>>>(g.deg, marks, S);
    NLP_LAUNCHED(h);
    NLP_TRY(exclusive_scan<uint32_t>(h, marks, S, (unsigned long long*)n_off.p, &M2));
  }
  NLP_CUDA(h, cudaMemcpyAsync((unsigned long long*)n_off.p + S, &M2, 8, cudaMemcpyHostToDevice, h->stream));
  if (M) {
    const uint32_t ntiles = (uint32_t)((M + DEL_TILE - 1) / DEL_TILE);
    NLP_TRY(ensure(h, h->oc_counts, (size_t)ntiles * 4));
    NLP_TRY(ensure(h, h->oc_off, (size_t)ntiles * 8));
    k_del_count<<<ntiles, DEL_THREADS, 0, h->stream>>>((const uint8_t*)h->del_bits.p, M, (uint32_t*)h->oc_counts.p);
    NLP_LAUNCHED(h);
    NLP_TRY(exclusive_scan<uint32_t>(h, (const uint32_t*)h->oc_counts.p, ntiles, (unsigned long long*)h->oc_off.p, nullptr));
    k_del_write<<<ntiles, DEL_THREADS, 0, h->stream>>>(g.keys, (const uint8_t*)h->del_bits.p, M,
                                                       (const unsigned long long*)h->oc_off.p, (uint32_t*)n_keys.p);
    NLP_LAUNCHED(h);
  }
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));       // &M2 above is a stack address
  if (cur_is_own) { std::swap(h->own_off, h->spare_off); std::swap(h->own_keys, h->spare_keys); }
  h->d_off = (const uint64_t*)h->own_off.p;
  h->d_keys = (const uint32_t*)h->own_keys.p;
  h->has_graph = false;
  // the result of a validated graph is valid by construction; its symmetry follows from the base's
  return finish_graph(h, true, base_sym == 1 ? (asym ? 0 : 1) : base_sym == 2 ? 0 : 0);
}

int nlp_graph_checkpoint(nlp_handle* h) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_graph_checkpoint: no graph set");
  NLP_CUDA(h, cudaSetDevice(h->device));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->d_off == (const uint64_t*)h->own_off.p && h->own_off.p) {
    // the handle's own arrays become the base (no copy); nlp_apply_deletions then writes elsewhere
    release(h->base_off); release(h->base_keys);
    h->base_off = h->own_off; h->base_keys = h->own_keys;
    h->own_off = DevBuf(); h->own_keys = DevBuf();
  } else if (h->d_off == (const uint64_t*)h->base_off.p && h->base_off.p) {
    // already the base
  } else {
    release(h->base_off); release(h->base_keys);     // a borrowed graph: only the pointers are remembered
  }
  h->base_d_off = h->d_off; h->base_d_keys = h->d_keys; h->base_S = h->S;
  h->base_sym = h->sym_state; h->base_maxmult = h->maxmult; h->base_id = h->graph_id;
  h->has_base = true;
  return NLP_OK;
}

int nlp_graph_rollback(nlp_handle* h) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_base) return fail(h, NLP_ERR_NO_GRAPH, "nlp_graph_rollback: no checkpoint");
  NLP_CUDA(h, cudaSetDevice(h->device));
  if (h->base_sym == 0 && h->base_id) {               // learnt since the checkpoint?
    auto known = h->known_sym.find(h->base_id);
    if (known != h->known_sym.end()) h->base_sym = known->second;
  }
  h->d_off = h->base_d_off; h->d_keys = h->base_d_keys; h->S = h->base_S;
  h->has_graph = false;
  h->maxmult = h->base_maxmult;
  const int rc = finish_graph(h, true, h->base_sym);   // validated when it was set: no second pass
  if (rc == NLP_OK) h->graph_id = h->base_id;
  return rc;
}

int nlp_graph_size(nlp_handle* h, uint32_t* span, uint64_t* entries) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_graph_size: no graph set");
  if (span) *span = h->S;
  if (entries) *entries = h->M;
  return NLP_OK;
}

int nlp_fetch_graph(nlp_handle* h, uint64_t* offsets, uint32_t* keys) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_graph) return fail(h, NLP_ERR_NO_GRAPH, "nlp_fetch_graph: no graph set");
  NLP_CUDA(h, cudaSetDevice(h->device));
  if (offsets) NLP_CUDA(h, cudaMemcpyAsync(offsets, h->d_off, ((size_t)h->S + 1) * 8, cudaMemcpyDefault, h->stream));
  if (keys && h->M) NLP_CUDA(h, cudaMemcpyAsync(keys, h->d_keys, (size_t)h->M * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLP_OK;
}

int nlp_ingest_mtx(nlp_handle* h, const char* text, uint64_t bytes, uint32_t flags, uint32_t* span, uint64_t* entries) {
  if (!h) return NLP_ERR_ARG;
  if (!text || !bytes) return fail(h, NLP_ERR_ARG, "nlp_ingest_mtx: no text");
  NLP_CUDA(h, cudaSetDevice(h->device));
  bool coordinate = false, symmetric = false;
  uint64_t rows = 0, cols = 0, size = 0;
  const uint64_t body0 = nlp::mtx_header(text, bytes, &coordinate, &symmetric, &rows, &cols, &size);
  if (!coordinate) return fail(h, NLP_ERR_ARG, "nlp_ingest_mtx: not a Matrix Market coordinate file");
  const uint64_t n = std::max(rows, cols);
  if (n >= 0xfffffffeull) return fail(h, NLP_ERR_CAPACITY, "nlp_ingest_mtx: more than 2^32 - 3 vertices");
  const uint32_t S = (uint32_t)n + 1u;
  // second pair of a line: 1 = the reader stores it too (symmetric banner), 2 = symmetrizeOmp adds it
  const int second = symmetric ? 1 : ((flags & NLP_INGEST_SYMMETRIZE) ? 2 : 0);
  const int drop_self = (flags & NLP_INGEST_DROP_SELF_LOOPS) ? 1 : 0;
  h->has_graph = false; h->has_result = false; h->has_base = false;
  h->S = S;                                             // the pair sort reads its digit count from the span
  // the text on the device
  const uint64_t body = bytes - body0;
  DevBuf d_text, d_tiles, d_at, d_ymin, d_xmax;
  auto drop = [&](int rc) { release(d_text); release(d_tiles); release(d_at); release(d_ymin); release(d_xmax); return rc; };
  uint64_t L = 0, P = 0, M = 0;
  int rc = NLP_OK;
  if (body) {
    if ((rc = ensure(h, d_text, bytes + 16)) != NLP_OK) return drop(rc);
    NLP_CUDA(h, cudaMemcpyAsync(d_text.p, text, bytes, cudaMemcpyHostToDevice, h->stream));
    const uint64_t ntiles = (body + MTX_TILE - 1) / MTX_TILE;
    if (ntiles >= 0xffffffffull) return drop(fail(h, NLP_ERR_CAPACITY, "nlp_ingest_mtx: text too large"));
    if ((rc = ensure(h, h->oc_counts, (size_t)ntiles * 4)) != NLP_OK) return drop(rc);
    if ((rc = ensure(h, d_tiles, (size_t)ntiles * 8)) != NLP_OK) return drop(rc);
    k_mtx_count<<<(unsigned)ntiles, MTX_THREADS, 0, h->stream>>>((const uint8_t*)d_text.p, body0, bytes, (uint32_t*)h->oc_counts.p);
    NLP_LAUNCHED(h);
    if ((rc = exclusive_scan<uint32_t>(h, (const uint32_t*)h->oc_counts.p, ntiles, (unsigned long long*)d_tiles.p, &L)) != NLP_OK) return drop(rc);
    if (2 * L >= 0xfffffff0ull) return drop(fail(h, NLP_ERR_CAPACITY, "nlp_ingest_mtx: too many lines for one pass"));
    if ((rc = ensure_candidates(h, 2 * L + 16)) != NLP_OK) return drop(rc);
    uint32_t* eu = (uint32_t*)h->cu[1].p; uint32_t* ev = (uint32_t*)h->cv[1].p; uint32_t* cnt = (uint32_t*)h->cs[1].p;
    if (L) {
      if ((rc = ensure(h, h->sym_flag, 32)) != NLP_OK) return drop(rc);
      NLP_CUDA(h, cudaMemsetAsync(h->sym_flag.p, 0, 4, h->stream));
      k_mtx_parse<<<(unsigned)ntiles, MTX_THREADS, 0, h->stream>>>((const uint8_t*)d_text.p, body0, bytes, (uint32_t)n,
                                                                   (const unsigned long long*)d_tiles.p, eu, ev, (unsigned int*)h->sym_flag.p);
      NLP_LAUNCHED(h);
      unsigned int bad = 0;
      NLP_CUDA(h, cudaMemcpyAsync(&bad, h->sym_flag.p, 4, cudaMemcpyDeviceToHost, h->stream));
      k_ing_count<<<grid_for(L, 256, h->num_sms * 16), 256, 0, h->stream>>>(eu, L, second, cnt);
      NLP_LAUNCHED(h);
      if ((rc = ensure(h, d_at, (size_t)L * 8)) != NLP_OK) return drop(rc);
      if ((rc = exclusive_scan<uint32_t>(h, cnt, L, (unsigned long long*)d_at.p, &P)) != NLP_OK) return drop(rc);   // synchronises: `bad` is valid
      if (bad & 1u) return drop(fail(h, NLP_ERR_ARG, "nlp_ingest_mtx: a vertex id is outside 1..max(rows, cols)"));
      if (P) {
        k_ing_emit<<<grid_for(L, 256, h->num_sms * 16), 256, 0, h->stream>>>(eu, ev, L, cnt, (const unsigned long long*)d_at.p, second,
                                                                            (uint32_t*)h->cu[0].p, (uint32_t*)h->cv[0].p, (uint32_t*)h->cs[0].p);
        NLP_LAUNCHED(h);
      }
    }
  }
  // sort by (u, v) with the tag as payload, copies per distinct pair, rows -> offsets
  int sb = 0;
  uint32_t* copies = nullptr;
  if (P) {
    if ((rc = radix_sort_pairs(h, 0, P, true, &sb)) != NLP_OK) return drop(rc);
    copies = (uint32_t*)h->cs[sb ^ 1].p;
    if ((rc = ensure(h, d_ymin, (size_t)S * 4)) != NLP_OK) return drop(rc);
    if ((rc = ensure(h, d_xmax, (size_t)S * 4)) != NLP_OK) return drop(rc);
    NLP_CUDA(h, cudaMemsetAsync(d_ymin.p, 0xff, (size_t)S * 4, h->stream));
    NLP_CUDA(h, cudaMemsetAsync(d_xmax.p, 0, (size_t)S * 4, h->stream));
    const uint32_t* pu = (const uint32_t*)h->cu[sb].p; const uint32_t* pv = (const uint32_t*)h->cv[sb].p;
    k_ing_class<<<grid_for(P, 256, h->num_sms * 16), 256, 0, h->stream>>>(pu, pv, (const uint32_t*)h->cs[sb].p, P, copies,
                                                                          (uint32_t*)d_ymin.p, (uint32_t*)d_xmax.p);
    NLP_LAUNCHED(h);
    k_ing_copies<<<grid_for(P, 256, h->num_sms * 16), 256, 0, h->stream>>>(pu, pv, P, (const uint32_t*)d_ymin.p, (const uint32_t*)d_xmax.p,
                                                                           drop_self, copies);
    NLP_LAUNCHED(h);
    if ((rc = ensure(h, d_at, (size_t)P * 8)) != NLP_OK) return drop(rc);
    if ((rc = exclusive_scan<uint32_t>(h, copies, P, (unsigned long long*)d_at.p, &M)) != NLP_OK) return drop(rc);
  }
  // a graph that is being replaced may live in own_*: the new one goes to the spare pair, as in nlp_apply_deletions
  const bool cur_is_own = h->d_off == (const uint64_t*)h->own_off.p && h->own_off.p != nullptr;
  DevBuf& n_off = cur_is_own ? h->spare_off : h->own_off;
  DevBuf& n_keys = cur_is_own ? h->spare_keys : h->own_keys;
  if ((rc = ensure(h, n_off, ((size_t)S + 1) * 8)) != NLP_OK) return drop(rc);
  if ((rc = ensure(h, n_keys, (size_t)std::max<uint64_t>(M, 1) * 4)) != NLP_OK) return drop(rc);
  if (P) {
    k_ing_write<<<grid_for(P, 256, h->num_sms * 16), 256, 0, h->stream>>>((const uint32_t*)h->cu[sb].p, (const uint32_t*)h->cv[sb].p, P,
                                                                          copies, (const unsigned long long*)d_at.p,
                                                                          (uint32_t*)n_keys.p, (unsigned long long*)n_off.p);
    NLP_LAUNCHED(h);
  }
  k_ing_tail<<<grid_for((uint64_t)S + 1, 256, h->num_sms * 8), 256, 0, h->stream>>>((const uint32_t*)h->cu[sb].p, P, S, (unsigned long long)M,
                                                                                    (unsigned long long*)n_off.p);
  NLP_LAUNCHED(h);
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  drop(NLP_OK);
  if (cur_is_own) { std::swap(h->own_off, h->spare_off); std::swap(h->own_keys, h->spare_keys); }
  h->d_off = (const uint64_t*)h->own_off.p;
  h->d_keys = (const uint32_t*)h->own_keys.p;
  // validated like any other graph (fingerprint, largest entry multiplicity); whether the rows are
  // symmetric as multisets is found out when a path needs to know (the reference's duplicated
  // entries need not be mirrored)
  NLP_TRY(finish_graph(h));
  if (span) *span = S;
  if (entries) *entries = h->M;
  return NLP_OK;
}

int nlp_fetch_deletions(nlp_handle* h, uint32_t* u, uint32_t* v, uint64_t capacity) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_deletions) return fail(h, NLP_ERR_NO_RESULT, "nlp_fetch_deletions: no batch generated");
  const uint64_t n = std::min<uint64_t>(capacity, h->del_n);
  if (!n) return NLP_OK;
  if (!u || !v) return fail(h, NLP_ERR_ARG, "nlp_fetch_deletions: null output");
  NLP_CUDA(h, cudaSetDevice(h->device));
  NLP_CUDA(h, cudaMemcpyAsync(u, h->del_u.p, n * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaMemcpyAsync(v, h->del_v.p, n * 4, cudaMemcpyDefault, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  return NLP_OK;
}

int nlp_deletions_device(nlp_handle* h, const uint32_t** d_u, const uint32_t** d_v, uint64_t* count) {
  if (!h) return NLP_ERR_ARG;
  if (!h->has_deletions) return fail(h, NLP_ERR_NO_RESULT, "nlp_deletions_device: no batch generated");
  if (d_u) *d_u = (const uint32_t*)h->del_u.p;
  if (d_v) *d_v = (const uint32_t*)h->del_v.p;
  if (count) *count = h->del_n;
  return NLP_OK;
}

int nlp_set_truth(nlp_handle* h, const uint32_t* u, const uint32_t* v, uint64_t n) {
  if (!h) return NLP_ERR_ARG;
  if (n && (!u || !v)) return fail(h, NLP_ERR_ARG, "nlp_set_truth: null input");
  NLP_CUDA(h, cudaSetDevice(h->device));
  h->has_truth = false;
  NLP_TRY(ensure(h, h->truth_key, n * 8));
  NLP_TRY(ensure(h, h->truth_tmp, n * 8));
  NLP_TRY(ensure(h, h->eval_ctr, 16));
  NLP_CUDA(h, cudaMemsetAsync(h->eval_ctr.p, 0, 16, h->stream));
  if (n) {
    uint32_t* du = (uint32_t*)h->truth_tmp.p;
    uint32_t* dv = du + n;
    NLP_CUDA(h, cudaMemcpyAsync(du, u, n * 4, cudaMemcpyDefault, h->stream));
    NLP_CUDA(h, cudaMemcpyAsync(dv, v, n * 4, cudaMemcpyDefault, h->stream));
    k_truth_pack<<<grid_for(n, 256, h->num_sms * 8), 256, 0, h->stream>>>(
        du, dv, n, (unsigned long long*)h->truth_key.p, (unsigned int*)h->eval_ctr.p + 2);
    NLP_LAUNCHED(h);
  }
  unsigned int unsorted = 0;
  NLP_CUDA(h, cudaMemcpyAsync(&unsorted, (unsigned int*)h->eval_ctr.p + 2, 4, cudaMemcpyDeviceToHost, h->stream));
  NLP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (unsorted) return fail(h, NLP_ERR_ARG, "nlp_set_truth: the edge list is not sorted ascending by (u, v)");
  h->truth_n = n;
  h->has_truth = true;
  return NLP_OK;
}

int nlp_evaluate(nlp_handle* h, nlp_evaluation* out) {
  if (!h) return NLP_ERR_ARG;
  if (!out) return fail(h, NLP_ERR_ARG, "nlp_evaluate: null output");
  if (!h->has_result) return fail(h, NLP_ERR_NO_RESULT, "nlp_evaluate: no result (run nlp_predict / nlp_merge first)");
  if (!h->has_truth) return fail(h, NLP_ERR_NO_TRUTH, "nlp_evaluate: no held-back edges (call nlp_set_truth first)");
  NLP_CUDA(h, cudaSetDevice(h->device));
  memset(out, 0, sizeof *out);
  const uint64_t count = h->res_count, n = h->truth_n;
  unsigned long long common = 0;
  NLP_CUDA(h, cudaEventRecord(h->ev_eval0, h->stream));
  if (count && n) {
    NLP_CUDA(h, cudaMemsetAsync(h->eval_ctr.p, 0, 8, h->stream));
    const int b = h->res_buf;
    k_evaluate<<<grid_for(count, 256, h->num_sms * 8), 256, 0, h->stream>>>(
        (const uint32_t*)h->cu[b].p, (const uint32_t*)h->cv[b].p, count, (const unsigned long long*)h->truth_key.p, n,
        (unsigned long long*)h->eval_ctr.p);
    NLP_LAUNCHED(h);
    NLP_CUDA(h, cudaMemcpyAsync(&common, h->eval_ctr.p, 8, cudaMemcpyDeviceToHost, h->stream));
  }
  NLP_CUDA(h, cudaEventRecord(h->ev_eval1, h->stream));
  NLP_CUDA(h, cudaEventSynchronize(h->ev_eval1));
  NLP_CUDA(h, cudaEventElapsedTime(&out->ms, h->ev_eval0, h->ev_eval1));
  out->predicted = 2 * count;                                  // main.cxx:51-54
  out->truth = n;
  out->common = common;                                        // main.cxx:55
  out->precision = (double)common / (double)std::max<uint64_t>(out->predicted, 1);   // main.cxx:201
  out->recall    = (double)common / (double)std::max<uint64_t>(out->truth, 1);       // main.cxx:202
  return NLP_OK;
}

uint64_t nlp_launch_count(const nlp_handle* h) { return h ? h->launches : 0; }

void* nlp_stream(const nlp_handle* h) { return h ? (void*)h->stream : nullptr; }

const char* nlp_last_error(const nlp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

const char* nlp_version(void) { return "nlp_b200 0.1 (sm_100a)"; }

}  // extern "C"
