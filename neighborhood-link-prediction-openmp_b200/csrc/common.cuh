// Shared device-side types and helpers of the B200 IHub/LHub link-prediction path.
// Hand-written for sm_100a; no CPU fallback, no multi-backend dispatch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NLP_FULL 0xffffffffu
#define NLP_EMPTY 0xffffffffu     // hash-slot sentinel (vertex ids are < span <= 2^32-1)
#define NLP_NO_SCORE 0xffffffffu  // "no pair here" in a record-aligned score array (a NaN pattern no kept
                                  // score can have; desc_key maps it to the worst key 0xffffffff)

namespace nlp {

enum Measure { M_CN = 0, M_JC, M_SI, M_SC, M_HP, M_HD, M_LHN, M_AA, M_RA };

enum { NBINS = 7 };
// Rows longer than LONG_ROW first-hop entries are cut into CHUNK-entry pieces by the frontier
// pass (one block per piece), so that a hub's row never serialises a single warp.
constexpr uint32_t LONG_ROW = 2048;
constexpr uint32_t CHUNK = 2048;
// Source bins (north_star kernel (a)): which path a source vertex takes.
//   0: 8-lane sub-warp groups   (work <= 8)        1: 32-lane warp (work <= 32)
//   2: smem hash, 1K slots      (bound <= 768)     3: smem hash, 4K slots (bound <= 3072)
//   4: smem hash, 16K slots     (bound <= 12288)   5: global dense spill table
//   6: windowed shared-memory counters (k_range): hub-heavy sources of the count measures whose
//      work pays for walking the vertex range in windows of C counters
// work(u)  = sum of deg(w) over eligible first-hop entries w of u   (wedges the reference scans)
// bound(u) = min(work(u), S-1-u) >= number of distinct v > u the source can touch
__host__ __device__ inline uint32_t bin_slots(int bin) { return bin == 2 ? 1024u : bin == 3 ? 4096u : 16384u; }
__host__ __device__ inline uint32_t bin_limit(int bin) { return bin_slots(bin) / 4u * 3u; }

struct DevGraph {
  const uint64_t* off;   // [S+1]
  const uint32_t* keys;  // [M], rows sorted ascending
  const uint32_t* deg;   // [S]
  uint32_t        S;
};

// Device-resident counters, zeroed at the start of every scoring repeat.
struct Counters {
  unsigned long long cursor;            // entries written to the candidate buffer
  unsigned long long reserved;          // admission control: cursor + bounds of in-flight sources
  unsigned long long first_hop;
  unsigned long long eligible_first_hop;
  unsigned long long wedges;
  unsigned long long candidates;
  unsigned long long kept;
  unsigned long long frontier;
  unsigned long long bin_count[8];
  unsigned long long bin_bound[8];      // sum of bound(u) (work(u) for bins 0,1) per bin
  unsigned long long deferred[8];       // sources pushed to the next pass (buffer full)
  unsigned long long queue[8];          // persistent-kernel work queues
  unsigned long long max_bound;         // largest bound in the dense bin
  unsigned long long overflow;          // must stay 0 (a table or the buffer overflowed)
};

// Pruning threshold: once the buffer held >= K candidates, only pairs strictly better than the
// current K-th one (canonical order: score desc, u asc, v asc) can still reach the result.
struct Threshold {
  uint32_t active;
  uint32_t key;   // desc_key(score)
  uint32_t u, v;
};

struct Params {
  DevGraph g;
  uint32_t D;          // MINDEGREE1 (0 = IHub)
  uint32_t F2;         // MAXFACTOR2
  uint32_t coop;       // count measures: block-cooperative wedge streaming (maxdeg small enough)
  uint32_t range_half; // k_range: sources with deg < range_half count in half words (windows twice as wide); 0 = never
  uint32_t range_quarter; // ... and those with deg < range_quarter in bytes (four times as wide); 0 = never
  int      measure;
  float    min_score;
  const uint32_t* elig;     // LHub eligibility bitmask (bit w = deg(w) <= D), null for IHub
  const double*   gtable;   // Adamic-Adar: gtable[d] = 1.0 / log((double)d), host libm values
  const uint32_t* work;     // [S] saturated work(u)
  // LHub: eligible first-hop entries (deg(w) <= D, deg(w) > 0) compacted by the frontier pass.
  // Row u's entries start at ekeys[off[u]]; a short row holds ecount[u] of them, a long row
  // (deg(u) > LONG_ROW) holds chunk_cnt[chunk_base[u] + c] at ekeys[off[u] + c * CHUNK].
  // Null for IHub (every first-hop entry is eligible: the row itself is the list).
  const uint32_t* ekeys;
  const uint32_t* ecount;
  const unsigned long long* chunk_base;
  const uint32_t* chunk_cnt;
  uint32_t* cu; uint32_t* cv; float* cs;   // candidate buffer (SoA)
  unsigned long long cap;
  unsigned long long soft_cap;   // admission of a pass stops once this many candidates are written
  Counters* ctr;
  const Threshold* thr;
};

// ---- order-preserving float key ------------------------------------------------------------
// ascending desc_key <=> descending score (IEEE total order on non-NaN floats)
__host__ __device__ inline uint32_t desc_key(uint32_t bits) {
  uint32_t asc = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);
  return ~asc;
}

__device__ __forceinline__ bool eligible(const Params& p, uint32_t w) {
  return p.elig ? ((__ldg(p.elig + (w >> 5)) >> (w & 31)) & 1u) != 0u : true;
}

__device__ __forceinline__ uint32_t hash32(uint32_t v) { return v * 0x9E3779B1u; }

// The nine score functions, same type chain as the reference lambdas
// (inc/predict.hxx:521,559,597,635,673,711,749,789,829): degrees are size_t, the count is
// uint32 (float for AA/RA), W = float; every operation IEEE round-to-nearest, no FMA.
__device__ __forceinline__ float score_fn(int measure, uint64_t du, uint64_t dv, uint32_t n, float nf) {
  const float N = __uint2float_rn(n);
  switch (measure) {
    case M_CN:  return N;
    case M_JC:  return __fdiv_rn(N, __ull2float_rn(du + dv - (uint64_t)n));
    case M_SI:  return __fdiv_rn(N, __ull2float_rn(du + dv));
    case M_SC:  return __double2float_rn(__ddiv_rn((double)N, __dsqrt_rn(__ull2double_rn(du * dv))));
    case M_HP:  return __fdiv_rn(N, __ull2float_rn(du < dv ? du : dv));
    case M_HD:  return __fdiv_rn(N, __ull2float_rn(du > dv ? du : dv));
    case M_LHN: return __fdiv_rn(N, __ull2float_rn(du * dv));
    default:    return nf;
  }
}

__device__ __forceinline__ bool measure_needs_dv(int measure, uint32_t F2) {
  return F2 != 0 || (measure >= M_JC && measure <= M_LHN);
}

// Per-warp statistics kept in registers and flushed once per kernel.
struct Tally {
  unsigned long long candidates, kept;
  __device__ Tally() : candidates(0), kept(0) {}
  __device__ void flush(Counters* c) {
    unsigned long long a = candidates, b = kept;
    #pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      a += __shfl_xor_sync(NLP_FULL, a, d);
      b += __shfl_xor_sync(NLP_FULL, b, d);
    }
    if ((threadIdx.x & 31) == 0) {
      if (a) atomicAdd(&c->candidates, a);
      if (b) atomicAdd(&c->kept, b);
    }
  }
};

// Score one touched pair (kernel (d), fused epilogue): returns true when the pair is kept
// (score > min_score, inc/predict.hxx:311) and passes the pruning threshold; tallies candidates
// and kept pairs.
__device__ __forceinline__ bool score_pair(const Params& p, bool has, uint32_t u, uint64_t du, uint32_t v, uint32_t n,
                                           float nf, Tally& t, float* score_out) {
  float score = 0.0f;
  bool keep = false;
  if (has) {
    uint64_t dv = 0;
    if (measure_needs_dv(p.measure, p.F2)) dv = __ldg(p.g.deg + v);
    // MAXFACTOR2 (inc/predict.hxx:292-296) depends on (u, v) only, so filtering the touched
    // pair here equals filtering every wedge: the pair is simply never touched.
    if (p.F2 && !(du <= (uint64_t)p.F2 * du && dv <= (uint64_t)p.F2 * du)) has = false;
    if (has) {
      score = score_fn(p.measure, du, dv, n, nf);
      keep = score > p.min_score;                                   // inc/predict.hxx:311
    }
  }
  t.candidates += has ? 1u : 0u;
  t.kept += keep ? 1u : 0u;
  if (keep && p.thr->active) {
    const uint32_t k = desc_key(__float_as_uint(score));
    const Threshold T = *p.thr;
    keep = (k < T.key) || (k == T.key && (u < T.u || (u == T.u && v < T.v)));
  }
  *score_out = score;
  return keep;
}

// Score one touched pair and append it to the candidate buffer.
// Must be called by all 32 lanes of a warp (`has` = this lane holds a touched pair).
// Returns the number of pairs the warp appended.
__device__ __forceinline__ uint32_t score_and_emit(const Params& p, bool has, uint32_t u, uint64_t du,
                                                   uint32_t v, uint32_t n, float nf, Tally& t) {
  float score;
  const bool keep = score_pair(p, has, u, du, v, n, nf, t, &score);
  const unsigned m = __ballot_sync(NLP_FULL, keep);
  if (m == 0) return 0;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(&p.ctr->cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(NLP_FULL, base, leader);
  if (keep) {
    const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
    if (pos < p.cap) { p.cu[pos] = u; p.cv[pos] = v; p.cs[pos] = score; }
    else atomicAdd(&p.ctr->overflow, 1ull);
  }
  return __popc(m);
}

}  // namespace nlp
