// LHub pair path: the w-centric formulation of inc/predict.hxx:298-304 for graphs whose rows are
// symmetric (entry multiplicity of w in row u == multiplicity of u in row w; checked on the
// device by k_symmetry, otherwise the source-centric kernels of wedge.cuh are used).
//
// The reference visits, for every source u, every eligible first-hop entry w (deg(w) <= D) and
// every entry v > u of row w.  On a symmetric graph the same wedges are found from w's side: each
// entry x_i = u of an eligible row w, paired with every later entry x_j = v > u of the same row.
// So instead of scanning all |E| first-hop entries for the hub test (the reference's dominant
// LHub cost, inc/predict.hxx:301) only the rows of eligible w are read:
//
//   k_pair_rows    items(w) = deg(w) if 2 <= deg(w) <= D        (one "item" = one entry of the row)
//   k_pair_items   per item: u, number of wedges cnt = #{j : x_j > u}, pointer to the first v
//   k_pair_emit    warp per 32 items, wedges packed back to back -> coalesced (u, v[, deg w]) records
//                  at deterministic positions (two exclusive scans), so records are in ascending w
//   radix sort     stable LSD sort of the records by (u, v)   (select.cuh)
//   k_pair_reduce  run-length count of equal (u, v) (or the ordered float fold for AA / RA --
//                  stable sort keeps ascending w, the reference's accumulation order), existing-
//                  edge exclusion by binary search in row u, fused scoring (inc/predict.hxx:306-311)
#pragma once
#include "common.cuh"
#include "frontier.cuh"
#include "wedge.cuh"

namespace nlp {

// Rows are symmetric <=> asym == 0 and dir[0] == dir[1] afterwards.  Only the entries (u, w) with
// u < w look up their mirror image (multiplicity of u in row w); dir[0] / dir[1] count the entries
// with u < w / u > w.  If every looked-up pair matches, the u > w side holds exactly the mirrored
// entries plus the entries nobody looked up, so equal counts mean there are none of those.
__global__ void __launch_bounds__(256) k_symmetry(DevGraph g, uint64_t M, unsigned int* __restrict__ asym,
                                                  unsigned long long* __restrict__ dir) {
  const uint32_t* __restrict__ keys = g.keys;
  const int lane = threadIdx.x & 31;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long n_up = 0, n_down = 0;
  for (uint64_t base = warp0 * 32u; base < M; base += nwarps * 32u) {
    // row of the warp's first entry by bisection (same loads in every lane: broadcast), then
    // every lane walks forward to its own row -- 32 consecutive entries span few rows
    uint32_t lo = 0, hi = g.S;                       // off[lo] <= base < off[hi]
    while (lo + 1 < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(g.off + mid) <= base) lo = mid; else hi = mid;
    }
    const uint64_t e = base + lane;
    if (e >= M) continue;
    uint32_t u = lo;
    while (__ldg(g.off + u + 1) <= e) ++u;
    const uint32_t w = __ldg(keys + e);
    if (w >= g.S) { atomicOr(asym, 1u); continue; }
    if (w == u) continue;                                      // a self-loop is its own mirror image
    if (w < u) { ++n_down; continue; }
    ++n_up;
    const uint64_t ub = __ldg(g.off + u), ue = __ldg(g.off + u + 1);
    if (e > ub && __ldg(keys + e - 1) == w) continue;          // looked up at the first entry of the run
    uint32_t mult = 1;
    while (e + mult < ue && __ldg(keys + e + mult) == w) ++mult;
    const uint64_t wb = __ldg(g.off + w);
    const uint32_t dw = (uint32_t)(__ldg(g.off + w + 1) - wb);
    uint32_t a = 0, b = dw;
    while (a < b) {
      const uint32_t mid = (a + b) >> 1;
      if (__ldg(keys + wb + mid) < u) a = mid + 1; else b = mid;
    }
    uint32_t c = 0;
    while (a + c < dw && __ldg(keys + wb + a + c) == u) ++c;
    if (c != mult) atomicOr(asym, 1u);
  }
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    n_up   += __shfl_xor_sync(NLP_FULL, n_up, k);
    n_down += __shfl_xor_sync(NLP_FULL, n_down, k);
  }
  if (lane == 0) {
    if (n_up)   atomicAdd(dir + 0, n_up);
    if (n_down) atomicAdd(dir + 1, n_down);
  }
}

// One thread per vertex w.
__global__ void __launch_bounds__(256) k_pair_rows(DevGraph g, uint32_t D, int rank, int world,
                                                   uint32_t* __restrict__ items, Counters* ctr) {
  unsigned long long t_first = 0, t_elig = 0, t_wedges = 0;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < g.S; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = g.deg[w];
    if (owns_row_block(w >> 5, rank, world)) t_first += d;     // first-hop entries of source w
    uint32_t it = 0;
    if (d == 1u) {   // eligible, but a single entry makes no wedge with v > u: count it here
      if (world <= 1 || owns_row_block((uint64_t)__ldg(g.keys + __ldg(g.off + w)) >> 5, rank, world)) { t_elig += 1; t_wedges += 1; }
    } else if (d >= 2u && d <= D) {
      it = d;
    }
    items[w] = it;
  }
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    t_first  += __shfl_xor_sync(NLP_FULL, t_first, k);
    t_elig   += __shfl_xor_sync(NLP_FULL, t_elig, k);
    t_wedges += __shfl_xor_sync(NLP_FULL, t_wedges, k);
  }
  if ((threadIdx.x & 31) == 0) {
    if (t_first)  atomicAdd(&ctr->first_hop, t_first);
    if (t_elig)   atomicAdd(&ctr->eligible_first_hop, t_elig);
    if (t_wedges) atomicAdd(&ctr->wedges, t_wedges);
  }
}

struct PairItems {
  uint32_t* u;              // [E] the source vertex this entry stands for
  uint32_t* cnt;            // [E] wedges it emits (0 when another rank owns u)
  uint32_t* dw;             // [E] deg(w) (term of the float measures)
  unsigned long long* ptr;  // [E] index into keys of its first second-hop entry
};

// One thread per eligible row w: fills the descriptors of its deg(w) items.
__global__ void __launch_bounds__(256) k_pair_items(DevGraph g, const uint32_t* __restrict__ items,
                                                    const unsigned long long* __restrict__ item_off, int rank, int world,
                                                    PairItems o, Counters* ctr) {
  const uint32_t* __restrict__ keys = g.keys;
  unsigned long long t_elig = 0, t_wedges = 0;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < g.S; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = items[w];
    if (!d) continue;
    const unsigned long long base = item_off[w];
    const uint64_t wb = __ldg(g.off + w);
    uint32_t j = 0;                                  // first entry > keys[wb + i] (rows are sorted multisets)
    for (uint32_t i = 0; i < d; ++i) {
      const uint32_t u = __ldg(keys + wb + i);
      if (j <= i) j = i + 1;
      while (j < d && __ldg(keys + wb + j) == u) ++j;
      uint32_t c = 0;
      if (owns_row_block(u >> 5, rank, world)) { c = d - j; t_elig += 1; t_wedges += d; }
      o.u[base + i] = u; o.cnt[base + i] = c; o.dw[base + i] = d; o.ptr[base + i] = wb + j;
    }
  }
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    t_elig   += __shfl_xor_sync(NLP_FULL, t_elig, k);
    t_wedges += __shfl_xor_sync(NLP_FULL, t_wedges, k);
  }
  if ((threadIdx.x & 31) == 0) {
    if (t_elig)   atomicAdd(&ctr->eligible_first_hop, t_elig);
    if (t_wedges) atomicAdd(&ctr->wedges, t_wedges);
  }
}

// One warp per 32 consecutive items; record k of the warp's tile lands at pair_off[first item] + k.
template <bool FLT>
__global__ void __launch_bounds__(256) k_pair_emit(const uint32_t* __restrict__ keys, uint64_t E, PairItems it,
                                                   const unsigned long long* __restrict__ pair_off, unsigned long long base,
                                                   uint32_t* __restrict__ pu, uint32_t* __restrict__ pv, uint32_t* __restrict__ pw) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t tile = warp0; tile * 32u < E; tile += nwarps) {
    const uint64_t e = tile * 32u + lane;
    uint32_t cnt = 0, u = 0, dw = 0;
    unsigned long long ptr = 0;
    if (e < E) { cnt = it.cnt[e]; u = it.u[e]; ptr = it.ptr[e]; if (FLT) dw = it.dw[e]; }
    const unsigned long long out0 = pair_off[tile * 32u] - base;   // base: first record of this launch's item range
    uint32_t inc = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
      if (lane >= d) inc += t;
    }
    const uint32_t tot = __shfl_sync(NLP_FULL, inc, 31);
    for (uint32_t sb = 0; sb < tot; sb += 32u) {
      const uint32_t idx = sb + lane;
      int j = 0;                                      // smallest j with inc[j] > idx
      #pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t x = __shfl_sync(NLP_FULL, inc, j + step - 1);
        if (x <= idx) j += step;
      }
      const uint32_t incj = __shfl_sync(NLP_FULL, inc, j);
      const uint32_t cntj = __shfl_sync(NLP_FULL, cnt, j);
      const uint32_t uj   = __shfl_sync(NLP_FULL, u, j);
      const unsigned long long ptrj = __shfl_sync(NLP_FULL, ptr, j);
      uint32_t dwj = 0;
      if (FLT) dwj = __shfl_sync(NLP_FULL, dw, j);
      if (idx < tot) {
        const uint32_t v = __ldg(keys + ptrj + (idx - (incj - cntj)));
        pu[out0 + idx] = uj; pv[out0 + idx] = v;
        if (FLT) pw[out0 + idx] = dwj;
      }
    }
  }
}

// Records sorted by (u, v): one thread per record, the first record of every run reduces the run.
// Output stays ALIGNED with the records: score_bits[i] = score of the pair whose run starts at
// record i, NLP_NO_SCORE everywhere else (inside runs, dropped pairs).  The kept pairs are thus
// still in ascending (u, v) order, which the ordered top-K (select.cuh) exploits: it only has
// to sort by score.
template <bool FLT>
__global__ void __launch_bounds__(256) k_pair_reduce(Params p, const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv,
                                                     const uint32_t* pw, uint64_t n, uint32_t* __restrict__ score_bits,
                                                     uint32_t* cnt_out = nullptr) {
  Tally tally;
  const uint64_t n32 = (n + 31u) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n32; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t u = 0, v = 0, cnt = 0;
    uint64_t du = 0;
    float acc = 0.0f;
    bool head = false;
    if (i < n) {
      u = pu[i]; v = pv[i];
      head = i == 0 || pu[i - 1] != u || pv[i - 1] != v;
    }
    if (head) {
      if (FLT) {   // inc/predict.hxx:788,828: acc = float(double(acc) + term), ascending w
        uint64_t j = i;
        do {
          acc = __double2float_rn(__dadd_rn((double)acc, flt_term(p, pw[j])));
          ++j;
        } while (j < n && pu[j] == u && pv[j] == v);
      } else {     // run length by galloping + binary search (the records are sorted)
        uint64_t lo = i, step = 1;
        while (lo + step < n && pu[lo + step] == u && pv[lo + step] == v) { lo += step; step <<= 1; }
        uint64_t hi = lo + step < n ? lo + step : n;          // first index known not to match (or n)
        while (lo + 1 < hi) {
          const uint64_t mid = lo + ((hi - lo) >> 1);
          if (pu[mid] == u && pv[mid] == v) lo = mid; else hi = mid;
        }
        cnt = (uint32_t)(hi - i);
      }
      const uint64_t ub = __ldg(p.g.off + u);
      du = __ldg(p.g.deg + u);
      // existing edges keep their candidate slot with value 0 (inc/predict.hxx:306-307)
      if (row_contains(p.g.keys, ub, (uint32_t)du, v)) { cnt = 0; acc = 0.0f; }
    }
    float score;
    const bool keep = score_pair(p, head, u, du, v, cnt, acc, tally, &score);
    if (i < n) {
      score_bits[i] = keep ? __float_as_uint(score) : NLP_NO_SCORE;
      // count measures, on request: the count after the exclusion, aligned like the scores
      // (0xffffffff = no pair starts here); may alias pw, which the count measures do not read
      if (!FLT && cnt_out) cnt_out[i] = head ? cnt : 0xffffffffu;
    }
  }
  // with cnt_out the caller scores these pairs again from the counts (k_score, reuse store) and the
  // candidate / kept counters are taken there -- counting here as well would count them twice
  if (FLT || !cnt_out) tally.flush(p.ctr);
}

}  // namespace nlp
