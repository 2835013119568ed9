// Kernels (b) (c) (d) of the north star: 2-hop wedge enumeration, common-neighbour
// accumulation and fused scoring, for the three source classes the frontier produces.
//
//   k_tiny<G,FLT>   G-lane sub-warp group per source (work <= G wedges): every lane holds one
//                   wedge, duplicates of v are found with match.any, no table at all
//   k_hash<FLT>     one team (= thread block) per source, open-addressing hash table in shared
//                   memory (1K / 4K / 16K slots of {key,value})
//   k_dense<FLT>    hub-heavy sources: one team per source on a span-sized dense table in
//                   global memory (per-block spill table) + touched list
//
// Replaces inc/predict.hxx:297-311 (clear, first-hop loop with hub cutoff, wedge scan,
// exclusion of u and N(u), scoring, min-score filter).  COUNT measures use integer atomics
// (exact in any order).  FLT measures (Adamic-Adar, resource allocation) reproduce the
// reference's float accumulation order -- ascending first-hop entry, each step
// acc = float(double(acc) + term) (inc/predict.hxx:788,828) -- bit for bit: one warp owns a
// source, the wedges of a batch sit in lanes in reference order, and the lowest lane of each
// match.any group folds its group's terms sequentially.
#pragma once
#include "common.cuh"

namespace nlp {

// ---------------------------------------------------------------------------------------------
// Accumulation term of the FLT measures (inc/predict.hxx:788, 828): a double.
__device__ __forceinline__ double flt_term(const Params& p, uint32_t dw) {
  return p.measure == M_AA ? __ldg(p.gtable + dw) : __ddiv_rn(1.0, (double)dw);
}

// First position in the sorted row [wb, wb+dw) whose key is > u.
__device__ __forceinline__ uint32_t skip_le(const uint32_t* __restrict__ keys, uint64_t wb, uint32_t dw, uint32_t u) {
  uint32_t lo = 0, hi = dw;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(keys + wb + mid) <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Stream all wedges (u, w, v) of source u through `sink.wedge(ok, v, g)`; every call is made by
// all 32 lanes of the calling warp (ok = lane holds a wedge with v > u).  The first-hop list is
// cut into chunks of 32 entries dealt round-robin to the team's warps; within a chunk the
// second-hop rows are packed back to back so lanes stay busy for short rows (LHub) and loads
// stay coalesced for long ones (IHub).  Order inside one warp = reference order.
template <bool FLT, class Sink>
__device__ __forceinline__ void stream_wedges(const Params& p, uint32_t u, uint64_t ub, uint32_t du,
                                              int warp_in_team, int team_warps, Sink& sink) {
  const int lane = threadIdx.x & 31;
  const uint32_t* __restrict__ keys = p.g.keys;
  for (uint64_t base = (uint64_t)warp_in_team * 32u; base < du; base += (uint64_t)team_warps * 32u) {
    const uint64_t i = base + lane;
    uint64_t wb = 0;
    uint32_t dw = 0;
    double g = 0.0;
    if (i < du) {
      const uint32_t w = __ldg(keys + ub + i);
      if (eligible(p, w)) {                               // hub cutoff, inc/predict.hxx:301
        wb = __ldg(p.g.off + w);
        dw = (uint32_t)(__ldg(p.g.off + w + 1) - wb);
        if (FLT && dw) g = flt_term(p, dw);
        if (dw > 32u) {                                   // sorted row: jump over v <= u
          const uint32_t s = skip_le(keys, wb, dw, u);
          wb += s; dw -= s;
        }
      }
    }
    if (!__any_sync(NLP_FULL, dw > (1u << 26))) {
      uint32_t inc = dw;                                  // inclusive scan of the row lengths
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
        if (lane >= d) inc += t;
      }
      const uint32_t tot = __shfl_sync(NLP_FULL, inc, 31);
      for (uint32_t sb = 0; sb < tot; sb += 32u) {
        const uint32_t idx = sb + lane;
        bool ok = idx < tot;
        int j = 0;                                        // smallest j with inc[j] > idx
        #pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const uint32_t x = __shfl_sync(NLP_FULL, inc, j + step - 1);
          if (x <= idx) j += step;
        }
        const uint32_t incj = __shfl_sync(NLP_FULL, inc, j);
        const uint32_t dwj  = __shfl_sync(NLP_FULL, dw, j);
        const uint64_t wbj  = __shfl_sync(NLP_FULL, wb, j);
        double gj = 0.0;
        if (FLT) gj = __shfl_sync(NLP_FULL, g, j);
        uint32_t v = 0;
        if (ok) { v = __ldg(keys + wbj + (idx - (incj - dwj))); ok = v > u; }   // inc/predict.hxx:292-296
        sink.wedge(ok, v, gj);
      }
    } else {   // a row longer than 2^26 entries: one row at a time (keeps the 32-bit scan exact)
      for (int j = 0; j < 32; ++j) {
        const uint32_t dwj = __shfl_sync(NLP_FULL, dw, j);
        const uint64_t wbj = __shfl_sync(NLP_FULL, wb, j);
        double gj = 0.0;
        if (FLT) gj = __shfl_sync(NLP_FULL, g, j);
        for (uint32_t k = 0; k < dwj; k += 32u) {
          const uint32_t kk = k + lane;
          bool ok = kk < dwj;
          uint32_t v = 0;
          if (ok) { v = __ldg(keys + wbj + kk); ok = v > u; }
          sink.wedge(ok, v, gj);
        }
      }
    }
  }
}

// Ordered fold of the FLT measures inside one warp: lanes hold wedges in reference order;
// `m` is the match.any group of this lane's v, `leader` its lowest lane.  Returns, in the
// leader lane, acc after adding the group's terms in lane order.
__device__ __forceinline__ float ordered_fold(float acc, bool mine, unsigned m, unsigned active, double g) {
  for (unsigned rest = active; rest; rest &= rest - 1u) {      // warp-uniform loop
    const int j = __ffs(rest) - 1;
    const double gj = __shfl_sync(NLP_FULL, g, j);
    if (mine && ((m >> j) & 1u)) acc = __double2float_rn(__dadd_rn((double)acc, gj));
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// Tiny path.  G lanes per source; requires work(u) <= G (all wedges of u fit the group).
template <int G, bool FLT>
__global__ void __launch_bounds__(256) k_tiny(Params p, const uint32_t* __restrict__ list, uint32_t n) {
  constexpr int NG = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gbase = lane - gl, grp = lane / G;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t sbase = warp0 * NG; sbase < n; sbase += nwarps * NG) {
    const uint64_t si = sbase + grp;
    const bool has = si < n;
    const uint32_t u = has ? __ldg(list + si) : 0u;
    uint64_t ub = 0;
    uint32_t du = 0;
    if (has) { ub = __ldg(p.g.off + u); du = (uint32_t)(__ldg(p.g.off + u + 1) - ub); }
    // assign wedge #gl of the source to lane gl
    bool mine = false;
    uint64_t my_ptr = 0;
    double my_g = 0.0;
    uint32_t filled = 0;
    const uint32_t nch = (du + G - 1) / G;
    const uint32_t maxch = __reduce_max_sync(NLP_FULL, nch);
    for (uint32_t c = 0; c < maxch; ++c) {
      const uint32_t i = c * G + gl;
      uint64_t wb = 0;
      uint32_t dw = 0;
      double g = 0.0;
      if (i < du) {
        const uint32_t w = __ldg(keys + ub + i);
        if (eligible(p, w)) {
          wb = __ldg(p.g.off + w);
          dw = (uint32_t)(__ldg(p.g.off + w + 1) - wb);
          if (FLT && dw) g = flt_term(p, dw);
        }
      }
      uint32_t inc = dw;
      #pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d, G);
        if (gl >= d) inc += t;
      }
      const uint32_t tot = __shfl_sync(NLP_FULL, inc, G - 1, G);
      const uint32_t rel = (uint32_t)gl - filled;                 // wraps when gl < filled
      const bool tgt = (uint32_t)gl >= filled && rel < tot;
      int j = 0;
      #pragma unroll
      for (int step = G / 2; step >= 1; step >>= 1) {
        const uint32_t x = __shfl_sync(NLP_FULL, inc, gbase + j + step - 1);
        if (tgt && x <= rel) j += step;
      }
      const uint32_t incj = __shfl_sync(NLP_FULL, inc, gbase + j);
      const uint32_t dwj  = __shfl_sync(NLP_FULL, dw, gbase + j);
      const uint64_t wbj  = __shfl_sync(NLP_FULL, wb, gbase + j);
      double gj = 0.0;
      if (FLT) gj = __shfl_sync(NLP_FULL, g, gbase + j);
      if (tgt) { mine = true; my_ptr = wbj + (rel - (incj - dwj)); my_g = gj; }
      filled += tot;                                              // <= work(u) <= G by binning
    }
    uint32_t v = 0;
    bool ok = false;
    if (mine) { v = __ldg(keys + my_ptr); ok = v > u; }
    // duplicates of v inside the group = common neighbours
    const unsigned long long mk = ok ? (((unsigned long long)(grp + 1) << 32) | v)
                                     : (0x8000000000000000ull | (unsigned)lane);
    const unsigned m = __match_any_sync(NLP_FULL, mk);
    const bool lead = ok && (__ffs(m) - 1) == lane;
    uint32_t cnt = __popc(m);
    float acc = 0.0f;
    if (FLT) {
      const unsigned active = __ballot_sync(NLP_FULL, ok);
      acc = ordered_fold(0.0f, lead, m, active, my_g);
    }
    // exclusion of existing edges (inc/predict.hxx:306-307): entries of N(u) keep their slot
    // but their value is zeroed, so they score 0
    if (lead) {
      uint32_t lo = 0, hi = du;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(keys + ub + mid) < v) lo = mid + 1; else hi = mid;
      }
      if (lo < du && __ldg(keys + ub + lo) == v) { cnt = 0; acc = 0.0f; }
    }
    score_and_emit(p, lead, u, du, v, cnt, acc, tally);
  }
  tally.flush(p.ctr);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory hash path.
struct HashSinkCount {
  uint2* slots; uint32_t mask; int shift;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double) {
    if (!ok) return;
    uint32_t s = hash32(v) >> shift;
    volatile uint32_t* kp = reinterpret_cast<volatile uint32_t*>(slots);
    while (true) {
      uint32_t k = kp[2 * s];
      if (k == NLP_EMPTY) k = atomicCAS(&slots[s].x, NLP_EMPTY, v);
      if (k == NLP_EMPTY || k == v) { atomicAdd(&slots[s].y, 1u); return; }
      s = (s + 1u) & mask;
    }
  }
};

struct HashSinkFlt {   // single warp owns the table
  uint2* slots; uint32_t mask; int shift;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double g) {
    const unsigned active = __ballot_sync(NLP_FULL, ok);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long mk = ok ? (unsigned long long)v : (0x8000000000000000ull | (unsigned)lane);
    const unsigned m = __match_any_sync(NLP_FULL, mk);
    const bool lead = ok && (__ffs(m) - 1) == lane;
    uint32_t s = 0;
    float acc = 0.0f;
    if (lead) {
      s = hash32(v) >> shift;
      volatile uint32_t* kp = reinterpret_cast<volatile uint32_t*>(slots);
      while (true) {
        uint32_t k = kp[2 * s];
        if (k == NLP_EMPTY) k = atomicCAS(&slots[s].x, NLP_EMPTY, v);
        if (k == NLP_EMPTY || k == v) break;
        s = (s + 1u) & mask;
      }
      acc = __uint_as_float(kp[2 * s + 1]);
    }
    acc = ordered_fold(acc, lead, m, active, g);
    if (lead) slots[s].y = __float_as_uint(acc);
    __syncwarp();
  }
};

// One team (= block) per source; blockDim = 32 * team_warps (team_warps = 1 for FLT).
template <bool FLT, bool ADMIT>
__global__ void k_hash(Params p, const uint32_t* __restrict__ list, uint32_t n, int bin,
                       uint32_t* __restrict__ deferred, int log2_slots) {
  extern __shared__ uint2 slots[];
  __shared__ unsigned long long s_next;
  __shared__ int s_go;
  __shared__ unsigned int s_emitted;
  const uint32_t nslots = 1u << log2_slots, mask = nslots - 1u;
  const int shift = 32 - log2_slots;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  for (uint32_t s = tid; s < nslots; s += blockDim.x) slots[s] = make_uint2(NLP_EMPTY, 0u);
  if (tid == 0) s_emitted = 0;
  __syncthreads();
  while (true) {
    if (tid == 0) s_next = atomicAdd(&p.ctr->queue[bin], 1ull);
    __syncthreads();
    const unsigned long long qi = s_next;
    if (qi >= n) break;
    const uint32_t u = __ldg(list + qi);
    uint32_t need = 0;
    if (ADMIT) {
      if (tid == 0) {
        const uint32_t w = p.work[u], room = p.g.S - 1u - u;
        need = w < room ? w : room;
        const unsigned long long r = atomicAdd(&p.ctr->reserved, (unsigned long long)need);
        if (r + need > p.cap) {
          atomicAdd(&p.ctr->reserved, 0ull - (unsigned long long)need);
          deferred[atomicAdd(&p.ctr->deferred[bin], 1ull)] = u;
          s_go = 0;
        } else s_go = 1;
      }
      __syncthreads();
      const int go = s_go;
      __syncthreads();
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    if (FLT) { HashSinkFlt sink{slots, mask, shift};   stream_wedges<true>(p, u, ub, du, warp, nw, sink); }
    else     { HashSinkCount sink{slots, mask, shift}; stream_wedges<false>(p, u, ub, du, warp, nw, sink); }
    __syncthreads();
    // zero the entries of N(u) (inc/predict.hxx:307); u itself is never inserted (v > u)
    for (uint32_t i = tid; i < du; i += blockDim.x) {
      const uint32_t v = __ldg(keys + ub + i);
      if (v <= u) continue;
      uint32_t s = hash32(v) >> shift;
      while (true) {
        const uint32_t k = slots[s].x;
        if (k == v) { slots[s].y = 0u; break; }
        if (k == NLP_EMPTY) break;
        s = (s + 1u) & mask;
      }
    }
    __syncthreads();
    // score every touched slot, append survivors, and leave the table empty for the next source
    uint32_t emitted = 0;
    for (uint32_t sb = warp * 32u; sb < nslots; sb += nw * 32u) {
      const uint2 e = slots[sb + lane];
      const bool has = e.x != NLP_EMPTY;
      if (has) slots[sb + lane] = make_uint2(NLP_EMPTY, 0u);
      emitted += score_and_emit(p, has, u, du, e.x, FLT ? 0u : e.y, FLT ? __uint_as_float(e.y) : 0.0f, tally);
    }
    if (ADMIT) {
      if (lane == 0 && emitted) atomicAdd(&s_emitted, emitted);
      __syncthreads();
      if (tid == 0) {
        atomicAdd(&p.ctr->reserved, (unsigned long long)s_emitted - (unsigned long long)need);
        s_emitted = 0;
      }
    }
    __syncthreads();
  }
  tally.flush(p.ctr);
}

// ---------------------------------------------------------------------------------------------
// Global dense spill path: table[v] for all v < S in this block's slice of HBM (L2-resident for
// small S), touched list beside it.  Values are reset by the scoring pass, so the table is all
// zero between sources.
struct DenseSinkCount {
  uint32_t* table; uint32_t* touched; unsigned int* s_cnt;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double) {
    bool first = false;
    if (ok) first = atomicAdd(table + v, 1u) == 0u;                // inc/predict.hxx:157-158
    const unsigned m = __ballot_sync(NLP_FULL, first);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(s_cnt, (unsigned)__popc(m));
    base = __shfl_sync(NLP_FULL, base, leader);
    if (first) touched[base + __popc(m & ((1u << lane) - 1u))] = v;
  }
};

struct DenseSinkFlt {   // single warp owns the table
  uint32_t* table; uint32_t* touched; unsigned int* s_cnt;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double g) {
    const unsigned active = __ballot_sync(NLP_FULL, ok);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long mk = ok ? (unsigned long long)v : (0x8000000000000000ull | (unsigned)lane);
    const unsigned m = __match_any_sync(NLP_FULL, mk);
    const bool lead = ok && (__ffs(m) - 1) == lane;
    float acc = 0.0f;
    bool first = false;
    if (lead) { acc = __uint_as_float(__ldcg(table + v)); first = !(acc != 0.0f); }   // inc/predict.hxx:176
    acc = ordered_fold(acc, lead, m, active, g);
    if (lead) __stcg(table + v, __float_as_uint(acc));
    const unsigned fm = __ballot_sync(NLP_FULL, first);
    if (fm) {
      unsigned int base = 0;
      const int leader = __ffs(fm) - 1;
      if (lane == leader) base = atomicAdd(s_cnt, (unsigned)__popc(fm));
      base = __shfl_sync(NLP_FULL, base, leader);
      if (first) touched[base + __popc(fm & ((1u << lane) - 1u))] = v;
    }
    __syncwarp();
  }
};

template <bool FLT, bool ADMIT>
__global__ void k_dense(Params p, const uint32_t* __restrict__ list, uint32_t n, int bin,
                        uint32_t* __restrict__ deferred, uint32_t* __restrict__ tables,
                        uint32_t* __restrict__ touched_all, uint64_t touched_cap) {
  __shared__ unsigned long long s_next;
  __shared__ int s_go;
  __shared__ unsigned int s_cnt, s_emitted;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  uint32_t* table = tables + (uint64_t)blockIdx.x * p.g.S;
  uint32_t* touched = touched_all + (uint64_t)blockIdx.x * touched_cap;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  if (tid == 0) { s_cnt = 0; s_emitted = 0; }
  __syncthreads();
  while (true) {
    if (tid == 0) s_next = atomicAdd(&p.ctr->queue[bin], 1ull);
    __syncthreads();
    const unsigned long long qi = s_next;
    if (qi >= n) break;
    const uint32_t u = __ldg(list + qi);
    uint32_t need = 0;
    if (ADMIT) {
      if (tid == 0) {
        const uint32_t w = p.work[u], room = p.g.S - 1u - u;
        need = w < room ? w : room;
        const unsigned long long r = atomicAdd(&p.ctr->reserved, (unsigned long long)need);
        if (r + need > p.cap) {
          atomicAdd(&p.ctr->reserved, 0ull - (unsigned long long)need);
          deferred[atomicAdd(&p.ctr->deferred[bin], 1ull)] = u;
          s_go = 0;
        } else s_go = 1;
      }
      __syncthreads();
      const int go = s_go;
      __syncthreads();
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    if (FLT) { DenseSinkFlt sink{table, touched, &s_cnt};   stream_wedges<true>(p, u, ub, du, warp, nw, sink); }
    else     { DenseSinkCount sink{table, touched, &s_cnt}; stream_wedges<false>(p, u, ub, du, warp, nw, sink); }
    __syncthreads();
    for (uint32_t i = tid; i < du; i += blockDim.x)                 // inc/predict.hxx:307
      __stcg(table + __ldg(keys + ub + i), 0u);
    __syncthreads();
    const uint32_t cnt = s_cnt;
    uint32_t emitted = 0;
    for (uint32_t sb = warp * 32u; sb < cnt; sb += nw * 32u) {
      const uint32_t i = sb + lane;
      const bool has = i < cnt;
      uint32_t v = 0, val = 0;
      if (has) { v = touched[i]; val = __ldcg(table + v); __stcg(table + v, 0u); }
      emitted += score_and_emit(p, has, u, du, v, FLT ? 0u : val, FLT ? __uint_as_float(val) : 0.0f, tally);
    }
    __syncthreads();
    if (tid == 0) s_cnt = 0;
    if (ADMIT) {
      if (lane == 0 && emitted) atomicAdd(&s_emitted, emitted);
      __syncthreads();
      if (tid == 0) {
        atomicAdd(&p.ctr->reserved, (unsigned long long)s_emitted - (unsigned long long)need);
        s_emitted = 0;
      }
    }
    __syncthreads();
  }
  tally.flush(p.ctr);
}

}  // namespace nlp
