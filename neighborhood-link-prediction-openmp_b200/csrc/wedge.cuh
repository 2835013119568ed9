// Kernels (b) (c) (d) of the north star: 2-hop wedge enumeration, common-neighbour
// accumulation and fused scoring, for the source classes the frontier produces.
//
//   k_tiny<G,FLT>   G-lane sub-warp group per source (work <= G wedges): every lane holds one
//                   wedge, duplicates of v are found with match.any, no table at all
//   k_hash<FLT>     one team (= thread block) per source, open-addressing hash table in shared
//                   memory (1K / 4K / 16K slots of {key,value}) + list of the claimed slots
//   k_dense<FLT>    hub-heavy sources: one team per source on a span-sized dense table in
//                   global memory (per-block spill table) + touched list
//
// Replaces inc/predict.hxx:297-311 (clear, first-hop loop with hub cutoff, wedge scan,
// exclusion of u and N(u), scoring, min-score filter).  The first-hop lists these kernels read
// are the eligible lists the frontier pass compacted (LHub) or the rows themselves (IHub).
// COUNT measures use integer atomics (exact in any order).  FLT measures (Adamic-Adar, resource
// allocation) reproduce the reference's float accumulation order -- ascending first-hop entry,
// each step acc = float(double(acc) + term) (inc/predict.hxx:788,828) -- bit for bit: one warp
// owns a source, the wedges of a batch sit in lanes in reference order, and the lowest lane of
// each match.any group folds its group's terms sequentially.
#pragma once
#include <cooperative_groups.h>
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

// Is v an entry of the sorted row [ub, ub+du)?
__device__ __forceinline__ bool row_contains(const uint32_t* __restrict__ keys, uint64_t ub, uint32_t du, uint32_t v) {
  uint32_t lo = 0, hi = du;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(keys + ub + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo < du && __ldg(keys + ub + lo) == v;
}

// The (compacted) first-hop list of a source, as a sequence of pieces.
struct FirstHop {
  const uint32_t* base;      // piece c starts at base + c * CHUNK
  uint32_t npieces;
  uint32_t single_count;     // entries of the only piece when npieces == 1
  const uint32_t* piece_cnt; // per-piece counts when npieces > 1
};

__device__ __forceinline__ FirstHop first_hop(const Params& p, uint32_t u, uint64_t ub, uint32_t du) {
  FirstHop f;
  if (!p.ekeys) { f.base = p.g.keys + ub; f.npieces = 1; f.single_count = du; f.piece_cnt = nullptr; }
  else if (du <= LONG_ROW) { f.base = p.ekeys + ub; f.npieces = 1; f.single_count = __ldg(p.ecount + u); f.piece_cnt = nullptr; }
  else { f.base = p.ekeys + ub; f.npieces = (du + CHUNK - 1) / CHUNK; f.single_count = 0; f.piece_cnt = p.chunk_cnt + __ldg(p.chunk_base + u); }
  return f;
}

// Stream the wedges behind 32 consecutive first-hop entries (lane i holds entry i of the chunk,
// or nothing) through `sink.wedge(ok, v, g)`; every call is made by all 32 lanes (ok = lane
// holds a wedge with v > u).  The second-hop rows are packed back to back so lanes stay busy for
// short rows (LHub) and loads stay coalesced for long ones (IHub); order = reference order.
template <bool FLT, class Sink>
__device__ __forceinline__ void stream_chunk(const Params& p, uint32_t u, bool has, uint32_t w, Sink& sink) {
  const int lane = threadIdx.x & 31;
  const uint32_t* __restrict__ keys = p.g.keys;
  uint64_t wb = 0;
  uint32_t dw = 0;
  double g = 0.0;
  if (has) {
    wb = __ldg(p.g.off + w);
    dw = (uint32_t)(__ldg(p.g.off + w + 1) - wb);
    if (FLT && dw) g = flt_term(p, dw);
    if (dw > 32u) {                                   // sorted row: jump over v <= u
      const uint32_t s = skip_le(keys, wb, dw, u);
      wb += s; dw -= s;
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

// All wedges of source u for one team: a single-piece list is cut into 32-entry chunks dealt
// round-robin to the team's warps; a multi-piece (long-row) list deals whole pieces.
template <bool FLT, class Sink>
__device__ __forceinline__ void stream_wedges(const Params& p, uint32_t u, uint64_t ub, uint32_t du,
                                              int warp_in_team, int team_warps, Sink& sink) {
  const int lane = threadIdx.x & 31;
  const FirstHop f = first_hop(p, u, ub, du);
  if (f.npieces == 1) {
    for (uint32_t base = (uint32_t)warp_in_team * 32u; base < f.single_count; base += (uint32_t)team_warps * 32u) {
      const uint32_t i = base + lane;
      const bool has = i < f.single_count;
      const uint32_t w = has ? __ldg(f.base + i) : 0u;
      stream_chunk<FLT>(p, u, has, w, sink);
    }
  } else {
    for (uint32_t c = warp_in_team; c < f.npieces; c += team_warps) {
      const uint32_t cnt = __ldg(f.piece_cnt + c);
      const uint32_t* pb = f.base + (uint64_t)c * CHUNK;
      for (uint32_t base = 0; base < cnt; base += 32u) {
        const uint32_t i = base + lane;
        const bool has = i < cnt;
        const uint32_t w = has ? __ldg(pb + i) : 0u;
        stream_chunk<FLT>(p, u, has, w, sink);
      }
    }
  }
}

// Block-wide inclusive scan of one u32 per thread into s_inc[0 .. blockDim.x) (blockDim.x a
// multiple of 32, at most 1024; s_wsum holds 32 words).
__device__ __forceinline__ void block_scan_u32(uint32_t x, uint32_t* s_inc, uint32_t* s_wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t inc = x;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < nw ? s_wsum[lane] : 0u;
    uint32_t winc = w;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(NLP_FULL, winc, d);
      if (lane >= d) winc += t;
    }
    s_wsum[lane] = winc - w;
  }
  __syncthreads();
  s_inc[threadIdx.x] = inc + s_wsum[warp];
  __syncthreads();
}

// Count measures, block-cooperative: thread t of the team takes first-hop entry t of a batch, the
// second-hop rows of the whole batch are laid end to end (block scan) and dealt to ALL threads, so
// a hub row inside the batch is shared by the team instead of stalling the one warp that drew it.
// Order does not matter for integer counts.  Requires blockDim.x * maxdeg < 2^32 (host checks).
template <class Sink>
__device__ __forceinline__ void stream_wedges_block(const Params& p, uint32_t u, uint64_t ub, uint32_t du, Sink& sink,
                                                    uint32_t* s_inc, unsigned long long* s_wb, uint32_t* s_wsum) {
  const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
  const uint32_t* __restrict__ keys = p.g.keys;
  const FirstHop f = first_hop(p, u, ub, du);
  for (uint32_t c = 0; c < f.npieces; ++c) {
    const uint32_t pc = f.npieces == 1 ? f.single_count : __ldg(f.piece_cnt + c);
    const uint32_t* pb = f.base + (uint64_t)c * CHUNK;
    for (uint32_t base = 0; base < pc; base += nt) {
      const uint32_t i = base + tid;
      uint64_t wb = 0;
      uint32_t dw = 0;
      if (i < pc) {
        const uint32_t w = __ldg(pb + i);
        wb = __ldg(p.g.off + w);
        dw = (uint32_t)(__ldg(p.g.off + w + 1) - wb);
        if (dw > 32u) {                                 // sorted row: jump over v <= u
          const uint32_t sk = skip_le(keys, wb, dw, u);
          wb += sk; dw -= sk;
        }
      }
      s_wb[tid] = wb;
      block_scan_u32(dw, s_inc, s_wsum);
      const uint32_t tot = s_inc[nt - 1];
      for (uint32_t b2 = (uint32_t)(tid & ~31); b2 < tot; b2 += nt) {      // warp-uniform trip count
        const uint32_t idx = b2 + lane;
        bool ok = idx < tot;
        uint32_t v = 0;
        if (ok) {
          uint32_t lo = 0, hi = nt - 1;                 // smallest j with s_inc[j] > idx
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_inc[mid] <= idx) lo = mid + 1; else hi = mid;
          }
          const uint32_t before = lo ? s_inc[lo - 1] : 0u;
          v = __ldg(keys + s_wb[lo] + (idx - before));
          ok = v > u;                                   // inc/predict.hxx:292-296
        }
        sink.wedge(ok, v, 0.0);
      }
      __syncthreads();
    }
  }
}

// Ordered fold of the FLT measures inside one warp: lanes hold wedges in reference order;
// `m` is the match.any group of this lane's v, `mine` marks its lowest lane, `dups` the lanes
// that sit in a group of two or more.  Returns, in the leader lane, acc after adding the group's
// terms in lane order.
__device__ __forceinline__ float ordered_fold(float acc, bool mine, unsigned m, unsigned dups, double g) {
  const int lane = threadIdx.x & 31;
  if (mine && !((dups >> lane) & 1u)) acc = __double2float_rn(__dadd_rn((double)acc, g));   // singleton group
  for (unsigned rest = dups; rest; rest &= rest - 1u) {      // warp-uniform loop over duplicated lanes
    const int j = __ffs(rest) - 1;
    const double gj = __shfl_sync(NLP_FULL, g, j);
    if (mine && ((m >> j) & 1u)) acc = __double2float_rn(__dadd_rn((double)acc, gj));
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// Tiny path.  G lanes per source; requires work(u) <= G (all wedges of u fit the group) and a
// single-piece first-hop list.
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
    uint32_t du = 0, fc = 0;
    const uint32_t* fb = keys;
    if (has) {
      ub = __ldg(p.g.off + u); du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
      if (p.ekeys) { fb = p.ekeys + ub; fc = __ldg(p.ecount + u); } else { fb = keys + ub; fc = du; }
    }
    // assign wedge #gl of the source to lane gl
    bool mine = false;
    uint64_t my_ptr = 0;
    double my_g = 0.0;
    uint32_t filled = 0;
    const uint32_t nch = (fc + G - 1) / G;
    const uint32_t maxch = __reduce_max_sync(NLP_FULL, nch);
    for (uint32_t c = 0; c < maxch; ++c) {
      const uint32_t i = c * G + gl;
      uint64_t wb = 0;
      uint32_t dw = 0;
      double g = 0.0;
      if (i < fc) {
        const uint32_t w = __ldg(fb + i);
        wb = __ldg(p.g.off + w);
        dw = (uint32_t)(__ldg(p.g.off + w + 1) - wb);
        if (FLT && dw) g = flt_term(p, dw);
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
      const unsigned dups = __ballot_sync(NLP_FULL, ok && cnt > 1);
      acc = ordered_fold(0.0f, lead, m, dups, my_g);
    }
    // exclusion of existing edges (inc/predict.hxx:306-307): entries of N(u) keep their slot
    // but their value is zeroed, so they score 0
    if (lead && row_contains(keys, ub, du, v)) { cnt = 0; acc = 0.0f; }
    score_and_emit(p, lead, u, du, v, cnt, acc, tally);
  }
  tally.flush(p.ctr);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory hash path.  Slots are {key, value}; the slots a source claims are also pushed
// on a list so scoring touches only those.
struct HashTable {
  uint2* slots; uint16_t* list; unsigned int* count; uint32_t mask; int shift;
  __device__ __forceinline__ void note_claims(bool claimed, uint32_t s) {
    const unsigned cm = __ballot_sync(NLP_FULL, claimed);
    if (!cm) return;
    const int lane = threadIdx.x & 31, leader = __ffs(cm) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(cm));
    base = __shfl_sync(NLP_FULL, base, leader);
    if (claimed) list[base + __popc(cm & ((1u << lane) - 1u))] = (uint16_t)s;
  }
};

struct HashSinkCount {
  HashTable t;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double) {
    bool claimed = false;
    uint32_t s = 0;
    if (ok) {
      s = hash32(v) >> t.shift;
      volatile uint32_t* kp = reinterpret_cast<volatile uint32_t*>(t.slots);
      while (true) {
        uint32_t k = kp[2 * s];
        if (k == NLP_EMPTY) { k = atomicCAS(&t.slots[s].x, NLP_EMPTY, v); claimed = k == NLP_EMPTY; }
        if (k == NLP_EMPTY || k == v) { atomicAdd(&t.slots[s].y, 1u); break; }
        s = (s + 1u) & t.mask;
      }
    }
    t.note_claims(claimed, s);
  }
};

struct HashSinkFlt {   // single warp owns the table
  HashTable t;
  __device__ __forceinline__ void wedge(bool ok, uint32_t v, double g) {
    const unsigned active = __ballot_sync(NLP_FULL, ok);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long mk = ok ? (unsigned long long)v : (0x8000000000000000ull | (unsigned)lane);
    const unsigned m = __match_any_sync(NLP_FULL, mk);
    const bool lead = ok && (__ffs(m) - 1) == lane;
    const unsigned dups = __ballot_sync(NLP_FULL, ok && __popc(m) > 1);
    uint32_t s = 0;
    float acc = 0.0f;
    bool claimed = false;
    if (lead) {
      s = hash32(v) >> t.shift;
      volatile uint32_t* kp = reinterpret_cast<volatile uint32_t*>(t.slots);
      while (true) {
        uint32_t k = kp[2 * s];
        if (k == NLP_EMPTY) { k = atomicCAS(&t.slots[s].x, NLP_EMPTY, v); claimed = k == NLP_EMPTY; }
        if (k == NLP_EMPTY || k == v) break;
        s = (s + 1u) & t.mask;
      }
      acc = __uint_as_float(kp[2 * s + 1]);
    }
    acc = ordered_fold(acc, lead, m, dups, g);
    if (lead) t.slots[s].y = __float_as_uint(acc);
    t.note_claims(claimed, s);
    __syncwarp();
  }
};

// Admission control shared by the persistent kernels (only when the candidate buffer cannot
// hold every candidate): reserve the source's bound or push it to the next pass.
__device__ __forceinline__ bool admit_source(const Params& p, uint32_t u, int bin, uint32_t* deferred, uint32_t* need_out) {
  const uint32_t w = p.work[u], room = p.g.S - 1u - u;
  const uint32_t need = w < room ? w : room;
  *need_out = need;
  for (;;) {
    // Stop admitting once this pass has written enough: the host then cuts the buffer to the
    // best K, which tightens the pruning threshold for everything that follows.  (The buffer
    // itself is much larger: its size bounds the reservations of the sources in flight.)
    if (*reinterpret_cast<volatile unsigned long long*>(&p.ctr->cursor) >= p.soft_cap) break;
    // optimistic reservation: one atomic in the common case
    const unsigned long long r = atomicAdd(&p.ctr->reserved, (unsigned long long)need);
    if (r + need <= p.cap) return true;
    atomicAdd(&p.ctr->reserved, 0ull - (unsigned long long)need);
    // No room because of what is already WRITTEN -> next pass.  No room only because of the
    // (pessimistic: bound, not count) reservations of sources still in flight -> wait for them;
    // every in-flight source belongs to a resident block that never waits itself, so this ends.
    // Waiting blocks only READ the counter (a transient reservation of every waiter would keep
    // all of them out for ever) and retry when it shows room.
    for (;;) {
      const unsigned long long written = *reinterpret_cast<volatile unsigned long long*>(&p.ctr->cursor);
      if (written + need > p.cap) goto defer;
      if (*reinterpret_cast<volatile unsigned long long*>(&p.ctr->reserved) + need <= p.cap) break;
      __nanosleep(2000);
    }
  }
defer:
  deferred[atomicAdd(&p.ctr->deferred[bin], 1ull)] = u;
  return false;
}

// One team (= block) per source; blockDim = 32 * team_warps (team_warps = 1 for FLT).
// Dynamic shared memory: slots[2^log2_slots] then the claimed-slot list (3/4 of that, u16).
template <bool FLT, bool ADMIT>
__global__ void k_hash(Params p, const uint32_t* __restrict__ list, uint32_t n, int bin,
                       uint32_t* __restrict__ deferred, int log2_slots) {
  extern __shared__ uint2 slots[];
  __shared__ unsigned long long s_wb[FLT ? 32 : 512];
  __shared__ uint32_t s_inc[FLT ? 32 : 512];
  __shared__ uint32_t s_wsum[32];
  __shared__ int s_go;
  __shared__ unsigned int s_emitted, s_count;
  const uint32_t nslots = 1u << log2_slots, mask = nslots - 1u;
  uint16_t* slist = reinterpret_cast<uint16_t*>(slots + nslots);
  const int shift = 32 - log2_slots;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  for (uint32_t s = tid; s < nslots; s += blockDim.x) slots[s] = make_uint2(NLP_EMPTY, 0u);
  if (tid == 0) { s_emitted = 0; s_count = 0; }
  __syncthreads();
  for (uint32_t qi = blockIdx.x; qi < n; qi += gridDim.x) {
    const uint32_t u = __ldg(list + qi);
    uint32_t need = 0;
    if (ADMIT) {
      if (tid == 0) s_go = admit_source(p, u, bin, deferred, &need) ? 1 : 0;
      __syncthreads();
      const int go = s_go;
      __syncthreads();
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    const HashTable table{slots, slist, &s_count, mask, shift};
    if (FLT) { HashSinkFlt sink{table};   stream_wedges<true>(p, u, ub, du, warp, nw, sink); }
    else if (p.coop) { HashSinkCount sink{table}; stream_wedges_block(p, u, ub, du, sink, s_inc, s_wb, s_wsum); }
    else     { HashSinkCount sink{table}; stream_wedges<false>(p, u, ub, du, warp, nw, sink); }
    __syncthreads();
    const uint32_t nt = s_count;
    // exclusion of N(u) (inc/predict.hxx:307; u itself is never inserted since v > u): either
    // zero the table entries of the row, or -- when the row is much longer than the table is
    // full -- look every touched key up in the sorted row while scoring
    const bool zero_pass = du <= 4u * nt + 64u;
    if (zero_pass) {
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
    }
    // score every touched slot, append survivors, and leave the table empty for the next source
    uint32_t emitted = 0;
    for (uint32_t sb = warp * 32u; sb < nt; sb += nw * 32u) {
      const uint32_t i = sb + lane;
      const bool has = i < nt;
      uint2 e = make_uint2(0u, 0u);
      if (has) {
        const uint32_t s = slist[i];
        e = slots[s];
        slots[s] = make_uint2(NLP_EMPTY, 0u);
        if (!zero_pass && row_contains(keys, ub, du, e.x)) e.y = 0u;
      }
      emitted += score_and_emit(p, has, u, du, e.x, FLT ? 0u : e.y, FLT ? __uint_as_float(e.y) : 0.0f, tally);
    }
    __syncthreads();
    if (tid == 0) s_count = 0;
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
    const unsigned dups = __ballot_sync(NLP_FULL, ok && __popc(m) > 1);
    float acc = 0.0f;
    bool first = false;
    if (lead) { acc = __uint_as_float(__ldcg(table + v)); first = !(acc != 0.0f); }   // inc/predict.hxx:176
    acc = ordered_fold(acc, lead, m, dups, g);
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
  __shared__ unsigned long long s_wb[FLT ? 32 : 512];
  __shared__ uint32_t s_inc[FLT ? 32 : 512];
  __shared__ uint32_t s_wsum[32];
  __shared__ int s_go;
  __shared__ unsigned int s_cnt, s_emitted;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  uint32_t* table = tables + (uint64_t)blockIdx.x * p.g.S;
  uint32_t* touched = touched_all + (uint64_t)blockIdx.x * touched_cap;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  if (tid == 0) { s_cnt = 0; s_emitted = 0; }
  __syncthreads();
  for (uint32_t qi = blockIdx.x; qi < n; qi += gridDim.x) {
    const uint32_t u = __ldg(list + qi);
    uint32_t need = 0;
    if (ADMIT) {
      if (tid == 0) s_go = admit_source(p, u, bin, deferred, &need) ? 1 : 0;
      __syncthreads();
      const int go = s_go;
      __syncthreads();
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    if (FLT) { DenseSinkFlt sink{table, touched, &s_cnt};   stream_wedges<true>(p, u, ub, du, warp, nw, sink); }
    else if (p.coop) { DenseSinkCount sink{table, touched, &s_cnt}; stream_wedges_block(p, u, ub, du, sink, s_inc, s_wb, s_wsum); }
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

// ---------------------------------------------------------------------------------------------
// Range path (count measures, hub-heavy sources): the candidate set of such a source is a large
// fraction of the vertex set (1.6e5 of 2.6e5 vertices at R-MAT 18), so instead of hashing, the
// block keeps DIRECT-ADDRESSED u32 counters in shared memory for a window [lo, lo + C) of v and
// walks the windows lo = u+1, u+1+C, ...  Rows are sorted, so every pass reads only the part of
// each second-hop row that falls in the window (two binary searches per row and pass); counting is
// one shared-memory atomic per wedge, nothing leaves the SM until scoring.  Replaces the
// reference's per-thread dense |V| array (inc/predict.hxx:117-122, 157-158), which the global
// spill tables of k_dense emulate in HBM at the price of one 32-byte sector RMW per wedge.
// Counter value 0 = untouched; RANGE_ZEROED = touched, then zeroed because v is in N(u)
// (inc/predict.hxx:306-307: such pairs stay candidates with value 0).
constexpr uint32_t RANGE_ZEROED = 0x80000000u;
enum { RANGE_THREADS = 1024, RANGE_RUN = 8 };

__device__ __forceinline__ uint32_t lower_bound_row(const uint32_t* __restrict__ keys, uint64_t b, uint32_t d, uint32_t x) {
  uint32_t lo = 0, hi = d;                            // first position with key >= x
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(keys + b + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// First position in [a, e) (absolute indices into keys, sorted row) whose key is >= x: gallop from
// a, then bisect.  Window after window the cursor only moves forward, so a step costs O(log of
// the segment it skips) dependent loads instead of O(log of the row).  KNOWN_LESS: the caller
// already knows keys[a] < x (and a < e).
template <bool KNOWN_LESS>
__device__ __forceinline__ unsigned long long gallop_to(const uint32_t* __restrict__ keys, unsigned long long a,
                                                        unsigned long long e, uint32_t x) {
  if (!KNOWN_LESS && (a >= e || __ldg(keys + a) >= x)) return a;
  unsigned long long lo = a, step = 1;                 // keys[lo] < x
  while (lo + step < e && __ldg(keys + lo + step) < x) { lo += step; step <<= 1; }
  unsigned long long hi = lo + step < e ? lo + step : e;   // keys[hi] >= x, or hi == e
  while (lo + 1 < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (__ldg(keys + mid) < x) lo = mid; else hi = mid;
  }
  return hi;
}

// PACK counters per 32-bit word (1, 2 or 4; see k_range): the fields cannot carry into each other
// because a count never reaches the field's top bit, which is the "touched, then zeroed" mark.
template <int PACK> struct RangeField {
  static constexpr uint32_t BITS = 32u / PACK, SHIFT = PACK == 4 ? 2u : (PACK == 2 ? 1u : 0u);
  static constexpr uint32_t MASK = PACK == 1 ? 0xffffffffu : ((1u << BITS) - 1u);
  static constexpr uint32_t MARK = 1u << (BITS - 1u), VALUE = MARK - 1u;
};

template <int PACK>
__device__ __forceinline__ bool range_count(uint32_t* cnt, uint32_t x) {   // true: first wedge that reaches this vertex
  typedef RangeField<PACK> F;
  if (PACK == 1) return atomicAdd(cnt + x, 1u) == 0u;   // inc/predict.hxx:156-158
  const uint32_t sh = (x & (uint32_t)(PACK - 1)) * F::BITS;
  return ((atomicAdd(cnt + (x >> F::SHIFT), 1u << sh) >> sh) & F::MASK) == 0u;
}

// Wedges behind up to RANGE_THREADS first-hop entries (thread t holds entry t) whose v lies in
// [vlo, vhi): counter[v - wbase] += 1.  The window parts of all rows are laid end to end and dealt
// to the threads of the WHOLE block, so one hub row among the entries is shared by 1024 threads
// instead of stalling the one warp that drew it (ncu: 54 % of the first version's stall samples
// sat at the barrier behind such warps).
//
// Windows are ALIGNED to multiples of the cell size C (window j of a source that packs PACK
// counters per word covers the cells [j * PACK, (j + 1) * PACK)), so where the part of a row in a
// window begins and ends does not depend on the source.  Long rows carry a fence table, built once
// per graph (k_fence_fill): fence[c] = first position of the row whose key is >= c * C.  A visit
// of such a row costs two independent loads.  At R-MAT 24 a hub-heavy source walks 80-320 windows
// and the hub rows hold most of the wedges: the gallop of the first version (22 dependent loads
// per visit of a 4e5-entry row) was the critical path of every window.
// Short rows keep a 16-byte record in rec[] (position, entries left, key at the position), carried
// from window to window: the first window finds the first key > u by bisection, later windows
// read the record and know from the cached key whether the row has anything in the window at all.
constexpr uint32_t RANGE_FENCED = 0x80000000u;

__device__ __forceinline__ uint4 range_pack(unsigned long long pos, unsigned long long e, uint32_t next) {
  return make_uint4((uint32_t)pos, (uint32_t)(pos >> 32), (uint32_t)(e - pos), next);
}

struct RangeFences {
  const uint32_t* slot_of;   // [S] fence-table slot of a row, 0xffffffff = none (null: no tables at all)
  const uint32_t* fence;     // [slots][ncell + 1]
  uint32_t ncell;
};

template <int PACK>
__device__ __forceinline__ void range_batch(const Params& p, const RangeFences& fx, bool has, bool first, uint32_t w,
                                            uint32_t u, uint32_t cell0, uint32_t wbase, uint32_t vhi,
                                            uint4* rec, uint32_t* cnt, uint32_t* touched, uint32_t* s_tn,
                                            uint32_t* s_inc, unsigned long long* s_wb, uint32_t* s_wsum) {
  const uint32_t* __restrict__ keys = p.g.keys;
  unsigned long long a = 0;
  uint32_t dw = 0;
  if (has) {
    const uint32_t cell1 = cell0 + (uint32_t)PACK < fx.ncell ? cell0 + (uint32_t)PACK : fx.ncell;
    if (first) {
      const unsigned long long wb = __ldg(p.g.off + w);
      const unsigned long long e = __ldg(p.g.off + w + 1);
      const uint32_t fs = fx.slot_of ? __ldg(fx.slot_of + w) : 0xffffffffu;
      if (fs != 0xffffffffu) {
        const uint32_t* f = fx.fence + (uint64_t)fs * (fx.ncell + 1u);
        const uint32_t s0 = __ldg(f + cell0), s1 = __ldg(f + cell1);
        a = wb + s0 + lower_bound_row(keys, wb + s0, s1 - s0, u + 1u);
        dw = (uint32_t)(wb + s1 - a);
        *rec = make_uint4((uint32_t)wb, (uint32_t)(wb >> 32), RANGE_FENCED | fs, 0u);
      } else {
        a = wb + lower_bound_row(keys, wb, (uint32_t)(e - wb), u + 1u);
        uint32_t next = a < e ? __ldg(keys + a) : 0xffffffffu;
        unsigned long long b = a;
        if (next < vhi) {
          b = vhi >= p.g.S ? e : gallop_to<true>(keys, a, e, vhi);
          next = b < e ? __ldg(keys + b) : 0xffffffffu;
        }
        *rec = range_pack(b, e, next);
        dw = (uint32_t)(b - a);
      }
    } else {
      const uint4 r = *rec;
      const unsigned long long pos = (unsigned long long)r.x | ((unsigned long long)r.y << 32);
      if (r.z & RANGE_FENCED) {
        const uint32_t* f = fx.fence + (uint64_t)(r.z & ~RANGE_FENCED) * (fx.ncell + 1u);
        const uint32_t s0 = __ldg(f + cell0), s1 = __ldg(f + cell1);
        a = pos + s0;
        dw = s1 - s0;
      } else {
        a = pos;
        if (r.w < vhi) {                                // the row has entries in this window (so r.z > 0)
          const unsigned long long e = a + r.z;
          const unsigned long long b = vhi >= p.g.S ? e : gallop_to<true>(keys, a, e, vhi);
          *rec = range_pack(b, e, b < e ? __ldg(keys + b) : 0xffffffffu);
          dw = (uint32_t)(b - a);
        }
      }
    }
  }
  s_wb[threadIdx.x] = a;
  block_scan_u32(dw, s_inc, s_wsum);                  // k_range is only used when 1024 * maxdeg < 2^32
  // Every thread takes RANGE_RUN consecutive wedges of the concatenation: ONE bisection finds the
  // row of the first, the others follow by walking s_inc forward.  (A bisection per wedge, the
  // first version, made this loop ~100 instructions per wedge and the whole kernel issue bound:
  // ncu, R-MAT 18 IHub, 69 % issue slots busy at 360 warp instructions per atomic instruction.)
  // Two variants measured slower at R-MAT 22 IHub and were dropped (profiles/r02_summary.md): one
  // warp-aggregated append to touched[] per step instead of an atomic per first touch (+8 %), and
  // a separate warp-coalesced pass over the long row parts (+43 %: two phases, two tails); a fast
  // path for runs that lie in one row (no per-wedge row stepping) gained nothing either (+7 %).
  const uint32_t tot = s_inc[RANGE_THREADS - 1];
  for (uint32_t c0 = threadIdx.x * (uint32_t)RANGE_RUN; c0 < tot; c0 += RANGE_THREADS * (uint32_t)RANGE_RUN) {
    uint32_t lo = 0, hi = RANGE_THREADS - 1;          // smallest j with s_inc[j] > c0
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_inc[mid] <= c0) lo = mid + 1; else hi = mid;
    }
    uint32_t row = lo;
    uint32_t rend = s_inc[row];                        // wedges [.., rend) belong to rows <= row
    unsigned long long base = s_wb[row] - (row ? s_inc[row - 1] : 0u);   // key of wedge idx = keys[base + idx]
    #pragma unroll
    for (int h = 0; h < RANGE_RUN; h += 4) {           // four addresses, four loads in flight, four atomics
      unsigned long long addr[4];
      #pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t idx = c0 + (uint32_t)(h + k);
        addr[k] = 0;
        if (idx < tot) {
          while (idx >= rend) { ++row; base = s_wb[row] - rend; rend = s_inc[row]; }   // also steps over empty rows
          addr[k] = base + idx;
        }
      }
      uint32_t v[4];
      #pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (c0 + (uint32_t)(h + k) < tot) ? __ldg(keys + addr[k]) : 0u;
      #pragma unroll
      for (int k = 0; k < 4; ++k)     // the first wedge that reaches a vertex also lists it for the scoring phase
        if (c0 + (uint32_t)(h + k) < tot && range_count<PACK>(cnt, v[k] - wbase)) touched[atomicAdd(s_tn, 1u)] = v[k];
    }
  }
  __syncthreads();
}

// All windows of one source.  PACK = 2: 16-bit counters, two per word, so a window spans 2 * C
// vertices and the source needs half the windows (half the (row, window) visits); PACK = 4: 8-bit
// counters, a quarter of the windows.  Valid while no count can reach the field's top bit: a count
// is at most deg(u) x (largest multiplicity of an entry in a row), so the host allows PACK = 2 for
// deg(u) < 2^15 / that multiplicity and PACK = 4 for deg(u) < 2^7 / that multiplicity
// (Params::range_half, range_quarter; the reference counts entries, and rows may be multisets).
// The top bit of a field is the "touched, then zeroed" mark.  Returns what this thread emitted.
template <int PACK>
__device__ __forceinline__ uint32_t range_source(const Params& p, const RangeFences& fx, uint32_t u, uint64_t ub, uint32_t du, const FirstHop& f,
                                                 uint32_t C, uint4* rec, uint32_t* cnt, uint32_t* touched, uint32_t* s_tn,
                                                 uint32_t* s_inc, unsigned long long* s_wb, uint32_t* s_wsum, Tally& tally) {
  typedef RangeField<PACK> F;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const uint32_t* __restrict__ keys = p.g.keys;
  const uint64_t span = (uint64_t)PACK * C;            // vertices per window
  uint32_t emitted = 0;
  // N(u) is walked along with the windows (every thread keeps the same cursor): xa = first entry of
  // row u not below the current window, xnext = its key
  uint32_t xa = lower_bound_row(keys, ub, du, u + 1u);
  uint32_t xnext = xa < du ? __ldg(keys + ub + xa) : 0xffffffffu;
  bool first = true;
  for (uint64_t j = ((uint64_t)u + 1) / span; j * span < p.g.S; ++j, first = false) {
    const uint32_t wbase = (uint32_t)(j * span);
    const uint32_t vhi = (uint32_t)((j + 1) * span < p.g.S ? (j + 1) * span : p.g.S);
    const uint32_t cell0 = (uint32_t)j * (uint32_t)PACK;
    for (uint32_t c = 0; c < f.npieces; ++c) {
      const uint32_t pc = f.npieces == 1 ? f.single_count : __ldg(f.piece_cnt + c);
      const uint32_t* pb = f.base + (uint64_t)c * CHUNK;
      for (uint32_t base = 0; base < pc; base += RANGE_THREADS) {
        const uint32_t i = base + tid;
        const bool has = i < pc;
        const uint64_t ci = (uint64_t)c * CHUNK + i;
        range_batch<PACK>(p, fx, has, first, (has && first) ? __ldg(pb + i) : 0u, u, cell0, wbase, vhi, rec + ci, cnt, touched, s_tn, s_inc, s_wb, s_wsum);
      }
    }
    if (xnext < vhi) {   // exclusion of N(u) inside the window (inc/predict.hxx:307)
      const uint32_t a = xa;
      const uint32_t b = vhi >= p.g.S ? du : (uint32_t)(gallop_to<true>(keys, ub + a, ub + du, vhi) - ub);
      for (uint32_t i = a + tid; i < b; i += blockDim.x) {
        const uint32_t x = __ldg(keys + ub + i) - wbase;
        if (PACK > 1) {   // neighbours share words: atomics, one field each
          const uint32_t sh = (x & (uint32_t)(PACK - 1)) * F::BITS;
          if ((cnt[x >> F::SHIFT] >> sh) & F::MASK) {
            atomicAnd(cnt + (x >> F::SHIFT), ~(F::VALUE << sh));
            atomicOr(cnt + (x >> F::SHIFT), F::MARK << sh);
          }
        } else {
          if (cnt[x] != 0u) cnt[x] = RANGE_ZEROED;
        }
      }
      xa = b;
      xnext = b < du ? __ldg(keys + ub + b) : 0xffffffffu;
    }
    __syncthreads();
    // Scoring walks the list of touched vertices (appended by the first wedge that reached each),
    // 32 at a time with all lanes busy, and reads the count back from the counter.  (Scanning the
    // window's counters instead -- 12 % of them touched at R-MAT 18 IHub -- was a third of the
    // kernel's instructions, and scoring straight from the scan, 4 of 32 lanes busy, 60 %: ncu.)
    {
      const uint32_t tn = *s_tn;
      for (uint32_t sb = (uint32_t)warp * 32u; sb < tn; sb += (uint32_t)nw * 32u) {
        const uint32_t i = sb + lane;
        const bool has = i < tn;
        uint32_t v = 0, c = 0;
        if (has) {
          v = __ldcg(touched + i);
          const uint32_t x = v - wbase;
          c = PACK == 1 ? (cnt[x] & ~RANGE_ZEROED) : ((cnt[x >> F::SHIFT] >> ((x & (uint32_t)(PACK - 1)) * F::BITS)) & F::VALUE);
        }
        emitted += score_and_emit(p, has, u, du, v, c, 0.0f, tally);
      }
      __syncthreads();                                                      // every warp has read its counts
      // clear: through the touched list when the window is sparsely hit (at R-MAT 24 a window sees a
      // few thousand wedges for 53 248 words), else the whole window
      const uint32_t slots = (vhi - wbase + (uint32_t)PACK - 1u) >> F::SHIFT;
      if (tn < (slots >> 2)) {
        for (uint32_t i = tid; i < tn; i += blockDim.x) cnt[(__ldcg(touched + i) - wbase) >> F::SHIFT] = 0u;
      } else {
        for (uint32_t i = tid; i < slots; i += blockDim.x) cnt[i] = 0u;
      }
      if (tid == 0) *s_tn = 0u;
    }
    __syncthreads();
  }
  return emitted;
}

template <bool ADMIT>
__global__ void __launch_bounds__(RANGE_THREADS, 1) k_range(Params p, RangeFences fx, const uint32_t* __restrict__ list, uint32_t n, int bin,
                                                             uint32_t* __restrict__ deferred, uint32_t C,
                                                             unsigned long long* __restrict__ cursors, uint64_t cursor_stride,
                                                             uint32_t* __restrict__ touched_all) {
  extern __shared__ uint32_t cnt[];                   // C words: C, 2 * C or 4 * C counters
  uint32_t* touched = touched_all + (uint64_t)blockIdx.x * 4u * C;          // vertices of the current window with a count
  uint4* rec = reinterpret_cast<uint4*>(cursors + (uint64_t)blockIdx.x * 2 * cursor_stride);   // [cursor_stride] row records
  __shared__ unsigned long long s_wb[RANGE_THREADS];
  __shared__ uint32_t s_inc[RANGE_THREADS];
  __shared__ uint32_t s_wsum[32];
  __shared__ int s_go;
  __shared__ uint32_t s_qi;
  __shared__ unsigned int s_emitted;
  __shared__ uint32_t s_tn;                           // entries of touched[]
  const int tid = threadIdx.x, lane = tid & 31;
  Tally tally;
  for (uint32_t i = tid; i < C; i += blockDim.x) cnt[i] = 0u;
  if (tid == 0) { s_emitted = 0; s_tn = 0u; }
  __syncthreads();
  for (;;) {
    if (tid == 0) s_qi = (uint32_t)atomicAdd(&p.ctr->queue[bin], 1ull);     // dynamic: sources differ by 1000x in work
    __syncthreads();
    const uint32_t qi = s_qi;
    __syncthreads();
    if (qi >= n) break;
    const uint32_t u = __ldg(list + qi);
    uint32_t need = 0;
    if (ADMIT) {
      if (tid == 0) s_go = admit_source(p, u, bin, deferred, &need) ? 1 : 0;
      __syncthreads();
      const int go = s_go;
      __syncthreads();
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    const FirstHop f = first_hop(p, u, ub, du);
    uint32_t emitted;
    if (du < p.range_quarter)   emitted = range_source<4>(p, fx, u, ub, du, f, C, rec, cnt, touched, &s_tn, s_inc, s_wb, s_wsum, tally);
    else if (du < p.range_half) emitted = range_source<2>(p, fx, u, ub, du, f, C, rec, cnt, touched, &s_tn, s_inc, s_wb, s_wsum, tally);
    else                        emitted = range_source<1>(p, fx, u, ub, du, f, C, rec, cnt, touched, &s_tn, s_inc, s_wb, s_wsum, tally);
    if (ADMIT) {
      if (lane == 0 && emitted) atomicAdd(&s_emitted, emitted);
      __syncthreads();
      if (tid == 0) {
        atomicAdd(&p.ctr->reserved, (unsigned long long)s_emitted - (unsigned long long)need);
        s_emitted = 0;
      }
      __syncthreads();
    }
  }
  tally.flush(p.ctr);
}

// Fence tables of the long rows (see range_batch): one warp per listed row, lanes over the cells.
__global__ void __launch_bounds__(256) k_fence_mark(DevGraph g, uint32_t min_deg, uint32_t* __restrict__ flag) {
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < g.S; w += (uint64_t)gridDim.x * blockDim.x)
    flag[w] = g.deg[w] >= min_deg ? 1u : 0u;
}

// scan[] = exclusive scan of the marks (slot of a marked row); rows that are not marked get
// slot 0xffffffff, marked rows fill their table.
__global__ void __launch_bounds__(256) k_fence_fill(DevGraph g, uint32_t min_deg, uint32_t C, uint32_t ncell,
                                                    const unsigned long long* __restrict__ scan,
                                                    uint32_t* __restrict__ slot_of, uint32_t* __restrict__ fence) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t w = warp; w < g.S; w += nwarps) {
    const uint32_t dw = g.deg[w];
    const bool marked = dw >= min_deg;
    const uint32_t slot = (uint32_t)scan[w];
    if (lane == 0) slot_of[w] = marked ? slot : 0xffffffffu;
    if (!marked) continue;
    const uint64_t wb = g.off[w];
    uint32_t* f = fence + (uint64_t)slot * (ncell + 1u);
    for (uint32_t c = lane; c <= ncell; c += 32) {
      const uint64_t x = (uint64_t)c * C;
      f[c] = x >= g.S ? dw : lower_bound_row(g.keys, wb, dw, (uint32_t)x);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Range path of the FLOAT measures (Adamic-Adar, resource allocation), hub-heavy sources.
// The reference accumulates acc(v) = float(double(acc(v)) + g(w)) over the first-hop entries w of u
// in ascending order (inc/predict.hxx:788, 828), and float addition does not commute, so the
// block-wide atomic counters of k_range cannot be used.  Here every WARP owns a source and a small
// window of float accumulators in shared memory (32 warps x 1664 floats = 208 KB per block), and
// walks the windows lo = u+1, u+1+1664, ...: per window the first-hop rows are visited 32 at a
// time (lane = row: cursor record, "anything in this window?" from the cached next key, gallop to
// the window's end), then the rows that have something are taken ONE AFTER THE OTHER in ascending
// order, the window part of a row spread over the lanes -- the entries of one row are distinct
// vertices, so the lanes never collide, and the row order is the reference's accumulation order:
// bit-exact without atomics or sorting.  (Rows that repeat an entry -- multiset rows -- would
// collide inside a row: such graphs keep the single-warp dense-table path.)
// A window's accumulator is 0 = untouched, RFLT_ZEROED = touched, then zeroed because v is in N(u)
// (inc/predict.hxx:306-307), else the sum (> 0: every term is positive).
constexpr uint32_t RFLT_WIN = 1664;
constexpr float RFLT_ZEROED = -1.0f;
enum { RFLT_WARPS = 32 };

// A warp is one dependent chain, so one source per warp would leave the longest source (a hub: 3e4
// rows x 157 windows at R-MAT 18) as the critical path of the whole kernel.  The windows of a source
// are independent, so heavy sources are cut into ITEMS = (source, range of windows) of about
// RFLT_ITEM_VISITS row visits each; an item is also the unit of admission when the candidate
// buffer is pruned (its bound: min(work, vertices in its windows)).
constexpr uint32_t RFLT_ITEM_VISITS = 1u << 16;

struct FltItems {
  uint32_t* u;        // [n] source vertex
  uint32_t* w0;       // [n] first window
  uint32_t* w1;       // [n] one past the last window
};

__device__ __forceinline__ uint32_t flt_item_chunks(uint32_t S, uint32_t u, uint32_t du) {
  const uint32_t nwin = (S - 1u - u + RFLT_WIN - 1u) / RFLT_WIN;
  if (nwin <= 1u) return 1u;
  const unsigned long long visits = (unsigned long long)du * nwin;
  unsigned long long c = (visits + RFLT_ITEM_VISITS - 1u) / RFLT_ITEM_VISITS;
  if (c < 1ull) c = 1ull;
  return (uint32_t)(c < nwin ? c : nwin);
}

__global__ void __launch_bounds__(256) k_flt_item_counts(DevGraph g, const uint32_t* __restrict__ list, uint32_t n, uint32_t* __restrict__ cnt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = list[i];
    cnt[i] = flt_item_chunks(g.S, u, g.deg[u]);
  }
}

__global__ void __launch_bounds__(256) k_flt_item_fill(DevGraph g, const uint32_t* __restrict__ list, uint32_t n,
                                                       const unsigned long long* __restrict__ off, FltItems it, uint32_t* __restrict__ ids) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = list[i];
    const uint32_t nwin = (g.S - 1u - u + RFLT_WIN - 1u) / RFLT_WIN;
    const uint32_t c = flt_item_chunks(g.S, u, g.deg[u]);
    const unsigned long long base = off[i];
    for (uint32_t k = 0; k < c; ++k) {
      it.u[base + k] = u;
      it.w0[base + k] = (uint32_t)((unsigned long long)nwin * k / c);
      it.w1[base + k] = (uint32_t)((unsigned long long)nwin * (k + 1) / c);
      ids[base + k] = (uint32_t)(base + k);
    }
  }
}

// Admission of one item (pruned candidate buffer): as admit_source, with the item's own bound.
__device__ __forceinline__ bool admit_item(const Params& p, uint32_t id, uint32_t need, int bin, uint32_t* deferred) {
  for (;;) {
    if (*reinterpret_cast<volatile unsigned long long*>(&p.ctr->cursor) >= p.soft_cap) break;
    const unsigned long long r = atomicAdd(&p.ctr->reserved, (unsigned long long)need);
    if (r + need <= p.cap) return true;
    atomicAdd(&p.ctr->reserved, 0ull - (unsigned long long)need);
    for (;;) {
      const unsigned long long written = *reinterpret_cast<volatile unsigned long long*>(&p.ctr->cursor);
      if (written + need > p.cap) goto defer;
      if (*reinterpret_cast<volatile unsigned long long*>(&p.ctr->reserved) + need <= p.cap) break;
      __nanosleep(2000);
    }
  }
defer:
  deferred[atomicAdd(&p.ctr->deferred[bin], 1ull)] = id;
  return false;
}

template <bool ADMIT>
__global__ void __launch_bounds__(RFLT_WARPS * 32, 1) k_range_flt(Params p, FltItems items, const uint32_t* __restrict__ list, uint32_t n, int bin,
                                                                  uint32_t* __restrict__ deferred, uint4* __restrict__ rec_all,
                                                                  double* __restrict__ g_all, uint64_t rec_stride, uint32_t* __restrict__ tlist_all) {
  extern __shared__ float facc[];                     // [RFLT_WARPS][RFLT_WIN]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc = facc + (uint32_t)warp * RFLT_WIN;
  uint4* rec = rec_all + ((uint64_t)blockIdx.x * RFLT_WARPS + warp) * rec_stride;
  double* grow = g_all + ((uint64_t)blockIdx.x * RFLT_WARPS + warp) * rec_stride;   // per first-hop row: its term g(w)
  // vertices of the current window that have a sum, in the order they were first reached: scoring
  // walks this list with all lanes busy instead of scanning the window (12 % of it is touched at
  // R-MAT 18), and only these accumulators need clearing
  uint32_t* tlist = tlist_all + ((uint64_t)blockIdx.x * RFLT_WARPS + warp) * RFLT_WIN;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t* __restrict__ keys = p.g.keys;
  Tally tally;
  for (uint32_t i = lane; i < RFLT_WIN; i += 32) acc[i] = 0.0f;
  __syncwarp();
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = (uint32_t)atomicAdd(&p.ctr->queue[bin], 1ull);      // dynamic: items still differ in work
    qi = __shfl_sync(NLP_FULL, qi, 0);
    if (qi >= n) break;
    const uint32_t id = __ldg(list + qi);
    const uint32_t u = __ldg(items.u + id), w0 = __ldg(items.w0 + id), w1 = __ldg(items.w1 + id);
    uint32_t need = 0;
    if (ADMIT) {
      const unsigned long long verts = (unsigned long long)(w1 - w0) * RFLT_WIN;
      const uint32_t wk = p.work[u];
      need = (unsigned long long)wk < verts ? wk : (uint32_t)verts;
      int go = 0;
      if (lane == 0) go = admit_item(p, id, need, bin, deferred) ? 1 : 0;
      go = __shfl_sync(NLP_FULL, go, 0);
      if (!go) continue;
    }
    const uint64_t ub = __ldg(p.g.off + u);
    const uint32_t du = (uint32_t)(__ldg(p.g.off + u + 1) - ub);
    const FirstHop f = first_hop(p, u, ub, du);
    uint32_t emitted = 0;
    for (uint32_t win = w0; win < w1; ++win) {
      const uint64_t lo64 = (uint64_t)u + 1 + (uint64_t)win * RFLT_WIN;
      if (lo64 >= p.g.S) break;
      const uint32_t vlo = (uint32_t)lo64;
      const uint32_t vhi = (uint32_t)(lo64 + RFLT_WIN < p.g.S ? lo64 + RFLT_WIN : p.g.S);
      const bool first = win == w0;                    // the item's first window: the rows are bisected, later windows continue
      uint32_t tn = 0;                                 // entries of tlist (warp-uniform)
      for (uint32_t c = 0; c < f.npieces; ++c) {
        const uint32_t pc = f.npieces == 1 ? f.single_count : __ldg(f.piece_cnt + c);
        const uint32_t* pb = f.base + (uint64_t)c * CHUNK;
        for (uint32_t base = 0; base < pc; base += 32) {
          const uint32_t i = base + lane;
          const bool has = i < pc;
          const uint64_t ci = (uint64_t)c * CHUNK + i;
          unsigned long long a = 0;
          uint32_t cntw = 0;
          double g = 0.0;
          if (has) {
            unsigned long long e;
            uint32_t next;
            bool moved = false;
            if (first) {
              const uint32_t w = __ldg(pb + i);
              const unsigned long long wb = __ldg(p.g.off + w);
              e = __ldg(p.g.off + w + 1);
              const uint32_t dw = (uint32_t)(e - wb);
              g = flt_term(p, dw);                     // inc/predict.hxx:788, 828: a double; kept per row for the later windows
              grow[ci] = g;
              a = wb + lower_bound_row(keys, wb, dw, vlo);
              next = a < e ? __ldg(keys + a) : 0xffffffffu;
              moved = true;
            } else {
              const uint4 r = rec[ci];
              a = (unsigned long long)r.x | ((unsigned long long)r.y << 32);
              e = a + r.z;
              next = r.w;
            }
            unsigned long long b = a;
            if (next < vhi) {                         // the row has entries in this window (so a < e)
              b = vhi >= p.g.S ? e : gallop_to<true>(keys, a, e, vhi);
              next = b < e ? __ldg(keys + b) : 0xffffffffu;
              moved = true;
              if (!first) g = grow[ci];
            }
            if (moved) rec[ci] = range_pack(b, e, next);
            cntw = (uint32_t)(b - a);
          }
          // Rows with wedges in the window, in ascending first-hop order, FOUR at a time: the key
          // loads of four rows are issued together (a warp is a single dependent chain here, so the
          // load latency per row is what bounds this kernel), the accumulations then follow one row
          // after the other.
          unsigned m = __ballot_sync(NLP_FULL, cntw != 0u);
          while (m) {
            int rr[4];
            #pragma unroll
            for (int j = 0; j < 4; ++j) { rr[j] = m ? __ffs(m) - 1 : -1; if (m) m &= m - 1u; }
            unsigned long long ar[4];
            uint32_t cr[4], kv[4];
            double gr[4];
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = rr[j] < 0 ? 0 : rr[j];
              ar[j] = __shfl_sync(NLP_FULL, a, r);
              cr[j] = rr[j] < 0 ? 0u : __shfl_sync(NLP_FULL, cntw, r);
              gr[j] = __shfl_sync(NLP_FULL, g, r);
              kv[j] = (uint32_t)lane < cr[j] ? __ldg(keys + ar[j] + lane) : 0u;
            }
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (rr[j] < 0) break;                    // warp-uniform
              for (uint32_t k0 = 0; k0 < cr[j]; k0 += 32) {      // warp-uniform trip count
                const uint32_t k = k0 + lane;
                bool fresh = false;
                uint32_t v = 0;
                if (k < cr[j]) {
                  v = k0 == 0 ? kv[j] : __ldg(keys + ar[j] + k);
                  const uint32_t x = v - vlo;
                  const float old = acc[x];
                  acc[x] = __double2float_rn(__dadd_rn((double)old, gr[j]));
                  fresh = old == 0.0f;                 // first wedge that reaches this vertex: list it
                }
                const unsigned fm = __ballot_sync(NLP_FULL, fresh);
                if (fresh) tlist[tn + __popc(fm & lt)] = v;
                tn += __popc(fm);
              }
              __syncwarp();                            // the next row may reach the same vertices
            }
          }
        }
      }
      {   // exclusion of N(u) inside the window (inc/predict.hxx:307)
        const uint32_t a0 = lower_bound_row(keys, ub, du, vlo);
        const uint32_t b0 = a0 + lower_bound_row(keys, ub + a0, du - a0, vhi);
        for (uint32_t i = a0 + lane; i < b0; i += 32) {
          const uint32_t x = __ldg(keys + ub + i) - vlo;
          if (acc[x] != 0.0f) acc[x] = RFLT_ZEROED;
        }
        __syncwarp();
      }
      __syncwarp();
      for (uint32_t sb = 0; sb < tn; sb += 32) {       // score the touched vertices and clear their accumulators
        const uint32_t i = sb + lane;
        const bool has = i < tn;
        uint32_t v = 0;
        float val = 0.0f;
        if (has) {
          v = tlist[i];
          val = acc[v - vlo];
          acc[v - vlo] = 0.0f;
          if (val < 0.0f) val = 0.0f;                  // touched, then zeroed: a candidate with value 0
        }
        emitted += score_and_emit(p, has, u, du, v, 0u, val, tally);
      }
      __syncwarp();
    }
    if (ADMIT && lane == 0) atomicAdd(&p.ctr->reserved, (unsigned long long)emitted - (unsigned long long)need);
  }
  tally.flush(p.ctr);
}

}  // namespace nlp
