// Kernel (a) of the north star: the work-balanced frontier.
//   k_degrees      deg[u] = off[u+1]-off[u]; number of 2048-entry chunks of every long row
//                  (once per graph)
//   k_chunk_fill   chunk -> source row table for the long rows (once per graph)
//   k_elig         LHub eligibility bitmask, bit w = deg(w) <= D (replaces the dependent random
//                  degree load of inc/predict.hxx:301 by an L2-resident 1-bit lookup)
//   k_work_short   one coalesced pass over the adjacency of all short rows: per source
//   k_work_long    work(u) = sum_{eligible w in N(u)} deg(w) (wedges the reference would scan for
//                  u, inc/predict.hxx:298-304) and, for LHub, the COMPACTED list of eligible
//                  first-hop entries, so the wedge kernels never touch a hub's row again.
//                  Long rows are cut into 2048-entry chunks (one block each) so a 10^5..10^6-degree
//                  hub does not serialise one warp.
//   k_bin          drop zero-work sources (inc/predict.hxx:287-289 visits them all) and bin the
//                  rest by work into the sub-warp / block-hash / dense-spill paths
#pragma once
#include "common.cuh"

namespace nlp {

// maxdeg[0] = largest degree; maxdeg[1] |= 1 when the offsets are not non-decreasing or a row is
// longer than 2^32 - 1 entries (validation of the caller's CSR, see k_validate_entries)
__global__ void __launch_bounds__(256) k_degrees(const uint64_t* __restrict__ off, uint32_t S,
                                                 uint32_t* __restrict__ deg, unsigned long long* __restrict__ nchunks,
                                                 uint32_t* maxdeg) {
  uint32_t local = 0;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < S; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t a = off[u], b = off[u + 1];
    if (b < a || b - a > 0xffffffffull) atomicOr(maxdeg + 1, 1u);
    const uint32_t d = b < a ? 0u : (uint32_t)(b - a);
    deg[u] = d;
    nchunks[u] = d > LONG_ROW ? (d + CHUNK - 1) / CHUNK : 0u;
    local = max(local, d);
  }
  local = __reduce_max_sync(NLP_FULL, local);
  if ((threadIdx.x & 31) == 0 && local) atomicMax(maxdeg, local);
}

// chunk_base = exclusive scan of nchunks; chunk_src[chunk_base[u] + c] = u
__global__ void __launch_bounds__(256) k_chunk_fill(const uint32_t* __restrict__ deg, const unsigned long long* __restrict__ chunk_base,
                                                    uint32_t S, uint32_t* __restrict__ chunk_src) {
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < S; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = deg[u];
    if (d > LONG_ROW) {
      const uint32_t n = (d + CHUNK - 1) / CHUNK;
      const unsigned long long b = chunk_base[u];
      for (uint32_t c = 0; c < n; ++c) chunk_src[b + c] = (uint32_t)u;
    }
  }
}

// Largest number of equal entries in one row (1 for a simple graph; rows are sorted, so equal
// entries are adjacent).  The reference counts ENTRIES, so with multiset rows a pair's count can
// reach deg(u) times this number -- k_range's half-word counters need the bound.  One warp per 32
// consecutive entries; the row of the first by bisection, the lanes walk forward from there.
// Validation of the caller's CSR (include/nlp_b200.h promises NLP_ERR_ARG): every key below span
// (flag 2), every row non-decreasing (flag 4) -- the kernels index deg[], the eligibility mask and
// the counters with unchecked keys and bisect rows, so a bad input must not get past
// nlp_set_graph -- and, in the same pass, the largest entry multiplicity (out[1]) and a 64-bit
// content fingerprint (sum over the entries of a hash of (position, row, key)): a graph that is
// bound again unchanged (the base graph of the next batch) is recognised by it and keeps what is
// already known about it (symmetric rows or not).
// A warp takes VAL_CHUNK consecutive entries: ONE bisection of the offsets finds the row of the
// first, after that the rows are walked forward.
enum { VAL_CHUNK = 4096 };

__device__ __forceinline__ unsigned long long val_mix(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256) k_validate_entries(DevGraph g, uint64_t M, unsigned int* __restrict__ flags,
                                                          unsigned int* __restrict__ maxmult, unsigned long long* __restrict__ fingerprint) {
  const uint32_t* __restrict__ keys = g.keys;
  const int lane = threadIdx.x & 31;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t best = 1, bad = 0;
  unsigned long long fp = 0;
  for (uint64_t c0 = warp0 * VAL_CHUNK; c0 < M; c0 += nwarps * VAL_CHUNK) {
    uint32_t lo = 0, hi = g.S;                       // off[lo] <= c0 < off[hi]
    while (lo + 1 < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(g.off + mid) <= c0) lo = mid; else hi = mid;
    }
    uint32_t urow = lo;                              // row of the batch's first entry (warp-uniform)
    const uint64_t cend = c0 + VAL_CHUNK < M ? c0 + VAL_CHUNK : M;
    for (uint64_t base = c0; base < cend; base += 32u) {
      while (urow + 1 < g.S && __ldg(g.off + urow + 1) <= base) ++urow;
      const uint64_t e = base + lane;
      if (e >= cend) continue;
      uint32_t u = urow;
      while (u + 1 < g.S && __ldg(g.off + u + 1) <= e) ++u;
      const uint64_t ub = __ldg(g.off + u), ue = __ldg(g.off + u + 1);
      const uint32_t w = __ldg(keys + e);
      fp += val_mix(e * 0x9E3779B97F4A7C15ull + ((unsigned long long)u << 32) + w);
      if (w >= g.S) bad |= 2u;
      if (e > ub) {
        const uint32_t prev = __ldg(keys + e - 1);
        if (prev > w) bad |= 4u;
        if (prev == w) continue;                                 // multiplicity measured at the first entry of the run
      }
      if (e + 1 < ue && __ldg(keys + e + 1) == w) {
        uint32_t mult = 2;
        while (e + mult < ue && __ldg(keys + e + mult) == w) ++mult;
        best = mult > best ? mult : best;
      }
    }
  }
  best = __reduce_max_sync(NLP_FULL, best);
  bad = __reduce_or_sync(NLP_FULL, bad);
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) fp += __shfl_xor_sync(NLP_FULL, fp, k);
  if (lane == 0) {
    if (best > 1u) atomicMax(maxmult, best);
    if (bad) atomicOr(flags, bad);
    if (fp) atomicAdd(fingerprint, fp);
  }
}

// One thread builds one 32-bit word of the mask.
__global__ void __launch_bounds__(256) k_elig(const uint32_t* __restrict__ deg, uint32_t S, uint32_t D,
                                              uint32_t* __restrict__ bits) {
  const uint32_t nwords = (S + 31u) >> 5;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t base = (uint32_t)w << 5;
    uint32_t m = 0;
    #pragma unroll 8
    for (uint32_t i = 0; i < 32; ++i) {
      const uint32_t v = base + i;
      if (v < S && deg[v] <= D) m |= 1u << i;
    }
    bits[w] = m;
  }
}

// Rows are owned in blocks of 32 consecutive vertices, dealt round-robin to the ranks.
__host__ __device__ inline bool owns_row_block(uint64_t rb, int rank, int world) {
  return world <= 1 || (int)(rb % (uint64_t)world) == rank;
}

struct WorkOut {
  unsigned long long* work64;   // [S]   exact work(u)
  uint32_t* ecount;             // [S]   compacted eligible entries of a short row
  uint32_t* ekeys;              // [M]   compacted eligible first-hop entries, row u at off[u]
  uint32_t* chunk_cnt;          // [NC]  compacted entries of every long-row chunk
};

// One warp streams the adjacency of the short rows among 32 consecutive rows, packed back to
// back (coalesced), looks up eligibility of every first-hop entry, segment-sums deg(w) per row
// and (LHub) writes the eligible entries with deg(w) > 0 compacted to the front of the row's
// slot in `ekeys`, order preserved.
template <bool LHUB>
__global__ void __launch_bounds__(256) k_work_short(DevGraph g, const uint32_t* __restrict__ elig, int rank, int world,
                                                    WorkOut o, Counters* ctr) {
  __shared__ unsigned long long acc[8][32];
  __shared__ uint32_t cnt[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const uint64_t nrb = ((uint64_t)g.S + 31u) >> 5;
  unsigned long long t_first = 0, t_elig = 0, t_wedges = 0;
  for (uint64_t rb = blockIdx.x * 8ull + wib; rb < nrb; rb += gridDim.x * 8ull) {
    if (!owns_row_block(rb, rank, world)) continue;
    const uint64_t row = rb * 32u + lane;
    uint64_t b0 = 0;
    uint32_t d = 0;
    if (row < g.S) { b0 = g.off[row]; d = (uint32_t)(g.off[row + 1] - b0); }
    const uint32_t sl = d > LONG_ROW ? 0u : d;
    uint32_t inc = sl;
    #pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
      const uint32_t t = __shfl_up_sync(NLP_FULL, inc, k);
      if (lane >= k) inc += t;
    }
    const uint32_t T = __shfl_sync(NLP_FULL, inc, 31);
    acc[wib][lane] = 0;
    cnt[wib][lane] = 0;
    __syncwarp();
    for (uint32_t base = 0; base < T; base += 32) {
      const uint32_t t = base + lane;
      const bool valid = t < T;
      int j = 0;                                   // smallest j with inc[j] > t
      #pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t x = __shfl_sync(NLP_FULL, inc, j + step - 1);
        if (x <= t) j += step;
      }
      const uint32_t incj = __shfl_sync(NLP_FULL, inc, j);
      const uint32_t slj  = __shfl_sync(NLP_FULL, sl, j);
      const uint64_t b0j  = __shfl_sync(NLP_FULL, b0, j);
      uint32_t w = 0, c = 0;
      bool el = false;
      if (valid) {
        w = __ldg(g.keys + b0j + (t - (incj - slj)));
        el = LHUB ? (((__ldg(elig + (w >> 5)) >> (w & 31)) & 1u) != 0u) : true;
        if (el) c = __ldg(g.deg + w);
      }
      t_first += valid ? 1u : 0u;
      t_elig += el ? 1u : 0u;
      t_wedges += c;
      const int j0 = __shfl_sync(NLP_FULL, j, 0);
      const bool uni = __all_sync(NLP_FULL, !valid || j == j0);
      if (uni) {   // whole batch inside one row: one shared-memory update
        const uint32_t lo = __reduce_add_sync(NLP_FULL, c & 0xffffu);
        const uint32_t hi = __reduce_add_sync(NLP_FULL, c >> 16);
        if (lane == 0) acc[wib][j0] += (unsigned long long)lo + ((unsigned long long)hi << 16);
      } else if (c) {
        atomicAdd(&acc[wib][j], (unsigned long long)c);
      }
      if (LHUB) {
        const bool keep = c != 0;
        const unsigned km = __ballot_sync(NLP_FULL, keep);
        if (km) {
          const unsigned m = uni ? __ballot_sync(NLP_FULL, valid) : __match_any_sync(NLP_FULL, valid ? (unsigned)j : 32u + lane);
          const uint32_t before = valid ? cnt[wib][j] : 0u;
          if (keep) o.ekeys[b0j + before + __popc(m & km & lt)] = w;
          __syncwarp();
          if (valid && (__ffs(m) - 1) == lane) cnt[wib][j] = before + __popc(m & km);
        }
      }
      __syncwarp();
    }
    if (row < g.S) {
      o.work64[row] = acc[wib][lane];       // 0 for long rows: k_work_long adds to it
      if (LHUB) o.ecount[row] = cnt[wib][lane];
    }
    __syncwarp();
  }
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    t_first  += __shfl_xor_sync(NLP_FULL, t_first, k);
    t_elig   += __shfl_xor_sync(NLP_FULL, t_elig, k);
    t_wedges += __shfl_xor_sync(NLP_FULL, t_wedges, k);
  }
  if (lane == 0) {
    if (t_first)  atomicAdd(&ctr->first_hop, t_first);
    if (t_elig)   atomicAdd(&ctr->eligible_first_hop, t_elig);
    if (t_wedges) atomicAdd(&ctr->wedges, t_wedges);
  }
}

// One block per 2048-entry chunk of a long row.
template <bool LHUB>
__global__ void __launch_bounds__(256) k_work_long(DevGraph g, const uint32_t* __restrict__ elig, int rank, int world,
                                                   const uint32_t* __restrict__ chunk_src,
                                                   const unsigned long long* __restrict__ chunk_base, uint32_t nchunks,
                                                   WorkOut o, Counters* ctr) {
  __shared__ uint32_t warp_keep[8];
  __shared__ unsigned long long warp_sum[8][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const uint32_t u = chunk_src[c];
    if (!owns_row_block(u >> 5, rank, world)) continue;
    const uint32_t ci = (uint32_t)(c - chunk_base[u]);
    const uint64_t rb = g.off[u], re = g.off[u + 1];
    const uint64_t cb = rb + (uint64_t)ci * CHUNK;
    const uint64_t ce = cb + CHUNK < re ? cb + CHUNK : re;
    unsigned long long t_first = 0, t_elig = 0, t_wedges = 0;
    uint32_t written = 0;                            // block-uniform running count of kept entries
    for (uint64_t base = cb; base < ce; base += 256) {
      const uint64_t e = base + tid;
      const bool valid = e < ce;
      uint32_t w = 0, cw = 0;
      bool el = false;
      if (valid) {
        w = __ldg(g.keys + e);
        el = LHUB ? (((__ldg(elig + (w >> 5)) >> (w & 31)) & 1u) != 0u) : true;
        if (el) cw = __ldg(g.deg + w);
      }
      t_first += valid ? 1u : 0u;
      t_elig += el ? 1u : 0u;
      t_wedges += cw;
      if (LHUB) {
        const bool keep = cw != 0;
        const unsigned km = __ballot_sync(NLP_FULL, keep);
        if (lane == 0) warp_keep[warp] = __popc(km);
        __syncthreads();
        uint32_t before = 0, total = 0;
        #pragma unroll
        for (int k = 0; k < 8; ++k) { const uint32_t x = warp_keep[k]; if (k < warp) before += x; total += x; }
        if (keep) o.ekeys[cb + written + before + __popc(km & lt)] = w;
        written += total;
        __syncthreads();
      }
    }
    #pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
      t_first  += __shfl_xor_sync(NLP_FULL, t_first, k);
      t_elig   += __shfl_xor_sync(NLP_FULL, t_elig, k);
      t_wedges += __shfl_xor_sync(NLP_FULL, t_wedges, k);
    }
    if (lane == 0) { warp_sum[warp][0] = t_first; warp_sum[warp][1] = t_elig; warp_sum[warp][2] = t_wedges; }
    __syncthreads();
    if (tid == 0) {
      unsigned long long a = 0, b = 0, w = 0;
      for (int k = 0; k < 8; ++k) { a += warp_sum[k][0]; b += warp_sum[k][1]; w += warp_sum[k][2]; }
      if (w) atomicAdd(&o.work64[u], w);
      if (LHUB) o.chunk_cnt[c] = written;
      atomicAdd(&ctr->first_hop, a);
      if (b) atomicAdd(&ctr->eligible_first_hop, b);
      if (w) atomicAdd(&ctr->wedges, w);
    }
    __syncthreads();
  }
}

struct BinLists { uint32_t* list[NBINS]; };

// range_c: counters per window of k_range (0 = that path is off, e.g. for the float measures,
// whose hub-heavy sources go to the dense spill tables); room = S-1-u
__device__ __forceinline__ int choose_bin(uint32_t work, uint32_t bound, uint32_t du, uint32_t range_c, uint32_t range_fixed,
                                          uint32_t range_div, uint32_t half_deg, uint32_t quarter_deg, uint32_t room) {
  if (du <= LONG_ROW) {
    if (work <= 8u) return 0;
    if (work <= 32u) return 1;
  }
  if (bound <= bin_limit(2)) return 2;
  if (bound <= bin_limit(3)) return 3;
  if (bound <= bin_limit(4)) return 4;
  if (range_c) {
    // Every window costs a scan of the counters plus a cursor step per first-hop row; the dense
    // table pays an HBM sector update per wedge (~47 SM-cycles against ~4 for a shared-memory
    // atomic, R-MAT 18/20 IHub).  range_fixed + du / range_div is the cost of one window of one
    // source (count measures: 256 + du / NLP_B200_RANGE_DIV; float measures, whose windows are
    // one warp's 1664 accumulators: 32 + du / 4).
    // Sources with deg < half_deg count in half words: their windows are twice as wide (deg <
    // quarter_deg: bytes, four times).
    const unsigned long long rc = (unsigned long long)range_c << (du < quarter_deg ? 2 : (du < half_deg ? 1 : 0));
    const unsigned long long passes = ((unsigned long long)room + rc - 1) / rc;
    if ((unsigned long long)work >= passes * ((unsigned long long)range_fixed + du / range_div)) return 6;
  }
  return 5;
}

__global__ void __launch_bounds__(256) k_bin(DevGraph g, const unsigned long long* __restrict__ work64, int rank, int world,
                                             uint32_t range_c, uint32_t range_fixed, uint32_t range_div, uint32_t half_deg, uint32_t quarter_deg, uint32_t* __restrict__ work, BinLists bl,
                                             Counters* ctr) {
  __shared__ unsigned long long s_cnt[NBINS], s_sum[NBINS], s_base[NBINS], s_max;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < NBINS) { s_cnt[threadIdx.x] = 0; s_sum[threadIdx.x] = 0; }
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  // every block handles one contiguous tile of vertices, so it needs one global atomic per bin
  const uint64_t per = (((uint64_t)g.S + gridDim.x - 1) / gridDim.x + 255u) & ~255ull;
  const uint64_t lo = blockIdx.x * per, hi = lo + per < g.S ? lo + per : g.S;
  // pass 1: classify + count
  for (uint64_t base = lo; base < hi; base += 256) {
    const uint64_t u = base + threadIdx.x;
    int bin = -1;
    uint32_t need = 0;
    if (u < hi && owns_row_block(u >> 5, rank, world)) {
      const unsigned long long w64 = work64[u];
      const uint32_t w = w64 > 0xffffffffull ? 0xffffffffu : (uint32_t)w64;
      work[u] = w;
      if (w) {
        const uint32_t room = g.S - 1u - (uint32_t)u;
        const uint32_t bound = w < room ? w : room;
        if (bound) { bin = choose_bin(w, bound, g.deg[u], range_c, range_fixed, range_div, half_deg, quarter_deg, room); need = bin < 2 ? w : bound; }
      }
    }
    #pragma unroll
    for (int b = 0; b < NBINS; ++b) {
      const unsigned m = __ballot_sync(NLP_FULL, bin == b);
      if (!m) continue;
      unsigned long long sum = (bin == b) ? need : 0, mx = sum;
      #pragma unroll
      for (int k = 16; k >= 1; k >>= 1) {
        sum += __shfl_xor_sync(NLP_FULL, sum, k);
        const unsigned long long o = __shfl_xor_sync(NLP_FULL, mx, k);
        mx = o > mx ? o : mx;
      }
      if (lane == 0) {
        atomicAdd(&s_cnt[b], (unsigned long long)__popc(m));
        atomicAdd(&s_sum[b], sum);
        if (b >= 3) atomicMax(&s_max, mx);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < NBINS) {
    const unsigned long long c = s_cnt[threadIdx.x];
    s_base[threadIdx.x] = c ? atomicAdd(&ctr->bin_count[threadIdx.x], c) : 0;
    if (c) atomicAdd(&ctr->bin_bound[threadIdx.x], s_sum[threadIdx.x]);
    s_cnt[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0 && s_max) atomicMax(&ctr->max_bound, s_max);
  __syncthreads();
  // pass 2: same classification (work[] was just written by this block), now append
  unsigned long long frontier = 0;
  for (uint64_t base = lo; base < hi; base += 256) {
    const uint64_t u = base + threadIdx.x;
    int bin = -1;
    if (u < hi && owns_row_block(u >> 5, rank, world)) {
      const uint32_t w = work[u];
      if (w) {
        const uint32_t room = g.S - 1u - (uint32_t)u;
        const uint32_t bound = w < room ? w : room;
        if (bound) bin = choose_bin(w, bound, g.deg[u], range_c, range_fixed, range_div, half_deg, quarter_deg, room);
      }
    }
    #pragma unroll
    for (int b = 0; b < NBINS; ++b) {
      const unsigned m = __ballot_sync(NLP_FULL, bin == b);
      if (!m) continue;
      unsigned long long at = 0;
      if (lane == 0) at = atomicAdd(&s_cnt[b], (unsigned long long)__popc(m));
      at = __shfl_sync(NLP_FULL, at, 0);
      if (bin == b) bl.list[b][s_base[b] + at + __popc(m & ((1u << lane) - 1u))] = (uint32_t)u;
      if (lane == 0) frontier += __popc(m);
    }
  }
  if (lane == 0 && frontier) atomicAdd(&ctr->frontier, frontier);
}

}  // namespace nlp
