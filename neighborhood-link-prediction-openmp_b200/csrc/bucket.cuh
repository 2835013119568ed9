// LHub bucket path (NLP_PATH_PAIR): the w-centric wedge enumeration of pairs.cuh, regrouped BY
// SOURCE so that counting needs no global sort at all.
//
// The reference counts, for every source u, the entries v > u of the rows of its eligible
// first-hop neighbours w (inc/predict.hxx:298-304, 153-160).  On a graph with symmetric rows the
// eligible first-hop entries of u are exactly the entries "u" of the eligible rows w, so they can
// be listed from w's side (only rows with deg <= D are read -- the reference's dominant LHub cost,
// the hub test on all |E| first-hop entries of inc/predict.hxx:301, disappears) and then grouped
// by u once per (graph, D).  That grouping is a pure function of the resident graph and the
// threshold: it is the PLAN, built on the device the first time a threshold is used and kept
// until the graph changes.
//
//   plan (once per graph and D; k_plan_*):
//     item      = one eligible first-hop entry (u, w) that has at least one wedge with v > u:
//                 cnt = #entries of row w above u, ptr = index of the first of them, dw = deg(w)
//     items are sorted by u (stable: ascending w inside a source, the reference's accumulation
//     order), so every source owns a contiguous run of items and a contiguous range of wedge
//     records; roff[source] = first record slot of the source in the record-aligned output.
//     small sources (<= CAP/2 records) are cut into BUCKETS: windows of CAP/2 records in the
//     running record count, every source belongs to the window it starts in, so a bucket holds at
//     most CAP records and at most CAP/2 sources.  The few big sources take the global sort of
//     pairs.cuh (k_pair_emit -> radix sort -> k_pair_reduce) and are copied into their slots.
//
//   per prediction (k_bucket, one thread block per bucket):
//     gather    the bucket's wedge records (v, local source index[, deg w]) straight from the CSR
//               rows into shared memory -- they never exist in HBM
//     sort      stable LSD radix sort in shared memory by (local source, v), 8-bit digits, ranks by
//               warp-private digit masks (the k_scatter scheme of select.cuh)
//     reduce    run length = common-neighbour count (or the ordered float fold of AA / RA: the
//               stable sort kept ascending w), existing-edge exclusion by binary search in row u,
//               fused scoring (inc/predict.hxx:306-311)
//     output    record-aligned like k_pair_reduce: score bits at the slot of a run's first record,
//               NLP_NO_SCORE elsewhere, so the kept pairs lie in ascending (u, v) order and the
//               ordered top-K of select.cuh only has to sort by score.
#pragma once
#include "common.cuh"
#include "select.cuh"
#include "wedge.cuh"

namespace nlp {

enum { BK_THREADS = 256, BK_WARPS = 8, BK_GATHER = 4 };
constexpr uint32_t BK_CAP_COUNT = 4096;   // records per bucket (6 B per record x 2 buffers: three blocks per SM; the same plan
                                          // serves all nine measures).  NLP_B200_BUCKET_CAP=8192: two blocks per SM, fewer big
                                          // sources -- measured 2 % slower per step and 6 ms slower end to end (second plan)
constexpr uint32_t BK_CAP_FLT = 4096;     // float measures carry deg(w) (10 B per record x 2 buffers)
constexpr uint32_t BK_NONE = 0xffffffffu;   // aligned count array: no pair starts at this slot
constexpr uint32_t BK_DONE = 0xfffffffeu;   // ... the slot's score is already final (big sources)

// ---- plan ------------------------------------------------------------------------------------
// One thread per row w.  items[w] = entries of an eligible row that have a later, larger entry
// (rows are sorted multisets: everything below the last key).  Also the graph-only counters of
// a prediction: first-hop entries, eligible first-hop entries (= sum of deg(w) over eligible w,
// by symmetry), wedges W(D) = sum of deg(w)^2 (inc/predict.hxx:155).
__global__ void __launch_bounds__(256) k_plan_rows(DevGraph g, uint32_t D, uint32_t* __restrict__ items, Counters* ctr) {
  unsigned long long t_first = 0, t_elig = 0, t_wedges = 0;
  for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < g.S; w += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = g.deg[w];
    t_first += d;
    uint32_t it = 0;
    if (d >= 1u && d <= D) {
      t_elig += d;
      t_wedges += (unsigned long long)d * d;
      if (d >= 2u) {
        const uint64_t wb = __ldg(g.off + w);
        const uint32_t last = __ldg(g.keys + wb + d - 1);
        uint32_t j = d - 1;
        while (j > 0 && __ldg(g.keys + wb + j - 1) == last) --j;
        it = j;
      }
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

// Descriptors of the items of every eligible row, in (w, position) order, plus the (u, item index)
// pairs the sort by u works on.  A warp takes 32 consecutive rows (lane = row for the row data) and
// then walks the rows one after the other with all lanes on the row's entries: thresholds like
// D = 1024 make rows of a thousand items eligible, which one thread per row would serialise.
__global__ void __launch_bounds__(256) k_plan_items(DevGraph g, const uint32_t* __restrict__ items,
                                                    const unsigned long long* __restrict__ item_off,
                                                    uint32_t* __restrict__ su, uint32_t* __restrict__ sidx,
                                                    uint32_t* __restrict__ it_cnt, uint32_t* __restrict__ it_dw,
                                                    unsigned long long* __restrict__ it_ptr) {
  const uint32_t* __restrict__ keys = g.keys;
  const int lane = threadIdx.x & 31;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t w0 = warp0 * 32u; w0 < g.S; w0 += nwarps * 32u) {
    const uint64_t w = w0 + lane;
    uint32_t n = 0, d = 0;
    unsigned long long base = 0, wb = 0;
    if (w < g.S) { n = items[w]; if (n) { d = g.deg[w]; base = item_off[w]; wb = __ldg(g.off + w); } }
    unsigned m = __ballot_sync(NLP_FULL, n != 0u);
    while (m) {
      const int r = __ffs(m) - 1;
      m &= m - 1u;
      const uint32_t nr = __shfl_sync(NLP_FULL, n, r), dr = __shfl_sync(NLP_FULL, d, r);
      const unsigned long long br = __shfl_sync(NLP_FULL, base, r), wr = __shfl_sync(NLP_FULL, wb, r);
      for (uint32_t i = lane; i < nr; i += 32) {
        const uint32_t u = __ldg(keys + wr + i);
        uint32_t j = i + 1;                            // first entry > u (rows are sorted multisets)
        while (j < dr && __ldg(keys + wr + j) == u) ++j;
        su[br + i] = u; sidx[br + i] = (uint32_t)(br + i);
        it_cnt[br + i] = dr - j; it_dw[br + i] = dr; it_ptr[br + i] = wr + j;
      }
    }
  }
}

// After the stable sort by u: wedge count of every item in sorted order, and the flag "first item
// of its source".
__global__ void __launch_bounds__(256) k_plan_heads(const uint32_t* __restrict__ su, const uint32_t* __restrict__ sidx,
                                                    const uint32_t* __restrict__ it_cnt, uint64_t E,
                                                    uint32_t* __restrict__ g_cnt, uint32_t* __restrict__ head) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < E; j += (uint64_t)gridDim.x * blockDim.x) {
    g_cnt[j] = it_cnt[sidx[j]];
    head[j] = (j == 0 || su[j - 1] != su[j]) ? 1u : 0u;
  }
}

// Per source (at its first item): first record slot.  src_roff has nsrc + 1 entries.
__global__ void __launch_bounds__(256) k_plan_sources(const uint32_t* __restrict__ head, const unsigned long long* __restrict__ hs,
                                                      const unsigned long long* __restrict__ rc, uint64_t E, uint64_t nsrc,
                                                      uint64_t P, unsigned long long* __restrict__ src_roff) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < E; j += (uint64_t)gridDim.x * blockDim.x)
    if (head[j]) src_roff[hs[j]] = rc[j];
  if (blockIdx.x == 0 && threadIdx.x == 0) src_roff[nsrc] = P;
}

// Per item: does its source fit a bucket (records <= half)?  Flags for the three compactions.
__global__ void __launch_bounds__(256) k_plan_class(const uint32_t* __restrict__ head, const unsigned long long* __restrict__ hs,
                                                    const unsigned long long* __restrict__ src_roff, const uint32_t* __restrict__ g_cnt,
                                                    uint64_t E, uint32_t half, uint32_t* __restrict__ f_item,
                                                    uint32_t* __restrict__ f_cnt, uint32_t* __restrict__ f_src) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < E; j += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = hs[j] + head[j] - 1ull;
    const bool small = src_roff[k + 1] - src_roff[k] <= (unsigned long long)half;
    f_item[j] = small ? 1u : 0u;
    f_cnt[j] = small ? g_cnt[j] : 0u;
    if (head[j]) f_src[k] = small ? 1u : 0u;
  }
}

struct BucketPlanDev {
  // small sources, ascending u; ns + 1 entries where noted
  const uint32_t* sm_u;                    // [ns]
  const uint32_t* sm_item;                 // [ns + 1] first item (index into the s_* arrays)
  const unsigned long long* sm_soff;       // [ns + 1] first record in the bucket (small-only) record space
  const unsigned long long* sm_roff;       // [ns]     first record slot in the aligned output
  // items of the small sources, source-major, ascending w inside a source
  const uint32_t* s_cnt;                   // [Es]
  const uint32_t* s_dw;                    // [Es]
  const unsigned long long* s_ptr;         // [Es]
  const uint32_t* s_src;                   // [Es] index of the item's source in sm_*
  const unsigned long long* s_loff;        // [Es] first record of the item in the bucket record space
  const uint32_t* bk_first;                // [nb + 1] first small source of every bucket (window of `half` records)
  uint32_t ns;
  uint32_t half;                           // bucket window = half records; CAP = 2 * half
};

struct PlanScatterOut {
  uint32_t* sm_u; uint32_t* sm_item; unsigned long long* sm_soff; unsigned long long* sm_roff;
  uint32_t* s_cnt; uint32_t* s_dw; unsigned long long* s_ptr; uint32_t* s_src; unsigned long long* s_loff;
  uint32_t* b_u; uint32_t* b_cnt; uint32_t* b_dw; unsigned long long* b_ptr; unsigned long long* b_off;
  unsigned long long* bg_first;            // [nbig + 1] first record of a big source in the big record space
  unsigned long long* bg_roff;             // [nbig]     its first slot in the aligned output
  uint32_t* bg_item;                       // [nbig + 1] its first item in the b_* arrays
};

// Split the sorted items into the small-source arrays and the big-source arrays.
__global__ void __launch_bounds__(256) k_plan_scatter(const uint32_t* __restrict__ su, const uint32_t* __restrict__ sidx,
                                                      const uint32_t* __restrict__ it_dw, const unsigned long long* __restrict__ it_ptr,
                                                      const uint32_t* __restrict__ g_cnt, const uint32_t* __restrict__ head,
                                                      const unsigned long long* __restrict__ hs, const unsigned long long* __restrict__ rc,
                                                      const uint32_t* __restrict__ f_item, const unsigned long long* __restrict__ si,
                                                      const unsigned long long* __restrict__ sr, const unsigned long long* __restrict__ ks,
                                                      uint64_t E, uint64_t Es, uint64_t Ps, uint64_t ns, uint64_t nbig, uint64_t P,
                                                      PlanScatterOut o) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < E; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t idx = sidx[j], u = su[j], cnt = g_cnt[j];
    const bool hd = head[j] != 0u;
    const unsigned long long k = hs[j] + (hd ? 1ull : 0ull) - 1ull;
    if (f_item[j]) {
      const unsigned long long t = si[j], kk = ks[k];
      o.s_cnt[t] = cnt; o.s_dw[t] = it_dw[idx]; o.s_ptr[t] = it_ptr[idx]; o.s_src[t] = (uint32_t)kk; o.s_loff[t] = sr[j];
      if (hd) { o.sm_u[kk] = u; o.sm_item[kk] = (uint32_t)t; o.sm_soff[kk] = sr[j]; o.sm_roff[kk] = rc[j]; }
    } else {
      const unsigned long long t = j - si[j], kb = k - ks[k], boff = rc[j] - sr[j];
      o.b_u[t] = u; o.b_cnt[t] = cnt; o.b_dw[t] = it_dw[idx]; o.b_ptr[t] = it_ptr[idx]; o.b_off[t] = boff;
      if (hd) { o.bg_first[kb] = boff; o.bg_roff[kb] = rc[j]; o.bg_item[kb] = (uint32_t)t; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    o.sm_item[ns] = (uint32_t)Es; o.sm_soff[ns] = Ps;
    o.bg_first[nbig] = P - Ps; o.bg_item[nbig] = (uint32_t)(E - Es);
  }
}

// ---- per prediction ----------------------------------------------------------------------------

__host__ __device__ constexpr uint32_t bucket_smem_bytes(bool flt, uint32_t cap) {
  // key[2][cap] u32, tag[2][cap] u16, (pay[2][cap] u32), mask[8][256] u32, hist[2][8][256] u16, scan scratch
  return cap * 4u * 2u + cap * 2u * 2u + (flt ? cap * 4u * 2u : 0u) + BK_WARPS * 256u * 4u + 2u * BK_WARPS * 256u * 2u + 64u;
}

// first index i in [0, n] with a[i] >= x (a ascending, n entries)
__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long* __restrict__ a, uint32_t n, unsigned long long x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Plan: first small source of every bucket window (one thread per window; nb + 1 entries).
__global__ void __launch_bounds__(256) k_plan_buckets(const unsigned long long* __restrict__ sm_soff, uint32_t ns, uint32_t half,
                                                      uint64_t nb, uint32_t* __restrict__ bk_first) {
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= nb; b += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t k = lower_bound_u64(sm_soff, ns + 1u, b * (unsigned long long)half);
    bk_first[b] = k < ns ? k : ns;
  }
}

// Multi-GPU: rank r of `world` owns the buckets [b0, b1) and with them an ascending range of
// source vertices [ulo, uhi); the big sources inside that range are its too, and its slots of the
// aligned output are one contiguous range.  out = {kb0, kb1, slot_lo, slot_hi} (single thread).
__global__ void k_plan_partition(BucketPlanDev pl, const uint32_t* __restrict__ b_u, const uint32_t* __restrict__ bg_item,
                                 const unsigned long long* __restrict__ bg_roff, uint32_t nbig, unsigned long long P,
                                 unsigned long long b0, unsigned long long b1, int first, int last,
                                 unsigned long long* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  const unsigned long long INF = 1ull << 32;
  const uint32_t k0 = first ? 0u : lower_bound_u64(pl.sm_soff, pl.ns + 1u, b0 * pl.half);
  const uint32_t k1 = last ? pl.ns : lower_bound_u64(pl.sm_soff, pl.ns + 1u, b1 * pl.half);
  const unsigned long long ulo = first ? 0ull : (k0 < pl.ns ? (unsigned long long)pl.sm_u[k0] : INF);
  const unsigned long long uhi = last ? INF : (k1 < pl.ns ? (unsigned long long)pl.sm_u[k1] : INF);
  unsigned long long kb[2];
  const unsigned long long bound[2] = {ulo, uhi};
  for (int t = 0; t < 2; ++t) {                      // first big source with u >= bound
    uint32_t lo = 0, hi = nbig;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if ((unsigned long long)b_u[bg_item[mid]] < bound[t]) lo = mid + 1; else hi = mid;
    }
    kb[t] = lo;
  }
  const unsigned long long s0 = k0 < pl.ns ? pl.sm_roff[k0] : P, s1 = k1 < pl.ns ? pl.sm_roff[k1] : P;
  const unsigned long long g0 = kb[0] < nbig ? bg_roff[kb[0]] : P, g1 = kb[1] < nbig ? bg_roff[kb[1]] : P;
  out[0] = kb[0]; out[1] = kb[1];
  out[2] = s0 < g0 ? s0 : g0;
  out[3] = s1 < g1 ? s1 : g1;
}

// One stable LSD pass over the n records in shared memory (warp w ranks records
// [w * 32 * rounds, (w + 1) * 32 * rounds): warp-major order = record order).  Ranks inside a warp
// come from warp-private digit masks (one atomicOr and one load per record, select.cuh); the masks
// clean themselves, and the (warp, digit) counters of the NEXT pass (`s_next`) are zeroed by the
// digit owners while they scan this pass's, so a pass costs three block barriers and no clearing loop.
template <bool FLT, int MAXR>
__device__ __forceinline__ void bucket_sort_pass(const uint32_t* __restrict__ kin, const uint16_t* __restrict__ tin, const uint32_t* __restrict__ pin,
                                                 uint32_t* __restrict__ kout, uint16_t* __restrict__ tout, uint32_t* __restrict__ pout,
                                                 uint32_t n, uint32_t rounds, bool by_tag, int shift,
                                                 uint16_t (*s_hist)[256], uint16_t (*s_next)[256], uint32_t (*s_mask)[256], uint32_t* s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t pre[MAXR];
  uint32_t* my_mask = s_mask[warp];
  uint16_t* my_hist = s_hist[warp];
  #pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if ((uint32_t)r < rounds) {                      // block-uniform
      const uint32_t idx = (uint32_t)warp * (32u * rounds) + (uint32_t)r * 32u + lane;
      const bool valid = idx < n;
      uint32_t d = 0;
      if (valid) d = by_tag ? (((uint32_t)tin[idx] >> shift) & 255u) : ((kin[idx] >> shift) & 255u);
      if (valid) atomicOr(my_mask + d, 1u << lane);
      __syncwarp();
      const unsigned m = valid ? *reinterpret_cast<volatile uint32_t*>(my_mask + d) : 0u;
      const uint32_t before = valid ? *reinterpret_cast<volatile uint16_t*>(my_hist + d) : 0u;
      pre[r] = (d << 16) | (before + __popc(m & lt));
      __syncwarp();
      if (valid && (__ffs(m) - 1) == lane) { my_hist[d] = (uint16_t)(before + __popc(m)); my_mask[d] = 0u; }
      __syncwarp();
    }
  }
  __syncthreads();
  {   // thread t owns digit t: exclusive scan over (digit, warp)
    uint32_t c[BK_WARPS], total = 0;
    #pragma unroll
    for (int w = 0; w < BK_WARPS; ++w) { c[w] = s_hist[w][tid]; total += c[w]; s_next[w][tid] = 0; }
    uint32_t run = sort_block_exclusive(total, s_warp);
    #pragma unroll
    for (int w = 0; w < BK_WARPS; ++w) { s_hist[w][tid] = (uint16_t)run; run += c[w]; }
  }
  __syncthreads();
  #pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if ((uint32_t)r < rounds) {
      const uint32_t idx = (uint32_t)warp * (32u * rounds) + (uint32_t)r * 32u + lane;
      if (idx < n) {
        const uint32_t pos = (uint32_t)s_hist[warp][pre[r] >> 16] + (pre[r] & 0xffffu);
        kout[pos] = kin[idx]; tout[pos] = tin[idx];
        if (FLT) pout[pos] = pin[idx];
      }
    }
  }
  __syncthreads();
}

// One thread block per bucket (window `first_bucket + blockIdx.x` of the bucket record space).
template <bool FLT, uint32_t CAP>
__global__ void __launch_bounds__(BK_THREADS, (CAP * (FLT ? 20u : 12u) + 16384u <= 74000u) ? 3 : 2)
k_bucket(Params p, BucketPlanDev pl, uint64_t first_bucket, int key_passes,
         uint32_t* __restrict__ al_u, uint32_t* __restrict__ al_v, uint32_t* __restrict__ al_c) {
  constexpr int MAXR = CAP / BK_THREADS;
  static_assert(BK_WARPS == SORT_WARPS && BK_THREADS == SORT_THREADS, "sort_block_exclusive is shared with select.cuh");
  extern __shared__ __align__(16) uint32_t bsm[];
  uint32_t* key0 = bsm;
  uint32_t* key1 = key0 + CAP;
  uint32_t* pay0 = key1 + CAP;                                     // FLT only
  uint32_t* pay1 = pay0 + (FLT ? CAP : 0);
  uint16_t* tag0 = reinterpret_cast<uint16_t*>(pay1 + (FLT ? CAP : 0));
  uint16_t* tag1 = tag0 + CAP;
  uint32_t (*s_mask)[256] = reinterpret_cast<uint32_t (*)[256]>(tag1 + CAP);
  uint16_t (*s_hist0)[256] = reinterpret_cast<uint16_t (*)[256]>(s_mask + BK_WARPS);
  uint16_t (*s_hist1)[256] = s_hist0 + BK_WARPS;
  uint32_t* s_warp = reinterpret_cast<uint32_t*>(s_hist1 + BK_WARPS);   // [8] scan scratch, then [8..9] k0, k1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* __restrict__ keys = p.g.keys;

  // ---- the bucket: small sources [k0, k1) that start inside the window --------------------------
  const uint32_t k0 = __ldg(pl.bk_first + first_bucket + blockIdx.x), k1 = __ldg(pl.bk_first + first_bucket + blockIdx.x + 1);
  if (k0 >= k1) return;                                            // block-uniform
  for (int i = tid; i < BK_WARPS * 256; i += BK_THREADS) { (&s_mask[0][0])[i] = 0; (&s_hist0[0][0])[i] = 0; }
  __syncthreads();
  const uint32_t i0 = __ldg(pl.sm_item + k0), i1 = __ldg(pl.sm_item + k1);
  const unsigned long long base = __ldg(pl.sm_soff + k0);
  const uint32_t n = (uint32_t)(__ldg(pl.sm_soff + k1) - base);    // <= CAP by construction
  const uint32_t ns = k1 - k0;
  if (n > CAP || ns > CAP / 2u) {                                  // cannot happen with a consistent plan
    if (tid == 0) atomicAdd(&p.ctr->overflow, 1ull);
    return;
  }

  // ---- gather: warp per 32 items, wedges packed back to back (as k_pair_emit, into shared memory)
  for (uint32_t tile = warp; tile * 32u < i1 - i0; tile += BK_WARPS) {
    const uint32_t e = i0 + tile * 32u + lane;
    uint32_t cnt = 0, tg = 0, dw = 0;
    unsigned long long ptr = 0, loff = 0;
    if (e < i1) {
      cnt = __ldg(pl.s_cnt + e); ptr = __ldg(pl.s_ptr + e); tg = __ldg(pl.s_src + e) - k0; loff = __ldg(pl.s_loff + e);
      if (FLT) dw = __ldg(pl.s_dw + e);
    }
    const uint32_t out0 = (uint32_t)(__shfl_sync(NLP_FULL, loff, 0) - base);
    uint32_t inc = cnt;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
      if (lane >= d) inc += t;
    }
    const uint32_t tot = __shfl_sync(NLP_FULL, inc, 31);
    for (uint32_t sb = 0; sb < tot; sb += 32u * BK_GATHER) {   // BK_GATHER independent key loads in flight per lane
      uint32_t v[BK_GATHER], tgs[BK_GATHER], dws[BK_GATHER];
      #pragma unroll
      for (int q = 0; q < BK_GATHER; ++q) {
        const uint32_t idx = sb + q * 32u + lane;
        int j = 0;                                    // smallest j with inc[j] > idx
        #pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const uint32_t x = __shfl_sync(NLP_FULL, inc, j + step - 1);
          if (x <= idx) j += step;
        }
        const uint32_t incj = __shfl_sync(NLP_FULL, inc, j);
        const uint32_t cntj = __shfl_sync(NLP_FULL, cnt, j);
        tgs[q] = __shfl_sync(NLP_FULL, tg, j);
        const unsigned long long ptrj = __shfl_sync(NLP_FULL, ptr, j);
        dws[q] = FLT ? __shfl_sync(NLP_FULL, dw, j) : 0u;
        v[q] = idx < tot ? __ldg(keys + ptrj + (idx - (incj - cntj))) : 0u;
      }
      #pragma unroll
      for (int q = 0; q < BK_GATHER; ++q) {
        const uint32_t idx = sb + q * 32u + lane;
        const uint32_t o = out0 + idx;
        if (idx < tot && o < CAP) {
          key0[o] = v[q]; tag0[o] = (uint16_t)tgs[q];
          if (FLT) pay0[o] = dws[q];
        }
      }
    }
  }
  __syncthreads();

  // ---- stable LSD radix sort by (local source, v) ---------------------------------------------
  const uint32_t rounds = (n + BK_THREADS - 1) / BK_THREADS;
  uint32_t *ka = key0, *kb = key1, *pa = pay0, *pb = pay1;
  uint16_t *ta = tag0, *tb = tag1;
  uint16_t (*ha)[256] = s_hist0, (*hb)[256] = s_hist1;
  const int tag_passes = ns > 256u ? 2 : (ns > 1u ? 1 : 0);
  for (int ps = 0; ps < key_passes + tag_passes; ++ps) {
    const bool by_tag = ps >= key_passes;
    bucket_sort_pass<FLT, MAXR>(ka, ta, pa, kb, tb, pb, n, rounds, by_tag, (by_tag ? ps - key_passes : ps) * 8, ha, hb, s_mask, s_warp);
    uint32_t* t1 = ka; ka = kb; kb = t1; uint16_t* t2 = ta; ta = tb; tb = t2; t1 = pa; pa = pb; pb = t1;
    uint16_t (*t3)[256] = ha; ha = hb; hb = t3;
  }
  // (ka, ta, pa) hold the sorted records; the other key / tag buffers are free now and take the
  // per-source caches: source vertex, and (aligned slot of the source's first record) - (its
  // local index) as two 32-bit halves.  ns <= CAP / 2.
  uint32_t* c_u = kb;                                              // [CAP / 2]
  uint32_t* c_dlo = kb + CAP / 2;                                  // [CAP / 2]
  uint32_t* c_dhi = reinterpret_cast<uint32_t*>(tb);               // [CAP / 2] (CAP u16 = CAP / 2 u32)
  for (uint32_t i = tid; i < ns; i += BK_THREADS) {
    c_u[i] = __ldg(pl.sm_u + k0 + i);
    const unsigned long long delta = __ldg(pl.sm_roff + k0 + i) - (__ldg(pl.sm_soff + k0 + i) - base);
    c_dlo[i] = (uint32_t)delta; c_dhi[i] = (uint32_t)(delta >> 32);
  }
  __syncthreads();

  // ---- reduce runs: record-aligned (u, v, count) at the slot of a run's first record --------------
  // Exclusion and scoring need a binary search in row u and degree gathers in global memory; with
  // the few warps a shared-memory-heavy block leaves resident they would be latency bound here,
  // so they run in k_score with one thread per slot.
  for (uint32_t r = 0; r < rounds; ++r) {
    const uint32_t idx = r * BK_THREADS + tid;
    if (idx >= n) break;
    const uint32_t v = ka[idx], tg = ta[idx];
    const bool head = idx == 0 || ka[idx - 1] != v || ta[idx - 1] != tg;
    uint32_t c = BK_NONE;
    if (head) {
      if (FLT) {   // inc/predict.hxx:788,828: acc = float(double(acc) + term), ascending w
        float acc = 0.0f;
        uint32_t j = idx;
        do {
          acc = __double2float_rn(__dadd_rn((double)acc, flt_term(p, pa[j])));
          ++j;
        } while (j < n && ka[j] == v && ta[j] == tg);
        c = __float_as_uint(acc);
      } else {
        uint32_t j = idx + 1;
        while (j < n && ka[j] == v && ta[j] == tg) ++j;
        c = j - idx;
      }
    }
    const unsigned long long slot = (((unsigned long long)c_dhi[tg] << 32) | c_dlo[tg]) + idx;
    al_c[slot] = c;
    if (head) { al_u[slot] = c_u[tg]; al_v[slot] = v; }
  }
}

// Is v an entry of the sorted row [ub, ub + du)?  Quaternary search: three probes per round are
// independent loads, so the dependent chain is log4 instead of log2 of the row length (k_score is
// bound by exactly this latency chain: ncu, long-scoreboard stall 24.6 per issue).
__device__ __forceinline__ bool row_contains_k4(const uint32_t* __restrict__ keys, uint64_t ub, uint32_t du, uint32_t v) {
  uint32_t lo = 0, hi = du;                           // the first entry >= v, if any, lies in [lo, hi)
  while (hi - lo > 4u) {
    const uint32_t q = (hi - lo) >> 2;
    const uint32_t m1 = lo + q, m2 = lo + 2u * q, m3 = lo + 3u * q;
    const uint32_t k1 = __ldg(keys + ub + m1), k2 = __ldg(keys + ub + m2), k3 = __ldg(keys + ub + m3);
    if (k2 < v) { if (k3 < v) lo = m3 + 1u; else { lo = m2 + 1u; hi = m3 + 1u; } }
    else        { if (k1 < v) { lo = m1 + 1u; hi = m2 + 1u; } else hi = m1 + 1u; }
  }
  bool found = false;
  #pragma unroll
  for (uint32_t i = 0; i < 4u; ++i)
    if (lo + i < hi) found |= __ldg(keys + ub + lo + i) == v;
  return found;
}

// One thread per slot of the aligned arrays: existing-edge exclusion by binary search in row u
// (inc/predict.hxx:306-307: such pairs keep their candidate slot with value 0), fused scoring
// (inc/predict.hxx:309-311).  al_c: count (or float accumulator bits) at the first record of a
// run, BK_NONE where there is no pair, BK_DONE where the score is already in al_s (big sources).
// exclude = false: the counts already went through the exclusion (reuse store, nlp_set_reuse);
// capture: write the count after the exclusion back (it is about to be stored).
// (A two-slots-per-thread variant with interleaved searches measured 0.9 ms per step SLOWER on
// BASELINE configs[1] -- profiles/r02_summary.md -- and was dropped.)
template <bool FLT>
__global__ void __launch_bounds__(256) k_score(Params p, const uint32_t* __restrict__ al_u, const uint32_t* __restrict__ al_v,
                                               uint32_t* al_c, uint32_t* __restrict__ al_s,
                                               uint64_t lo, uint64_t hi, Select11* sel, bool exclude, bool capture) {
  // first level of the top-K select (select.cuh, Select11): histogram of the top 11 bits of every
  // kept score, accumulated here so that the select does not have to read the scores again
  __shared__ uint32_t s_h[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) s_h[i] = 0;
  __syncthreads();
  Tally tally;
  const uint64_t n = hi - lo, n32 = (n + 31u) & ~31ull;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n32; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = lo + t;
    const uint32_t c = t < n ? al_c[i] : BK_DONE;
    const bool head = c != BK_NONE && c != BK_DONE;
    uint32_t u = 0, v = 0, cnt = 0;
    uint64_t du = 0;
    float acc = 0.0f;
    if (head) {
      u = al_u[i]; v = al_v[i];
      if (FLT) acc = __uint_as_float(c); else cnt = c;
      const uint64_t ub = __ldg(p.g.off + u);
      du = __ldg(p.g.deg + u);
      if (exclude && row_contains_k4(p.g.keys, ub, (uint32_t)du, v)) { cnt = 0; acc = 0.0f; }
      if (!FLT && capture) al_c[i] = cnt;              // the reuse store keeps the count after the exclusion
    }
    float score;
    const bool keep = score_pair(p, head, u, du, v, cnt, acc, tally, &score);
    uint32_t sb = keep ? __float_as_uint(score) : NLP_NO_SCORE;
    if (c != BK_DONE) al_s[i] = sb;
    else if (t < n) sb = al_s[i];                      // big source: scored by k_pair_reduce already
    if (sb != NLP_NO_SCORE) atomicAdd(&s_h[desc_key(sb) >> 21], 1u);
  }
  tally.flush(p.ctr);
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (s_h[i]) atomicAdd(&sel->hist[i], (unsigned long long)s_h[i]);
}

// Big sources: their records were sorted and reduced by the global machinery of pairs.cuh in the
// candidate buffers (record j of the big record space, local index j - first); copy them into
// their slots of the aligned output.
__global__ void __launch_bounds__(256) k_big_place(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv,
                                                   const uint32_t* __restrict__ ps, const uint32_t* __restrict__ pc, uint64_t n, uint64_t first,
                                                   const unsigned long long* __restrict__ bg_first,
                                                   const unsigned long long* __restrict__ bg_roff, uint32_t kb0, uint32_t kb1,
                                                   uint32_t* __restrict__ al_u, uint32_t* __restrict__ al_v, uint32_t* __restrict__ al_s,
                                                   uint32_t* __restrict__ al_c) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long pos = first + j;
    uint32_t lo = kb0, hi = kb1;                     // last big source with bg_first <= pos
    while (lo + 1 < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (__ldg(bg_first + mid) <= pos) lo = mid; else hi = mid;
    }
    const unsigned long long slot = __ldg(bg_roff + lo) + (pos - __ldg(bg_first + lo));
    const uint32_t s = ps[j];
    al_s[slot] = s;
    const uint32_t c = pc ? pc[j] : BK_DONE;           // with the counts (reuse store) k_score scores the slot again
    al_c[slot] = c;
    if (s != NLP_NO_SCORE || (pc && c != BK_NONE)) { al_u[slot] = pu[j]; al_v[slot] = pv[j]; }
  }
}

}  // namespace nlp
