// Random edge removal on the device: the step main.cxx runs before every
// sweep of predictions -- generateEdgeDeletions + tidyBatchUpdateU, inc/batch.hxx:99-112, 200-208,
// called from runBatches, main.cxx:165-168 -- reproduced draw for draw for a given
// std::default_random_engine seed (parity anchors: oracle/batch_oracle.c, tests/batch_parallel.py).
//
// The reference draws the batch from ONE sequential stream, and a deletion consumes a
// data-dependent amount of it: u = K(1 + (span-1) * dis(rnd)); an isolated u fails without a
// second draw and is retried (five tries); otherwise vi = K(dis(rnd) * deg(u)) picks the vi-th
// entry of row u.  Two facts make it parallel:
//   * default_random_engine is the Lehmer generator x <- 16807 x mod (2^31 - 1), so the engine
//     output number k is seed * 16807^k mod (2^31 - 1): any position on its own, one modular
//     exponentiation; a double of uniform_real_distribution (libstdc++ generate_canonical<double,
//     53>) is made from two consecutive outputs -- "slot" p uses outputs 2p+1 and 2p+2;
//   * what a deletion does is a function of the slot it starts at.  k_batch_slots / k_batch_next
//     compute that function for every slot, the starts of the batch are the orbit of slot 0
//     (pointer doubling: k_batch_extend / k_batch_square), k_batch_emit draws the edges.
// The pairs are then sorted by (u, v) with the radix sort of select.cuh and made unique.
#pragma once
#include "common.cuh"

namespace nlp {

constexpr uint64_t LEHMER_M = 2147483647ull;          // 2^31 - 1
constexpr uint64_t LEHMER_A = 16807ull;
constexpr uint32_t BATCH_NONE = 0xffffffffu;

__device__ __forceinline__ uint64_t lehmer_pow(uint64_t e) {          // 16807^e mod (2^31 - 1)
  uint64_t r = 1, b = LEHMER_A;
  while (e) {
    if (e & 1ull) r = (r * b) % LEHMER_M;
    b = (b * b) % LEHMER_M;
    e >>= 1;
  }
  return r;
}

// Slot p: the double generate_canonical<double, 53> makes from engine outputs 2p+1 and 2p+2
// (bits/random.tcc: sum = (x1 - min) + (x2 - min) * r with r = max - min + 1 = 2147483646, the
// product rounded before the add, divided by double(r * r) = 2^62 - 2^33; >= 1 becomes
// nextafter(1, 0)), and the vertex an attempt starting here draws: K(1 + (span - 1) * d)
// (inc/batch.hxx:56), BATCH_NONE when that vertex has no entries (inc/batch.hxx:32).
__global__ void __launch_bounds__(256) k_batch_slots(const uint32_t* __restrict__ deg, uint32_t S, uint32_t seed0, uint64_t P,
                                                     double* __restrict__ dval, uint32_t* __restrict__ uat) {
  for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t x1 = ((uint64_t)seed0 * lehmer_pow(2ull * p + 1ull)) % LEHMER_M;
    const uint64_t x2 = (x1 * LEHMER_A) % LEHMER_M;
    const double prod = __dmul_rn((double)(x2 - 1ull), 2147483646.0);
    const double sum = __dadd_rn((double)(x1 - 1ull), prod);
    double d = __ddiv_rn(sum, 4611686009837453312.0);
    if (d >= 1.0) d = __longlong_as_double(0x3fefffffffffffffll);
    dval[p] = d;
    const uint32_t u = (uint32_t)__dadd_rn(1.0, __dmul_rn((double)(S - 1u), d));
    uat[p] = (u < S && deg[u] != 0u) ? u : BATCH_NONE;
  }
}

// A deletion that starts at slot p: hit[p] = slot of its first successful vertex draw among
// p .. p+4 (retry(..., 5), inc/_utility.hxx:198-203), BATCH_NONE if all five fail;
// next[p] = slot where the following deletion starts (a success also uses the slot after it).
__global__ void __launch_bounds__(256) k_batch_next(const uint32_t* __restrict__ uat, uint64_t P,
                                                    uint32_t* __restrict__ hit, uint32_t* __restrict__ next) {
  for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h = BATCH_NONE;
    #pragma unroll
    for (int k = 4; k >= 0; --k)
      if (p + k < P && uat[p + k] != BATCH_NONE) h = (uint32_t)(p + k);
    hit[p] = h;
    const uint64_t n = h != BATCH_NONE ? (uint64_t)h + 2ull : p + 5ull;
    next[p] = (uint32_t)(n < P ? n : P - 1ull);        // clamped slots lie beyond anything the batch can reach
  }
}

// starts[known + l] = jump[starts[l]] for l < take, where jump = next^known.
__global__ void __launch_bounds__(256) k_batch_extend(uint32_t* __restrict__ starts, uint64_t known, uint64_t take,
                                                      const uint32_t* __restrict__ jump) {
  for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < take; l += (uint64_t)gridDim.x * blockDim.x)
    starts[known + l] = jump[starts[l]];
}

__global__ void __launch_bounds__(256) k_batch_square(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t P) {
  for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += (uint64_t)gridDim.x * blockDim.x)
    out[p] = in[in[p]];
}

// Deletion l: the vi-th entry of row u, vi = K(dis(rnd) * deg(u)) from the slot after the vertex
// draw (inc/batch.hxx:33-38); both directions (inc/batch.hxx:105-106).  A deletion whose five
// tries all failed writes two BATCH_NONE pairs, which sort to the end and are dropped.
__global__ void __launch_bounds__(256) k_batch_emit(DevGraph g, const uint32_t* __restrict__ starts, uint64_t B,
                                                    const uint32_t* __restrict__ hit, const uint32_t* __restrict__ uat,
                                                    const double* __restrict__ dval,
                                                    uint32_t* __restrict__ pu, uint32_t* __restrict__ pv) {
  for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < B; l += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t h = hit[starts[l]];
    uint32_t u = BATCH_NONE, v = BATCH_NONE;
    if (h != BATCH_NONE) {
      u = uat[h];
      const uint32_t d = g.deg[u];
      const uint32_t vi = (uint32_t)__dmul_rn(dval[h + 1], (double)d);
      v = __ldg(g.keys + __ldg(g.off + u) + (vi < d ? vi : d - 1u));
    }
    pu[2 * l] = u;     pv[2 * l] = v;
    pu[2 * l + 1] = v; pv[2 * l + 1] = u;
  }
}

// Pairs sorted by (u, v): flag the first of every run of equal pairs (uniqueEdgesU,
// inc/batch.hxx:186-193), not the BATCH_NONE fillers.
__global__ void __launch_bounds__(256) k_batch_heads(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, uint64_t n,
                                                     uint32_t* __restrict__ head) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = pu[i], v = pv[i];
    head[i] = (u != BATCH_NONE && (i == 0 || pu[i - 1] != u || pv[i - 1] != v)) ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256) k_batch_compact(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, uint64_t n,
                                                       const uint32_t* __restrict__ head, const unsigned long long* __restrict__ pos,
                                                       uint32_t* __restrict__ ou, uint32_t* __restrict__ ov) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (head[i]) { ou[pos[i]] = pu[i]; ov[pos[i]] = pv[i]; }
}


// ---- apply a batch of deletions to the resident CSR (inc/batch.hxx:239-247, main.cxx:169) --------
// applyBatchUpdateOmpU calls removeEdge(u, v) for every pair of the (unique) list and updates the
// graph; one stored copy goes per request (a duplicated entry of a multiset row survives its own
// removal, _algorithm.hxx:132-139).  Here: mark the first stored copy of every request by binary
// search in row u (one bit per entry), count the marks per row, scan the new degrees into the
// new offsets, and compact the keys tile by tile.
enum { DEL_THREADS = 256, DEL_PER_THREAD = 8, DEL_TILE = DEL_THREADS * DEL_PER_THREAD };

__global__ void __launch_bounds__(256) k_del_mark(DevGraph g, const uint32_t* __restrict__ du, const uint32_t* __restrict__ dv,
                                                  uint64_t n, uint32_t* __restrict__ bits, uint32_t* __restrict__ marks) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = du[i], v = dv[i];
    if (u >= g.S) continue;
    const uint64_t ub = __ldg(g.off + u);
    const uint32_t d = __ldg(g.deg + u);
    uint32_t lo = 0, hi = d;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(g.keys + ub + mid) < v) lo = mid + 1; else hi = mid;
    }
    if (lo >= d || __ldg(g.keys + ub + lo) != v) continue;            // not stored: ignored, as removeEdge does
    const uint64_t e = ub + lo;
    const uint32_t bit = 1u << (e & 31u);
    const uint32_t old = atomicOr(bits + (e >> 5), bit);
    if (!(old & bit)) atomicAdd(marks + u, 1u);                       // a repeated request removes nothing more
  }
}

__global__ void __launch_bounds__(256) k_del_newdeg(const uint32_t* __restrict__ deg, uint32_t* __restrict__ marks, uint32_t S) {
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < S; u += (uint64_t)gridDim.x * blockDim.x)
    marks[u] = deg[u] - marks[u];
}

// kept entries per tile; a thread owns 8 consecutive entries = one byte of the bit array
__global__ void __launch_bounds__(DEL_THREADS) k_del_count(const uint8_t* __restrict__ bits, uint64_t M, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t s_warp[DEL_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * DEL_TILE + (uint64_t)threadIdx.x * DEL_PER_THREAD;
  uint32_t c = 0;
  if (base < M) {
    const uint32_t left = (uint32_t)(M - base < DEL_PER_THREAD ? M - base : DEL_PER_THREAD);
    const uint32_t gone = bits[base >> 3] & ((1u << left) - 1u);
    c = left - __popc(gone);
  }
  c = __reduce_add_sync(NLP_FULL, c);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    #pragma unroll
    for (int w = 0; w < DEL_THREADS / 32; ++w) t += s_warp[w];
    tile_counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(DEL_THREADS) k_del_write(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ bits, uint64_t M,
                                                            const unsigned long long* __restrict__ tile_off, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_warp[DEL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * DEL_TILE + (uint64_t)threadIdx.x * DEL_PER_THREAD;
  uint32_t left = 0, gone = 0, c = 0;
  if (base < M) {
    left = (uint32_t)(M - base < DEL_PER_THREAD ? M - base : DEL_PER_THREAD);
    gone = bits[base >> 3] & ((1u << left) - 1u);
    c = left - __popc(gone);
  }
  uint32_t inc = c;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
  #pragma unroll
  for (int w = 0; w < DEL_THREADS / 32; ++w) before += w < warp ? s_warp[w] : 0u;
  unsigned long long pos = tile_off[blockIdx.x] + before + inc - c;
  for (uint32_t k = 0; k < left; ++k)
    if (!((gone >> k) & 1u)) out[pos++] = keys[base + k];
}


// First stored copy of b in the sorted row a; false when b is not stored.
__device__ __forceinline__ bool del_first_copy(const DevGraph& g, uint32_t a, uint32_t b, uint64_t* e) {
  const uint64_t ab = __ldg(g.off + a);
  const uint32_t d = __ldg(g.deg + a);
  uint32_t lo = 0, hi = d;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(g.keys + ab + mid) < b) lo = mid + 1; else hi = mid;
  }
  *e = ab + lo;
  return lo < d && __ldg(g.keys + ab + lo) == b;
}

// Does the batch keep symmetric rows symmetric?  For every request (u, v) that removes an entry, the
// first stored copy of the mirror entry (v, u) must be marked too (then both rows lose one copy of
// the pair).  Runs after k_del_mark on a graph whose rows ARE symmetric; *asym != 0 afterwards = no.
__global__ void __launch_bounds__(256) k_del_mirror(DevGraph g, const uint32_t* __restrict__ du, const uint32_t* __restrict__ dv,
                                                    uint64_t n, const uint32_t* __restrict__ bits, unsigned int* __restrict__ asym) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = du[i], v = dv[i];
    if (u >= g.S || v >= g.S) continue;
    uint64_t e1 = 0, e2 = 0;
    if (!del_first_copy(g, u, v, &e1)) continue;                     // the request removes nothing
    if (!del_first_copy(g, v, u, &e2) || !((bits[e2 >> 5] >> (e2 & 31u)) & 1u)) atomicOr(asym, 1u);
  }
}

}  // namespace nlp
