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

}  // namespace nlp
