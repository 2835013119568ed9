// On-device evaluation of a prediction against the held-back edges: the step main.cxx runs on
// the host after every prediction (PREDICT_LINKS, main.cxx:48-57):
//
//   insertions1 = both directions of every predicted edge, sorted, unique      (main.cxx:51-54, 97-104)
//   common1     = set_intersection(insertions0, insertions1)                   (main.cxx:55, 126-133)
//   precision   = |common1| / max(|insertions1|, 1),  recall = |common1| / max(|insertions0|, 1)
//                                                                              (main.cxx:201-202)
//
// insertions0 is the sorted directed list of the removed edges (main.cxx:206-207).  The predicted
// pairs are distinct with u < v, so insertions1 holds 2 x count distinct directed edges and the
// intersection is the number of those found in insertions0 (a directed edge that insertions0
// lists several times still matches once: insertions1 has it once).  The result never leaves the
// GPU: one binary search per direction in the packed truth keys.
#pragma once
#include "common.cuh"

namespace nlp {

// key[i] = u[i] << 32 | v[i]; *unsorted != 0 afterwards when the list is not ascending by (u, v).
__global__ void __launch_bounds__(256) k_truth_pack(const uint32_t* __restrict__ tu, const uint32_t* __restrict__ tv, uint64_t n,
                                                    unsigned long long* __restrict__ key, unsigned int* __restrict__ unsorted) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = ((unsigned long long)tu[i] << 32) | (unsigned long long)tv[i];
    key[i] = k;
    if (i) {
      const unsigned long long prev = ((unsigned long long)tu[i - 1] << 32) | (unsigned long long)tv[i - 1];
      if (prev > k) atomicOr(unsorted, 1u);
    }
  }
}

__device__ __forceinline__ bool truth_contains(const unsigned long long* __restrict__ key, uint64_t n, unsigned long long x) {
  uint64_t lo = 0, hi = n;                            // first position with key >= x
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(key + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(key + lo) == x;
}

// common += #{(u, v) predicted : (u, v) in truth} + #{... : (v, u) in truth}
__global__ void __launch_bounds__(256) k_evaluate(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, uint64_t count,
                                                  const unsigned long long* __restrict__ key, uint64_t n,
                                                  unsigned long long* __restrict__ common) {
  unsigned int c = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long u = pu[i], v = pv[i];
    c += truth_contains(key, n, (u << 32) | v) ? 1u : 0u;
    c += truth_contains(key, n, (v << 32) | u) ? 1u : 0u;
  }
  c = __reduce_add_sync(NLP_FULL, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(common, (unsigned long long)c);
}

}  // namespace nlp
