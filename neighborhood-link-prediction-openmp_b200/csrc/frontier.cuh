// Kernel (a) of the north star: the work-balanced frontier.
//   k_degrees : deg[u] = off[u+1]-off[u]                       (once per graph)
//   k_elig    : LHub eligibility bitmask, bit w = deg(w) <= D  (replaces the dependent random
//               degree load of inc/predict.hxx:301 by an L2-resident 1-bit lookup)
//   k_work    : work(u) = sum_{eligible w in N(u)} deg(w)      (wedges the reference would scan
//               for u, inc/predict.hxx:298-304) -- one coalesced pass over the adjacency
//   k_bin     : drop zero-work sources (inc/predict.hxx:287-289 visits them all) and bin the
//               rest by work into the sub-warp / warp / block-hash / dense-spill paths
#pragma once
#include "common.cuh"

namespace nlp {

__global__ void __launch_bounds__(256) k_degrees(const uint64_t* __restrict__ off, uint32_t S,
                                                 uint32_t* __restrict__ deg, uint32_t* maxdeg) {
  uint32_t local = 0;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < S; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t d = (uint32_t)(off[u + 1] - off[u]);
    deg[u] = d;
    local = max(local, d);
  }
  local = __reduce_max_sync(NLP_FULL, local);
  if ((threadIdx.x & 31) == 0 && local) atomicMax(maxdeg, local);
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

// One warp streams the adjacency of 32 consecutive rows (one contiguous, coalesced range of
// `keys`), looks up eligibility of every first-hop entry and segment-sums deg(w) per row.
template <bool LHUB>
__global__ void __launch_bounds__(256) k_work(DevGraph g, const uint32_t* __restrict__ elig, int rank, int world,
                                              uint32_t* __restrict__ work, Counters* ctr) {
  __shared__ unsigned long long acc[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t nrb = ((uint64_t)g.S + 31u) >> 5;
  unsigned long long t_first = 0, t_elig = 0, t_wedges = 0;
  for (uint64_t rb = blockIdx.x * 8ull + wib; rb < nrb; rb += gridDim.x * 8ull) {
    if (!owns_row_block(rb, rank, world)) continue;
    const uint64_t row = rb * 32u + lane;
    const uint64_t r0 = row < g.S ? row : g.S;
    const uint64_t r1 = row < g.S ? row + 1 : g.S;
    const uint64_t b0 = g.off[r0];
    const uint64_t e_end = g.off[r1];
    const uint64_t B = __shfl_sync(NLP_FULL, b0, 0);
    const uint64_t E = __shfl_sync(NLP_FULL, e_end, 31);
    acc[wib][lane] = 0;
    __syncwarp();
    for (uint64_t base = B; base < E; base += 32) {
      const uint64_t idx = base + lane;
      const bool valid = idx < E;
      uint32_t c = 0;
      bool el = false;
      if (valid) {
        const uint32_t w = __ldg(g.keys + idx);
        el = LHUB ? (((__ldg(elig + (w >> 5)) >> (w & 31)) & 1u) != 0u) : true;
        if (el) c = __ldg(g.deg + w);
      }
      // row of this entry: smallest j with e_end[j] > idx
      int j = 0;
      #pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const uint64_t x = __shfl_sync(NLP_FULL, e_end, j + step - 1);
        if (x <= idx) j += step;
      }
      t_first += valid ? 1u : 0u;
      t_elig += el ? 1u : 0u;
      t_wedges += c;
      const int j0 = __shfl_sync(NLP_FULL, j, 0);
      const bool uni = __all_sync(NLP_FULL, !valid || j == j0);
      if (uni) {   // whole batch inside one (long) row: one shared-memory update
        const uint32_t lo = __reduce_add_sync(NLP_FULL, c & 0xffffu);
        const uint32_t hi = __reduce_add_sync(NLP_FULL, c >> 16);
        if (lane == 0) acc[wib][j0] += (unsigned long long)lo + ((unsigned long long)hi << 16);
      } else if (c) {
        atomicAdd(&acc[wib][j], (unsigned long long)c);
      }
      __syncwarp();
    }
    const unsigned long long a = acc[wib][lane];
    if (row < g.S) work[row] = a > 0xffffffffull ? 0xffffffffu : (uint32_t)a;
    __syncwarp();
  }
  #pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    t_first  += __shfl_xor_sync(NLP_FULL, t_first, d);
    t_elig   += __shfl_xor_sync(NLP_FULL, t_elig, d);
    t_wedges += __shfl_xor_sync(NLP_FULL, t_wedges, d);
  }
  if (lane == 0) {
    if (t_first)  atomicAdd(&ctr->first_hop, t_first);
    if (t_elig)   atomicAdd(&ctr->eligible_first_hop, t_elig);
    if (t_wedges) atomicAdd(&ctr->wedges, t_wedges);
  }
}

struct BinLists { uint32_t* list[NBINS]; };

__device__ __forceinline__ int choose_bin(uint32_t work, uint32_t bound, uint32_t du) {
  if (work <= 8u && du <= 64u) return 0;
  if (work <= 32u && du <= 256u) return 1;
  if (bound <= bin_limit(2)) return 2;
  if (bound <= bin_limit(3)) return 3;
  if (bound <= bin_limit(4)) return 4;
  return 5;
}

__global__ void __launch_bounds__(256) k_bin(DevGraph g, const uint32_t* __restrict__ work, int rank, int world,
                                             BinLists bl, Counters* ctr) {
  const int lane = threadIdx.x & 31;
  const uint64_t S32 = ((uint64_t)g.S + 31u) & ~31ull;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < S32; u += (uint64_t)gridDim.x * blockDim.x) {
    int bin = -1;
    uint32_t need = 0;
    if (u < g.S && owns_row_block(u >> 5, rank, world)) {
      const uint32_t w = work[u];
      if (w) {
        const uint32_t room = g.S - 1u - (uint32_t)u;
        const uint32_t bound = w < room ? w : room;
        bin = choose_bin(w, bound, g.deg[u]);
        need = bin < 2 ? w : bound;
        if (bound == 0) bin = -1;    // last vertex: no v > u exists
      }
    }
    const unsigned any = __ballot_sync(NLP_FULL, bin >= 0);
    if (!any) continue;
    #pragma unroll
    for (int b = 0; b < NBINS; ++b) {
      const unsigned m = __ballot_sync(NLP_FULL, bin == b);
      if (!m) continue;
      const int leader = __ffs(m) - 1;
      unsigned long long sum = (bin == b) ? need : 0;
      unsigned long long mx = (bin == b) ? need : 0;
      #pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        sum += __shfl_xor_sync(NLP_FULL, sum, d);
        const unsigned long long o = __shfl_xor_sync(NLP_FULL, mx, d);
        mx = o > mx ? o : mx;
      }
      unsigned long long base = 0;
      if (lane == leader) {
        base = atomicAdd(&ctr->bin_count[b], (unsigned long long)__popc(m));
        atomicAdd(&ctr->bin_bound[b], sum);
        if (b == 5) atomicMax(&ctr->max_bound, mx);
      }
      base = __shfl_sync(NLP_FULL, base, leader);
      if (bin == b) bl.list[b][base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)u;
    }
    if (lane == 0) atomicAdd(&ctr->frontier, (unsigned long long)__popc(any));
  }
}

}  // namespace nlp
