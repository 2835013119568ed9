// Device-wide exclusive prefix sum (hand-written, three launches: tile sums, spine, tile scan).
// Used for the long-row chunk table of the frontier and for the output offsets of the pair path
// (both replace what the reference gets for free from its serial loops, inc/predict.hxx:287-338).
#pragma once
#include "common.cuh"

namespace nlp {

enum { SCAN_THREADS = 256, SCAN_PER_THREAD = 8, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD };

__device__ __forceinline__ unsigned long long warp_inclusive_u64(unsigned long long x) {
  const int lane = threadIdx.x & 31;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(NLP_FULL, x, d);
    if (lane >= d) x += t;
  }
  return x;
}

// Block-wide exclusive scan of one value per thread (blockDim.x == SCAN_THREADS); returns the
// exclusive prefix of this thread, *total receives the block sum.
__device__ __forceinline__ unsigned long long block_exclusive_u64(unsigned long long x, unsigned long long* total) {
  __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
  __shared__ unsigned long long s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long inc = warp_inclusive_u64(x);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0ull;
    const unsigned long long winc = warp_inclusive_u64(w);
    if (lane < SCAN_THREADS / 32) s_warp[lane] = winc - w;
    if (lane == SCAN_THREADS / 32 - 1) s_total = winc;
  }
  __syncthreads();
  const unsigned long long r = s_warp[warp] + inc - x;
  *total = s_total;
  __syncthreads();
  return r;
}

template <class TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const TIn* __restrict__ in, uint64_t n,
                                                             unsigned long long* __restrict__ tile_sums) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_PER_THREAD;
  unsigned long long s = 0;
  #pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k)
    if (base + k < n) s += (unsigned long long)in[base + k];
  unsigned long long total;
  block_exclusive_u64(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single block: exclusive scan of the tile sums in place; the grand total goes to *total_out.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_spine(unsigned long long* __restrict__ tile_sums, uint32_t ntiles,
                                                             unsigned long long* __restrict__ total_out) {
  unsigned long long run = 0;
  for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long x = i < ntiles ? tile_sums[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_u64(x, &total);
    if (i < ntiles) tile_sums[i] = run + ex;
    run += total;
  }
  if (threadIdx.x == 0) *total_out = run;
}

// out[i] = sum of in[0..i) ; in and out may alias when TIn is 64-bit.
template <class TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TIn* in, uint64_t n,
                                                             const unsigned long long* __restrict__ tile_sums,
                                                             unsigned long long* out) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_PER_THREAD;
  unsigned long long x[SCAN_PER_THREAD];
  unsigned long long s = 0;
  #pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    x[k] = base + k < n ? (unsigned long long)in[base + k] : 0ull;
    s += x[k];
  }
  unsigned long long total;
  unsigned long long run = tile_sums[blockIdx.x] + block_exclusive_u64(s, &total);
  #pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    if (base + k < n) out[base + k] = run;
    run += x[k];
  }
}

}  // namespace nlp
