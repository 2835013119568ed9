// Graph ingest on the device: the text of a Matrix Market coordinate file
// -> the CSR of the graph main.cxx:243-245 builds on the host with
//   readMtxOmpW (inc/mtx.hxx:151-188: strtoull per line, both directions for a symmetric header),
//   symmetrizeOmp (inc/symmetrize.hxx:71-82: add the reverse of every edge),
//   removeSelfLoopsOmpU (inc/selfLoop.hxx:117-124),
// on a DiGraph: repeated lines collapse when the reader's rows are first sorted, but the merge of the
// reverse edges leaves some entries twice (k_ing_class) -- entry for entry the reference's graph.
//
//   k_mtx_count    line starts per 8 KB tile of the text
//   k_mtx_parse    one thread per 32 bytes: every line that starts there is parsed (two unsigned
//                  integers; the rest of the line -- a weight -- is ignored) into slot
//                  (tile base + rank in the tile) of the edge arrays; blank / comment lines leave u = 0
//   k_ing_count    directed pairs a line contributes (0 for skipped lines, 2 when both directions are
//                  wanted) -> scan -> k_ing_emit writes them with a tag: stored by the reader / added
//                  by symmetrizeOmp
//   radix sort     stable LSD sort of the pairs by (u, v), tag as payload   (select.cuh)
//   k_ing_class    class of every distinct pair (stored, added, both) + per-row statistics
//   k_ing_copies   copies of the pair in the reference's graph: its merge routine duplicates some of
//                  the pairs that are in both lists (see k_ing_class), self-loop removal takes one away
//                  -> scan -> k_ing_write: keys + row offsets
// The header (banner, comments, size line) is a few hundred bytes and is parsed on the host by the
// caller (nlp_ingest_mtx).
#pragma once
#include "common.cuh"

namespace nlp {

constexpr int MTX_TILE = 8192, MTX_THREADS = 256, MTX_PER_THREAD = MTX_TILE / MTX_THREADS;

__device__ __forceinline__ bool mtx_line_start(const uint8_t* __restrict__ t, uint64_t i, uint64_t body0) {
  return i == body0 || __ldg(t + i - 1) == '\n';
}

__global__ void __launch_bounds__(MTX_THREADS) k_mtx_count(const uint8_t* __restrict__ text, uint64_t body0, uint64_t bytes,
                                                           uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint64_t lo = body0 + (uint64_t)blockIdx.x * MTX_TILE + (uint64_t)threadIdx.x * MTX_PER_THREAD;
  uint32_t c = 0;
  for (int k = 0; k < MTX_PER_THREAD; ++k)
    if (lo + k < bytes && mtx_line_start(text, lo + k, body0)) ++c;
  #pragma unroll
  for (int k = 16; k >= 1; k >>= 1) c += __shfl_xor_sync(NLP_FULL, c, k);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_n, c);
  __syncthreads();
  if (threadIdx.x == 0) counts[blockIdx.x] = s_n;
}

// Unsigned decimal at text[p...]; false when there is no digit.  Values that do not fit 32 bits
// saturate (and are then reported as out of range by the caller).
__device__ __forceinline__ bool mtx_uint(const uint8_t* __restrict__ t, uint64_t& p, uint64_t bytes, uint32_t* out) {
  while (p < bytes) {                                  // blanks (strtoull skips white space, not newlines here: a line is a line)
    const uint8_t c = __ldg(t + p);
    if (c == ' ' || c == '\t' || c == '\r') ++p; else break;
  }
  unsigned long long v = 0;
  int digits = 0;
  while (p < bytes) {
    const uint8_t c = __ldg(t + p);
    if (c < '0' || c > '9') break;
    if (v < (1ull << 40)) v = v * 10ull + (unsigned long long)(c - '0');
    ++p; ++digits;
  }
  *out = v > 0xffffffffull ? 0xffffffffu : (uint32_t)v;
  return digits > 0;
}

// flags[0] |= 1: a vertex id outside 1..n; the edge arrays get u = 0 for lines without an edge.
__global__ void __launch_bounds__(MTX_THREADS) k_mtx_parse(const uint8_t* __restrict__ text, uint64_t body0, uint64_t bytes, uint32_t n,
                                                           const unsigned long long* __restrict__ tile_base,
                                                           uint32_t* __restrict__ eu, uint32_t* __restrict__ ev, unsigned int* __restrict__ flags) {
  __shared__ uint32_t s_warp[MTX_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t lo = body0 + (uint64_t)blockIdx.x * MTX_TILE + (uint64_t)threadIdx.x * MTX_PER_THREAD;
  uint32_t c = 0;
  for (int k = 0; k < MTX_PER_THREAD; ++k)
    if (lo + k < bytes && mtx_line_start(text, lo + k, body0)) ++c;
  uint32_t inc = c;                                     // inclusive scan over the block
  #pragma unroll
  for (int k = 1; k < 32; k <<= 1) { const uint32_t o = __shfl_up_sync(NLP_FULL, inc, k); if (lane >= k) inc += o; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  unsigned long long slot = tile_base[blockIdx.x] + before + inc - c;
  if (!c) return;
  for (int k = 0; k < MTX_PER_THREAD; ++k) {
    const uint64_t i = lo + k;
    if (i >= bytes || !mtx_line_start(text, i, body0)) continue;
    uint64_t p = i;
    uint32_t u = 0, v = 0;
    bool ok = mtx_uint(text, p, bytes, &u) && mtx_uint(text, p, bytes, &v);   // a blank line, a comment: no edge
    if (ok && (u < 1u || u > n || v < 1u || v > n)) { atomicOr(flags, 1u); ok = false; }
    eu[slot] = ok ? u : 0u;
    ev[slot] = v;
    ++slot;
  }
}

// Directed pairs of the lines.  Every line (u, v) gives the entry v of row u (tag 0: what the
// reader stores).  `second`: 1 = also (v, u) with tag 0 (a symmetric banner: the READER stores both,
// inc/mtx.hxx:181), 2 = also (v, u) with tag 1 (symmetrizeOmp adds the reverse of every stored edge
// LATER, inc/symmetrize.hxx:77 -- the tags matter, see k_ing_class).
__global__ void __launch_bounds__(256) k_ing_count(const uint32_t* __restrict__ eu, uint64_t L, int second, uint32_t* __restrict__ cnt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < L; i += (uint64_t)gridDim.x * blockDim.x)
    cnt[i] = eu[i] == 0u ? 0u : (second ? 2u : 1u);
}

__global__ void __launch_bounds__(256) k_ing_emit(const uint32_t* __restrict__ eu, const uint32_t* __restrict__ ev, uint64_t L,
                                                  const uint32_t* __restrict__ cnt, const unsigned long long* __restrict__ at, int second,
                                                  uint32_t* __restrict__ pu, uint32_t* __restrict__ pv, uint32_t* __restrict__ pt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < L; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cnt[i];
    if (!c) continue;
    const unsigned long long a = at[i];
    const uint32_t u = eu[i], v = ev[i];
    pu[a] = u; pv[a] = v; pt[a] = 0u;
    if (c == 2u) { pu[a + 1] = v; pv[a + 1] = u; pt[a + 1] = second == 2 ? 1u : 0u; }
  }
}

// Class of every run of equal (u, v) in the sorted pairs, written at its first pair (0 elsewhere):
// 1 = stored by the reader only (x), 2 = added by symmetrizeOmp only (y), 3 = both.
// Why it matters: the reference merges the added entries y into the stored row x with its own
// set_union_last_inplace (inc/_algorithm.hxx:177-221), and that routine writes an entry that is in
// BOTH lists twice once it has switched to its deque loop -- which happens at the first y that is
// not in x while x still holds a larger entry.  A row of the reference's graph is therefore a sorted
// MULTISET: x united with y, plus a second copy of every common entry above that first inserted y.
// (Files that list both (u, v) and (v, u) of reciprocal links -- most "general" web graphs -- get
// such rows; the prediction counts entries, so the copies change the scores.)  Reproduced here, not
// corrected: row_ymin[u] = smallest class-2 entry, row_xmax[u] = largest stored entry.
__global__ void __launch_bounds__(256) k_ing_class(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, const uint32_t* __restrict__ pt,
                                                   uint64_t P, uint32_t* __restrict__ cls, uint32_t* __restrict__ row_ymin, uint32_t* __restrict__ row_xmax) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t u = pu[i], v = pv[i];
    if (i != 0 && pu[i - 1] == u && pv[i - 1] == v) { cls[i] = 0u; continue; }
    uint32_t c = 0;
    for (uint64_t j = i; j < P && pu[j] == u && pv[j] == v; ++j) c |= pt[j] ? 2u : 1u;
    cls[i] = c;
    if (c == 2u) atomicMin(row_ymin + u, v); else atomicMax(row_xmax + u, v);
  }
}

// Copies of every distinct pair in the reference's graph (see k_ing_class); removeSelfLoopsOmpU takes
// ONE copy of (u, u) away (set_difference_inplace with a single request, inc/_algorithm.hxx:114-143).
__global__ void __launch_bounds__(256) k_ing_copies(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, uint64_t P,
                                                    const uint32_t* __restrict__ row_ymin, const uint32_t* __restrict__ row_xmax,
                                                    int drop_self, uint32_t* __restrict__ cls) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = cls[i];
    if (!c) continue;
    const uint32_t u = pu[i], v = pv[i];
    const uint32_t yk = row_ymin[u];
    uint32_t n = (c == 3u && yk < row_xmax[u] && v > yk) ? 2u : 1u;
    if (drop_self && u == v) --n;
    cls[i] = n;
  }
}

// keys of the pairs (copies[i] of them at pos[i]); the first pair of a row also writes the offsets
// of its row and of the empty rows before it.  Rows after the last pair: k_ing_tail.
__global__ void __launch_bounds__(256) k_ing_write(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv, uint64_t P,
                                                   const uint32_t* __restrict__ copies, const unsigned long long* __restrict__ pos,
                                                   uint32_t* __restrict__ keys, unsigned long long* __restrict__ off) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < P; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long at = pos[i];
    const uint32_t u = pu[i];
    for (uint32_t k = 0; k < copies[i]; ++k) keys[at + k] = pv[i];
    if (i == 0 || pu[i - 1] != u) {
      const uint32_t first = i == 0 ? 0u : pu[i - 1] + 1u;
      for (uint32_t x = first; x <= u; ++x) off[x] = at;
    }
  }
}

__global__ void __launch_bounds__(256) k_ing_tail(const uint32_t* __restrict__ pu, uint64_t P, uint32_t S, unsigned long long M,
                                                  unsigned long long* __restrict__ off) {
  const uint64_t first = P ? (uint64_t)pu[P - 1] + 1u : 0u;
  for (uint64_t x = first + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x <= S; x += (uint64_t)gridDim.x * blockDim.x) off[x] = M;
}

}  // namespace nlp
