// Kernel (e) of the north star: global top-K selection over the candidate buffer, replacing
// the per-thread heaps and the serial T-way merge of inc/predict.hxx:313-336, 431-460.
//
// Candidates are ordered by the 96-bit key (desc_key(score), u, v) -- the canonical
// (score desc, u asc, v asc) order -- with a hand-written stable LSD radix sort (8-bit digits,
// digits that are constant over the whole buffer are skipped).  When the buffer holds more than
// K candidates, an MSD radix select first narrows it to the K best (plus the tie bucket of the
// K-th) so only ~K entries are sorted.
#pragma once
#include "common.cuh"

namespace nlp {

enum { SORT_THREADS = 256, SORT_WARPS = 8, SORT_ROUNDS = 16, SORT_TILE = SORT_THREADS * SORT_ROUNDS };

// word 0 = v, 1 = u, 2 = desc_key(score); LSD passes run word 0 digit 0 ... word 2 digit 3
__device__ __forceinline__ uint32_t sort_word(int word, uint32_t u, uint32_t v, uint32_t s) {
  return word == 0 ? v : word == 1 ? u : desc_key(s);
}

// hist[12][256]: digit histograms of all 12 passes in one read of the buffer
__global__ void __launch_bounds__(256) k_prehist(const uint32_t* __restrict__ cu, const uint32_t* __restrict__ cv,
                                                 const uint32_t* __restrict__ cs, uint64_t n,
                                                 unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[12 * 256];
  for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t w0 = cv[i], w1 = cu[i], w2 = desc_key(cs[i]);
    #pragma unroll
    for (int d = 0; d < 4; ++d) {
      atomicAdd(&sh[(0 + d) * 256 + ((w0 >> (8 * d)) & 255u)], 1u);
      atomicAdd(&sh[(4 + d) * 256 + ((w1 >> (8 * d)) & 255u)], 1u);
      atomicAdd(&sh[(8 + d) * 256 + ((w2 >> (8 * d)) & 255u)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// counts[d * nblocks + b] = number of items of tile b whose digit is d
__global__ void __launch_bounds__(SORT_THREADS) k_tilehist(const uint32_t* __restrict__ cu, const uint32_t* __restrict__ cv,
                                                            const uint32_t* __restrict__ cs, uint64_t n, int word, int shift,
                                                            uint32_t* __restrict__ counts, uint32_t nblocks) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
  const uint32_t* __restrict__ src = word == 0 ? cv : word == 1 ? cu : cs;
  #pragma unroll 4
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const uint64_t i = base + (uint64_t)r * SORT_THREADS + threadIdx.x;
    if (i < n) {
      uint32_t x = src[i];
      if (word == 2) x = desc_key(x);
      atomicAdd(&sh[(x >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  counts[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// One block per digit: in-place exclusive scan of that digit's row of tile counts.
__global__ void __launch_bounds__(256) k_rowscan(uint32_t* __restrict__ counts, uint32_t nblocks,
                                                 uint32_t* __restrict__ totals) {
  __shared__ uint32_t part[256];
  uint32_t* row = counts + (uint64_t)blockIdx.x * nblocks;
  const uint32_t per = (nblocks + 255u) / 256u;
  const uint32_t lo = min(threadIdx.x * per, nblocks), hi = min(lo + per, nblocks);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; ++i) s += row[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {          // Hillis-Steele inclusive scan
    uint32_t t = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;
  for (uint32_t i = lo; i < hi; ++i) { const uint32_t c = row[i]; row[i] = run; run += c; }
  if (threadIdx.x == 255) totals[blockIdx.x] = part[255];
}

// Exclusive scan of one u32 per thread over the SORT_THREADS threads of the block.
__device__ __forceinline__ uint32_t sort_block_exclusive(uint32_t x, uint32_t* s_warp /*[SORT_WARPS]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = x;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(NLP_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
  #pragma unroll
  for (int w = 0; w < SORT_WARPS; ++w) before += w < warp ? s_warp[w] : 0u;
  __syncthreads();
  return before + inc - x;
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // make the init visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Shared-memory plan of k_scatter (all dynamic): the raw tile (2 or 3 arrays of SORT_TILE words,
// filled by TMA bulk copies), the tile permutation, warp-private digit counters, per-digit bases.
__host__ __device__ constexpr uint32_t scatter_smem_bytes(bool has3) {
  return SORT_TILE * 4u * (has3 ? 3u : 2u) + SORT_TILE * 2u + SORT_WARPS * 256u * 4u + 256u * 4u + SORT_WARPS * 4u + 16u;
}

// Stable scatter of one tile.  The tile's arrays arrive in shared memory by TMA bulk copies (no
// registers, whole tile in flight at once; the buffers are padded to a multiple of SORT_TILE so a
// full tile can always be read).  Ranks inside a warp come from digit masks: every lane ORs its
// bit into a warp-private word per digit (shared memory) and reads the word back -- one atomic and
// one load per record.  (The first version built the mask from nine ballots and spent 60 % of the
// kernel's instructions -- LOP3 / ISETP / SEL / SHF / VOTE on the ncu source page -- doing so;
// MATCH.ANY itself measured ~270 cycles of issue per SM sub-partition on B200.)  Warps are ordered
// by a shared-memory scan, tiles by the scanned count matrix.  A tile-local permutation sorts the
// tile by digit, so records leave in tile order and every digit's records form one contiguous
// run: a warp store touches the 2-3 runs it straddles instead of 32 different lines.
template <bool HAS3>
__global__ void __launch_bounds__(SORT_THREADS, HAS3 ? 3 : 4)
k_scatter(const uint32_t* __restrict__ iu, const uint32_t* __restrict__ iv, const uint32_t* __restrict__ is,
          uint32_t* __restrict__ ou, uint32_t* __restrict__ ov, uint32_t* __restrict__ os, uint64_t n,
          int word, int shift, const uint32_t* __restrict__ counts, uint32_t nblocks, const uint32_t* __restrict__ totals) {
  extern __shared__ __align__(128) uint32_t smem[];
  constexpr int NW = HAS3 ? 3 : 2;
  uint32_t* raw = smem;                                            // [NW][SORT_TILE]
  uint16_t* s_perm = reinterpret_cast<uint16_t*>(raw + NW * SORT_TILE);   // [SORT_TILE]
  uint32_t (*s_hist)[256] = reinterpret_cast<uint32_t (*)[256]>(raw + NW * SORT_TILE + SORT_TILE / 2);
  uint32_t* s_gbase = &s_hist[0][0] + SORT_WARPS * 256;
  uint32_t* s_warp = s_gbase + 256;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_warp + SORT_WARPS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tbase = (uint64_t)blockIdx.x * SORT_TILE;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    mbar_expect_tx(s_bar, NW * SORT_TILE * 4);
    bulk_g2s(raw, iu + tbase, SORT_TILE * 4, s_bar);
    bulk_g2s(raw + SORT_TILE, iv + tbase, SORT_TILE * 4, s_bar);
    if (HAS3) bulk_g2s(raw + 2 * SORT_TILE, is + tbase, SORT_TILE * 4, s_bar);
  }
  // digit masks of the ranking phase; the permutation (written after that phase) reuses the space
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_perm);          // [SORT_WARPS][256] == SORT_TILE * 2 bytes
  static_assert(SORT_WARPS * 256 * 4 == SORT_TILE * 2, "digit masks alias the tile permutation");
  for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) { (&s_hist[0][0])[i] = 0; s_mask[i] = 0; }
  __syncthreads();
  mbar_wait(s_bar, 0);
  const uint32_t tile_n = (uint32_t)(n - tbase < (uint64_t)SORT_TILE ? n - tbase : (uint64_t)SORT_TILE);
  // the word this pass sorts on (sort_word: 0 = v, 1 = u, 2 = score)
  const uint32_t* kw = word == 0 ? raw + SORT_TILE : word == 1 ? raw : raw + 2 * SORT_TILE;
  uint32_t pre[SORT_ROUNDS];
  const unsigned lt = (1u << lane) - 1u;
  uint32_t* my_mask = s_mask + warp * 256;
  #pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const uint32_t idx = (uint32_t)warp * (32 * SORT_ROUNDS) + r * 32 + lane;
    uint32_t x = kw[idx];
    if (word == 2) x = desc_key(x);
    const bool valid = idx < tile_n;
    const uint32_t d = (x >> shift) & 255u;
    if (valid) atomicOr(my_mask + d, 1u << lane);
    __syncwarp();
    const unsigned m = valid ? *reinterpret_cast<volatile uint32_t*>(my_mask + d) : 0u;   // lanes holding digit d
    const uint32_t before = valid ? s_hist[warp][d] : 0u;
    pre[r] = (d << 16) | (before + __popc(m & lt));                // digit and rank inside (warp, digit)
    __syncwarp();
    if (valid && (__ffs(m) - 1) == lane) { s_hist[warp][d] = before + __popc(m); my_mask[d] = 0u; }
    __syncwarp();
  }
  __syncthreads();
  {   // thread t owns digit t
    uint32_t c[SORT_WARPS], tile_count = 0;
    #pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) { c[w] = s_hist[w][tid]; tile_count += c[w]; }
    const uint32_t tile_start = sort_block_exclusive(tile_count, s_warp);       // records of smaller digits in this tile
    const uint32_t digit_start = sort_block_exclusive(totals[tid], s_warp);     // ... in the whole buffer
    uint32_t run = tile_start;
    #pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) { s_hist[w][tid] = run; run += c[w]; }
    s_gbase[tid] = digit_start + counts[(uint64_t)tid * nblocks + blockIdx.x] - tile_start;
  }
  __syncthreads();
  #pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const uint32_t idx = (uint32_t)warp * (32 * SORT_ROUNDS) + r * 32 + lane;
    if (idx < tile_n) s_perm[s_hist[warp][pre[r] >> 16] + (pre[r] & 0xffffu)] = (uint16_t)idx;
  }
  __syncthreads();
  #pragma unroll 4
  for (int k = 0; k < SORT_ROUNDS; ++k) {
    const uint32_t lp = (uint32_t)k * SORT_THREADS + tid;
    if (lp < tile_n) {
      const uint32_t src = s_perm[lp];
      uint32_t x = kw[src];
      if (word == 2) x = desc_key(x);
      const uint32_t gp = s_gbase[(x >> shift) & 255u] + lp;
      ou[gp] = raw[src]; ov[gp] = raw[SORT_TILE + src];
      if (HAS3) os[gp] = raw[2 * SORT_TILE + src];
    }
  }
}

// ---- MSD radix select ------------------------------------------------------------------------
// State of the narrowing: items whose 96-bit key has `prefix` in its top `bits` bits are still
// undecided; `above` items are already known to rank before them.
struct SelectState {
  uint32_t pre_hi, pre_mid, pre_lo;   // prefix words (desc_key(score), u, v), masked
  uint32_t bits;                      // resolved leading bits (multiple of 8)
  unsigned long long above;           // items strictly before the undecided bucket
  unsigned long long bucket;          // items in the undecided bucket
  uint32_t done;                      // narrowing finished: later hist/step launches are no-ops
  uint32_t pad;
  unsigned long long hist[256];
};

__device__ __forceinline__ bool select_matches(const SelectState* st, uint32_t w2, uint32_t w1, uint32_t w0) {
  const uint32_t b = st->bits;
  if (b == 0) return true;
  if (b <= 32) return (w2 >> (32 - b)) == (st->pre_hi >> (32 - b));
  if (w2 != st->pre_hi) return false;
  if (b <= 64) return (w1 >> (64 - b)) == (st->pre_mid >> (64 - b));
  if (w1 != st->pre_mid) return false;
  return (w0 >> (96 - b)) == (st->pre_lo >> (96 - b));
}

__device__ __forceinline__ uint32_t select_digit(uint32_t bits, uint32_t w2, uint32_t w1, uint32_t w0) {
  // next 8 bits after `bits` leading ones
  const uint32_t word = bits < 32 ? w2 : bits < 64 ? w1 : w0;
  const uint32_t sh = 24 - (bits & 31u);
  return (word >> sh) & 255u;
}

__global__ void __launch_bounds__(256) k_select_hist(const uint32_t* __restrict__ cu, const uint32_t* __restrict__ cv,
                                                     const uint32_t* __restrict__ cs, uint64_t n, SelectState* st) {
  __shared__ uint32_t sh[256];
  if (st->done) return;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t bits = st->bits;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t w2 = desc_key(cs[i]);
    uint32_t w1 = 0, w0 = 0;
    if (bits >= 32) w1 = cu[i];
    if (bits >= 64) w0 = cv[i];
    if (select_matches(st, w2, w1, w0)) atomicAdd(&sh[select_digit(bits, w2, w1, w0)], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// Single thread: pick the digit bucket that holds the K-th item, extend the prefix by 8 bits.
// The narrowing is finished (st->done) once the survivors (above + bucket) fit K + slack or
// max_bits are resolved; the host queues several levels back to back and reads the state once.
__global__ void k_select_step(SelectState* st, unsigned long long K, unsigned long long slack, uint32_t max_bits) {
  if (st->done) return;
  unsigned long long above = st->above;
  int d = 0;
  for (; d < 255; ++d) {
    if (above + st->hist[d] >= K) break;
    above += st->hist[d];
  }
  const uint32_t bits = st->bits;
  const uint32_t sh = 24 - (bits & 31u);
  if (bits < 32) st->pre_hi |= (uint32_t)d << sh;
  else if (bits < 64) st->pre_mid |= (uint32_t)d << sh;
  else st->pre_lo |= (uint32_t)d << sh;
  st->above = above;
  st->bucket = st->hist[d];
  st->bits = bits + 8;
  if (above + st->hist[d] <= K + slack || bits + 8 >= max_bits) st->done = 1;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// Keep every item whose key prefix is <= the selected prefix (i.e. `above` + the tie bucket).
__global__ void __launch_bounds__(256) k_select_compact(const uint32_t* __restrict__ cu, const uint32_t* __restrict__ cv,
                                                        const uint32_t* __restrict__ cs, uint64_t n, const SelectState* st,
                                                        uint32_t* __restrict__ ou, uint32_t* __restrict__ ov,
                                                        uint32_t* __restrict__ os, unsigned long long* cursor) {
  const uint32_t b = st->bits;
  const int lane = threadIdx.x & 31;
  const uint64_t n32 = (n + 31u) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n32; i += (uint64_t)gridDim.x * blockDim.x) {
    bool keep = false;
    uint32_t u = 0, v = 0, s = 0;
    if (i < n) {
      u = cu[i]; v = cv[i]; s = cs[i];
      const uint32_t w2 = desc_key(s);
      // compare the top b bits of (w2, u, v) with the prefix
      int cmp = 0;   // -1 before, 0 equal, 1 after
      const uint32_t b2 = b < 32 ? b : 32;
      const uint32_t x2 = b2 ? (w2 >> (32 - b2)) : 0u, p2 = b2 ? (st->pre_hi >> (32 - b2)) : 0u;
      cmp = x2 < p2 ? -1 : x2 > p2 ? 1 : 0;
      if (cmp == 0 && b > 32) {
        const uint32_t b1 = b < 64 ? b - 32 : 32;
        const uint32_t x1 = u >> (32 - b1), p1 = st->pre_mid >> (32 - b1);
        cmp = x1 < p1 ? -1 : x1 > p1 ? 1 : 0;
        if (cmp == 0 && b > 64) {
          const uint32_t b0 = b - 64;
          const uint32_t x0 = v >> (32 - b0), p0 = st->pre_lo >> (32 - b0);
          cmp = x0 < p0 ? -1 : x0 > p0 ? 1 : 0;
        }
      }
      keep = cmp <= 0;
    }
    const unsigned m = __ballot_sync(NLP_FULL, keep);
    if (!m) continue;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(NLP_FULL, base, leader);
    if (keep) {
      const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
      ou[pos] = u; ov[pos] = v; os[pos] = s;
    }
  }
}

// ---- ordered top-K (pair path) ----------------------------------------------------------------
// Input: n records (u, v) sorted by (u, v) and an aligned array of score bits (NLP_NO_SCORE =
// nothing there).  The best K are found by an MSD radix select on the score alone, copied out IN
// ORDER (tile counts -> scan -> write), and then only need a stable sort by score: ties keep the
// ascending (u, v) order, which is the canonical order.
enum { OC_THREADS = 256, OC_PER_THREAD = 8, OC_TILE = OC_THREADS * OC_PER_THREAD };

// mode 0: every valid score; mode 1: score key prefix <= selected prefix (st->bits in [8, 32])
__device__ __forceinline__ bool ordered_keep(const SelectState* st, uint32_t sbits, int mode) {
  if (sbits == NLP_NO_SCORE) return false;
  if (mode == 0) return true;
  const uint32_t b = st->bits;
  return (desc_key(sbits) >> (32 - b)) <= (st->pre_hi >> (32 - b));
}

__global__ void __launch_bounds__(OC_THREADS) k_ordered_count(const uint32_t* __restrict__ sbits, uint64_t n,
                                                              const SelectState* __restrict__ st, int mode,
                                                              uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t s_warp[OC_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * OC_TILE + (uint64_t)threadIdx.x * OC_PER_THREAD;
  uint32_t c = 0;
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k)
    if (base + k < n && ordered_keep(st, sbits[base + k], mode)) ++c;
  c = __reduce_add_sync(NLP_FULL, c);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    #pragma unroll
    for (int w = 0; w < OC_THREADS / 32; ++w) t += s_warp[w];
    tile_counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(OC_THREADS) k_ordered_write(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv,
                                                              const uint32_t* __restrict__ sbits, uint64_t n,
                                                              const SelectState* __restrict__ st, int mode,
                                                              const unsigned long long* __restrict__ tile_off,
                                                              uint32_t* __restrict__ ou, uint32_t* __restrict__ ov,
                                                              uint32_t* __restrict__ os) {
  __shared__ uint32_t s_warp[OC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * OC_TILE + (uint64_t)threadIdx.x * OC_PER_THREAD;
  uint32_t sb[OC_PER_THREAD];
  uint32_t c = 0, keepmask = 0;
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k) {
    sb[k] = base + k < n ? sbits[base + k] : NLP_NO_SCORE;
    if (ordered_keep(st, sb[k], mode)) { keepmask |= 1u << k; ++c; }
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
  for (int w = 0; w < OC_THREADS / 32; ++w) before += w < warp ? s_warp[w] : 0u;
  unsigned long long pos = tile_off[blockIdx.x] + before + inc - c;
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k) {
    if ((keepmask >> k) & 1u) { ou[pos] = pu[base + k]; ov[pos] = pv[base + k]; os[pos] = sb[k]; ++pos; }
  }
}


// ---- ordered top-K, exact (bucket path) --------------------------------------------------------
// The kept pairs lie in ascending (u, v) order in record-aligned arrays.  The K best are
//   every pair whose score is better than the K-th score, plus
//   the first `need` pairs IN ARRAY ORDER of those tied with it (canonical order inside a tie class
//   is ascending (u, v)),
// so after an MSD radix select on the 32 score bits (digits of 11 + 11 + 10 bits; the first
// histogram is accumulated by the scoring kernel itself) an in-order compaction leaves exactly K
// survivors and a stable sort by score finishes the canonical order.  With several ranks the
// histograms are all-reduced, so every rank narrows to the same global cutoff.
struct Select11 {
  uint32_t prefix;                    // resolved leading bits of desc_key(score), left-aligned
  uint32_t bits;                      // 0, 11, 22 or 32
  unsigned long long above;           // pairs strictly better than the undecided bucket
  unsigned long long bucket;          // pairs in the undecided bucket
  unsigned long long need;            // valid when done: how many of the bucket belong to the result
  uint32_t done, pad;
  uint32_t key_or, key_nor;           // OR of the survivors' score keys / of their complements: a bit varies iff set in both
                                      // (the final sort skips the digits in which no bit varies; several ranks OR theirs on the host)
  unsigned long long hist[2048];
};

__device__ __forceinline__ int sel11_width(uint32_t bits) { return bits < 22u ? 11 : 10; }

// 0: not a survivor, 1: better than the bucket, 2: in the bucket
__device__ __forceinline__ int sel11_class(uint32_t prefix, uint32_t bits, uint32_t sbits) {
  if (sbits == NLP_NO_SCORE) return 0;
  if (bits == 0u) return 2;
  const uint32_t k = desc_key(sbits) >> (32u - bits), p = prefix >> (32u - bits);
  return k < p ? 1 : (k == p ? 2 : 0);
}

// Level histogram of the undecided bucket (skipped once the select is done).
__global__ void __launch_bounds__(256) k_sel11_hist(const uint32_t* __restrict__ sbits, uint64_t n, Select11* st) {
  __shared__ uint32_t sh[2048];
  if (st->done) return;
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  const uint32_t bits = st->bits, prefix = st->prefix;
  const int width = sel11_width(bits);
  const uint32_t shift = 32u - bits - (uint32_t)width, mask = (1u << width) - 1u;
  // 16-byte loads where the array allows it (one 4-byte load in flight per thread left this pass
  // latency bound: 33 us for 64 MB); the scalar loop takes the tail, or everything when the
  // scores start at an odd offset (a rank's slot range of a multi-GPU prediction)
  const uint64_t nv = ((reinterpret_cast<uintptr_t>(sbits) & 15u) == 0u) ? n / 4u : 0u;
  const uint4* __restrict__ sv = reinterpret_cast<const uint4*>(sbits);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 q = sv[i];
    const uint32_t s4[4] = {q.x, q.y, q.z, q.w};
    #pragma unroll
    for (int k = 0; k < 4; ++k)
      if (sel11_class(prefix, bits, s4[k]) == 2) atomicAdd(&sh[(desc_key(s4[k]) >> shift) & mask], 1u);
  }
  for (uint64_t i = nv * 4u + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = sbits[i];
    if (sel11_class(prefix, bits, s) == 2) atomicAdd(&sh[(desc_key(s) >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&st->hist[i], (unsigned long long)sh[i]);
}

// The OC_PER_THREAD (= 8) consecutive scores of a thread: two 16-byte loads when the tile is whole
// and the array 16-byte aligned, else guarded scalar loads (NLP_NO_SCORE past the end).
__device__ __forceinline__ void oc_load8(const uint32_t* __restrict__ sbits, uint64_t base, uint64_t n, uint32_t* sb) {
  if (base + 8u <= n && (reinterpret_cast<uintptr_t>(sbits + base) & 15u) == 0u) {
    const uint4 a = *reinterpret_cast<const uint4*>(sbits + base), b = *reinterpret_cast<const uint4*>(sbits + base + 4);
    sb[0] = a.x; sb[1] = a.y; sb[2] = a.z; sb[3] = a.w; sb[4] = b.x; sb[5] = b.y; sb[6] = b.z; sb[7] = b.w;
  } else {
    #pragma unroll
    for (int k = 0; k < 8; ++k) sb[k] = base + k < n ? sbits[base + k] : NLP_NO_SCORE;
  }
}

// One block of 256 threads: the bin that holds the K-th pair, extend the prefix.  Done when all 32
// bits are resolved (the bucket is an exact tie class: need = K - above of it) or when the whole
// bucket belongs to the result anyway (above + bucket <= K).
__global__ void __launch_bounds__(256) k_sel11_step(Select11* st, unsigned long long K) {
  __shared__ unsigned long long s_part[256];
  __shared__ int s_bin;
  __shared__ unsigned long long s_before;
  if (st->done) return;
  const uint32_t bits = st->bits;
  const int width = sel11_width(bits), nb = 1 << width, tid = threadIdx.x;
  const int per = nb / 256;
  unsigned long long loc[8], sum = 0;
  for (int k = 0; k < per; ++k) { loc[k] = st->hist[tid * per + k]; sum += loc[k]; }
  s_part[tid] = sum;
  if (tid == 0) { s_bin = -1; s_before = 0; }
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {               // inclusive scan
    const unsigned long long t = tid >= d ? s_part[tid - d] : 0ull;
    __syncthreads();
    s_part[tid] += t;
    __syncthreads();
  }
  const unsigned long long above0 = st->above;
  unsigned long long before = above0 + s_part[tid] - sum;    // pairs before this thread's bins
  if (before < K && before + sum >= K) {             // exactly one thread (if any)
    for (int k = 0; k < per; ++k) {
      if (before + loc[k] >= K) { s_bin = tid * per + k; s_before = before; break; }
      before += loc[k];
    }
  }
  __syncthreads();
  if (tid == 0) {
    int bin = s_bin;
    unsigned long long bef = s_before;
    if (bin < 0) {                                    // fewer than K pairs in all: everything survives
      bin = nb - 1;
      bef = above0 + s_part[255] - st->hist[nb - 1];
    }
    const unsigned long long bucket = st->hist[bin];
    st->prefix |= (uint32_t)bin << (32u - bits - (uint32_t)width);
    st->above = bef;
    st->bucket = bucket;
    st->bits = bits + (uint32_t)width;
    if (bef + bucket <= K) { st->done = 1; st->need = bucket; }
    else if (bits + (uint32_t)width >= 32u) { st->done = 1; st->need = K - bef; }
  }
  __syncthreads();
  for (int i = tid; i < 2048; i += 256) st->hist[i] = 0;
}

// Per tile: survivors better than the bucket (high word) and in the bucket (low word); and, over
// all survivors, which bits of the score key vary (OR of the keys, OR of their complements).
__global__ void __launch_bounds__(OC_THREADS) k_ordered_count2(const uint32_t* __restrict__ sbits, uint64_t n, Select11* st,
                                                               unsigned long long* __restrict__ tile_counts) {
  __shared__ unsigned long long s_warp[OC_THREADS / 32];
  __shared__ uint32_t s_or[2];
  if (threadIdx.x < 2) s_or[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t prefix = st->prefix, bits = st->bits;
  const uint64_t base = (uint64_t)blockIdx.x * OC_TILE + (uint64_t)threadIdx.x * OC_PER_THREAD;
  unsigned long long c = 0;
  uint32_t o1 = 0, o0 = 0;
  uint32_t sb[OC_PER_THREAD];
  oc_load8(sbits, base, n, sb);
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k) {
    const int cls = sel11_class(prefix, bits, sb[k]);   // NLP_NO_SCORE (also past the end): class 0
    if (cls) {
      c += cls == 1 ? (1ull << 32) : 1ull;
      const uint32_t key = desc_key(sb[k]);
      o1 |= key; o0 |= ~key;
    }
  }
  #pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(NLP_FULL, c, d);
  o1 = __reduce_or_sync(NLP_FULL, o1);
  o0 = __reduce_or_sync(NLP_FULL, o0);
  if ((threadIdx.x & 31) == 0) {
    s_warp[threadIdx.x >> 5] = c;
    if (o1) atomicOr(&s_or[0], o1);
    if (o0) atomicOr(&s_or[1], o0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    #pragma unroll
    for (int w = 0; w < OC_THREADS / 32; ++w) t += s_warp[w];
    tile_counts[blockIdx.x] = t;
    if (s_or[0]) atomicOr(&st->key_or, s_or[0]);
    if (s_or[1]) atomicOr(&st->key_nor, s_or[1]);
  }
}

// Survivors in array order: every better pair, and the bucket pairs whose index inside the bucket
// (over the whole array, `tie_base` more on this rank's left) is below `need`.
__global__ void __launch_bounds__(OC_THREADS) k_ordered_write2(const uint32_t* __restrict__ pu, const uint32_t* __restrict__ pv,
                                                               const uint32_t* __restrict__ sbits, uint64_t n,
                                                               const Select11* __restrict__ st, unsigned long long need,
                                                               const unsigned long long* __restrict__ tile_off,
                                                               uint32_t* __restrict__ ou, uint32_t* __restrict__ ov,
                                                               uint32_t* __restrict__ os) {
  __shared__ unsigned long long s_warp[OC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t prefix = st->prefix, bits = st->bits;
  const uint64_t base = (uint64_t)blockIdx.x * OC_TILE + (uint64_t)threadIdx.x * OC_PER_THREAD;
  uint32_t sb[OC_PER_THREAD];
  int cls[OC_PER_THREAD];
  unsigned long long c = 0;
  oc_load8(sbits, base, n, sb);
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k) {
    cls[k] = sel11_class(prefix, bits, sb[k]);
    c += cls[k] == 1 ? (1ull << 32) : (cls[k] == 2 ? 1ull : 0ull);
  }
  unsigned long long inc = c;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(NLP_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned long long before = 0;
  #pragma unroll
  for (int w = 0; w < OC_THREADS / 32; ++w) before += w < warp ? s_warp[w] : 0ull;
  const unsigned long long at = tile_off[blockIdx.x] + before + inc - c;
  unsigned long long nbetter = at >> 32, ntie = at & 0xffffffffull;
  #pragma unroll
  for (int k = 0; k < OC_PER_THREAD; ++k) {
    if (cls[k] == 1 || (cls[k] == 2 && ntie < need)) {
      const unsigned long long pos = nbetter + (ntie < need ? ntie : need);
      ou[pos] = pu[base + k]; ov[pos] = pv[base + k]; os[pos] = sb[k];
    }
    nbetter += cls[k] == 1 ? 1ull : 0ull;
    ntie += cls[k] == 2 ? 1ull : 0ull;
  }
}


// Multi-GPU: the all-gather delivers every rank's survivors as [rank][3][width] words (u, v, score
// bits, padded to the widest rank); one pass packs them into contiguous (u, v, score) arrays in
// rank order.  off[r] = survivors of the ranks before r (off[W] = total), passed by value.
struct GatherOffsets { unsigned long long off[17]; int world; };

__global__ void __launch_bounds__(256) k_gather_unpack(const uint32_t* __restrict__ recv, unsigned long long width, GatherOffsets g,
                                                       uint32_t* __restrict__ ou, uint32_t* __restrict__ ov, uint32_t* __restrict__ os) {
  const unsigned long long total = g.off[g.world];
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    int r = 0;
    while (r + 1 < g.world && g.off[r + 1] <= i) ++r;
    const unsigned long long k = i - g.off[r];
    const uint32_t* src = recv + (unsigned long long)r * width * 3ull;
    ou[i] = src[k]; ov[i] = src[width + k]; os[i] = src[2ull * width + k];
  }
}

}  // namespace nlp
