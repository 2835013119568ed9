/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's random edge removal, the step
 * right before the prediction path, pinned against the compiled
 * reference (oracle/ref_batch_driver.cxx, tests/test_batch_oracle.py).  Nothing in the product
 * path uses it yet: it is the parity anchor for moving batch generation onto the GPU.
 *
 * What the reference does (all file:line into /root/reference):
 *   runBatches                main.cxx:157-179   deletions = generateEdgeDeletions(rnd, y, size_t(d*|E|/2), 1, span-1, true)
 *   generateEdgeDeletions     inc/batch.hxx:99-112    batchSize times: retry(removeRandomEdge, 5)
 *   retry                     inc/_utility.hxx:198-203  up to N calls until one returns true
 *   removeRandomEdge          inc/batch.hxx:52-58     u = K(i + n * dis(rnd))
 *   removeRandomEdgeFrom      inc/batch.hxx:29-40     deg(u) == 0 -> false (no second draw);
 *                                                     vi = K(dis(rnd) * deg(u)); the vi-th entry v
 *                                                     of row u; push (u, v) and (v, u)
 *   tidyBatchUpdateU          inc/batch.hxx:200-208   keep existing edges, sort by (u, v), unique
 * The graph is NOT changed while the batch is drawn, so an edge can be drawn twice (collapsed by
 * `unique`), and the sampling is uniform over (vertex, then incident entry), not over edges.
 *
 * The random stream: std::default_random_engine = minstd_rand0 (x <- 16807 x mod 2^31-1), and
 * uniform_real_distribution<double>(0,1) = libstdc++'s generate_canonical<double, 53>: two engine
 * words per double, sum = (x1 - 1) + (x2 - 1) * 2147483646.0 (double arithmetic, the product
 * rounded before the add), divided by double(2147483646^2) = 2^62 - 2^33; a result >= 1 becomes
 * nextafter(1, 0).  Each step below is one IEEE double operation, as in the reference binary.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t state; } minstd0;

static void minstd0_seed(minstd0* g, uint32_t seed) {
  g->state = seed % 2147483647u;           /* linear_congruential_engine::seed: 0 maps to 1 */
  if (g->state == 0) g->state = 1;
}

static uint32_t minstd0_next(minstd0* g) {
  g->state = (uint32_t)(((uint64_t)g->state * 16807u) % 2147483647u);
  return g->state;
}

/* generate_canonical<double, 53>(minstd_rand0): bits/random.tcc */
static double canonical(minstd0* g) {
  const double r = 2147483646.0;            /* max() - min() + 1 */
  volatile double sum = 0.0, tmp = 1.0, prod; /* volatile: no contraction of the product into the add */
  prod = (double)(minstd0_next(g) - 1u) * tmp; sum = sum + prod; tmp = (double)((long double)tmp * (long double)r);
  prod = (double)(minstd0_next(g) - 1u) * tmp; sum = sum + prod; tmp = (double)((long double)tmp * (long double)r);
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}

static int cmp_pair(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

/* offsets[span+1], keys[M] (rows sorted ascending -- forEachEdge order of the reference's graphs).
 * Draws `batch_size` deletions with seed `seed` exactly as generateEdgeDeletions(rnd, x, batch_size,
 * 1, span-1, true) followed by tidyBatchUpdateU.  Returns the sorted unique DIRECTED list (both
 * directions of every removed edge) in *out_u / *out_v (malloc'ed, *count entries), and the
 * number of engine words consumed in *words (the next draw of the caller's stream starts there).
 * Returns 0, or -1 when out of memory. */
int nlp_oracle_edge_deletions(const uint64_t* offsets, const uint32_t* keys, uint32_t span, uint32_t seed,
                              uint64_t batch_size, uint32_t** out_u, uint32_t** out_v, uint64_t* count, uint64_t* words) {
  minstd0 g;
  minstd0_seed(&g, seed);
  uint64_t cap = 2 * batch_size + 2, n = 0, used = 0;
  uint64_t* pairs = (uint64_t*)malloc(cap * sizeof(uint64_t));
  if (!pairs) return -1;
  const double first = 1.0, range = (double)(size_t)(span - 1);      /* i = 1, n = span - 1 (main.cxx:166) */
  for (uint64_t l = 0; l < batch_size && span > 1; ++l) {
    for (int attempt = 0; attempt < 5; ++attempt) {                  /* retry(..., 5) */
      volatile double scaled = range * canonical(&g); used += 2;
      const uint32_t u = (uint32_t)(first + scaled);                 /* K(i + n*dis(rnd)) */
      const uint64_t deg = u < span ? offsets[u + 1] - offsets[u] : 0;
      if (deg == 0) continue;                                        /* removeRandomEdgeFrom: false, no second draw */
      const uint32_t vi = (uint32_t)(canonical(&g) * (double)deg); used += 2;
      if (vi < deg) {                                                /* the vi-th entry (always true: canonical < 1) */
        const uint32_t v = keys[offsets[u] + vi];
        pairs[n++] = ((uint64_t)u << 32) | v;
        pairs[n++] = ((uint64_t)v << 32) | u;
        break;
      }
      /* vi == deg cannot happen (canonical < 1), but the reference would return false and retry */
    }
  }
  qsort(pairs, n, sizeof(uint64_t), cmp_pair);                       /* sortEdgesByIdU */
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i)                                   /* uniqueEdgesU */
    if (i == 0 || pairs[i] != pairs[i - 1]) pairs[m++] = pairs[i];
  uint32_t* ou = (uint32_t*)malloc((m ? m : 1) * sizeof(uint32_t));
  uint32_t* ov = (uint32_t*)malloc((m ? m : 1) * sizeof(uint32_t));
  if (!ou || !ov) { free(pairs); free(ou); free(ov); return -1; }
  for (uint64_t i = 0; i < m; ++i) { ou[i] = (uint32_t)(pairs[i] >> 32); ov[i] = (uint32_t)pairs[i]; }
  free(pairs);
  *out_u = ou; *out_v = ov; *count = m;
  if (words) *words = used;
  return 0;
}

void nlp_oracle_batch_free(void* p) { free(p); }
