/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's IHub/LHub
 * neighbourhood link-prediction hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may build, load or call this file; the product path
 * (the CUDA library behind include/nlp_b200.h) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement
 * against the unmodified reference templates compiled from /root/reference
 * (oracle/_ref/libnlpref.so, see oracle/ref_driver.cxx) and against the committed
 * golden vectors in tests/golden/ that were generated from that same library
 * (tests/golden/make_golden.py), including the 7-vertex known-answer case of
 * SURVEY.md section 8c.
 *
 * What it restates (all citations are into the reference checkout):
 *   - sequential main loop                       inc/predict.hxx:214-265
 *   - wedge scan + first-touch list              inc/predict.hxx:153-160, 172-179
 *   - clear of touched slots                     inc/predict.hxx:187-192
 *   - hub cutoff (LHub)                          inc/predict.hxx:227
 *   - v>u / MAXFACTOR2 filter                    inc/predict.hxx:218-222
 *   - self / first-order-neighbour exclusion     inc/predict.hxx:231-233
 *   - the nine score functions                   inc/predict.hxx:504,542,580,618,656,694,732,770-771,810-811
 *   - min-score filter                           inc/predict.hxx:237
 *   - result = top-maxEdges, score descending    inc/predict.hxx:239-263, 369-372
 * The reference breaks score ties by heap accident; this oracle (and the GPU path)
 * use the canonical total order (score desc, u asc, v asc) required by BASELINE.json.
 * The reference's per-thread heaps are replaced by "collect, then sort + truncate",
 * which selects the same multiset of scores and is exact under the canonical order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { NLP_CN = 0, NLP_JC, NLP_SI, NLP_SC, NLP_HP, NLP_HD, NLP_LHN, NLP_AA, NLP_RA };

typedef struct {
  uint64_t first_hop;           /* first-hop entries scanned (predict.hxx:224)                  */
  uint64_t eligible_first_hop;  /* of those, entries that pass the hub cutoff (predict.hxx:227)  */
  uint64_t wedges;              /* second-hop entries visited, W(D) (predict.hxx:155)            */
  uint64_t wedges_vgtu;         /* of those, entries that pass the ft filter (predict.hxx:156)   */
  uint64_t candidates;          /* sum over u of |vedgs_u| (predict.hxx:157)                     */
  uint64_t kept;                /* candidates with score > minScore (predict.hxx:237)            */
} nlp_oracle_stats;

typedef struct { uint32_t u, v; float score; } cand_t;

typedef struct { cand_t* a; size_t n, cap; } cand_vec;

static int cand_cmp(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return  1;
  if (a->u != b->u) return a->u < b->u ? -1 : 1;
  if (a->v != b->v) return a->v < b->v ? -1 : 1;
  return 0;
}

static int cand_push(cand_vec* c, uint32_t u, uint32_t v, float s, uint64_t max_edges) {
  if (c->n == c->cap) {
    /* bounded memory: once 2K (+slack) entries are buffered, keep the best K only */
    if (max_edges != UINT64_MAX && c->n >= 2 * max_edges + 4096) {
      qsort(c->a, c->n, sizeof(cand_t), cand_cmp);
      c->n = (size_t)max_edges;
    } else {
      size_t ncap = c->cap ? c->cap * 2 : 4096;
      cand_t* na = (cand_t*)realloc(c->a, ncap * sizeof(cand_t));
      if (!na) return -1;
      c->a = na; c->cap = ncap;
    }
  }
  c->a[c->n].u = u; c->a[c->n].v = v; c->a[c->n].score = s; c->n++;
  return 0;
}

/* predict.hxx:504,542,580,618,656,694,732,771,811 -- same type chain as the C++ lambdas:
 * degrees are size_t, the count is uint32 (float for AA/RA), W = float.                */
static inline float score_fn(int measure, uint64_t du, uint64_t dv, uint32_t n_int, float n_flt) {
  float N = (float)n_int;
  switch (measure) {
    case NLP_CN:  return N;
    case NLP_JC:  return N / (float)(uint64_t)(du + dv - (uint64_t)n_int);
    case NLP_SI:  return N / (float)(uint64_t)(du + dv);
    case NLP_SC:  return (float)((double)N / sqrt((double)(uint64_t)(du * dv)));
    case NLP_HP:  return N / (float)(du < dv ? du : dv);
    case NLP_HD:  return N / (float)(du > dv ? du : dv);
    case NLP_LHN: return N / (float)(uint64_t)(du * dv);
    default:      return n_flt;  /* AA, RA: the accumulated float itself */
  }
}

/*
 * Returns 0 on success.  Output arrays are malloc'ed here (free with nlp_oracle_free);
 * *out_count = min(max_edges, #kept) edges in canonical order.
 */
int nlp_oracle_predict(const uint64_t* offsets, const uint32_t* keys, uint32_t span,
                       int measure, uint32_t min_degree1, uint32_t max_factor2,
                       uint64_t max_edges, float min_score, int threads,
                       uint32_t** out_u, uint32_t** out_v, float** out_score,
                       uint64_t* out_count, nlp_oracle_stats* stats) {
  const size_t S = span;
  const int custom = (measure == NLP_AA || measure == NLP_RA);
  int T = 1;
#ifdef _OPENMP
  T = threads > 0 ? threads : omp_get_max_threads();
#endif
  (void)threads;
  cand_vec* lists = (cand_vec*)calloc((size_t)T, sizeof(cand_vec));
  nlp_oracle_stats* st = (nlp_oracle_stats*)calloc((size_t)T, sizeof(nlp_oracle_stats));
  int failed = 0;
  if (!lists || !st) return -1;

  if (max_edges > 0) {   /* predict.hxx:367: nothing is scanned when maxEdges == 0 */
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      /* predict.hxx:116-122: touched-key list + dense span-sized accumulator */
      uint32_t* vedgs  = (uint32_t*)malloc((S ? S : 1) * sizeof(uint32_t));
      uint32_t* veoutI = custom ? NULL : (uint32_t*)calloc(S ? S : 1, sizeof(uint32_t));
      float*    veoutF = custom ? (float*)calloc(S ? S : 1, sizeof(float)) : NULL;
      size_t    nv = 0;
      nlp_oracle_stats* s = &st[t];
      cand_vec* c = &lists[t];
      int ok = vedgs && (veoutI || veoutF);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 2048)
#endif
      for (size_t u = 0; u < S; ++u) {
        if (!ok) continue;
        const uint64_t ub = offsets[u], ue = offsets[u + 1];
        const uint64_t du = ue - ub;
        /* predict.hxx:223 -> 187-192 */
        if (custom) for (size_t i = 0; i < nv; ++i) veoutF[vedgs[i]] = 0.0f;
        else        for (size_t i = 0; i < nv; ++i) veoutI[vedgs[i]] = 0;
        nv = 0;
        /* predict.hxx:224-230 */
        for (uint64_t e = ub; e < ue; ++e) {
          const uint32_t w = keys[e];
          const uint64_t wb = offsets[w], we = offsets[w + 1];
          const uint64_t dw = we - wb;
          s->first_hop++;
          if (min_degree1 && dw > (uint64_t)min_degree1) continue;   /* predict.hxx:227 */
          s->eligible_first_hop++;
          s->wedges += dw;
          double term = 0.0;
          if (measure == NLP_AA) term = 1.0 / log((double)dw);        /* predict.hxx:770 */
          if (measure == NLP_RA) term = 1.0 / (double)dw;             /* predict.hxx:810 */
          for (uint64_t f = wb; f < we; ++f) {
            const uint32_t v = keys[f];
            /* predict.hxx:218-222 (note: the reference's middle clause compares deg(u) with itself) */
            int pass = v > u;
            if (pass && max_factor2) {
              const uint64_t dv = offsets[v + 1] - offsets[v];
              pass = du <= (uint64_t)max_factor2 * du && dv <= (uint64_t)max_factor2 * du;
            }
            if (!pass) continue;                                        /* predict.hxx:156 */
            s->wedges_vgtu++;
            if (custom) {
              if (!(veoutF[v] != 0.0f)) vedgs[nv++] = v;                /* predict.hxx:176 */
              veoutF[v] = (float)((double)veoutF[v] + term);           /* predict.hxx:177: float += double */
            } else {
              if (!veoutI[v]) vedgs[nv++] = v;                          /* predict.hxx:157 */
              ++veoutI[v];                                              /* predict.hxx:158 */
            }
          }
        }
        /* predict.hxx:232-233 */
        if (custom) { veoutF[u] = 0.0f; for (uint64_t e = ub; e < ue; ++e) veoutF[keys[e]] = 0.0f; }
        else        { veoutI[u] = 0;    for (uint64_t e = ub; e < ue; ++e) veoutI[keys[e]] = 0; }
        /* predict.hxx:235-263 */
        s->candidates += nv;
        for (size_t i = 0; i < nv; ++i) {
          const uint32_t v = vedgs[i];
          const uint64_t dv = offsets[v + 1] - offsets[v];
          const float score = score_fn(measure, du, dv, custom ? 0u : veoutI[v], custom ? veoutF[v] : 0.0f);
          if (score <= min_score) continue;                             /* predict.hxx:237 */
          s->kept++;
          if (cand_push(c, (uint32_t)u, v, score, max_edges)) ok = 0;
        }
      }
      if (!ok) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
        failed = 1;
      }
      free(vedgs); free(veoutI); free(veoutF);
    }
  }

  /* merge (replaces predict.hxx:431-460 / 369-372) */
  size_t total = 0;
  for (int t = 0; t < T; ++t) total += lists[t].n;
  cand_t* all = (cand_t*)malloc((total ? total : 1) * sizeof(cand_t));
  if (!all) failed = 1;
  size_t p = 0;
  nlp_oracle_stats sum; memset(&sum, 0, sizeof sum);
  for (int t = 0; t < T; ++t) {
    if (all && lists[t].n) memcpy(all + p, lists[t].a, lists[t].n * sizeof(cand_t));
    p += lists[t].n; free(lists[t].a);
    sum.first_hop += st[t].first_hop; sum.eligible_first_hop += st[t].eligible_first_hop;
    sum.wedges += st[t].wedges; sum.wedges_vgtu += st[t].wedges_vgtu;
    sum.candidates += st[t].candidates; sum.kept += st[t].kept;
  }
  free(lists); free(st);
  if (failed) { free(all); return -1; }
  qsort(all, total, sizeof(cand_t), cand_cmp);
  size_t n = total;
  if (max_edges != UINT64_MAX && n > max_edges) n = (size_t)max_edges;
  uint32_t* ou = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t* ov = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  float*    os = (float*)   malloc((n ? n : 1) * sizeof(float));
  if (!ou || !ov || !os) { free(all); free(ou); free(ov); free(os); return -1; }
  for (size_t i = 0; i < n; ++i) { ou[i] = all[i].u; ov[i] = all[i].v; os[i] = all[i].score; }
  free(all);
  *out_u = ou; *out_v = ov; *out_score = os; *out_count = n;
  if (stats) *stats = sum;
  return 0;
}

void nlp_oracle_free(void* p) { free(p); }
