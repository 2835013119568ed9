/*
 * nlp_b200.h -- C ABI of the B200-native IHub/LHub neighbourhood link-prediction path.
 *
 * The reference (puzzlef/neighborhood-link-prediction-openmp) has no FFI: its hot path is the
 * C++ template family `predictLinks<Measure>[Omp]<MINDEGREE1,MAXFACTOR2,FORCEHEAP>(x, o)`
 * in inc/predict.hxx:502-831, entered from the PREDICT_LINKS macro at main.cxx:48-57.  This
 * header is the boundary a maintainer binds instead; the C++ shim
 * include/predict_b200.hxx re-creates the 18 template entry points and the two structs on top
 * of it, so main.cxx / batch.hxx use the GPU path unchanged.
 *
 * Plain pointers and sizes only; no C++ or torch types.  One handle drives one GPU and is
 * used by one host thread at a time (the reference call is blocking and serial as well).
 * Every function returns an nlp_status; nlp_last_error() gives the message.  There is no CPU
 * fallback: without a CUDA device nlp_create() fails with NLP_ERR_CUDA.
 */
#ifndef NLP_B200_H
#define NLP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nlp_handle nlp_handle;

typedef enum {
  NLP_OK            = 0,
  NLP_ERR_ARG       = 1,   /* bad argument (null pointer, unknown measure, unsorted offsets ...) */
  NLP_ERR_CUDA      = 2,   /* CUDA runtime error, or no usable device                           */
  NLP_ERR_NO_GRAPH  = 3,   /* nlp_predict before nlp_set_graph                                  */
  NLP_ERR_CAPACITY  = 4,   /* candidate buffer cannot hold an unbounded (max_edges = -1) result  */
  NLP_ERR_NO_RESULT = 5,   /* nlp_fetch before nlp_predict                                      */
  NLP_ERR_NO_TRUTH  = 6,   /* nlp_evaluate before nlp_set_truth                                 */
  NLP_ERR_COMM      = 7    /* NCCL error, or libnccl.so.2 not loadable                          */
} nlp_status;

/* Similarity measures, in the order main.cxx:212-220 runs them.
 * Score functions: inc/predict.hxx:521 (CN), 559 (JC), 597 (SI), 635 (SC), 673 (HP),
 * 711 (HD), 749 (LHN), 788-789 (AA), 828-829 (RA).                                            */
typedef enum {
  NLP_COMMON_NEIGHBORS     = 0,
  NLP_JACCARD_COEFFICIENT  = 1,
  NLP_SORENSEN_INDEX       = 2,
  NLP_SALTON_COSINE        = 3,
  NLP_HUB_PROMOTED         = 4,
  NLP_HUB_DEPRESSED        = 5,
  NLP_LEICHT_HOLME_NERMAN  = 6,
  NLP_ADAMIC_ADAR          = 7,
  NLP_RESOURCE_ALLOCATION  = 8,
  NLP_NUM_MEASURES         = 9
} nlp_measure;

#define NLP_UNBOUNDED UINT64_MAX   /* PredictLinkOptions::maxEdges default, size_t(-1) */

/* Replaces PredictLinkOptions<W> (inc/predict.hxx:33-55) plus the template arguments
 * <MINDEGREE1, MAXFACTOR2, FORCEHEAP> of inc/predict.hxx:214, which become runtime values. */
typedef struct {
  int32_t  measure;       /* nlp_measure                                                        */
  uint32_t min_degree1;   /* MINDEGREE1: 0 = IHub; D > 0 = LHub, first-hop w with deg(w) > D are
                             skipped as intermediates (inc/predict.hxx:301). Any value allowed. */
  uint32_t max_factor2;   /* MAXFACTOR2 (inc/predict.hxx:292-296); 0 = off, as in main.cxx       */
  int32_t  repeat;        /* PredictLinkOptions::repeat: scoring phase runs this many times and
                             scoring_ms is the mean (inc/predict.hxx:426-430); <= 0 is run once  */
  uint64_t max_edges;     /* PredictLinkOptions::maxEdges (NLP_UNBOUNDED = all candidates)       */
  float    min_score;     /* PredictLinkOptions::minScore: keep score > min_score (predict.hxx:311) */
} nlp_options;

/* Replaces the scalar part of PredictLinkResult<K,W> (inc/predict.hxx:65-102) and adds the
 * counters BASELINE.json's metric needs (wedges/s, roofline bytes).                             */
typedef struct {
  uint64_t count;               /* predicted edges = min(max_edges, #kept candidates)            */
  float    time_ms;             /* PredictLinkResult::time        = scoring + select/merge       */
  float    scoring_ms;          /* PredictLinkResult::scoringTime = mean over `repeat`           */
  float    select_ms;           /* time_ms - scoring_ms (inc/predict.hxx:431-460)                */
  float    frontier_ms;         /* part of scoring_ms spent building the work-balanced frontier  */
  uint64_t first_hop;           /* first-hop entries of the sources this rank owns               */
  uint64_t eligible_first_hop;  /* of those, entries passing the hub cutoff                      */
  uint64_t wedges;              /* W(D): second-hop entries the reference would visit            */
  uint64_t candidates;          /* distinct (u, v>u) pairs touched, incl. adjacent ones          */
  uint64_t kept;                /* pairs with score > min_score                                  */
  uint64_t emitted;             /* pairs written to the candidate buffer (after threshold prune) */
  uint64_t frontier_sources;    /* sources with non-zero work                                    */
  uint64_t bin_sources[8];      /* sources per path: [0]=8-lane [1]=32-lane [2..4]=hash 1K/4K/16K
                                   [5]=global dense spill [6]=windowed shared-memory counters
                                   (hub-heavy sources of the count measures), [7] reserved       */
  uint32_t passes;              /* candidate-buffer passes (1 unless the buffer had to be pruned)*/
  uint32_t path;                /* nlp_path the prediction ran on (NLP_PATH_SOURCE or NLP_PATH_PAIR) */
  float    phase_ms[8];         /* device time per phase of the LAST scoring repeat (CUDA events on
                                   the handle's stream).  NLP_PATH_SOURCE: [0] frontier
                                   (eligibility+work+binning) [1] hub-heavy sources (windowed counters + dense spill) [2] hash 16K
                                   [3] hash 4K [4] hash 1K [5] 32-lane [6] 8-lane; summed over the
                                   passes when the buffer had to be pruned (passes > 1; the top-K
                                   cuts between passes are not included).  NLP_PATH_PAIR:
                                   [0] eligible rows + item descriptors + scans [1] wedge-record
                                   emission [2] radix sort by (u, v) [3] run reduce + scoring.
                                   Both: [7] final select+sort.                                  */
  uint64_t pair_records;        /* NLP_PATH_PAIR: wedge records (u, v>u) emitted and sorted      */
} nlp_result;

/* Which kernels run the scoring phase.  Results are identical; the choice is about speed.
 *   NLP_PATH_SOURCE  one team per source vertex u: frontier over all first-hop entries, wedges
 *                    counted in shared-memory hash tables / global spill tables (IHub always).
 *   NLP_PATH_PAIR    LHub only, graphs with symmetric rows only (checked on the device): every
 *                    eligible row w emits its (u, v>u) pairs, the pairs are radix-sorted and
 *                    run-length reduced.  Falls back to NLP_PATH_SOURCE when not admissible.
 *   NLP_PATH_AUTO    NLP_PATH_PAIR when admissible, else NLP_PATH_SOURCE (default).             */
typedef enum { NLP_PATH_AUTO = 0, NLP_PATH_SOURCE = 1, NLP_PATH_PAIR = 2, NLP_PATH_PAIR_SORT = 3 } nlp_path;

/* Create a predictor bound to CUDA device `device` (one process per GPU). */
int nlp_create(nlp_handle** out, int device);
int nlp_destroy(nlp_handle* h);

/* CSR graph in: offsets[span+1] (64-bit, offsets[0] == 0, non-decreasing), keys[offsets[span]]
 * (32-bit vertex ids < span, each row sorted ascending; rows may hold duplicates -- entries are
 * counted, as the reference's forEachEdgeKey loops do).  HOST pointers; the library copies
 * them to its GPU.  Mirrors the read API the templates use on DiGraph / DiGraphCsr:
 * span(), degree(), hasVertex(), forEachEdgeKey() (inc/Graph.hxx:419,557,537,500).
 * The graph stays resident for any number of nlp_predict calls (main.cxx runs 99 per batch). */
int nlp_set_graph(nlp_handle* h, const uint64_t* offsets, const uint32_t* keys, uint32_t span);

/* Same, for arrays that already live in this GPU's memory (borrowed, not copied; must stay
 * valid until the next nlp_set_graph* / nlp_destroy).                                         */
int nlp_set_graph_device(nlp_handle* h, const uint64_t* d_offsets, const uint32_t* d_keys,
                         uint32_t span);

/* Multi-GPU: this handle processes the source vertices of partition `rank` of `world`: blocks
 * of 32 consecutive vertex ids are dealt round-robin to the ranks, which interleaves the
 * degree-correlated id ranges so every rank gets a near-equal share of each work bin.  The
 * graph itself is replicated.  Default (0, 1).  Replaces the OpenMP `schedule(dynamic,2048)`
 * split of inc/predict.hxx:287.                                                                */
int nlp_set_partition(nlp_handle* h, int rank, int world);

/* ---- multi-GPU inside the library ------------------------------------------
 * One process (and one handle) per GPU of a node, the CSR replicated on every GPU.  With a
 * communicator, nlp_predict itself merges across the ranks -- this is what replaces the serial
 * T-way heap merge of inc/predict.hxx:431-460:
 *   1. every rank scores the sources it owns: the LHub bucket path deals contiguous, wedge-work
 *      balanced source ranges (prefix sum of the wedge records per source, cut into equal
 *      parts); the source-centric kernels interleave blocks of 32 ids;
 *   2. the radix select's digit histograms are summed over the ranks with ncclAllReduce, so all
 *      ranks agree on the GLOBAL cutoff and each keeps only its candidates above it (~K/N);
 *   3. ONE ncclAllGather (preceded by an all-gather of the counts) moves the survivors;
 *   4. every rank runs the same final sort; the result (identical on all ranks, and identical to
 *      the single-GPU result: the canonical order does not depend on the partition) becomes the
 *      handle's result.
 * nlp_comm_unique_id: rank 0 creates the 128-byte id; the application hands it to the other
 * ranks over whatever it already has (MPI_Bcast, a torch.distributed broadcast, a file).
 * nlp_comm_init: collective over all ranks; also sets the partition (rank, world).  libnccl.so.2
 * is loaded on first use (NLP_ERR_COMM when it is absent).                                       */
#define NLP_COMM_ID_BYTES 128
int nlp_comm_unique_id(void* id);
int nlp_comm_init(nlp_handle* h, const void* id, int rank, int world);
int nlp_comm_destroy(nlp_handle* h);
/* Payload bytes this rank has received through the all-gathers so far. */
uint64_t nlp_comm_bytes(const nlp_handle* h);

/* Reuse across predictions on the same graph ("sweep fusion"; default off).
 * The nine measures share their common-neighbour counts, and main.cxx:212-220 asks for all of
 * them at the same thresholds: with reuse on, the pair path keeps the sorted wedge records of a
 * threshold (per D and partition, least-recently-used thresholds dropped beyond a quarter of the
 * scratch budget), and a later prediction at that threshold only runs the reduce / score / select
 * kernels.  Results are identical.  Every call (on or off) empties the store; nlp_set_graph* does
 * too.  Note that `repeat` > 1 then times one cold and repeat-1 warm scoring passes.             */
int nlp_set_reuse(nlp_handle* h, int on);

/* Force a scoring path (testing / measurement); default NLP_PATH_AUTO. */
int nlp_set_path(nlp_handle* h, int path);

/* Upper bound, in bytes, of GPU scratch (candidate buffer, spill tables) nlp_predict may use.
 * 0 = default (a fraction of the free memory at first use).                                    */
int nlp_set_scratch_limit(nlp_handle* h, uint64_t bytes);

/* Run one prediction.  Blocking.  The predicted edges stay in GPU memory, sorted by
 * (score desc, u asc, v asc), until the next nlp_predict / nlp_merge on this handle.           */
int nlp_predict(nlp_handle* h, const nlp_options* opt, nlp_result* res);

/* Copy the last result to caller arrays of at least `capacity` elements each (u < v); the
 * pointers may be host memory (the normal case) or memory of this GPU.
 * Copies min(capacity, count) edges.                                                           */
int nlp_fetch(nlp_handle* h, uint32_t* u, uint32_t* v, float* score, uint64_t capacity);

/* Same, without blocking: the result is first copied to a staging buffer on the GPU (so the next
 * nlp_predict may start at once), then transferred to the caller's arrays on a second stream.
 * The arrays (pinned host memory for a truly asynchronous transfer) must stay untouched until
 * nlp_fetch_wait() returns.  At most two transfers are in flight; a third call waits for the
 * oldest.  Use: predict, fetch_async, predict, fetch_async, ..., fetch_wait.                    */
int nlp_fetch_async(nlp_handle* h, uint32_t* u, uint32_t* v, float* score, uint64_t capacity);
int nlp_fetch_wait(nlp_handle* h);

/* Device pointers of the last result (count elements each), for on-device consumers and for
 * the multi-GPU all-gather.                                                                    */
int nlp_result_device(nlp_handle* h, const uint32_t** d_u, const uint32_t** d_v,
                      const float** d_score, uint64_t* count);

/* Multi-GPU merge step (replaces the T-way heap merge of inc/predict.hxx:431-460): given the
 * concatenated candidates of all ranks in THIS GPU's memory (n elements each), select the best
 * max_edges in canonical order; the result replaces the handle's last result.
 * select_ms (optional) receives the device time.                                               */
int nlp_merge(nlp_handle* h, const uint32_t* d_u, const uint32_t* d_v, const float* d_score,
              uint64_t n, uint64_t max_edges, float* select_ms);

/* ---- evaluation on the device ---------------------------------------
 * main.cxx scores every prediction against the edges the batch removed (PREDICT_LINKS,
 * main.cxx:48-57): both directions of the predicted edges are sorted and made unique
 * (main.cxx:51-54, directedInsertions main.cxx:97-104), intersected with the sorted directed list
 * of removed edges (commonEdges, main.cxx:126-133), and precision / recall are the size of the
 * intersection over the two list sizes (main.cxx:201-202).  With the result already in GPU
 * memory this is one kernel, and the K x 12 bytes of a prediction need not cross PCIe at all.   */

/* The held-back edges: n directed (u, v) pairs sorted ascending by (u, v) -- main.cxx's
 * `deletions0` (main.cxx:206-207; both directions of every removed undirected edge).  HOST (or
 * this GPU's) pointers; copied.  Independent of the graph; stays until the next nlp_set_truth.
 * NLP_ERR_ARG when the list is not sorted.                                                      */
int nlp_set_truth(nlp_handle* h, const uint32_t* u, const uint32_t* v, uint64_t n);

typedef struct {
  uint64_t predicted;   /* |insertions1| = 2 x predicted edges (main.cxx:51-54)                  */
  uint64_t truth;       /* |insertions0| = n of nlp_set_truth                                    */
  uint64_t common;      /* |commonEdges(insertions0, insertions1)| (main.cxx:55)                 */
  double   precision;   /* common / max(predicted, 1) (main.cxx:201)                             */
  double   recall;      /* common / max(truth, 1)     (main.cxx:202)                             */
  float    ms;          /* device time of the evaluation                                         */
} nlp_evaluation;

/* Evaluate the handle's last result (nlp_predict or nlp_merge) against the held-back edges.    */
int nlp_evaluate(nlp_handle* h, nlp_evaluation* out);

/* ---- batch generation on the device ------------------------------------
 * Before every sweep of predictions main.cxx removes random edges from the graph (runBatches,
 * main.cxx:165-168): generateEdgeDeletions (inc/batch.hxx:99-112: batch_size times, up to five
 * tries of "random vertex u in [1, span), then a random entry of row u"; the graph is not changed
 * while the batch is drawn) followed by tidyBatchUpdateU (inc/batch.hxx:200-208: sort by (u, v),
 * unique).  nlp_generate_deletions draws the same batch from the resident graph for
 * std::default_random_engine(seed) -- the same edges, draw for draw (the generator is a Lehmer
 * sequence, so every position of the stream is computed on its own; csrc/batch.cuh).
 * *count = directed pairs (both directions of every removed edge, sorted by (u, v), unique);
 * *words = engine outputs consumed (the stream position the reference's `rnd` is left at).
 * Uses the candidate buffers: the last prediction result is gone afterwards.                    */
int nlp_generate_deletions(nlp_handle* h, uint32_t seed, uint64_t batch_size, uint64_t* count,
                           uint64_t* words);

/* Copy the generated pairs to caller arrays (host or this GPU), min(capacity, count) of them.   */
int nlp_fetch_deletions(nlp_handle* h, uint32_t* u, uint32_t* v, uint64_t capacity);

/* Device pointers of the generated pairs, e.g. for nlp_set_truth(h, d_u, d_v, count): they are
 * exactly main.cxx's sorted `deletions0` (main.cxx:206-207).  Valid until the next
 * nlp_generate_deletions / nlp_destroy.                                                         */
int nlp_deletions_device(nlp_handle* h, const uint32_t** d_u, const uint32_t** d_v, uint64_t* count);

/* ---- apply a batch of deletions on the device (second half) -----------------
 * runBatches applies the tidied batch to a copy of the graph before the sweep of predictions
 * (applyBatchUpdateOmpU, inc/batch.hxx:239-247, main.cxx:164,169): removeEdge(u, v) for every
 * directed pair, then update().  nlp_apply_deletions does that to the RESIDENT graph without a host
 * round trip of the CSR: the first stored copy of every requested pair is marked by binary search,
 * the new degrees are scanned into new offsets and the keys are compacted.  n directed pairs
 * (host or this GPU's pointers -- e.g. nlp_deletions_device's), unique as tidyBatchUpdateU leaves
 * them; pairs that are not stored are ignored.  The result becomes the handle's graph (in memory
 * the handle owns: arrays lent with nlp_set_graph_device are not written, so the base graph can be
 * bound again for the next batch).  Uses the candidate buffers: the last result is gone.         */
int nlp_apply_deletions(nlp_handle* h, const uint32_t* del_u, const uint32_t* del_v, uint64_t n);

/* runBatches starts every batch from a copy of the loaded graph (main.cxx:164: y = duplicate(x)).
 * nlp_graph_checkpoint marks the resident graph as that base -- without copying: arrays the handle
 * owns are set aside (nlp_apply_deletions writes its result elsewhere), lent arrays are only
 * remembered -- and nlp_graph_rollback makes the base the resident graph again, without the
 * validation and symmetry passes a fresh nlp_set_graph* pays.                                    */
int nlp_graph_checkpoint(nlp_handle* h);
int nlp_graph_rollback(nlp_handle* h);

/* Size of the resident graph, and a copy of its CSR (offsets[span + 1], keys[entries]; host or
 * this GPU's pointers; either may be NULL).                                                      */
int nlp_graph_size(nlp_handle* h, uint32_t* span, uint64_t* entries);
int nlp_fetch_graph(nlp_handle* h, uint64_t* offsets, uint32_t* keys);

/* Graph ingest on the device.  Replaces main.cxx:243-245 --
 * readMtxOmpW (inc/mtx.hxx:151-188, header: inc/mtx.hxx:38-55), symmetrizeOmp
 * (inc/symmetrize.hxx:71-82), removeSelfLoopsOmpU (inc/selfLoop.hxx:117-124) -- entry for entry:
 * repeated lines collapse, but the reference's merge of the reverse edges (set_union_last_inplace,
 * inc/_algorithm.hxx:177-221) stores some entries that are in both directions of the file twice,
 * and the prediction counts entries, so that is reproduced (csrc/ingest.cuh).  text = the whole
 * Matrix Market coordinate file in HOST memory.  Body lines are "u v [weight]" with 1-based ids; the weight is ignored, blank and
 * '%' lines are skipped.  A "symmetric" / "skew-symmetric" banner stores both directions of every
 * line, as the reference's reader does; NLP_INGEST_SYMMETRIZE does it for a "general" file
 * (main.cxx:244), NLP_INGEST_DROP_SELF_LOOPS is main.cxx:245.  The CSR (span = max(rows, cols) + 1,
 * vertex 0 unused, rows sorted) is built on the GPU into arrays the handle owns
 * and becomes the resident graph; nlp_graph_size / nlp_fetch_graph return it.
 * NLP_ERR_ARG: not a coordinate matrix, or a vertex id outside 1..max(rows, cols).
 * NLP_ERR_CAPACITY: 2^32 - 16 or more directed pairs before deduplication.  Everything is resident
 * at once (about 100 bytes of GPU memory per line of the file).                                    */
#define NLP_INGEST_SYMMETRIZE      1u
#define NLP_INGEST_DROP_SELF_LOOPS 2u
int nlp_ingest_mtx(nlp_handle* h, const char* text, uint64_t bytes, uint32_t flags, uint32_t* span, uint64_t* entries);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
uint64_t nlp_launch_count(const nlp_handle* h);

/* The cudaStream_t (as void*) all work of this handle is issued on. */
void* nlp_stream(const nlp_handle* h);

/* Message for the most recent non-OK status of this handle (or of nlp_create when h == NULL). */
const char* nlp_last_error(const nlp_handle* h);

const char* nlp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NLP_B200_H */
