/*
 * cadence_dense.h -- C ABI of the B200-native dense-retrieval engine (libcadence_dense.so).
 *
 * This is the drop-in boundary for cadence-rag's /retrieve dense lane.  Every entry point
 * names the reference interface it replaces (paths relative to the reference repo).  The
 * reference reaches its dense lane through SQL sent to Postgres+pgvector; the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md (cadence_rag_b200/_ffi.py is
 * that stub, shipped).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns int32 status: CDR_OK (0) or a negative CDR_ERR_* code; a
 *     human-readable message for the calling thread is available from cdr_last_error().
 *   - "dev" pointers are CUDA device pointers on the store's device; "host" pointers are
 *     ordinary (ideally pinned) host memory.  The caller owns every query/output buffer.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless the name ends in `_host` (those synchronise).
 *   - rows are addressed by their position in the store ("row"); ids are the reference's
 *     BIGSERIAL chunk_id / artifact_chunk_id values (alembic/versions/0001_initial_schema.py:78,
 *     0006_add_artifact_chunks.py:22) and MUST be appended in strictly increasing order
 *     (the order `SELECT ... ORDER BY chunk_id` yields), so that "ties broken by chunk_id"
 *     equals "ties broken by row".
 *   - threading: entry points are re-entrant.  A store's kernel workspaces are per stream; calls that share a
 *     stream are serialised by stream order (the store lock is held while work is enqueued), calls on different
 *     streams overlap on the device.  The *_host entry points stage requests and responses in per-thread pinned
 *     and device buffers, so concurrent requests never share staging.  Mutating calls (append, update, finalize)
 *     take the same lock; searches issued after they return see the new rows.
 *   - ordering of every dense result: score descending, NaN scores last (a zero-norm vector
 *     gives a NaN cosine distance in pgvector, which PostgreSQL sorts last), ties by id
 *     ascending.  Fewer survivors than k => short list (SQL LIMIT semantics).
 */
#ifndef CADENCE_DENSE_H
#define CADENCE_DENSE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDR_ABI_VERSION 1

/* status codes */
#define CDR_OK                 0
#define CDR_ERR_INVALID       -1   /* bad argument */
#define CDR_ERR_CUDA          -2   /* CUDA runtime/driver error (see cdr_last_error) */
#define CDR_ERR_OOM           -3   /* device allocation failed / capacity exceeded */
#define CDR_ERR_UNSORTED_IDS  -4   /* ids not strictly increasing */
#define CDR_ERR_STATE         -5   /* store not in the right state for the call */
#define CDR_ERR_UNSUPPORTED   -6   /* dim / k / precision combination not built */
#define CDR_ERR_NO_DEVICE     -7   /* no CUDA device: there is NO CPU fallback */

/* store flags: which copies of the embeddings stay resident */
#define CDR_STORE_FP32  1u   /* fp32 rows (exact lane, fp64 re-score) */
#define CDR_STORE_BF16  2u   /* bf16 rows, L2-normalised (batched tensor-core lane) */

#define CDR_MAX_K 248        /* largest LIMIT served by the fused top-k paths */

typedef struct cdr_store cdr_store;

/* ---- library ------------------------------------------------------------------------------ */
int32_t     cdr_abi_version(void);
const char *cdr_last_error(void);                 /* thread-local, never NULL */
int32_t     cdr_device_count(int32_t *out_count); /* CDR_ERR_NO_DEVICE when none */

/* ---- resident corpus store: one per reference table ("chunks", "artifact_chunks") ----------
 * Replaces the `embedding vector(1024)` column + its filter columns
 * (alembic/versions/0001_initial_schema.py:78-87, 0006_add_artifact_chunks.py:22-33) as read by
 * app/retrieve.py:339-351 / 374-386. */
int32_t cdr_store_create(cdr_store **out, int32_t device, int64_t capacity_rows, int32_t dim,
                         uint32_t flags);
int32_t cdr_store_destroy(cdr_store *s);

/* Append n rows.  Replaces app/embedding_pipeline.py:149-168 (_update_embeddings: one
 * `UPDATE ... SET embedding = CAST(:e AS vector(D))` per row).  rows_f32 is [n, dim] row-major.
 * valid_u8 (nullable) marks rows whose embedding IS NOT NULL (1) / IS NULL (0; row content
 * ignored).  call_slot is the dictionary code of call_id; started_at_us is call_started_at in
 * microseconds since the Unix epoch; tag_bits is the per-call tag dictionary mask.
 * is_device != 0 => all pointers are device pointers. */
int32_t cdr_store_append(cdr_store *s, const float *rows_f32, const int64_t *ids,
                         const int32_t *call_slot, const int64_t *started_at_us,
                         const uint64_t *tag_bits, const uint8_t *valid_u8, int64_t n,
                         int32_t is_device, void *stream);

/* A sealed (finalized) store stays live:
 *  - cdr_store_append keeps working after cdr_store_finalize for rows whose ids continue the increasing
 *    sequence (new chunks of a newly ingested call); the order is verified before anything is written,
 *    so a rejected batch (CDR_ERR_UNSORTED_IDS) leaves the store untouched.
 *  - cdr_store_update_embeddings replaces app/embedding_pipeline.py:149-168 (_update_embeddings) for rows
 *    that already exist -- the backfill of rows ingested with `embedding IS NULL`:
 *        UPDATE <table> SET embedding = CAST(:embedding AS vector(D)) WHERE <id> = :row_id
 *    ids_host[n] are located by binary search on the device; rows_f32_host is [n, dim].  Each row's fp32
 *    copy, inverse norm, normalised bf16 copy and NOT NULL bit are rewritten.  All-or-nothing: if any id
 *    is unknown the call fails with CDR_ERR_INVALID and nothing is updated.  Ids must be distinct.
 * Both synchronise `stream`; searches issued afterwards on that stream see the new rows. */
int32_t cdr_store_update_embeddings(cdr_store *s, const int64_t *ids_host, const float *rows_f32_host,
                                    int64_t n, void *stream);

/* Append n synthetic rows generated on the device (global rows first_row..first_row+n-1 of the
 * counter-based corpus specified in oracle/synth_ref.c).  id = id_base + global
 * row, call_slot = global_row / rows_per_call, started_at = t0_us + call_slot * call_period_us,
 * tag_bits = two Philox-chosen tags of 16 per call. */
int32_t cdr_store_append_synthetic(cdr_store *s, uint64_t seed, int64_t first_row, int64_t n,
                                   int64_t id_base, int32_t rows_per_call, int64_t t0_us,
                                   int64_t call_period_us, void *stream);

/* Seal the store: checks id monotonicity, builds the valid bitmap / search workspaces. */
int32_t cdr_store_finalize(cdr_store *s, void *stream);

/* rows, dim, flags, n_valid (rows with embedding IS NOT NULL), device */
int32_t cdr_store_info(const cdr_store *s, int64_t *rows, int32_t *dim, uint32_t *flags,
                       int64_t *n_valid, int32_t *device);

/* Copy resident data back to the host (tests, snapshots).  Any out pointer may be NULL. */
int32_t cdr_store_read_rows(cdr_store *s, int64_t first_row, int64_t n, float *out_f32_host,
                            uint16_t *out_bf16_host, int64_t *out_ids_host,
                            int32_t *out_call_slot_host, int64_t *out_started_at_host,
                            uint64_t *out_tag_bits_host, float *out_inv_norm_host);

/* The same for device destinations (checkers that recompute scores on the GPU without a host round trip, e.g. the
 * independent matmul oracle of bench.py): rows [first_row, first_row+n) as stored -- fp32 and / or the normalised
 * bf16 copy -- into caller-owned DEVICE buffers, asynchronously on `stream`.  Either out pointer may be NULL. */
int32_t cdr_store_copy_rows_device(cdr_store *s, int64_t first_row, int64_t n, float *out_f32_dev,
                                   uint16_t *out_bf16_dev, void *stream);

/* Validity flags (embedding IS NOT NULL) of rows [first_row, first_row+n) as bytes (snapshots). */
int32_t cdr_store_read_valid(cdr_store *s, int64_t first_row, int64_t n, uint8_t *out_valid_u8_host);

/* ---- synthetic queries / rows into caller memory ---------------------------------------- */
int32_t cdr_synth_rows(float *out_dev, uint64_t seed, int64_t first_row, int64_t n, int32_t dim,
                       void *stream);

/* ---- K6 filter bitmap --------------------------------------------------------------------
 * Replaces app/retrieve.py:93-120 (_build_filter_clause) evaluated per row inside Postgres, and
 * app/retrieve.py:303-323 (_estimate_dense_candidates: exact COUNT(*) ... AND embedding IS NOT
 * NULL) via out_count.  Predicate per row:
 *   valid AND (call_slot_bitmap == NULL OR bit(call_slot))            -- call_id = ANY(:call_ids)
 *         AND (!has_date_from OR started_at >= date_from_us)          -- call_started_at >= :date_from
 *         AND (!has_date_to   OR started_at <= date_to_us)            -- call_started_at <= :date_to
 *         AND (!has_tag_filter OR (tag_bits & tag_any) != 0)          -- c.tags && :call_tags
 * call_slot_bitmap_host: nullable host bitmap over call slots (n_call_slots bits, uint32 words);
 * a non-NULL all-zero bitmap is `call_ids == []` (matches nothing).
 * out_allow_dev: uint32[ceil(CAPACITY/32)] device bitmap (capacity_rows of cdr_store_create), bit r => row r passes;
 *   bits of rows beyond the row count at the time of the call are written as 0.  Sizing by capacity is what makes a
 *   bitmap safe to use while the sealed store grows (cdr_store_append): a later scan that already sees the larger row
 *   count reads "not allowed" for the new rows instead of reading past the bitmap.  Every allow_dev argument below
 *   must cover ceil(capacity/32) words.
 * out_count_host: number of passing rows (the call synchronises the stream to return it). */
int32_t cdr_filter_build(cdr_store *s, const uint32_t *call_slot_bitmap_host,
                         int64_t n_call_slots, int32_t has_date_from, int64_t date_from_us,
                         int32_t has_date_to, int64_t date_to_us, int32_t has_tag_filter,
                         uint64_t tag_any, uint32_t *out_allow_dev, int64_t *out_count_host,
                         void *stream);

/* ---- K1 + K3/K4: exact fp32 cosine scan with fused top-k, fp64 re-score -------------------
 * Replaces the SQL at app/retrieve.py:339-351 (_fetch_chunks_dense) / 374-386
 * (_fetch_artifacts_dense) in mode "exact":
 *   SELECT id, 1 - (embedding <=> :q) AS score ... WHERE <filters> AND embedding IS NOT NULL
 *   ORDER BY embedding <=> :q LIMIT :k
 * q_dev: [nq, dim] fp32.  allow_dev: bitmap from cdr_filter_build or NULL (all valid rows).
 * out_score_dev f64[nq,k] (= 1 - cosine distance), out_id_dev i64[nq,k], out_n_dev i32[nq].
 * Unused tail slots are filled with score NaN / id -1. */
int32_t cdr_search_exact_f32(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                             const uint32_t *allow_dev, double *out_score_dev,
                             int64_t *out_id_dev, int32_t *out_n_dev, void *stream);

/* Same call with HOST buffers: copies the queries host->device, searches, copies the results
 * device->host and synchronises.  This is the call the reference-facing plugin makes per
 * request (the counterpart of conn.execute(...).mappings() at app/retrieve.py:340-353). */
int32_t cdr_search_exact_f32_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                  const uint32_t *allow_dev, double *out_score_host,
                                  int64_t *out_id_host, int32_t *out_n_host, void *stream);

/* Mode "ann" for a single query (or a few): the same HBM-bound scan over the store's bf16 copy of the rows -- half the
 * bytes of the fp32 scan -- with candidate lists twice as wide (128 for k <= 120), and the survivors re-scored exactly
 * (fp64 on the fp32 rows when they are resident, else on the bf16 rows).  Where the reference's planner says "ann" it
 * walks an HNSW index with ef_search = 80 (app/retrieve.py:291-298); this lane answers with recall ~1.0 against the
 * exact scan (tests/test_gpu_engine.py::test_bf16_scan_lane).  Needs CDR_STORE_BF16, dim in {256,512,768,1024}. */
int32_t cdr_search_scan_bf16(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                             const uint32_t *allow_dev, double *out_score_dev,
                             int64_t *out_id_dev, int32_t *out_n_dev, void *stream);
int32_t cdr_search_scan_bf16_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                  const uint32_t *allow_dev, double *out_score_host,
                                  int64_t *out_id_host, int32_t *out_n_host, void *stream);

/* The same lane for a BATCH of concurrent exact requests: every tile a CTA streams from HBM is scored
 * against several queries ("shared reads"): 3 held in registers, or -- for 7 or more queries at k <= 56 -- 8
 * streamed from shared memory while the rows sit in registers, so nq queries cost nq/3 .. nq/8 scans of the corpus.
 * cdr_search_exact_f32 above keeps one scan per query (the single-query GEMV of the reference's
 * one-query-per-request flow); per-query arithmetic is the same, and so is every bit of the result. */
int32_t cdr_search_exact_f32_shared(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                                    const uint32_t *allow_dev, double *out_score_dev,
                                    int64_t *out_id_dev, int32_t *out_n_dev, void *stream);
int32_t cdr_search_exact_f32_shared_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                         const uint32_t *allow_dev, double *out_score_host,
                                         int64_t *out_id_host, int32_t *out_n_host, void *stream);

/* ---- K2 + K3: batched bf16 tensor-core scan (tcgen05) with fused threshold top-k epilogue --
 * Serves mode "ann" of app/retrieve.py:290-298 (_configure_dense_session: HNSW ef_search) by
 * brute force on the tensor cores: the score matrix never reaches HBM; survivors are re-scored
 * exactly (fp64 accumulate on the fp32 rows when resident, else on the bf16 rows) and ordered
 * like the exact lane.  q_dev: [nq, dim] fp32 (converted to bf16 internally).
 * Asynchronous on `stream` like every device-pointer entry point: a candidate-list overflow (adversarial data) is
 * detected on the device and the affected queries are re-run on the exact lane (the bf16 scan lane for bf16-only
 * stores) by launches that read the query list from device memory -- no host round trip.  Rows whose score is NaN
 * (zero or non-finite embeddings) are admitted as in the exact lane: below every real score, in id order. */
int32_t cdr_search_batch_bf16(cdr_store *s, const float *q_dev, int32_t nq, int32_t k,
                              const uint32_t *allow_dev, double *out_score_dev,
                              int64_t *out_id_dev, int32_t *out_n_dev, void *stream);
int32_t cdr_search_batch_bf16_host(cdr_store *s, const float *q_host, int32_t nq, int32_t k,
                                   const uint32_t *allow_dev, double *out_score_host,
                                   int64_t *out_id_host, int32_t *out_n_host, void *stream);

/* ---- K4: k-way merge of per-shard top-k lists ---------------------------------------------
 * Row-sharded corpora (one store per GPU): after the NCCL all-gather of every rank's
 * [nq,k] (score,id,n) lists this merges R lists per query into the global top-k with the
 * same ordering rule.  scores f64[R,nq,k], ids i64[R,nq,k], n i32[R,nq].  Every input list must be in
 * result order (score desc, NaN last, id asc) -- as every search entry point returns it: the merge ranks
 * an entry by its own position plus a binary search in each of the other lists. */
int32_t cdr_topk_merge(const double *scores_dev, const int64_t *ids_dev, const int32_t *n_dev,
                       int32_t R, int32_t nq, int32_t k, double *out_score_dev,
                       int64_t *out_id_dev, int32_t *out_n_dev, void *stream);

/* ---- K4p: the same exchange + merge over NVLink peer memory, one kernel per rank --------------
 * Alternative transport for the step above when all ranks are processes on one node: every rank
 * owns a receive buffer that its peers map with CUDA IPC; cdr_peer_exchange_merge launches ONE
 * kernel that pushes this rank's [nq,k] lists into every rank's buffer (posted NVLink stores +
 * a system-scope release flag per query), waits for the peers' lists on local memory, and merges
 * them with the K4 ordering rule.  No NCCL call, no pack/unpack copies; identical result on every
 * rank.  Protocol (csrc/peer.cu): every rank issues the same sequence of calls on one stream.
 *   create  : allocates the buffer for batches of <= max_nq queries, lists of <= max_k; writes the
 *             64-byte IPC handle to out_handle.  Exchange the handles of all ranks (rank-major,
 *             e.g. torch.distributed.all_gather_object) and pass them to connect.
 *   connect : maps the peers' buffers.  CDR_ERR_UNSUPPORTED when a peer cannot be mapped (ranks on
 *             different nodes / no P2P): the caller then stays on the NCCL transport. */
#define CDR_PEER_MAX_RANKS 16
#define CDR_PEER_HANDLE_BYTES 64
typedef struct cdr_peer_group cdr_peer_group;
int32_t cdr_peer_group_create(cdr_peer_group **out, int32_t device, int32_t rank, int32_t world,
                              int32_t max_nq, int32_t max_k, void *out_handle);
int32_t cdr_peer_group_connect(cdr_peer_group *g, const void *all_handles);
int32_t cdr_peer_group_destroy(cdr_peer_group *g);
int32_t cdr_peer_exchange_merge(cdr_peer_group *g, const double *scores_dev, const int64_t *ids_dev,
                                const int32_t *n_dev, int32_t nq, int32_t k, double *out_score_dev,
                                int64_t *out_id_dev, int32_t *out_n_dev, void *stream);

/* One rank's whole step over a row-sharded table in ONE call: the local lane over this rank's shard followed by the
 * K4p exchange + merge, enqueued back to back on `stream` (the local lists live in the group's own buffers).  On a
 * small shard a single-query request is a few tens of microseconds of kernels; issuing its launches from one C call
 * keeps the GPU from idling between them while the host returns to Python and calls again.
 * lane: CDR_DENSE_LANE_EXACT_F32 (one scan per query), CDR_DENSE_LANE_EXACT_F32_SHARED (shared reads for a batch),
 * CDR_DENSE_LANE_SCAN_BF16 or CDR_DENSE_LANE_BATCH_BF16.  g == NULL (one rank): the lane alone.  Same arguments,
 * results and ordering as the lane's own entry point followed by cdr_peer_exchange_merge; k <= the group's max_k.
 * For the three scan lanes (<= 8 ranks, k <= 56 on fp32 rows / k <= 64 on bf16 rows) the exchange runs INSIDE the lane's
 * finalize kernel -- the CTA that ordered a query's list pushes it to the peers, waits for theirs and merges -- so there
 * is no separate exchange launch and no local list round trip; other shapes enqueue the K4p kernel behind the lane. */
#define CDR_DENSE_LANE_EXACT_F32_SHARED 3
int32_t cdr_search_sharded(cdr_store *s, cdr_peer_group *g, int32_t lane, const float *q_dev, int32_t nq, int32_t k,
                           const uint32_t *allow_dev, double *out_score_dev, int64_t *out_id_dev,
                           int32_t *out_n_dev, void *stream);

/* ---- K5: reciprocal-rank fusion -------------------------------------------------------------
 * Replaces app/retrieve.py:245-260 (_rrf_merge), bit-exact: for lanes in order (bm25,
 * tech_tokens, dense; app/retrieve.py:537-547) and ranks from 1,
 *   score[id] = score.get(id, 0.0) + 1.0 / (rrf_k + rank)           (IEEE fp64, no contraction)
 * then a stable descending sort, so ties keep first-seen order.
 * lane_ids_dev: all lanes of all queries concatenated; lane_offsets_dev: i32[nq*L + 1], lane l of
 * query q is lane_ids[off[q*L+l] .. off[q*L+l+1]).  At most CDR_RRF_MAX_ITEMS ids per query.
 * Outputs are [nq, max_out]: fused ids, fp64 scores, lane-hit bit masks (bit l = lane l),
 * and out_n[nq] = number of distinct ids. */
#define CDR_RRF_MAX_ITEMS 1024
int32_t cdr_rrf_merge(const int64_t *lane_ids_dev, const int32_t *lane_offsets_dev, int32_t nq,
                      int32_t L, int32_t rrf_k, int32_t max_out, int64_t *out_ids_dev,
                      double *out_scores_dev, uint32_t *out_lane_mask_dev, int32_t *out_n_dev,
                      void *stream);
int32_t cdr_rrf_merge_host(const int64_t *lane_ids_host, const int32_t *lane_offsets_host,
                           int32_t nq, int32_t L, int32_t rrf_k, int32_t max_out,
                           int64_t *out_ids_host, double *out_scores_host,
                           uint32_t *out_lane_mask_host, int32_t *out_n_host, void *stream);

/* ---- tech_tokens lexical lane, device resident -------------------------
 * Replaces the SQL of _fetch_chunks_tech / _fetch_artifacts_tech (app/retrieve.py:183-242):
 *   WHERE <filters> AND tech_tokens && :tokens ORDER BY call_started_at DESC, id ASC LIMIT :limit
 * The index is CSR postings over dictionary-encoded tokens (token t -> rows
 * post_rows[post_offsets[t] .. post_offsets[t+1]), any order on input) plus rank[row] = position of the row in the
 * order (call_started_at DESC, id ASC); cdr_tech_index_create keeps every list in rank order on the device, so a
 * query reads the head of each list instead of all of it.  Queries carry token ids (-1 = unknown token).  The filter
 * arguments are those of cdr_filter_build, without the `embedding IS NOT NULL` term (the tech
 * lane's WHERE has none).  Host buffers; synchronises.  out_ids [nq, limit] (unused slots -1). */
typedef struct cdr_tech_index cdr_tech_index;
int32_t cdr_tech_index_create(cdr_tech_index **out, cdr_store *s, const int64_t *post_offsets_host,
                              int32_t n_tokens, const uint32_t *post_rows_host,
                              const uint32_t *rank_host);
int32_t cdr_tech_index_destroy(cdr_tech_index *ix);
int32_t cdr_tech_lane_host(cdr_tech_index *ix, const int32_t *token_ids_host,
                           const int32_t *n_tokens_host, int32_t nq, int32_t max_tokens,
                           const uint32_t *call_slot_bitmap_host, int64_t n_call_slots,
                           int32_t has_date_from, int64_t date_from_us, int32_t has_date_to,
                           int64_t date_to_us, int32_t has_tag_filter, uint64_t tag_any,
                           int32_t limit, int64_t *out_ids_host, int32_t *out_n_host, void *stream);

/* ---- fused hybrid /retrieve for one table ---------------------------------------------------
 * One call = what retrieve_evidence runs per request and table between `with engine.connect()` and
 * the RRF (app/retrieve.py:445-550):
 *   _fetch_chunks_tech (:183-209)  ->  _estimate_dense_candidates (:303-323)  ->
 *   _fetch_chunks_dense mode "exact" (:326-354)  ->  _rrf_merge({bm25, tech_tokens, dense}) (:245-260)
 * for nq queries that share one filter, with ONE host->device copy of the packed request, ONE
 * device->host copy of the packed response and ONE stream synchronisation.  The BM25 lane
 * (pg_search; out of scope) is an input: bm25_ids_host holds each query's ranked ids back to back,
 * bm25_offsets_host[nq+1] delimits them (both NULL = empty lane).
 *   filter        NULL or all-empty = unscoped.  Applied to both lanes; only the dense lane also
 *                 requires `embedding IS NOT NULL` (as in the reference SQL).
 *   q_host        [nq, dim] fp32, or NULL = dense lane disabled (EmbeddingClientError path,
 *                 app/retrieve.py:426-432): the fusion then has the two lexical lanes only.
 *   tech_index    NULL = empty tech_tokens lane; token ids as for cdr_tech_lane_host.
 * Lane order in the fusion and in the lane-hit masks: bit 0 bm25, bit 1 tech_tokens, bit 2 dense.
 * Outputs (host): out_count = COUNT(*) of the dense lane's WHERE clause (n_valid when unscoped);
 * dense ids/scores [nq, dense_k] + n; tech ids [nq, tech_limit] + n; fused ids/scores/masks
 * [nq, max_out] + n.  Unused slots: id -1. */
typedef struct cdr_filter_spec {
    const uint32_t *call_slot_bitmap_host;   /* NULL = no call filter; all-zero = call_ids == [] */
    int64_t n_call_slots;
    int32_t has_date_from;
    int32_t has_date_to;
    int64_t date_from_us;
    int64_t date_to_us;
    int32_t has_tag_filter;
    int32_t dense_lane;                      /* CDR_DENSE_LANE_*: which kernel serves the group's dense lane */
    uint64_t tag_any;
} cdr_filter_spec;

/* Dense lane of a group.  EXACT_F32: the exact fp32 scan (shared reads inside the group) -- always right, and what
 * mode "exact" requires.  BATCH_BF16: the tensor-core lane of cdr_search_batch_bf16 (candidates from bf16 products,
 * re-scored in fp64 on the fp32 rows) for groups whose planner mode is "ann" (unscoped requests: the reference walks its
 * HNSW index there, app/retrieve.py:291-298); needs CDR_STORE_BF16, dense_k <= 192. */
#define CDR_DENSE_LANE_EXACT_F32 0
#define CDR_DENSE_LANE_BATCH_BF16 1
#define CDR_DENSE_LANE_SCAN_BF16 2   /* cdr_search_scan_bf16: one HBM-bound scan of the bf16 rows per query (mode "ann",
                                        single requests / small groups), exact re-score; dim in {256,512,768,1024} */

int32_t cdr_hybrid_retrieve_host(
    cdr_store *s, cdr_tech_index *tech_index, const cdr_filter_spec *filter, const float *q_host, int32_t nq,
    int32_t dense_k, const int32_t *token_ids_host, const int32_t *n_tokens_host, int32_t max_tokens,
    int32_t tech_limit, const int64_t *bm25_ids_host, const int32_t *bm25_offsets_host, int32_t rrf_k,
    int32_t max_out, int64_t *out_count_host, int64_t *out_dense_ids_host, double *out_dense_scores_host,
    int32_t *out_dense_n_host, int64_t *out_tech_ids_host, int32_t *out_tech_n_host, int64_t *out_fused_ids_host,
    double *out_fused_scores_host, uint32_t *out_fused_mask_host, int32_t *out_fused_n_host, void *stream);

/* The same call for requests with DIFFERENT filters: the nq queries are ordered so that queries sharing a
 * filter are consecutive; group g covers queries [group_offsets_host[g], group_offsets_host[g+1]) and uses
 * filters[g] (an all-empty spec = unscoped; filters == NULL = every group unscoped).  Each group runs its own
 * K6 / tech-lane / dense launches (shared corpus reads inside a group); lane assembly, RRF, the copies and the
 * single synchronisation cover the whole block.  out_count_host has n_groups entries.  This is what a
 * micro-batcher in front of concurrent /retrieve requests calls (cadence_rag_b200.retrieve.RequestBatcher). */
int32_t cdr_hybrid_retrieve_groups_host(
    cdr_store *s, cdr_tech_index *tech_index, const cdr_filter_spec *filters, const int32_t *group_offsets_host,
    int32_t n_groups, const float *q_host, int32_t nq, int32_t dense_k, const int32_t *token_ids_host,
    const int32_t *n_tokens_host, int32_t max_tokens, int32_t tech_limit, const int64_t *bm25_ids_host,
    const int32_t *bm25_offsets_host, int32_t rrf_k, int32_t max_out, int64_t *out_count_host,
    int64_t *out_dense_ids_host, double *out_dense_scores_host, int32_t *out_dense_n_host, int64_t *out_tech_ids_host,
    int32_t *out_tech_n_host, int64_t *out_fused_ids_host, double *out_fused_scores_host, uint32_t *out_fused_mask_host,
    int32_t *out_fused_n_host, void *stream);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* Number of this library's kernels launched by the calling process so far (bench.py's
 * gpu_launches claim). */
int64_t cdr_kernel_launch_count(void);
/* Average device time (ms) of the dominant scan kernel measured with CUDA events on `stream`
 * between cdr_prof_begin / cdr_prof_end (bench.py's roofline.achieved).  kind: 0 = K1 exact
 * scan, 1 = K2 batched bf16. */
int32_t cdr_prof_enable(int32_t on);
int32_t cdr_prof_read(int32_t kind, double *out_total_ms, int64_t *out_launches);
/* Per-launch device times (ms) of the recorded launches, oldest first; writes min(n, max_n). */
int32_t cdr_prof_read_launches(int32_t kind, double *out_ms, int64_t max_n, int64_t *out_n);

#ifdef __cplusplus
}
#endif
#endif /* CADENCE_DENSE_H */
