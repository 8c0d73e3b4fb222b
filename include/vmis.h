/* include/vmis.h — C ABI of the B200-native VMIS-kNN `predict_next` path.
 *
 * Drop-in boundary for bolcom/serenade's `src/vmisknn` hot path.  Every entry
 * point cites the reference interface it replaces (file:line under the
 * reference tree).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - Item ids are the reference's external u64 ids (io.rs:10 `ItemId = u64`).
 *   - Training session ids are the reference's u32 session indices
 *     (io.rs:9; position in `session_to_items_sorted`, vmis_index.rs:32).
 *   - Query batches are CSR: q_items[q_off[q] .. q_off[q+1]) is the evolving
 *     session of query q, oldest → newest (mod.rs:120 `evolving_session`).
 *   - All functions returning int return 0 on success, <0 on error; the message
 *     is available from vmis_last_error() (thread-local).  Nothing aborts the
 *     process (the reference `unwrap()`s / panics instead — mod.rs:138,157).
 *   - There is NO CPU fallback: query entry points fail with VMIS_ERR_CUDA if
 *     no sm_100 device is usable.
 *   - Query entry points are re-entrant: many host threads may call them
 *     concurrently on one index (mirrors actix workers sharing Arc<VMISIndex>,
 *     serving.rs:39,65).
 */
#ifndef VMIS_H_
#define VMIS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMIS_OK 0
#define VMIS_ERR_ARG (-1)      /* bad argument                                  */
#define VMIS_ERR_IO (-2)       /* file could not be read / parsed               */
#define VMIS_ERR_CUDA (-3)     /* CUDA runtime error or no usable device        */
#define VMIS_ERR_LIMIT (-4)    /* k / m / session length beyond kernel limits   */

/* Longest evolving session a query may carry (the reference's HPO grid stops at 100 items,
 * hyperparameter_search.rs:20; the session weight of mod.rs:110-116 is zero from position 100 on). */
#define VMIS_MAX_SESSION_LEN 128
/* vmis_predict_batch_device cannot fail a launch that is already queued: a query whose evolving session exceeds
 * VMIS_MAX_SESSION_LEN gets this value in out_counts (and no recommendations) instead of a count. */
#define VMIS_COUNT_TOO_LONG 0xFFFFFFFFu

/* device ordinal for a host-only handle: the index is built and the trait accessors work, but every
 * query entry point fails with VMIS_ERR_CUDA (used by CPU-side tests of the builders). */
#define VMIS_DEVICE_NONE (-1)

/* ProductAttributes (vmis_index.rs:23-26) packed in one byte. */
#define VMIS_ATTR_EXISTS 1u
#define VMIS_ATTR_FOR_SALE 2u
#define VMIS_ATTR_ADULT 4u

typedef struct vmis_index vmis_index_t; /* opaque: host mirror + device CSR arrays (VMISIndex, vmis_index.rs:28-35) */

typedef struct vmis_stats {
  uint64_t n_sessions;        /* all training sessions (un-pruned, vmis_index.rs:79)   */
  uint64_t n_sessions_kept;   /* sessions with len <= max_len (vmis_index.rs:452)      */
  uint64_t n_items;           /* distinct items in kept sessions                       */
  uint64_t n_pairs_kept;      /* (session,item) pairs kept = idf numerator (:509-512)  */
  uint64_t n_postings;        /* posting entries after truncation to m (:504)          */
  uint64_t max_len;           /* max_training_session_length (:67)                     */
  uint64_t m_build;           /* m_most_recent_sessions used at build                  */
  uint64_t device_bytes;      /* HBM bytes held by the index                           */
  double idf_weighting;
} vmis_stats_t;

/* Report of a pre-computed index load (vmis_index_from_avro / vmis_index_from_parts). */
typedef struct vmis_prebuilt_info {
  uint64_t item_files, session_files;      /* .avro files read under itemindex/ and sessionindex/              */
  uint64_t item_records, session_records;  /* records decoded (before "last record of a key wins")             */
  uint64_t lists_reordered;                /* posting lists not in (timestamp desc, session idx desc) order     */
  uint64_t duplicate_postings;             /* repeated session ids dropped from posting lists                   */
  uint64_t pruned_postings;                /* postings dropped because their session exceeds max_session_len    */
  uint32_t m_carry;                        /* m <= m_carry: first-match positions are carried through the list  */
                                           /* merges; larger m scans the item lists like mod.rs:133-138         */
  uint32_t prebuilt;                       /* 1 = the handle was made from pre-computed parts                   */
} vmis_prebuilt_info_t;

/* Per-query work counters written by the kernel when requested (bench: exact
 * algorithmic bytes per SURVEY.md §8d). */
typedef struct vmis_query_stats {
  uint32_t postings_visited;  /* sum_j min(df_j, m) over distinct known items  */
  uint32_t n_neighbors;       /* |N| <= k                                      */
  uint32_t neighbor_items;    /* sum of len_s over s in N                      */
  uint32_t n_out;             /* recommendations written                       */
} vmis_query_stats_t;

/* ---- construction ------------------------------------------------------- */

/* VMISIndex::new_from_csv(path, m_most_recent_sessions, idf_weighting) — vmis_index.rs:38-83
 * (read_from_file :591-752 + prepare_hashmap :422-528).  max_training_session_length is the
 * p99.5 of session lengths as in :67,:715 (t-digest restated).  device = CUDA ordinal. */
vmis_index_t* vmis_index_from_csv(const char* path, size_t m, double idf_weighting, int device);

/* Same, with max_training_session_length given explicitly (0 = compute p99.5). */
vmis_index_t* vmis_index_from_csv_ex(const char* path, size_t m, double idf_weighting, size_t max_len, int device);

/* read_from_file (vmis_index.rs:591-752) on its own: the training sessions of a TSV exactly as new_from_csv sees them
 * (session index = rank of the session id, items ascending, max timestamp; incl. the last-row quirk).  For callers that
 * build many indexes from one file — the HPO objective rebuilds the index for every trial (objective.rs:17): parse
 * once, then vmis_index_from_sessions(..., m, max_len = 0, idf_weighting, ...) per trial gives the same index as
 * vmis_index_from_csv.  The pointers of the view stay valid until vmis_sessions_free. */
typedef struct vmis_sessions vmis_sessions_t;
vmis_sessions_t* vmis_sessions_from_csv(const char* path);
int vmis_sessions_view(const vmis_sessions_t* sessions, const uint64_t** items, const uint64_t** sess_off,
                       const uint32_t** sess_ts, size_t* n_sessions);
void vmis_sessions_free(vmis_sessions_t* sessions);

/* prepare_hashmap(historical_sessions, timestamps, m, max_len, idf_weighting) — vmis_index.rs:422-528,
 * followed by the VMISIndex{..} assembly of :75-82.  items[sess_off[s]..sess_off[s+1]) are the item ids
 * of training session s (any order; duplicates not allowed), sess_ts[s] its max timestamp. */
vmis_index_t* vmis_index_from_sessions(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                       size_t n_sessions, size_t m, size_t max_len, double idf_weighting, int device);
/* The same with product attributes (item_to_product_attributes, vmis_index.rs:33,514-518): attr_items[i] gets the
 * flags attr_flags[i] (VMIS_ATTR_FOR_SALE | VMIS_ATTR_ADULT; 0 = "no attributes", which fails the business rules of
 * mod.rs:162-182).  Items not listed keep the CSV default {adult: false, for_sale: true}; attr_items == NULL is
 * vmis_index_from_sessions.  (The dense order of the items is internal to the index, so the attributes travel as
 * (external id, flags) pairs rather than as one positional array.) */
vmis_index_t* vmis_index_from_sessions_attrs(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                             size_t n_sessions, size_t m, size_t max_len, double idf_weighting,
                                             const uint64_t* attr_items, const uint8_t* attr_flags, size_t n_attrs,
                                             int device);

/* ---- item-sharded postings (BASELINE.json config 5; no counterpart in the reference, which replicates) ----
 * The item→sessions index is partitioned over the GPUs of one box: shard s of n holds the posting lists of the
 * items whose dense index % n == s; the session→items lists, idf and attributes are replicated.  Every process
 * builds its shard, exports it (CUDA IPC handle, 64 bytes), and attaches the shards of its peers; the predict
 * kernel then streams remote posting lists over NVLink inside the same launch (TMA bulk copies / loads on the
 * peer-mapped pointers), so results are bit-identical to the single-GPU index and there is no separate
 * exchange step.  Queries may be sharded over the ranks in any way. */
vmis_index_t* vmis_index_from_sessions_sharded(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                               size_t n_sessions, size_t m, size_t max_len, double idf_weighting,
                                               int device, uint32_t shard, uint32_t n_shards);
int vmis_index_export_shard(const vmis_index_t* index, void* handle64);
int vmis_index_attach_shard(vmis_index_t* index, uint32_t shard, const void* handle64);
/* same-process variant (several shards on GPUs of one process, or on one GPU in tests) */
int vmis_index_attach_shard_ptr(vmis_index_t* index, uint32_t shard, const void* device_ptr);
const void* vmis_index_shard_ptr(const vmis_index_t* index);

/* ---- on-device index build (prepare_hashmap, vmis_index.rs:422-528, as radix sorts + kernels) ------------
 * Training sessions already resident in HBM (e.g. produced by another GPU job); max_len must be explicit.
 * The resulting handle has no host mirror of the sessions: vmis_items_for_session / vmis_session_timestamp fail,
 * every other entry point works.  Results are bit-identical to the host-built index. */
vmis_index_t* vmis_index_from_device_sessions(const uint64_t* d_items, const uint64_t* d_sess_off, const uint32_t* d_sess_ts,
                                              size_t n_sessions, size_t m, size_t max_len, double idf_weighting, int device,
                                              uint32_t shard, uint32_t n_shards);
/* vmis_synth_sessions generated and indexed entirely on the device (BASELINE configs 4-5). */
vmis_index_t* vmis_index_synth(uint64_t seed, uint64_t n_items, uint64_t n_sessions, size_t m, size_t max_len,
                               double idf_weighting, int device, uint32_t shard, uint32_t n_shards);

/* VMISIndex::new(base_path) (vmis_index.rs:85-313): loads the production on-disk index — Avro object container
 * files under <base_path>/itemindex/ ({ItemId, session_indices_time_ordered, idf, ForSale, IsAdult}, :184-192) and
 * <base_path>/sessionindex/ ({SessionIndex, item_ids_asc, Time}, :249-255); codecs null / deflate / snappy.
 * Posting lists, idf values and attributes are taken as stored (nothing is recomputed from the sessions).
 * Lists are normalised to the canonical (timestamp desc, session idx desc) order; data on which the reference
 * would panic at query time (a posting naming a missing session, a session item without an itemindex record)
 * fails the load.  VMIS_DEVICE_NONE gives a host-only handle (accessors only). */
vmis_index_t* vmis_index_from_avro(const char* base_path, int device);
vmis_index_t* vmis_index_from_avro_sharded(const char* base_path, int device, uint32_t shard, uint32_t n_shards);
/* max_session_len > 0 additionally drops the sessions with more items from every posting list — what prepare_hashmap
 * does for a CSV index at the 99.5th length percentile (vmis_index.rs:67,452).  The offline Avro index keeps every
 * session (the reference's production statistics name sessions of 9408 events, vmis_index.rs:116-126); one such
 * session makes the worst-case score table of a query k x 9408 entries, so a serving index should be loaded with a
 * bound (0 = keep everything; queries then fail with VMIS_ERR_LIMIT and a message naming this option once
 * k x longest session exceeds the kernel's workspace bound). */
vmis_index_t* vmis_index_from_avro_ex(const char* base_path, int device, uint32_t shard, uint32_t n_shards,
                                      size_t max_session_len);
/* The same from arrays in memory: item i has ids item_ids[i], posting list post_sessions[post_off[i] ..
 * post_off[i+1]) (session indices), idf[i] and attr[i] (VMIS_ATTR_* bits; NULL = for sale, not adult); sessions
 * are dense by session index like session_to_items_sorted / session_to_max_time_stamp (:28-35). */
vmis_index_t* vmis_index_from_parts(const uint64_t* item_ids, const uint64_t* post_off, const uint32_t* post_sessions,
                                    const double* idf, const uint8_t* attr_or_null, size_t n_items, const uint64_t* items,
                                    const uint64_t* sess_off, const uint32_t* sess_ts, size_t n_sessions, int device,
                                    uint32_t shard, uint32_t n_shards);
vmis_index_t* vmis_index_from_parts_ex(const uint64_t* item_ids, const uint64_t* post_off, const uint32_t* post_sessions,
                                       const double* idf, const uint8_t* attr_or_null, size_t n_items, const uint64_t* items,
                                       const uint64_t* sess_off, const uint32_t* sess_ts, size_t n_sessions, int device,
                                       uint32_t shard, uint32_t n_shards, size_t max_session_len);
int vmis_index_prebuilt_info(const vmis_index_t* index, vmis_prebuilt_info_t* out);
/* The reverse direction: writes the index in that on-disk format (the reference computes it offline with Spark), so an
 * index built on the GPU can be served by the reference: <base_path>/itemindex/part-NNNNN.avro and
 * <base_path>/sessionindex/part-NNNNN.avro, codec "null" or "deflate" (NULL = deflate), n_files parts each.  Needs a
 * handle that still has its host mirror of the sessions (built from a TSV, session arrays or Avro). */
int vmis_index_to_avro(const vmis_index_t* index, const char* base_path, const char* codec, uint32_t n_files);

/* ---- serialised index blob (fast restart; the reference rebuilds from CSV or re-reads Avro at start-up,
 * serving.rs:37-52).  The blob holds the flat HBM arrays of one handle (one shard).  A loaded handle has no host
 * mirror of the sessions (vmis_items_for_session / vmis_session_timestamp fail); everything else works. */
int vmis_index_save(const vmis_index_t* index, const char* path);
vmis_index_t* vmis_index_load(const char* path, int device);

/* item_to_product_attributes (vmis_index.rs:34; Avro fields ForSale/IsAdult :184-192).  Replaces the
 * attributes of the listed items (flags = VMIS_ATTR_* bits; 0 removes the entry).  Call before serving. */
int vmis_index_set_attributes(vmis_index_t* index, const uint64_t* items, const uint8_t* flags, size_t n);

void vmis_index_free(vmis_index_t* index);

int vmis_index_stats(const vmis_index_t* index, vmis_stats_t* out);

/* ---- the hot path ------------------------------------------------------- */

/* predict(index, evolving_session, k, m, how_many, enable_business_logic) — mod.rs:118-215, batched.
 * Host buffers.  out_ids/out_scores are n_q × how_many, row q holds out_counts[q] recommendations in
 * `into_sorted_vec()` order (score descending, mod.rs:339-355; ties: item id ascending).  An empty evolving
 * session yields count 0 (the reference panics, mod.rs:157).  stream: cudaStream_t, or NULL for a
 * pooled internal stream.  The call returns after the results are in the host buffers. */
int vmis_predict_batch(const vmis_index_t* index, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                       uint32_t k, uint32_t m, uint32_t how_many, int enable_business_logic,
                       uint64_t* out_ids, double* out_scores, uint32_t* out_counts, void* stream);

/* Same computation with every buffer already resident on the index's device (no copies, no sync;
 * work is enqueued on `stream`; NULL = the CUDA default stream).  out_stats may be NULL.  The query lengths are on
 * the device, so the host cannot reject an over-long evolving session here: such a query comes back with
 * out_counts[q] == VMIS_COUNT_TOO_LONG and an empty row (vmis_predict_batch returns VMIS_ERR_LIMIT instead). */
int vmis_predict_batch_device(const vmis_index_t* index, const uint64_t* d_q_items, const uint32_t* d_q_off,
                              uint32_t n_q, uint32_t k, uint32_t m, uint32_t how_many, int enable_business_logic,
                              uint64_t* d_out_ids, double* d_out_scores, uint32_t* d_out_counts,
                              vmis_query_stats_t* d_out_stats, void* stream);

/* Single query convenience wrapper over vmis_predict_batch (the exact shape of mod.rs:118-125).
 * Returns the number of recommendations (>= 0) or a negative error. */
int vmis_predict(const vmis_index_t* index, const uint64_t* evolving_session, size_t len, size_t k, size_t m,
                 size_t how_many, int enable_business_logic, uint64_t* out_ids, double* out_scores);

/* SimilarityComputationNew::find_neighbors(evolving_session, k, m) — similarity_indexed.rs:13-22,
 * vmis_index.rs:325-415, batched.  out_sess/out_sim are n_q × k; row q holds out_counts[q] neighbours ordered
 * (similarity desc, timestamp desc, session id desc); ids are the reference's session indices. */
int vmis_find_neighbors_batch(const vmis_index_t* index, const uint64_t* q_items, const uint32_t* q_off,
                              uint32_t n_q, uint32_t k, uint32_t m, uint32_t* out_sess, double* out_sim,
                              uint32_t* out_counts, void* stream);

/* ---- trait accessors (host mirror; similarity_indexed.rs:8-24) ---------- */

/* items_for_session(&u32) -> &[u64] — vmis_index.rs:317-319.  Returns a pointer owned by the index
 * (valid until vmis_index_free) and writes the length; NULL if the session index is out of range. */
const uint64_t* vmis_items_for_session(const vmis_index_t* index, uint32_t session, size_t* len);

/* idf(&u64) -> f64 — vmis_index.rs:321-323.  0 on success; VMIS_ERR_ARG for an unknown item (reference panics). */
int vmis_idf(const vmis_index_t* index, uint64_t item, double* out);

/* find_attributes(&u64) -> Option<&ProductAttributes> — vmis_index.rs:417-419.
 * Returns VMIS_ATTR_* bits; 0 = None. */
int vmis_find_attributes(const vmis_index_t* index, uint64_t item);

/* item_to_top_sessions_ordered[item] (vmis_index.rs:29): the item's posting list, most recent first,
 * truncated to m, as reference session indices.  Returns the list length and copies up to cap entries.
 * Only host-only handles (VMIS_DEVICE_NONE) keep the postings in host memory. */
size_t vmis_postings(const vmis_index_t* index, uint64_t item, uint32_t* out, size_t cap);

/* session_to_max_time_stamp[session] — vmis_index.rs:30.  0 on success. */
int vmis_session_timestamp(const vmis_index_t* index, uint32_t session, uint32_t* out);

/* ---- synthetic workloads (BASELINE.json configs 2-5; no counterpart in the reference) ---------------- */

/* Deterministic generator: log-uniform (Zipf s=1) item popularity over n_items sparse u64 ids, session
 * lengths from the reference's empirical percentiles (vmis_index.rs:116-126, cap 34), items distinct
 * within a session, one unique u32 timestamp per session.  Call with items == NULL to get the total
 * number of interactions in *n_interactions; then call again with buffers of that size
 * (sess_off has n_sessions + 1 entries). */
int vmis_synth_sessions(uint64_t seed, uint64_t n_items, uint64_t n_sessions, uint64_t* items, uint64_t* sess_off,
                        uint32_t* sess_ts, uint64_t* n_interactions);

/* Evolving-session queries: a held-out session from the same generator, a random prefix of it, its last
 * max_items_in_session items (evaluator.rs:46-57).  q_items must hold n_q × max_items_in_session ids. */
int vmis_synth_queries(uint64_t seed, uint64_t n_items, uint32_t n_q, uint32_t max_items_in_session,
                       uint64_t* q_items, uint32_t* q_off);

/* ---- online call shape: micro-batching (recommend_resource.rs:56 called from actix workers, serving.rs:62-94) ----
 * Worker threads call vmis_batcher_predict() with ONE evolving session each (the exact shape of mod.rs:118-125) and
 * block; a dispatcher thread turns the waiting requests into one vmis_predict_batch call as soon as max_batch are
 * queued or the oldest has waited max_wait_us.  Returns the number of recommendations (>= 0) or a negative error. */
typedef struct vmis_batcher vmis_batcher_t;
vmis_batcher_t* vmis_batcher_create(const vmis_index_t* index, uint32_t k, uint32_t m, uint32_t how_many,
                                    int enable_business_logic, uint32_t max_batch, uint32_t max_wait_us);
int vmis_batcher_predict(vmis_batcher_t* batcher, const uint64_t* evolving_session, size_t len, uint64_t* out_ids,
                         double* out_scores);
int vmis_batcher_stats(vmis_batcher_t* batcher, uint64_t* n_batches, uint64_t* n_requests);
void vmis_batcher_destroy(vmis_batcher_t* batcher);
/* Open-loop load generator over the batcher (measurement aid for the online call shape): n_threads caller threads
 * replay the evolving sessions of a CSR batch at target_rps for duration_ms; request j is due at t0 + j/target_rps
 * and its latency (microseconds, into lat_us[0..cap)) counts from that due time.  Returns the number of completed
 * requests or a negative VMIS_ERR_*. */
long long vmis_batcher_load_test(vmis_batcher_t* batcher, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                                 uint32_t n_threads, double target_rps, uint32_t duration_ms, float* lat_us, size_t cap,
                                 double* achieved_rps);
/* Closed-loop lone caller (measurement aid for the reference's own call shape, mod.rs:118-125): n_calls vmis_predict()
 * calls from the calling thread, one evolving session of the CSR batch per call; per-call latencies (microseconds) into
 * lat_us[0..n_calls).  Returns n_calls or a negative VMIS_ERR_*. */
long long vmis_predict_latency_test(const vmis_index_t* index, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                                    uint32_t k, uint32_t m, uint32_t how_many, int enable_business_logic, uint32_t n_calls,
                                    float* lat_us);

/* ---- misc ---------------------------------------------------------------- */

/* ---- serving shell: GET /v1/recommend minus HTTP (recommend_resource.rs:20-65) ------------------------------
 * Evolving-session window over an in-process store with the semantics of RocksDBSessionStore (sessions/mod.rs:
 * key md5(session_id) as u128, sessions idle for more than max_session_idle_secs (default 20 min) start over,
 * entries older than session_ttl_secs (default 30 min, serving.rs:55-56) are dropped), then predict through a
 * micro-batcher.  Thread safe; call vmis_server_recommend from every worker thread. */
typedef struct vmis_server vmis_server_t;
vmis_server_t* vmis_server_create(const vmis_index_t* index, uint32_t k, uint32_t m, uint32_t how_many,
                                  uint32_t max_items_in_session, int enable_business_logic, uint32_t max_batch,
                                  uint32_t max_wait_us, uint64_t session_ttl_secs, uint64_t max_session_idle_secs);
/* v1_recommend: updates the session (user_consent != 0) or uses [item_id] alone, predicts, writes up to how_many
 * item ids best first (scores optional).  Returns the count or a negative VMIS_ERR_*. */
int vmis_server_recommend(vmis_server_t* server, const char* session_id, uint64_t item_id, int user_consent,
                          uint64_t* out_ids, double* out_scores_or_null);
/* The window step alone (recommend_resource.rs:39-54): returns the evolving session that would be predicted on. */
int vmis_server_session_window(vmis_server_t* server, const char* session_id, uint64_t item_id, int user_consent,
                               uint64_t* out_items, size_t cap);
/* get_session_items (sessions/mod.rs:37-57) without modifying the store. */
int vmis_server_stored_items(vmis_server_t* server, const char* session_id, uint64_t* out_items, size_t cap);
/* Tests: pin the store's clock to epoch_secs (0 = system clock again) and sweep expired entries. */
int vmis_server_set_clock(vmis_server_t* server, uint64_t epoch_secs);
int vmis_server_stats(vmis_server_t* server, uint64_t* n_sessions, uint64_t* n_batches, uint64_t* n_requests);
void vmis_server_destroy(vmis_server_t* server);
/* md5 digest used for the session key (recommend_resource.rs:27). */
void vmis_md5(const void* data, size_t len, uint8_t out16[16]);

const char* vmis_last_error(void);   /* message of the last failure on this thread */
int vmis_last_error_code(void);      /* its VMIS_ERR_* code (constructors return NULL on failure) */
const char* vmis_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VMIS_H_ */
