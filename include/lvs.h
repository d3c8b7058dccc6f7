/* lattice-b200 vector store: C ABI of the B200-native collection backend (liblattice_b200.so).
 *
 * This is the drop-in boundary for the semantic-search hot path of lattice (iAmLakshya/code-rag).  The
 * reference has no FFI of its own: the path sits behind the duck-typed Python class
 * `lattice.embeddings.client.QdrantManager` (reference src/lattice/embeddings/client.py:18-228, protocol
 * `VectorStore` in src/lattice/core/protocols.py:34-52), whose methods forward to qdrant-client.  Each entry
 * point below names the reference call it replaces; `code_rag_b200/client.py` binds them with ctypes and
 * INTEGRATION.md shows the stub a lattice maintainer would add.
 *
 * Conventions: plain C, no exceptions.  Every function returns LVS_OK (0) or a negative LVS_E* code;
 * lvs_last_error() gives a thread-local message.  The caller owns every host buffer; the library owns device
 * memory.  One process drives ONE GPU (one process per GPU under torchrun); a collection handle may be used from
 * any thread (calls on one handle are serialised internally).  Calls are synchronous on return.  "Rows" are GLOBAL
 * row numbers everywhere in this API: a shard created with row_base B owns rows B, B+1, ... (dense, chosen by the
 * caller; the Python adapter appends).
 */
#ifndef LATTICE_B200_LVS_H
#define LATTICE_B200_LVS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVS_OK 0
#define LVS_EINVAL (-1)     /* bad argument */
#define LVS_ECUDA (-2)      /* CUDA runtime error (message has the cudaError string) */
#define LVS_ENOMEM (-3)     /* device or host allocation failed */
#define LVS_ESTATE (-4)     /* library not initialised / no CUDA device */
#define LVS_ELIMIT (-5)     /* k, dim or batch beyond what the kernels support */
#define LVS_ENAN (-6)       /* a host query contains NaN (qdrant local mode refuses it too) */

#define LVS_STORAGE_F32 0
#define LVS_STORAGE_BF16 1
#define LVS_METRIC_COSINE 0 /* models.Distance.COSINE, client.py:99 */
#define LVS_METRIC_DOT 1
#define LVS_DT_F32 0
#define LVS_DT_F64 1
#define LVS_DT_BF16 2
#define LVS_MAX_FILTER_COLS 8
#define LVS_ANY 0xFFFFFFFFu        /* filter column not constrained */
#define LVS_NULL_CODE 0u           /* payload key missing or None: matches no MatchValue */
#define LVS_NO_MATCH 0xFFFFFFFEu   /* filter value never seen in this column: matches nothing */
#define LVS_MAX_K 224              /* largest `limit` one search call serves */
#define LVS_FLAG_UNPROVEN 1        /* out_flags bit: exactness bound not met even at the largest candidate set */
#define LVS_FLAG_EXCHANGE 2        /* out_flags bit (sharded searches): a peer's lists did not arrive in time; the result is not trustworthy */

typedef struct lvs_collection lvs_collection;

/* ---- library ---------------------------------------------------------------------------------------- */
int lvs_version(void);
const char* lvs_last_error(void);
/* Bind the process to one CUDA device.  Replaces QdrantManager.connect (client.py:32-45). Fails with LVS_ESTATE
 * when no sm_100 device is present: there is no CPU fallback. */
int lvs_init(int device);
int lvs_shutdown(void);                                   /* QdrantManager.close, client.py:47-55 */
int lvs_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem, int64_t* free_mem);

/* ---- collections (client.py:72-113 create_collections/_create_collection_with_indexes) ----------------- */
int lvs_collection_create(const char* name, int dim, int storage, int metric, int n_filter_cols,
                          int64_t capacity_rows, int64_t row_base, lvs_collection** out);
int lvs_collection_destroy(lvs_collection* c);            /* client.delete_collection, client.py:215 */
int lvs_collection_reserve(lvs_collection* c, int64_t capacity_rows);
int64_t lvs_rows(const lvs_collection* c);                /* high-water mark (rows ever written) */
int64_t lvs_count(const lvs_collection* c);               /* live rows: CollectionInfo.points_count, client.py:204-210 */
int64_t lvs_capacity(const lvs_collection* c);
uint64_t lvs_search_counter(const lvs_collection* c);     /* vector searches served so far (drives the replay; 64-bit, never wraps) */
/* Account n reference searches that this shard did not execute itself (a restored or re-balanced shard catching up with the
 * collection's history; tests of the counter's 2^32 boundary).  Local mode re-normalises every stored row on every search
 * (oracle/qdrant_local.py point 2), so the number of searches since a row was written is part of the row's state. */
int lvs_advance_search_counter(lvs_collection* c, uint64_t n);

/* ---- upsert (QdrantManager.upsert, client.py:115-130; K4 kernel) ----------------------------------------
 * vecs: n x dim row-major host array of `dtype`.  rows: n global row numbers (the shard grows as needed), or NULL to
 * append at lvs_rows().  codes: n x n_filter_cols dictionary codes (row-major) or NULL.  ties: n tie-break keys
 * (order among exactly equal scores is (tie asc, row asc)) or NULL => global row number. */
int lvs_upsert(lvs_collection* c, const void* vecs, int dtype, int64_t n, const int64_t* rows,
               const uint32_t* codes, const uint64_t* ties);
/* Same with DEVICE pointers (vectors produced on the GPU never visit the host); rows are row0..row0+n-1.
 * d_codes is n x n_filter_cols on the device or NULL.  stream: cudaStream_t or NULL for the collection's stream. */
int lvs_upsert_device(lvs_collection* c, const void* d_vecs, int dtype, int64_t n, int64_t row0,
                      const uint32_t* d_codes, const uint64_t* d_ties, void* stream);
/* (Re)write one filter column for n rows (rows NULL => row0..). Lets the adapter index a payload key lazily. */
int lvs_set_codes(lvs_collection* c, int col, const int64_t* rows, int64_t row0, int64_t n, const uint32_t* codes);

/* ---- delete (QdrantManager.delete, client.py:159-169) -------------------------------------------------- */
int lvs_delete_rows(lvs_collection* c, const int64_t* rows, int64_t n, int64_t* n_deleted);
/* Tombstone every live row matching `want` (n_filter_cols codes, LVS_ANY = unconstrained); the matching GLOBAL rows
 * are returned (unordered, at most cap) so the host can drop their payloads; *n_matched is the full count. */
int lvs_delete_where(lvs_collection* c, const uint32_t* want, int64_t* out_rows, int64_t cap, int64_t* n_matched);

/* ---- search (QdrantManager.search -> query_points, client.py:132-157; K1/K2 + exact rescoring) -----------
 * queries: Q x dim host array of `dtype` (LVS_DT_F32 / LVS_DT_F64).  want: n_filter_cols codes or NULL (client.py:171-176
 * _build_filter: conjunction of exact matches).  Outputs, Q x k each: scores (float64, local-mode arithmetic), GLOBAL
 * rows (-1 padded), tie keys; counts[Q] valid results; flags[Q] (LVS_FLAG_UNPROVEN).  Results are ordered
 * (score desc, tie asc, row asc).  Q searches are accounted as Q consecutive reference searches. */
int lvs_search(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want,
               double* out_scores, int64_t* out_rows, uint64_t* out_ties, uint32_t* out_counts, int32_t* out_flags);
/* Pipelined form of lvs_search for throughput: submit copies the queries (pinned staging, async H2D), enqueues the search
 * and the D2H of the result and returns a ticket; wait blocks until that search has finished and fills the outputs (same
 * meaning as lvs_search, including the repeat of flagged queries).  Up to 4 searches may be in flight per collection; they
 * execute in submission order and are accounted as consecutive reference searches. */
int lvs_search_submit(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want, int* ticket);
int lvs_search_wait(lvs_collection* c, int ticket, double* out_scores, int64_t* out_rows, uint64_t* out_ties,
                    uint32_t* out_counts, int32_t* out_flags);
/* Non-blocking: *done = 1 once the device work of `ticket` has finished (lvs_search_wait will then return without waiting, unless it
 * has flagged queries to repeat).  For event loops (the asyncio adapter polls between yields instead of parking a thread). */
int lvs_search_poll(lvs_collection* c, int ticket, int* done);
/* A search of up to 4 queries on the scan path is ONE kernel that stores its result block and then a completion word into the
 * slot's mapped pinned memory; lvs_search_wait polls that word (no event between consecutive kernels, so searches submitted back
 * to back keep overlapping on the GPU).  Batches and timing mode complete through a CUDA event as before. */
/* Device-pointer form used by the sharded path: queries and the four Q x k / Q outputs are DEVICE buffers, the work is
 * enqueued on `stream` (cudaStream_t, NULL = collection stream) and completed before return; flags are host. */
int lvs_search_device(lvs_collection* c, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                      double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                      int32_t* out_flags, void* stream);
/* Repeat of queries that an enqueue-only search left flagged: same meaning as lvs_search_device, but always on the exact scan
 * (never the tensor-core path) and accounted as the reference searches number search_no .. search_no + Q - 1 that it repeats
 * (1-based, <= lvs_search_counter), so that the scores are the ones the first attempt would have returned. */
int lvs_search_device_at(lvs_collection* c, uint64_t search_no, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                         double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                         int32_t* out_flags, void* stream);
/* Enqueue-only form for pipelined callers: nothing is synchronised, flags are written to the DEVICE buffer d_out_flags
 * and a flagged query is NOT repeated with a larger candidate set (the caller may re-issue it synchronously).
 * Consecutive searches enqueued back to back on one stream overlap (programmatic dependent launch: the next search streams
 * the shard while this one's last CTAs rescore and exchange; option "pdl", on by default, off while "timing" is on).  The
 * kernel reads the RAW query buffer before it waits for the previous kernel of the stream: d_queries must be produced by a
 * copy, by work on another stream ordered with an event, or by a kernel that does not signal programmatic launch completion
 * early - every ordinary producer qualifies; otherwise set "pdl" to 0. */
int lvs_search_device_async(lvs_collection* c, const void* d_queries, int dtype, int Q, int k, const uint32_t* want,
                            double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                            int32_t* d_out_flags, void* stream);
/* Filter-only lookup (search with query_vector=None, scroll, count: client.py:178-202, query/context/builder.py:111-119,
 * projects/cleanup.py:41-61): all live rows matching `want` (GLOBAL rows, unordered, at most cap), *n_matched = full count. */
int lvs_match_rows(lvs_collection* c, const uint32_t* want, int64_t* out_rows, int64_t cap, int64_t* n_matched);

/* ---- K5: merge of per-shard top-k lists after the all-gather ---------------------------------------------------
 * d_scores/d_rows/d_ties: G blocks of Q x k, `shard_stride` 8-byte elements apart (0 => dense, Q*k), device buffers. */
int lvs_merge_topk_device(const double* d_scores, const int64_t* d_rows, const uint64_t* d_ties, int64_t shard_stride,
                          int G, int Q, int k, double* d_out_scores, int64_t* d_out_rows, uint64_t* d_out_ties, uint32_t* d_out_counts,
                          void* stream);   /* enqueue only: ordered on `stream` */

/* ---- K5': exchange + merge over NVLink peer memory (replaces the NCCL all-gather + lvs_merge_topk_device) ----------------------
 * One lvs_exchange per process.  create allocates this rank's gather buffer and returns its 64-byte CUDA IPC handle; the
 * caller all-gathers the handles (any transport) and passes the world x 64 bytes to connect, which maps the peers' buffers.
 * Every rank must issue the same sequence of exchanges (lvs_exchange_merge_device and lvs_search_sharded_device_async calls).
 * lvs_exchange_merge_device: d_local = this rank's packed [3][Q][k] int64 block (float64 score bits | global rows | tie keys),
 * d_local_flags = its [Q] flags (or NULL); the kernel stores them into every rank's buffer with P2P stores, raises a system-scope
 * flag, waits for all ranks' flags and merges into d_out ([3][Q][k]) / d_out_counts ([Q]) / d_out_flags ([Q]: OR of every rank's
 * flags, | LVS_FLAG_EXCHANGE after a timeout; may be NULL).  Enqueue only (ordered on `stream`). */
typedef struct lvs_exchange lvs_exchange;
int lvs_exchange_create(int world, int rank, int max_q, int max_k, lvs_exchange** out, void* ipc_handle_out);
int lvs_exchange_connect(lvs_exchange* ex, const void* all_handles);
int lvs_exchange_merge_device(lvs_exchange* ex, const int64_t* d_local, const int32_t* d_local_flags, int Q, int k,
                              int64_t* d_out, uint32_t* d_out_counts, int32_t* d_out_flags, void* stream);
/* The sharded search as one call per rank: this shard's search + the exchange + the merge, enqueued on `stream`.
 * Every rank calls it with the same queries, k and filter; every rank gets the same merged d_out ([3][Q][k] int64 as above),
 * counts and flags (OR over the shards, so a query that one shard could not prove exact is flagged everywhere and the caller can
 * repeat it collectively).  Up to 4 queries on the scan path are ONE kernel per rank (scan + exact rescoring + exchange + merge);
 * batches on the tensor-core path add the exchange kernel behind the finalize kernel.  Enqueue only, no repeat of flagged queries. */
int lvs_search_sharded_device_async(lvs_collection* c, lvs_exchange* ex, const void* d_queries, int dtype, int Q, int k,
                                    const uint32_t* want, int64_t* d_out, uint32_t* d_out_counts, int32_t* d_out_flags, void* stream);
/* Host-buffer form of the sharded search, pipelined like lvs_search_submit: every rank submits the same queries; lvs_search_wait on
 * the ticket returns the MERGED lists and the merged flags (OR over the shards).  Flagged queries are not repeated here: every rank
 * sees the same flags and repeats them collectively (lvs_search_device_at + lvs_exchange_merge_device). */
int lvs_search_submit_sharded(lvs_collection* c, lvs_exchange* ex, const void* queries, int dtype, int Q, int k, const uint32_t* want,
                              int* ticket);
int lvs_exchange_error(lvs_exchange* ex);     /* 1 if a peer's flag ever timed out (synchronises) */
int lvs_exchange_destroy(lvs_exchange* ex);

/* ---- K3: fused hybrid ranking (HybridRanker.rank_results query/ranking/ranker.py:18-226 + ResultScorer scorer.py:9-126
 *      [mode 0]; ResultReranker.fuse_results / deduplicate / normalize_scores query/reranker.py:29-145 [mode 1]) -------------
 * Candidates of all queries are concatenated in the reference's insertion order (primary, callers, callees, methods,
 * parent_classes, child_classes, vector hits); offsets[q]..offsets[q+1] are query q's.  Strings are interned by the caller:
 * key_id / file_id are dense ids of f"{file_path}:{entity_name}:{start_line}" and of file_path.  All host pointers. */
typedef struct lvs_rank_batch {
    int32_t n_queries;
    const int32_t* offsets;        /* [n_queries + 1] */
    const uint8_t* kind;           /* 0 primary, 1 caller, 2 callee, 3 method/parent/child, 4 vector hit */
    const uint32_t* key_id;
    const uint32_t* file_id;
    const int32_t* depth;          /* callers/callees: metadata depth (0 is treated as 1, scorer.py:25) */
    const double* entity_match;    /* 1.0 exact name match, 0.5 substring, 0.0 none (scorer.py:31-35) */
    const int32_t* degree;         /* total_degree of the entity in the centrality dict, -1 if absent */
    const uint8_t* flags;          /* bit0 summary, bit1 docstring, bit2 signature, bit3 content present */
    const int32_t* content_len;    /* vector hits: len(content), -1 if none */
    const double* vscore;          /* vector hits: similarity score */
    const double* weights;         /* [n_queries][4]: graph, vector, centrality, context weight (models.py:59-91) */
} lvs_rank_batch;
/* Outputs (host, per query max_total rows): leader candidate index within the query, merged score, min-max normalised score,
 * 7 merged signals (RankingSignal order: graph_match, vector_similarity, centrality, query_entity_match,
 * relationship_relevance, code_quality, context_richness) + presence mask, source (0 graph, 1 vector, 2 hybrid);
 * out_leader[total candidates] = leader of every candidate (lets the host fill missing text fields in merge order). */
int lvs_rank_fuse(const lvs_rank_batch* in, int mode, int max_per_file, int max_total, double entity_bonus, double rel_bonus,
                  int32_t* out_count, int32_t* out_index, double* out_score, double* out_norm, double* out_signals,
                  uint8_t* out_sigmask, uint8_t* out_source, int32_t* out_leader, float* device_ms);

/* ---- fused search -> rank (the glue between QueryEngine._execute_vector_search query/engine.py:315-346
 *      and HybridRanker._process_vector_results ranking/ranker.py:150-169) ---------------------------------------------
 * Per-row ranking attributes are written once, at upsert: the interned ids of f"{file_path}:{entity_name}:{start_line}",
 * of file_path and of (graph_node_id or entity_name); the id of the lower-cased entity name in the collection's name
 * pool; len(content) (-1: none) and the presence flags (bit0 summary, bit3 content).  lvs_search_rank then runs the
 * search, builds the vector-hit candidates from those columns on the device (entity-name match scorer.py:91-96 included)
 * and ranks them together with the caller's graph candidates in ONE call: no host hop between top-k and blend. */
int lvs_rank_names_append(lvs_collection* c, const uint8_t* bytes, const uint32_t* lens, int n, uint32_t* first_id);
int lvs_rank_attrs_set(lvs_collection* c, const int64_t* rows, int n, const uint32_t* key_id, const uint32_t* file_id,
                       const uint32_t* cent_id, const uint32_t* name_id, const int32_t* content_len, const uint8_t* flags);
typedef struct lvs_rank_query_ctx {
    const int32_t* ent_off;        /* [Q + 1]: query q's entities are ent_off[q] .. ent_off[q+1] */
    const uint32_t* ent_str_off;   /* [n_entities + 1] byte offsets of the lower-cased UTF-8 entity names */
    const uint8_t* ent_bytes;
    const int32_t* cen_off;        /* [Q + 1]: query q's centrality entries */
    const uint32_t* cen_id;        /* interned (graph_node_id or entity_name) */
    const int32_t* cen_deg;        /* total_degree */
} lvs_rank_query_ctx;
/* `graph` holds ONLY the graph candidates (kinds 0..3; content_len / vscore ignored) and the weights; each query's vector hits
 * are appended behind its graph candidates, k slots per query: candidate index i >= n_graph(q) is hit slot i - n_graph(q).
 * out_leader has graph->offsets[Q] + Q*k entries in that combined order.  out_flags bit0: the search could not prove
 * exactness for the query with the default candidate set (repeat it through lvs_search + lvs_rank_fuse).
 * device_ms[2]: search, gather + rank. */
int lvs_search_rank(lvs_collection* c, const void* queries, int dtype, int Q, int k, const uint32_t* want,
                    const lvs_rank_batch* graph, const lvs_rank_query_ctx* ctx, int max_per_file, int max_total,
                    double entity_bonus, double rel_bonus, double* out_hit_scores, int64_t* out_hit_rows,
                    uint32_t* out_hit_counts, int32_t* out_flags, int32_t* out_count, int32_t* out_index, double* out_score,
                    double* out_signals, uint8_t* out_sigmask, uint8_t* out_source, int32_t* out_leader, float* device_ms);

/* Two collections in one call: QueryEngine._execute_vector_search (query/engine.py:331-344) appends, for five intents, the hits
 * of a `summaries` search (limit // 2) behind the code hits.  c2 is searched with the queries sel2[0..Q2) (increasing indices into the
 * batch, k2 hits each); a query's candidates are [its graph candidates][its hits in c][its hits in c2], compact; out_leader has
 * graph->offsets[Q] + Q*(k + k2) entries.  The interned key / file / centrality ids must be shared by the two collections.
 * c2 == NULL or Q2 == 0: exactly lvs_search_rank. */
typedef struct lvs_rank_hits { double* scores; int64_t* rows; uint32_t* counts; int32_t* flags; } lvs_rank_hits;   /* host outputs */
int lvs_search_rank2(lvs_collection* c, lvs_collection* c2, const void* queries, int dtype, int Q, int k, int k2,
                     const uint32_t* want, const uint32_t* want2, const int32_t* sel2, int Q2,
                     const lvs_rank_batch* graph, const lvs_rank_query_ctx* ctx, int max_per_file, int max_total,
                     double entity_bonus, double rel_bonus, const lvs_rank_hits* hits1, const lvs_rank_hits* hits2,
                     int32_t* out_count, int32_t* out_index, double* out_score, double* out_signals, uint8_t* out_sigmask,
                     uint8_t* out_source, int32_t* out_leader, float* device_ms);

/* ---- compaction: after a mass delete (projects/cleanup.py:38-73 removes a whole project) tombstones still
 *      cost scan bandwidth.  lvs_move_rows copies row src[i] (vector, tombstone, codes, tie key, write epoch, norm, ranking
 *      attributes) over row dst[i] and turns src[i] into a tombstone (GLOBAL rows; the two sets are disjoint); lvs_truncate then
 *      drops the trailing rows, which must all be tombstones.  Search results do not depend on where a row lives. */
int lvs_move_rows(lvs_collection* c, const int64_t* src, const int64_t* dst, int64_t n);
int lvs_truncate(lvs_collection* c, int64_t n_rows);

/* ---- snapshots: the shard's device arrays (vectors, tombstones, codes, tie keys, write epochs,
 *      norms, search counter, ranking attributes and name pool) to / from one file, so that an index survives a restart the
 *      way the Qdrant volume does (reference docker-compose.yml:42-43).  A loaded shard answers every search exactly as the
 *      saved one would have (the replay state is part of the snapshot).  Ids, payloads and dictionaries are the host's. */
int lvs_snapshot_save(lvs_collection* c, const char* path);
int lvs_snapshot_load(const char* path, const char* name, int64_t capacity_rows, lvs_collection** out);

/* ---- embedding on the GPUs that search -----------------------------------------------------------
 * The step in front of upsert: the reference embeds a chunk with UniXcoder - transformers' RobertaModel over the token ids with
 * bidirectional attention among the non-pad tokens, masked mean pooling (src/lattice/providers/unixcoder_provider.py:137-155) -
 * converts the vectors to python lists (:194-215) and passes them to QdrantManager.upsert (embeddings/indexer.py:77-86).  The
 * encoder below runs that forward pass on the device (tcgen05 GEMMs with fused bias / GELU / residual epilogues, attention on
 * mma.sync, LayerNorm and pooling kernels; bf16 activations, fp32 accumulation) and lvs_encoder_embed_upsert feeds the pooled
 * vectors straight into the shard's upsert kernel.  Tokenisation stays with the caller (the tokenizer is a host-side dictionary).
 * Parameters are loaded by their Hugging Face state-dict names ("embeddings.word_embeddings.weight",
 * "encoder.layer.3.attention.self.query.weight", ...; a leading "roberta." / "model." is ignored) as fp32 host arrays. */
typedef struct lvs_encoder lvs_encoder;
typedef struct lvs_encoder_config {
    int32_t vocab, hidden, n_layers, n_heads, intermediate, max_pos, pad_id;   /* RobertaConfig; hidden = 64 * n_heads */
    float ln_eps;
} lvs_encoder_config;
int lvs_encoder_create(const lvs_encoder_config* cfg, lvs_encoder** out);
int lvs_encoder_load(lvs_encoder* e, const char* name, const float* data, int64_t n);
/* ids: B x L token ids (host, pad_id-padded).  out: B x hidden float32 sentence embeddings (host). */
int lvs_encoder_embed(lvs_encoder* e, const int32_t* ids, int B, int L, float* out);
/* Embed and upsert in one call: the B vectors go from the pooling kernel to the upsert kernel without leaving HBM.  rows / codes /
 * ties as in lvs_upsert. */
int lvs_encoder_embed_upsert(lvs_encoder* e, lvs_collection* c, const int32_t* ids, int B, int L, const int64_t* rows,
                             const uint32_t* codes, const uint64_t* ties);
int lvs_encoder_last_ms(const lvs_encoder* e, float* ms);   /* device time of the last forward pass (CUDA events) */
int lvs_encoder_destroy(lvs_encoder* e);

/* ---- instrumentation --------------------------------------------------------------------------------- */
/* Device time (ms, CUDA events on the collection's stream) of the kernels of the last lvs_search* call on this handle:
 * [0] query prep  [1] scan / tensor-core kernel(s)  [2] finalize  [3] whole device section; n_launches = kernels launched. */
int lvs_last_search_timing(const lvs_collection* c, float* ms4, int* n_launches, int* kernel_kind);
/* Device time of the last (up to max_n, <= 256) scan-kernel launches, oldest first, from CUDA events recorded on the
 * launching stream, with the algorithmic bytes of each launch.  The stream must have been synchronised. */
int lvs_scan_times(lvs_collection* c, int max_n, float* out_ms, double* out_bytes, int* n);
/* Profiling aid (option "dbg_times" = 1): 16 globaltimer stamps (ns) of the phases of the last scan-kernel launch on this handle -
 * [0] first CTA starts, then the LATEST CTA to reach: [1] queries normalised, [2] shard scanned, [3] list written, [4] helpers see every
 * list, [5] candidates selected, [6] rescoring shares done, [7] result stored / merged; inside the selection: [8] list maxima and
 * list heads loaded, [9] threshold found, [10] keys gathered; after the rescoring: [11] the ordering CTA has every exact score,
 * [12] result ordered and proven, [13] completion word stored for the polling host;
 * inside the rescoring: [14] candidate rows staged, [15] re-normalisations replayed. */
int lvs_last_kernel_phases(lvs_collection* c, uint64_t* ns16);
/* Tunables: "stage_kb", "stages", "grid", "force_kpl" (0 = auto), "timing" (1 = record CUDA events around every scan launch for
 * lvs_last_search_timing / lvs_scan_times; default 0; it serialises consecutive searches), "pdl" (1 = programmatic dependent launch
 * of the scan kernel, default),
 * "gemm_min_q" (batch size from which the tensor-core path is used, default 3; fp32 shards at least 5), "path" (0 auto, 1 scan only,
 * 2 tensor-core whenever eligible), "gemm_stages", "gemm_stages_b" (CTA-pair form: corpus buffers of the split operand rings, default 5;
 * 0 = one ring of combined stages), "gemm_keep" (keys per K2 list, 4..16), "gemm_no_pair" (1 = never use the
 * cta_group::2 form), "gemm_no_tf32" (1 = fp32 shards stay on the scan), "gemm_no_unit" (1 = always scale by 1/||row||),
 * "gemm_dbg" (profiling switches, see GemmParams::dbg_mode).  Returns LVS_EINVAL for unknown names. */
int lvs_set_option(lvs_collection* c, const char* name, int value);
/* Copy rows back (debug / snapshots): out is n x dim float32 of the values a fresh reference collection would hold. */
int lvs_fetch_rows_f32(lvs_collection* c, const int64_t* rows, int64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif /* LATTICE_B200_LVS_H */
