/*
 * quiver_gpu.h — C ABI of libquivergpu.so, the B200 (sm_100a) implementation of
 * Quiver's exact-search hot path.
 *
 * This is the drop-in boundary: the functions below are what a cgo package
 * (pkg/gpu) binds so that a `gpu.Index` can satisfy
 * `core.Index` (reference pkg/core/collection.go:78-87) / `hybrid.Index`
 * (pkg/hybrid/types.go:196-214) and replace `hybrid.ExactIndex`
 * (pkg/hybrid/exact.go:14-133) behind `HybridIndex.searchWithStrategy`
 * (pkg/hybrid/hybrid_index.go:473-585).
 *
 * Conventions
 *   - plain C types only; every function returns 0 (QG_OK) or a qg_status code,
 *     and leaves a thread-local message readable through qg_last_error();
 *   - rows are dense int64 row indices in upload order; string IDs stay in the
 *     host language (the reference keeps them in Go maps, exact.go:16);
 *   - the caller owns every in/out buffer; the library owns device memory
 *     behind the opaque handles; no pointer is retained after a call returns;
 *   - qg_search_*, qg_batch_distance*, qg_filter_eval are re-entrant (they
 *     mirror the reference's shared RLock, exact.go:93); upload / tombstone /
 *     facet-column / destroy calls need external exclusion (the reference's
 *     exclusive Lock, exact.go:39,62);
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns QG_ERR_CUDA.
 */
#ifndef QUIVER_GPU_H
#define QUIVER_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_ABI_VERSION 1

typedef enum qg_status {
  QG_OK = 0,
  QG_ERR_INVALID = 1,   /* bad argument (null pointer, negative size, ...)          */
  QG_ERR_DIM = 2,       /* dimension mismatch  (exact.go:46,101)                    */
  QG_ERR_K = 3,         /* k <= 0              (exact.go:104-106)                   */
  QG_ERR_CUDA = 4,      /* CUDA runtime / driver failure, or no device              */
  QG_ERR_OOM = 5,       /* device or pinned-host allocation failed                  */
  QG_ERR_UNSUPPORTED = 6, /* predicate or option the device path does not implement */
  QG_ERR_RANGE = 7      /* row index out of range                                   */
} qg_status;

/* Distance semantics: reference pkg/vectortypes/distances.go (line per metric). */
typedef enum qg_metric {
  QG_COSINE = 0, /* distances.go:12-40  1 - dot/(|a||b|), zero norm => 1, sim clamped */
  QG_L2 = 1,     /* distances.go:43-55  sqrt(sum (a-b)^2), fp32 subtract, fp64 sum     */
  QG_DOT = 2,    /* distances.go:77-90  1 - dot                                        */
  QG_SQL2 = 3,   /* distances.go:60-72  sum (a-b)^2, sequential fp32                   */
  QG_L1 = 4      /* distances.go:93-104 sum |a-b|, fp32 subtract, fp64 sum             */
} qg_metric;

/* Which reference arithmetic the returned float32 distances reproduce. */
typedef enum qg_arith {
  QG_ARITH_VECTORTYPES = 0, /* float64 accumulators, pkg/vectortypes/distances.go      */
  QG_ARITH_HNSW_F32 = 1     /* sequential float32, pkg/hnsw/adapter.go:105-167
                               (cosine / l2 / dot only)                                */
} qg_arith;

typedef struct qg_config {
  int device;            /* CUDA device ordinal                                         */
  int arith;             /* qg_arith                                                    */
  int64_t reserve_rows;  /* capacity hint; 0 = grow on demand                           */
  int select_margin;     /* extra fp32-scan candidates re-ranked exactly; 0 = default   */
  int flags;             /* QG_FLAG_* bits, 0 = defaults                                */
} qg_config;
/* Do not keep the bf16 copy of the corpus (+50 % device memory for dim <= 512): batches of 8 and more then
 * stream the fp32 rows through the tensor cores as tf32, smaller ones take the flat fp32 scan. Results are
 * the same either way — every returned distance is recomputed from the fp32 rows. */
#define QG_FLAG_NO_BF16_COPY 1

typedef struct qg_index qg_index;
typedef struct qg_filter qg_filter;

/* ---- library ----------------------------------------------------------------------- */
int qg_abi_version(void);
const char* qg_last_error(void);
int qg_device_count(int* out_count);
/* Name and SM count of a device (buffer may be NULL). */
int qg_device_info(int device, char* name_buf, size_t name_len, int* out_sm_count,
                   int* out_cc_major, int* out_cc_minor);

/* ---- index lifecycle (replaces NewExactIndex / Insert / Delete / Size,
 *      exact.go:29-70,136-141) ------------------------------------------------------- */
int qg_index_create(qg_index** out, int dim, int metric, const qg_config* cfg /*nullable*/);
int qg_index_destroy(qg_index* idx);
/* Append n rows (row-major n x dim float32, host memory, pinned or pageable); rows are
 * copied (exact.go:53-54). *first_row receives the index of the first appended row. */
int qg_index_upload(qg_index* idx, const float* rows, int64_t n, int64_t* first_row /*nullable*/);
/* Same, source already on this index's device. */
int qg_index_upload_device(qg_index* idx, const void* d_rows, int64_t n, int64_t* first_row);
/* Append n synthetic rows generated on the device by the counter-based generator shared
 * with oracle/ (see oracle/synth.h): element (row, col) depends only on (kind, seed,
 * global_row0 + row, col). kind: 0 uniform[0,1), 1 SIFT-like integers 0..217,
 * 2 approx-normal (sum of 4 uniforms), 3 = kind 2 L2-normalised per row. */
int qg_index_upload_synthetic(qg_index* idx, int kind, uint64_t seed, int64_t global_row0,
                              int64_t n, int64_t* first_row);
/* Mark rows deleted (Delete, exact.go:61-70). Already-deleted rows are a no-op. */
int qg_index_tombstone(qg_index* idx, const int64_t* rows, int64_t n);
/* Squeeze the tombstoned rows out of every per-row array (vectors, bf16 copy, norms, facet
 * columns and their element lists): live rows keep their relative order and are renumbered
 * 0 .. size-1, the capacity shrinks to fit, qg_index_rows() == qg_index_size() afterwards.
 * The reference drops a deleted vector from its map at once (exact.go:61-70,
 * hybrid_index.go:244-290), so a Go index never scans dead entries; this call gives the device
 * index the same property after a burst of deletes. old_to_new (nullable) receives, for each of
 * the qg_index_rows() rows before the call, its new row or -1; the caller renumbers its
 * id <-> row tables with it. Works out of place (the live part of the index must fit a second
 * time; an error leaves the index untouched and the contents of old_to_new undefined). Needs
 * the same external exclusion as upload / tombstone; compiled filters stay valid and are
 * re-evaluated on their next use. */
int qg_index_compact(qg_index* idx, int64_t* old_to_new /*nullable*/, int64_t* out_rows /*nullable*/);
int64_t qg_index_size(const qg_index* idx);  /* live rows                    */
int64_t qg_index_rows(const qg_index* idx);  /* rows held (uploaded and not yet compacted away) */
int qg_index_dim(const qg_index* idx);
int qg_index_metric(const qg_index* idx);
/* Copy stored vectors back (IncludeVectors, collection.go:766-770). */
int qg_index_fetch(qg_index* idx, const int64_t* rows, int64_t n, float* out /* n x dim */);

/* ---- facet / metadata columns and predicates ---------------------------------------
 * One column per field. A row's value is described by:
 *   kind  : qg_value_kind
 *   num   : the float64 value when kind == QG_KIND_NUMBER
 *   scode : order-preserving dictionary code of the value's Go "%v" text
 *           (case-sensitive; core.valuesEqual / compareValues, collection.go:601-634);
 *           -1 when kind is MISSING
 *   fcode : dictionary code of the case-folded string (strings.EqualFold,
 *           facets.go:73-77) for STRING rows, else -1
 * Array / map values are QG_KIND_OTHER; the elements of arrays travel separately
 * (qg_facets_set_array_column) for the one predicate that looks inside them (QG_OP_ELEM_IN). */
typedef enum qg_value_kind {
  QG_KIND_MISSING = 0, /* field absent from the row's metadata / facets */
  QG_KIND_NULL = 1,    /* present, JSON null                            */
  QG_KIND_STRING = 2,
  QG_KIND_NUMBER = 3,
  QG_KIND_BOOL = 4,    /* num = 0/1                                     */
  QG_KIND_OTHER = 5,   /* array / map; flag bit 0x80 set = non-empty    */
  QG_KIND_NOROW = 6    /* the row has no metadata / facet entry at all  */
} qg_value_kind;

int qg_facets_set_column(qg_index* idx, int field, const uint8_t* kind, const double* num,
                         const int32_t* scode, const int32_t* fcode, int64_t n);

/* Elements of the array-valued rows of a column (facets.go:308-320: a SetFilter matches an array
 * facet when ANY element equals ANY filter value under valuesEqual). CSR layout: row r owns
 * elem_codes[offsets[r] .. offsets[r+1]); rows that are not arrays own nothing. The codes come from a
 * per-column dictionary kept by the host (equal codes <=> valuesEqual elements: numbers by float64
 * value, everything else by reflect.DeepEqual). Call after qg_facets_set_column for the same field;
 * n must equal that column's n. */
int qg_facets_set_array_column(qg_index* idx, int field, const int32_t* offsets /* n + 1 */,
                               const int32_t* elem_codes, int64_t n, int64_t n_elems);

/* A predicate is already lowered by the host-side compiler (quiver_b200/host) from the
 * reference's filter objects into this normal form; the device evaluates
 *   match(row) = AND_i pred_i(row)
 * with each pred_i = OR over its clauses. */
typedef enum qg_clause_op {
  QG_OP_FALSE = 0,
  QG_OP_TRUE = 1,
  QG_OP_KIND_IN = 2,      /* kind bitmask test: (1<<kind) & ia                          */
  QG_OP_NUM_EQ_TOL = 3,   /* kind NUMBER and |num - fa| <= fb    (collection.go:604)    */
  QG_OP_NUM_EQ = 4,       /* kind NUMBER/BOOL per mask ia and num == fa (facets.go:81)  */
  QG_OP_NUM_CMP = 5,      /* kind NUMBER and num (ia: 0 <,1 <=,2 >,3 >=) fa             */
  QG_OP_NUM_RANGE = 6,    /* kind NUMBER and lo/hi bounds, ia bit0 has_lo, bit1 incl_lo,
                             bit2 has_hi, bit3 incl_hi; fa = lo, fb = hi (facets.go:126) */
  QG_OP_SCODE_EQ = 7,     /* kinds in mask ib and scode == ia                           */
  QG_OP_SCODE_CMP = 8,    /* kinds in mask ib and scode (ic: 0 <,1 <=,2 >,3 >=) ia,
                             where ia is a rank in the column's sorted dictionary       */
  QG_OP_FCODE_EQ = 9,     /* kind STRING and fcode == ia                                */
  QG_OP_SCODE_IN = 10,    /* kinds in mask ib and scode in set[ia .. ia+ic)             */
  QG_OP_FCODE_IN = 11,    /* kind STRING and fcode in set[ia .. ia+ic)                  */
  QG_OP_NUM_IN_TOL = 12,  /* kind NUMBER and any |num - fset[ia+j]| <= fb, j < ic       */
  QG_OP_NUM_IN = 13,      /* kinds in mask ib and num == fset[ia+j] for some j < ic     */
  QG_OP_NUM_BITS_EQ = 14, /* kind NUMBER and bit pattern of num == bits(fa)             */
  QG_OP_ELEM_IN = 15,     /* kind OTHER and any element code of the row's array in
                             set[ia .. ia+ic) (SetFilter over array facets, facets.go:308-320;
                             needs qg_facets_set_array_column)                           */
  QG_OP_WHOLE_EQ = 16     /* kind OTHER and fcode == ia: the row's whole array / map value
                             equals the filter's under reflect.DeepEqual (EqualityFilter,
                             facets.go:85); for OTHER rows fcode is the id of the value in the
                             column's whole-value dictionary                              */
} qg_clause_op;

typedef struct qg_clause {
  int32_t op;      /* qg_clause_op */
  int32_t field;   /* column index */
  int32_t negate;  /* 1 = logical NOT of the clause result */
  int32_t ia, ib, ic;
  double fa, fb;
} qg_clause;

typedef struct qg_pred {
  int32_t first_clause; /* index into the clause array            */
  int32_t n_clauses;    /* OR of these; 0 clauses = false         */
  int32_t negate;       /* 1 = NOT(OR(...))                       */
  int32_t require_row;  /* 1 = rows of kind NOROW never match     */
} qg_pred;

int qg_filter_compile(qg_index* idx, const qg_pred* preds, int n_preds, const qg_clause* clauses,
                      int n_clauses, const int32_t* iset, int n_iset, const double* fset,
                      int n_fset, qg_filter** out);
/* Evaluate over all uploaded rows; tombstoned rows are NOT removed from this mask (the
 * search ANDs the live mask itself). mask_out (nullable) receives ceil(rows/64) words,
 * bit r%64 of word r/64 = row r matches. */
int qg_filter_eval(qg_index* idx, qg_filter* f, uint64_t* mask_out, int64_t* out_matches);
int qg_filter_destroy(qg_filter* f);

/* ---- search (replaces ExactIndex.Search exact.go:92-133, the exact branch of
 *      searchWithStrategy hybrid_index.go:515-570 and BatchSearch :677-811) ----------
 * For each of q queries: the k nearest live rows that pass `filter` (nullable), ascending
 * by (distance, row). out_dist / out_row are q x k, unused tail entries are +inf / -1;
 * out_count[i] = number of results of query i = min(k, matching live rows).
 * When `negatives` (q x dim, nullable) is given, out_negdist[i*k+j] = distance between
 * result row j and negatives[i] in the index metric (hybrid_index.go:536-546); the
 * caller applies `d - w*d_neg` and the (score, ID) order, which needs the string IDs.
 * Error order follows exact.go:96-106: an empty index returns 0 results and QG_OK even
 * when k <= 0; then the dimension is checked by the caller-supplied `dim`; then k. */
int qg_search_batch(qg_index* idx, const float* queries, int q, int dim, int k,
                    qg_filter* filter, const float* negatives, float* out_dist,
                    float* out_negdist, int64_t* out_row, int* out_count);

/* Same with every buffer resident on the index's device and the work enqueued on
 * `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream). Returns
 * after enqueueing; results are valid once the stream has been synchronised.
 * The scans only select candidates and a certificate proves the selection; a query whose
 * certificate failed (rare: the sampled threshold admitted too few rows, or adversarial data)
 * comes back with out_count = -1 and unspecified rows. qg_search_batch repeats such queries
 * through the flat scan and the exhaustive path before it returns; a caller of this
 * asynchronous form checks the counts and re-submits those queries one at a time. */
int qg_search_batch_device(qg_index* idx, const void* d_queries, int q, int dim, int k,
                           qg_filter* filter, const void* d_negatives, void* d_out_dist,
                           void* d_out_negdist, void* d_out_row, void* d_out_count,
                           void* stream);

/* The reference algorithm verbatim on the device (exact.go:114-129): the exact distance of EVERY live
 * row that passes `filter`, a full sort by (distance, row), the first k. Same arguments and results as
 * qg_search_batch; no candidate selection, no certificate, q full passes over the corpus. This is the
 * library's own last-resort path, exported so that full-size parity tests can use it as the GPU-side
 * oracle: it is checked against the CPU oracle on corpora the CPU finishes, and the fast regimes are
 * checked against it on the 10M / 100M-row configurations ("parity chain"). */
int qg_search_exhaustive(qg_index* idx, const float* queries, int q, int dim, int k, qg_filter* filter,
                         float* out_dist, int64_t* out_row, int* out_count);

/* ---- row-sharded search across GPUs (no reference counterpart) -----------
 * Per-shard top-k as packed 64-bit keys: high 32 bits = order-preserving image of the
 * exact float32 distance, low 32 bits = global row (row_base + local row). Missing
 * entries are all-ones. The keys of all shards are exchanged by the host layer (NCCL
 * all-gather) and merged by qg_merge_shard_keys_device. Uncertified queries (see above) are
 * repeated by the exhaustive path inside the call, which therefore synchronises `stream` once. */
int qg_search_shard_keys_device(qg_index* idx, const void* d_queries, int q, int dim, int k,
                                qg_filter* filter, int64_t row_base, void* d_out_keys /* q x k u64 */,
                                void* stream);
/* d_keys_gathered: world x q x k u64 (rank-major). Writes q x k distances / int64 global
 * rows / counts. `device` = device the buffers live on. */
int qg_merge_shard_keys_device(int device, const void* d_keys_gathered, int world, int q, int k,
                               void* d_out_dist, void* d_out_row, void* d_out_count,
                               void* stream);

/* ---- several GPUs of one box behind the C ABI (no reference counterpart: Quiver is one
 *      process with one in-memory index, pkg/core/db.go:707-845 -> hybrid_index.go:677-811) --------
 * The collective is NCCL over NVLink / NVSwitch, loaded at run time (libnccl.so.2) the first time one
 * of these entry points is used; the rest of the library has no NCCL dependency. Two shapes:
 *
 *  (1) one process per GPU (the launch shape of bench.py / torchrun): every rank holds a qg_comm made
 *      from a shared 128-byte id (qg_comm_unique_id on rank 0, handed to the others by the caller) and
 *      its own qg_index (shard or replica); qg_comm_search_* runs this rank's part of a batch, the
 *      exchange and the merge, all enqueued on the caller's stream.
 *  (2) one host process driving all GPUs (the shape a Go host has): qg_group_* owns one index, one
 *      stream, one NCCL communicator (ncclCommInitAll) and one worker thread per device.
 *
 * Layouts (qg_layout): QG_LAYOUT_ROWS — the corpus is row-sharded in contiguous blocks, every GPU
 * scans its shard for the whole batch, the per-shard top-k keys travel through one all-gather of
 * q*k*8 bytes per rank and are merged (north_star); QG_LAYOUT_QUERIES — every GPU holds the whole corpus
 * and answers a contiguous block of the batch (the reference's own structure: BatchSearch runs one
 * goroutine per query over one shared index). Both are exact, results equal the single-GPU run bit for bit. */
typedef enum qg_layout { QG_LAYOUT_ROWS = 0, QG_LAYOUT_QUERIES = 1, QG_LAYOUT_AUTO = 2 } qg_layout;
typedef struct qg_comm qg_comm;
#define QG_COMM_ID_BYTES 128
int qg_comm_unique_id(void* id_out /* QG_COMM_ID_BYTES */);
int qg_comm_create_rank(const void* id /* QG_COMM_ID_BYTES */, int world, int rank, int device, qg_comm** out);
int qg_comm_destroy(qg_comm* c);
int qg_comm_world(const qg_comm* c);
int qg_comm_rank(const qg_comm* c);
/* Row-sharded batch: `shard` holds rows [row_base, row_base + qg_index_rows) of the corpus; all buffers on
 * this rank's device; every rank receives the merged q x k result (global rows). Collective: every rank
 * of the communicator must call with the same q, dim, k. */
int qg_comm_search_rows_device(qg_comm* c, qg_index* shard, const void* d_queries, int q, int dim, int k,
                               qg_filter* filter, int64_t row_base, void* d_out_dist, void* d_out_row,
                               void* d_out_count, void* stream);
/* Query-split batch over replicas: d_queries holds all q queries on every rank, rank r answers the block
 * [r*ceil(q/world), ...). gather != 0: the result blocks are all-gathered and every rank ends with all q
 * results; gather == 0: no collective at all, the rank's block lands at its place of the q x k outputs
 * (the other rows are left untouched) — the caller copies only its own block to the host. */
int qg_comm_search_queries_device(qg_comm* c, qg_index* replica, const void* d_queries, int q, int dim, int k,
                                  qg_filter* filter, int gather, void* d_out_dist, void* d_out_row,
                                  void* d_out_count, void* stream);

typedef struct qg_group qg_group;
int qg_group_create(const int* devices, int n_devices, int dim, int metric, const qg_config* cfg /*nullable; .device ignored*/,
                    qg_group** out);
int qg_group_destroy(qg_group* g);
int qg_group_devices(const qg_group* g);
/* The per-device index (facet columns, tombstones, introspection); row numbering of a row-sharded group:
 * device i holds global rows [qg_group_row_base(g, i), + qg_index_rows). */
qg_index* qg_group_index(qg_group* g, int i);
int64_t qg_group_row_base(const qg_group* g, int i);
int64_t qg_group_rows(const qg_group* g);   /* rows of the corpus (not counting replicas) */
int qg_group_layout(const qg_group* g);     /* QG_LAYOUT_ROWS / QG_LAYOUT_QUERIES */
/* Load the corpus once (a group is filled by one upload call; `layout` AUTO: replicate when the fp32 rows
 * plus the bf16 copy stay under a quarter of one GPU's memory, else shard by rows). Host rows go through
 * each device's pinned staging, all devices in parallel. */
int qg_group_upload(qg_group* g, const float* rows, int64_t n, int layout);
int qg_group_upload_synthetic(qg_group* g, int kind, uint64_t seed, int64_t n, int layout);
/* qg_search_batch over the group: host buffers in, host buffers out, global rows. Row-sharded: every
 * device scans its shard, all-gather of the keys, device 0 merges and copies out. Replicated: every
 * device answers a block of the batch and copies it straight into the caller's buffers (no collective). */
int qg_group_search_batch(qg_group* g, const float* queries, int q, int dim, int k, float* out_dist,
                          int64_t* out_row, int* out_count);

/* ---- neighbour-distance batches for HNSW (replaces the per-pair computeDistance call
 *      in searchLayer, pkg/hnsw/hnsw.go:536-563 / :547) ------------------------------
 * out[i] = distance(query, row rows[i]) in the index metric and arithmetic;
 * rows[i] == 0xFFFFFFFF yields +inf. */
int qg_batch_distance(qg_index* idx, const float* query, int dim, const uint32_t* rows, int n,
                      float* out);
/* b queries at once, each with m candidate rows (row-major b x m, 0xFFFFFFFF = skip). */
int qg_batch_distance_multi(qg_index* idx, const float* queries, int b, int dim,
                            const uint32_t* rows, int m, float* out);

/* The same with the queries resident on the device: a graph walk issues hundreds of expansion
 * steps for one batch of queries, so the queries are uploaded once. rows is b x m (host), entry
 * [i*m + j] belongs to query i of the set; 0xFFFFFFFF = skip (+inf). */
typedef struct qg_queries qg_queries;
int qg_queries_upload(qg_index* idx, const float* queries, int b, int dim, qg_queries** out);
int qg_queries_destroy(qg_queries* qs);
int qg_batch_distance_queries(qg_index* idx, const qg_queries* qs, const uint32_t* rows, int m, float* out);

/* ---- hnsw.Search on the device (replaces the whole walk of pkg/hnsw/hnsw.go:471-580, 602-672) -----
 * The reference's graph as flat arrays, resident on the index's device; node id = row of the index
 * (insertion order): level[n] (-1 = deleted node), adj0[n x max_m0] level-0
 * connections in list order padded with 0xFFFFFFFF, upper_off[n + 1] / upper_adj = per node level[i]
 * blocks of m entries for the levels 1..level[i]. The arrays are copied. */
typedef struct qg_hnsw qg_hnsw;
int qg_hnsw_upload(qg_index* idx, int64_t n_nodes, int m, int max_m0, int entry_point, int current_level,
                   const int32_t* level, const uint32_t* adj0, const int64_t* upper_off, const uint32_t* upper_adj,
                   qg_hnsw** out);
int qg_hnsw_destroy(qg_hnsw* g);
/* Graph construction on the device for the rows the index holds (replaces the loop of hnsw.Insert /
 * connectNode calls, hnsw.go:266-468, that builds the reference's graph). Nodes are
 * inserted in batches: each node of a batch searches the graph of the committed nodes with
 * ef = ef_construction, links to its closest m (max_m0 on level 0) results — selectNeighbors' plain
 * closest-k, hnsw.go:583-599 — and the reverse links are merged afterwards with the reference's prune rule.
 * Levels follow randomLevel's law (p = 0.25 per level, at most min(max_level, 10) promotions, hnsw.go:716-738)
 * from the counter-based hash of (seed, node). A batched build is not step-identical to sequential inserts;
 * its bar is recall at equal efSearch against the host-built graph. max_batch <= 0 = default (8192). */
int qg_hnsw_build(qg_index* idx, int m, int max_m0, int ef_construction, int max_level, uint64_t seed, int max_batch,
                  qg_hnsw** out);
/* The graph as flat host arrays (the layout qg_hnsw_upload takes). upper_adj has qg_hnsw_upper_len entries. */
int64_t qg_hnsw_nodes(const qg_hnsw* g);
int64_t qg_hnsw_upper_len(const qg_hnsw* g);
int qg_hnsw_export(const qg_hnsw* g, int32_t* level, uint32_t* adj0, int64_t* upper_off, uint32_t* upper_adj,
                   int* entry_point, int* current_level);
/* One persistent kernel, a warp per query: ef = 1 descent through the upper layers, base layer with
 * ef = max(ef_search, k), the reference's heaps / visit order / stop and admit rules, distances in the
 * index's metric and arithmetic — step-identical to the reference's walk. out_idx / out_dist are q x k
 * (k already clamped to the node count by the caller, hnsw.go:615-617; unused entries 0xFFFFFFFF / +inf),
 * out_count[i] = results of the graph walk (may be < k: the caller runs the under-fill exact pass,
 * hnsw.go:676-710) or -1 when the query's candidate heap outgrew the kernel's shared-memory slice (the
 * caller repeats it with the host walk); out_evals (nullable) = distance evaluations per query. */
int qg_hnsw_search_batch(qg_index* idx, const qg_hnsw* g, const float* queries, int q, int dim, int k, int ef_search,
                         uint32_t* out_idx, float* out_dist, int* out_count, int64_t* out_evals);

/* ---- introspection for benchmarks and tests -----------------------------------------*/
typedef struct qg_scan_stats {
  int64_t rows_scanned;     /* rows whose vector bytes the last scan read            */
  int64_t bytes_algorithmic;/* rows_scanned*dim*4 + side columns read                */
  int32_t kernel_launches;  /* kernels launched by the last search call              */
  int32_t queries_per_pass; /* queries served by one pass over the corpus            */
  int32_t passes;           /* corpus passes of the last call                        */
  int32_t escalations;      /* selections repeated with a larger candidate set       */
  int32_t path;             /* 0 exhaustive, 1 flat scan, 2 gather scan, 3 tensor-core */
  int32_t reserved;
} qg_scan_stats;
int qg_last_scan_stats(const qg_index* idx, qg_scan_stats* out);

/* Per-kernel device time, for roofline reporting: while profiling is on, every scan / merge
 * launch of qg_search_*_device is bracketed by CUDA events on the caller's stream.
 * qg_index_read_profile waits for the recorded events, returns the sums and clears them. */
typedef struct qg_profile {
  double scan_ms;            /* sum over scan-kernel launches                       */
  double finalize_ms;        /* sum over merge / re-rank launches                   */
  int64_t scan_launches;
  int64_t finalize_launches;
  double prep_ms;            /* tensor-core regime: sample + threshold kernels      */
  int64_t prep_launches;
} qg_profile;
int qg_index_set_profiling(qg_index* idx, int on);
int qg_index_read_profile(qg_index* idx, qg_profile* out);

/* Development aid: run one tensor-core pass (sample -> threshold -> scan) for nq <= 256 host
 * queries and copy back the per-query admission thresholds, candidate counts and the raw
 * candidate keys (nq x *cap_out; high 32 bits = order-preserving image of the tf32 scan score,
 * low 32 bits = row). Nothing is re-ranked. */
int qg_debug_tc_pass(qg_index* idx, const float* queries, int nq, int k, float* tau_out, int* cnt_out,
                     uint64_t* cand_out, int* cap_out);

#ifdef __cplusplus
}
#endif
#endif /* QUIVER_GPU_H */
