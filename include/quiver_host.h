/*
 * quiver_host.h — C ABI of libquiverhost.so: the host side above libquivergpu.so.
 *
 * The reference host is Go; there is no Go toolchain in the build image, so the host logic a
 * `pkg/gpu` package would carry (string IDs, request validation, negative-example re-rank, the
 * predicate compiler) is written in C++ and mirrors the reference's names, argument meaning and
 * error text:
 *   qh_index       hybrid.HybridIndex in exact mode behind core.Index
 *                  (pkg/hybrid/hybrid_index.go:86-585, 677-811; pkg/core/collection.go:78-96)
 *   qh_collection  core.Collection: metadata, facet fields, Search / FluentSearch / SearchWithFacets
 *                  (pkg/core/collection.go:133-331, 637-807, 874-1108, 1141-1207)
 * Every function returns 0 or a qg_status code (include/quiver_gpu.h); qh_last_error() returns
 * the message the reference would have put in its `error`.
 */
#ifndef QUIVER_HOST_H
#define QUIVER_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qh_index qh_index;
typedef struct qh_collection qh_collection;
typedef struct qh_results qh_results;

const char* qh_last_error(void);

/* ---- result sets: one list of (id, distance) per query ------------------------------------ */
int qh_results_queries(const qh_results* r);
int qh_results_count(const qh_results* r, int query);
const char* qh_results_id(const qh_results* r, int query, int j);
float qh_results_distance(const qh_results* r, int query, int j);
/* types.SearchResultItem (pkg/types/search.go:31-42): Score = 1 - Distance (collection.go:763); Vector and
 * Metadata are present only in the results of qh_collection_search_request when its options asked for them
 * (NULL / length 0 otherwise, like the `omitempty` fields). */
float qh_results_score(const qh_results* r, int query, int j);
const float* qh_results_vector(const qh_results* r, int query, int j, int* out_len);
const char* qh_results_metadata(const qh_results* r, int query, int j);
int qh_results_free(qh_results* r);

/* ---- hybrid index, exact strategy ------------------------------------------------------------
 * distance: "cosine" | "euclidean" | "dot_product" | "manhattan" | "squared_euclidean" and the
 * HTTP aliases "l2", "dot", "cos", "" (pkg/api/handlers.go:66-72); unknown => cosine
 * (pkg/vectortypes/types.go:46-47). arith: 0 = vectortypes float64, 1 = hnsw float32. */
int qh_index_create(qh_index** out, int dim, const char* distance, int arith, int device);
int qh_index_destroy(qh_index* idx);
/* Insert (hybrid_index.go:86-131): dimension and duplicate-ID errors with the reference's text;
 * the vector is copied. */
int qh_index_insert(qh_index* idx, const char* id, const float* vec, int dim);
int qh_index_insert_batch(qh_index* idx, const char* const* ids, const float* vecs, int64_t n, int dim);
/* Delete (hybrid_index.go:241-290): a missing ID is an error (ExactIndex alone would not mind). */
int qh_index_delete(qh_index* idx, const char* id);
/* HybridIndex.DeleteBatch (hybrid_index.go:293-375): "some vectors not found: [a b]" and nothing deleted
 * when an id is missing; otherwise one tombstone launch for the whole batch. */
int qh_index_delete_batch(qh_index* idx, const char* const* ids, int64_t n);
int64_t qh_index_size(const qh_index* idx);
/* Drop the rows of deleted vectors from the device arrays (qg_index_compact) and renumber the
 * id <-> row tables. The reference frees a vector the moment it is deleted (exact.go:61-70,
 * hybrid_index.go:244-290); a wrapper calls this when deleted rows outnumber live ones, under the
 * write lock it already holds for Delete (collection.go:376). *out_removed (nullable) = rows dropped. */
int qh_index_compact(qh_index* idx, int64_t* out_removed);
/* Search(query, k) (hybrid_index.go:378; exact.go:92-133). */
int qh_index_search(qh_index* idx, const float* query, int dim, int k, qh_results** out);

/* SearchWithRequest / BatchSearch (hybrid_index.go:383-469, 677-811). force_strategy: "" or
 * "exact" run the exact path; "hnsw" is rejected here (the graph walk is not part of this index);
 * anything else is the reference's "invalid search strategy" error. negatives: nq x dim or NULL;
 * a negative example is used only when negative_weight > 0 (hybrid_index.go:417). */
int qh_index_batch_search(qh_index* idx, const float* queries, int nq, int dim, int k, const float* negatives,
                          int neg_dim, float negative_weight, const char* force_strategy, qh_results** out);

/* ---- collection ------------------------------------------------------------------------------- */
int qh_collection_create(qh_collection** out, const char* name, int dim, const char* distance, int device);
int qh_collection_destroy(qh_collection* c);
/* Add (collection.go:133-215): metadata_json may be NULL. */
int qh_collection_add(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json);
int qh_collection_add_batch(qh_collection* c, const char* const* ids, const float* vecs, int64_t n, int dim,
                            const char* const* metadata_json /* n entries, each may be NULL */);
int qh_collection_delete(qh_collection* c, const char* id);
/* Collection.DeleteBatch (collection.go:375-414): "vector not found: <id>" for the first missing id. */
int qh_collection_delete_batch(qh_collection* c, const char* const* ids, int64_t n);
/* Collection.Update (collection.go:417-466): vec nullable (keep the vector), metadata_json nullable / empty
 * (keep the metadata). A new vector is Delete + Insert under the same id. */
int qh_collection_update(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json);
/* Collection.UpdateBatch (collection.go:469-529): all vectors validated first ("no vectors provided for batch
 * update", "vector ID cannot be empty", "vector not found: <id>", "invalid vector dimension for vector <id>:
 * expected %d, got %d", "invalid metadata format for vector <id>: ..."), then one tombstone launch and one
 * upload for the batch. metadata_json nullable; a null / empty entry keeps that vector's metadata. */
int qh_collection_update_batch(qh_collection* c, const char* const* ids, const float* vecs, int64_t n, int dim,
                               const char* const* metadata_json);
int64_t qh_collection_count(const qh_collection* c);
/* qh_index_compact for the collection's index; the per-row metadata moves with the rows. */
int qh_collection_compact(qh_collection* c, int64_t* out_removed);
int qh_collection_set_facet_fields(qh_collection* c, const char* const* fields, int n);

/* types.Filter (pkg/types/search.go:45-52). value_json is the operand as JSON text; a number
 * written without '.', 'e', 'E' is a Go int, otherwise a float64 (this only matters for its "%v"
 * text). op: "=", "!=", ">", ">=", "<", "<=", "in", "not_in". */
typedef struct qh_filter {
  const char* field;
  const char* op;
  const char* value_json;
} qh_filter;
/* Collection.Search / FluentSearch.Execute (collection.go:637-807, 1094): ranks the rows that pass
 * all filters and returns the first k (the reference ranks all rows and filters afterwards; the
 * result is the same list). k is clamped to Count() like FluentSearch.validate (:924-926). */
int qh_collection_search(qh_collection* c, const float* query, int dim, int k, const qh_filter* filters, int n_filters,
                         qh_results** out);

/* types.SearchOptions + SearchRequest.NamespaceID (pkg/types/search.go:45-52, 84-85) as FluentSearch sets
 * them (IncludeVectors / IncludeMetadata / UseExactSearch / WithNamespace, collection.go:946-985). */
typedef struct qh_search_options {
  int include_vectors;
  int include_metadata;
  int exact_search;          /* carried, not consulted: this index is always exact (see host.cpp) */
  const char* namespace_id;  /* carried, not consulted by Collection.Search (nullable) */
} qh_search_options;
/* Collection.Search(types.SearchRequest) (collection.go:637-807): qh_collection_search + the decoration of
 * every result with its stored vector / metadata document (collection.go:758-779). opt nullable. */
int qh_collection_search_request(qh_collection* c, const float* query, int dim, int k, const qh_filter* filters,
                                 int n_filters, const qh_search_options* opt, qh_results** out);
/* persistence.Collection.Search / SearchWithFacets (pkg/persistence/collection.go:226-261, 327-378): the
 * prefilter path under that type's argument checks (limit <= 0 = every row). */
int qh_collection_persistence_search(qh_collection* c, const float* query, int dim, int limit,
                                     const struct qh_facet_filter* filters, int n_filters, qh_results** out);

/* facets.Filter implementations (pkg/facets/facets.go). type: 0 equality (value_json), 1 range
 * (min_json / max_json, NULL or "null" = open; include flags), 2 set (value_json = JSON array),
 * 3 exists (should_exist). */
typedef struct qh_facet_filter {
  int type;
  const char* field;
  const char* value_json;
  const char* min_json;
  const char* max_json;
  int include_min, include_max;
  int should_exist;
} qh_facet_filter;
/* Collection.SearchWithFacets (collection.go:1141-1207). */
int qh_collection_search_with_facets(qh_collection* c, const float* query, int dim, int k,
                                     const qh_facet_filter* filters, int n_filters, qh_results** out);
/* Row-pass bits of a predicate set over the rows in insertion order (bit-exactness checks):
 * mask_out receives count bytes (0/1); ids_out (nullable) the matching row ids are not returned —
 * use qh_collection_row_id. which: 0 = core filters, 1 = facet filters. */
int qh_collection_filter_mask(qh_collection* c, int which, const qh_filter* filters, const qh_facet_filter* ffilters,
                              int n_filters, uint8_t* mask_out, int64_t n_rows);
int64_t qh_collection_rows(const qh_collection* c);
const char* qh_collection_row_id(const qh_collection* c, int64_t row);

/* ---- HNSW search with GPU-batched neighbour distances -------------------------------------------
 * The graph is the reference's (pkg/hnsw/hnsw.go:44-82), viewed as flat arrays; node id = row of
 * the index (insertion order). The walk is hnsw.Search / searchLayer (hnsw.go:471-580, 602-713)
 * step for step — same heaps, same visit order, same stop / admit rules — except that the
 * distances of one expansion step (hnsw.go:536-563, up to MaxM0 = 32 neighbours) of ALL queries
 * of the batch are evaluated by one qg_batch_distance_queries call. */
typedef struct qh_hnsw_graph {
  int64_t n_nodes;
  int m, max_m0;             /* hnsw.Config M / MaxM0 (hnsw.go:16-25) */
  int entry_point, current_level, ef_search;
  const int32_t* level;      /* [n] node level, -1 = deleted node (nil in HNSW.Nodes) */
  const uint32_t* adj0;      /* [n x max_m0] level-0 connections in list order, 0xFFFFFFFF padded */
  const int64_t* upper_off;  /* [n+1] start of node i's upper-level block in upper_adj */
  const uint32_t* upper_adj; /* per node: level[i] blocks of m entries (levels 1..level[i]) */
} qh_hnsw_graph;
/* out_evals / out_steps (nullable): distance evaluations per query [nq] and lock-step rounds. */
int qh_hnsw_search_batch(qh_index* idx, const qh_hnsw_graph* g, const float* queries, int nq, int dim, int k,
                         qh_results** out, int64_t* out_evals, int64_t* out_steps);

/* The same search with the whole walk on the device: the graph is uploaded once (qh_hnsw_upload; the
 * caller keeps the host arrays of `g` alive for the lifetime of the handle), every search is one launch of
 * the persistent kernel behind qg_hnsw_search_batch (a warp per query, heaps in shared memory, the
 * reference's sift rules — step-identical), the under-fill exact pass of all affected queries is ONE
 * batched exact search. out_fallbacks (nullable): queries repeated by the lock-step walk above because
 * their candidate heap outgrew the kernel's shared-memory slice. */
typedef struct qh_hnsw_dev qh_hnsw_dev;
int qh_hnsw_upload(qh_index* idx, const qh_hnsw_graph* g, qh_hnsw_dev** out);
int qh_hnsw_dev_free(qh_hnsw_dev* d);
int qh_hnsw_search_device(qh_index* idx, qh_hnsw_dev* d, const float* queries, int nq, int dim, int k,
                          qh_results** out, int64_t* out_evals, int* out_fallbacks);
/* HNSWAdapter.SearchWithNegativeExample (pkg/hnsw/adapter.go:345-437): max(2k, 30) candidates from the graph
 * search (clamped to the index size), Distance - w * negDistance in float32 with w clamped to at most 1
 * (:375-377), stable (Distance, ID) order, the first k — the returned Distance is the adjusted score, as in
 * the reference. No negative example / w <= 0 / no more than k candidates: the plain search truncated to k. */
int qh_hnsw_search_negative(qh_index* idx, qh_hnsw_dev* d, const float* query, int dim, const float* negative,
                            int neg_dim, float negative_weight, int k, qh_results** out);

/* ---- development aids (CPU-only; no device needed) --------------------------------------------- */
/* fmt.Sprintf("%v", json value) into buf; returns the length or -1 on a JSON error. */
int qh_debug_sprint_v(const char* value_json, int typed_literals, char* buf, int buf_len);
/* strings.EqualFold(a, b) as the predicate compiler sees it. */
int qh_debug_equal_fold(const char* a, const char* b);

#ifdef __cplusplus
}
#endif
#endif /* QUIVER_HOST_H */
