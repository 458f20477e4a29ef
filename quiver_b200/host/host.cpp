// host.cpp — libquiverhost.so: the host side above the GPU C ABI (include/quiver_host.h).
// String IDs, request validation with the reference's error text, the negative-example re-rank
// (pkg/hybrid/hybrid_index.go:515-570) and the predicate pushdown for Collection.Search /
// SearchWithFacets (pkg/core/collection.go:679-752, 1141-1207).
#include "../../include/quiver_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/quiver_gpu.h"
#include "filter_compile.hpp"
#include "value.hpp"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int gpu_fail(int rc) { return fail(rc, qg_last_error()); }

int metric_of(const char* name) {
  const std::string s = name ? name : "";
  if (s == "euclidean" || s == "l2") return QG_L2;
  if (s == "dot_product" || s == "dot") return QG_DOT;
  if (s == "manhattan") return QG_L1;
  if (s == "squared_euclidean") return QG_SQL2;
  return QG_COSINE;  // "cosine", "cos", "" and anything unknown (types.go:46-47)
}

struct Hit {
  std::string id;
  float distance;
  // types.SearchResultItem (pkg/types/search.go:31-42): filled by qh_collection_search_request
  std::vector<float> vector;  // Options.IncludeVectors
  std::string metadata;       // Options.IncludeMetadata: the stored json.RawMessage
  bool has_metadata = false;
};

}  // namespace

struct qh_results {
  std::vector<std::vector<Hit>> lists;
};

// ---- hybrid index (exact strategy) ---------------------------------------------------------------
struct qh_index {
  qg_index* h = nullptr;
  int dim = 0;
  mutable std::shared_mutex mu;          // hybrid_index.go:42: searches share, mutations exclude
  std::vector<std::string> ids;          // row -> id
  std::unordered_map<std::string, int64_t> rows;  // live id -> row
  // Single-vector Inserts are write-combined: the vector is copied here (Insert copies, exact.go:53-54), its row
  // number is fixed at once, and the H2D upload happens for the whole run of them — before the next call that
  // reads device rows (search, delete, compact, a filter), or when the buffer holds `pending_cap` rows.
  std::vector<float> pending;
  int64_t pending_rows = 0;
  int64_t pending_cap = 4096;  // 0 = upload every Insert on its own (QH_INSERT_BUFFER=0)
};

namespace {

// Upload the write-combined Inserts (caller holds idx->mu exclusively). If the upload fails the buffered rows
// are dropped together with their ids (they are the last pending_rows entries of idx->ids), so that the id
// table and the device rows stay in step: the failing call reports the error, a retry of those Inserts works,
// and later searches are not poisoned by rows that can never be flushed.
int flush_locked(qh_index* idx) {
  if (idx->pending_rows == 0) return 0;
  int64_t first = 0;
  int rc = qg_index_upload(idx->h, idx->pending.data(), idx->pending_rows, &first);
  if (rc == 0 && first + idx->pending_rows != (int64_t)idx->ids.size()) rc = -1;
  if (rc != 0) {
    const int err = rc > 0 ? gpu_fail(rc) : fail(QG_ERR_CUDA, "insert buffer and device rows out of step");
    for (int64_t i = 0; i < idx->pending_rows && !idx->ids.empty(); ++i) {
      idx->rows.erase(idx->ids.back());
      idx->ids.pop_back();
    }
    idx->pending.clear();
    idx->pending_rows = 0;
    return err;
  }
  idx->pending.clear();
  idx->pending_rows = 0;
  return 0;
}

// Shared lock for a reader of device rows: taken only once nothing is pending (the flush needs the exclusive
// lock, so it is done in between; an Insert slipping in just repeats the round).
std::shared_lock<std::shared_mutex> lock_flushed(qh_index* idx, int* rc) {
  *rc = 0;
  for (;;) {
    std::shared_lock<std::shared_mutex> lk(idx->mu);
    if (idx->pending_rows == 0) return lk;
    lk.unlock();
    std::unique_lock<std::shared_mutex> wl(idx->mu);
    if ((*rc = flush_locked(idx))) return std::shared_lock<std::shared_mutex>();
  }
}

int index_insert_locked(qh_index* idx, const char* const* ids, const float* vecs, int64_t n, int dim) {
  if (dim != idx->dim)
    return fail(QG_ERR_DIM, "vector dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  std::unordered_map<std::string, int> batch;
  for (int64_t i = 0; i < n; ++i) {
    const std::string id = ids[i] ? ids[i] : "";
    if (idx->rows.count(id) || !batch.emplace(id, 1).second)
      return fail(QG_ERR_INVALID, "vector with ID " + id + " already exists");
  }
  if (n == 1 && idx->pending_cap > 0) {
    idx->pending.insert(idx->pending.end(), vecs, vecs + dim);
    idx->pending_rows++;
    const std::string id0 = ids[0] ? ids[0] : "";
    idx->rows[id0] = (int64_t)idx->ids.size();
    idx->ids.push_back(id0);
    return idx->pending_rows >= idx->pending_cap ? flush_locked(idx) : 0;
  }
  if (int rc = flush_locked(idx)) return rc;  // rows keep their insertion order
  int64_t first = 0;
  if (int rc = qg_index_upload(idx->h, vecs, n, &first)) return gpu_fail(rc);
  for (int64_t i = 0; i < n; ++i) {
    const std::string id = ids[i] ? ids[i] : "";
    idx->ids.push_back(id);
    idx->rows[id] = first + i;
  }
  return 0;
}

// searchWithStrategy, exact branch (hybrid_index.go:515-570) for a batch of queries.
int index_search_locked(qh_index* idx, const float* queries, int nq, int dim, int k, const float* negatives,
                        float neg_weight, qg_filter* filter, int64_t n_matching, qh_results* out) {
  out->lists.assign((size_t)nq, {});
  if (nq == 0) return 0;
  const bool has_neg = negatives != nullptr && neg_weight > 0.f;
  int retrieve = k;
  if (has_neg && k > 0) {
    retrieve = std::max(2 * k, 30);  // hybrid_index.go:517-522
    const int64_t cap = filter ? n_matching : (int64_t)idx->rows.size();
    if ((int64_t)retrieve > cap) retrieve = (int)std::max<int64_t>(cap, 1);
  }
  const int kk = std::max(retrieve, 1);
  std::vector<float> dist((size_t)nq * kk), negd(has_neg ? (size_t)nq * kk : 0);
  std::vector<int64_t> row((size_t)nq * kk);
  std::vector<int> cnt((size_t)nq);
  if (int rc = qg_search_batch(idx->h, queries, nq, dim, retrieve, filter, has_neg ? negatives : nullptr, dist.data(),
                               has_neg ? negd.data() : nullptr, row.data(), cnt.data()))
    return gpu_fail(rc);
  for (int i = 0; i < nq; ++i) {
    std::vector<Hit>& list = out->lists[(size_t)i];
    const int n = cnt[(size_t)i];
    list.reserve((size_t)n);
    for (int j = 0; j < n; ++j) {
      float d = dist[(size_t)i * kk + j];
      if (has_neg) {
        // result.Distance - negWeight*negDistance, in float32 like the Go expression (:552);
        // compiled with -ffp-contract=off so the multiply and the subtract are not fused
        const float prod = neg_weight * negd[(size_t)i * kk + j];
        d = d - prod;
      }
      list.push_back(Hit{idx->ids[(size_t)row[(size_t)i * kk + j]], d, {}, {}, false});
    }
    if (has_neg) {
      // sort.SliceStable by (Distance, ID) (:555-560), then the first k (:567-569)
      std::stable_sort(list.begin(), list.end(), [](const Hit& a, const Hit& b) {
        if (a.distance == b.distance) return a.id < b.id;
        return a.distance < b.distance;
      });
      if ((int)list.size() > k) list.resize((size_t)k);
    }
  }
  return 0;
}

}  // namespace

// ---- collection -------------------------------------------------------------------------------------
struct SharedFilter {
  qg_filter* f = nullptr;
  ~SharedFilter() { if (f) qg_filter_destroy(f); }
};
struct qh_collection : qh::ColumnSource {
  std::string name;
  int dim = 0;
  qh_index* index = nullptr;
  std::vector<qh::ValuePtr> metadata;      // per row: parsed metadata object, or nullptr
  std::vector<std::string> metadata_raw;   // per row: the document as stored (json.RawMessage, collection.go:160-168); "" = none
  std::vector<std::string> facet_fields;
  // device columns: one per (family, field); rebuilt lazily after a mutation
  struct DevCol {
    int index;
    uint64_t epoch;
    qh::Column col;
  };
  std::unordered_map<std::string, DevCol> cols;  // key: "m:" + field or "f:" + path
  uint64_t epoch = 1;
  int family = 0;  // 0 = metadata columns, 1 = facet columns (set before compiling)
  std::mutex col_mu;  // column (re)builds touch device state: one search at a time compiles predicates
  // Collection.mu (collection.go:100): Add / Update / Delete / compact / SetFacetFields hold it exclusively for
  // their whole span — id lookups, the index mutation and the metadata edit are one step for every reader —
  // Search / FluentSearch / SearchWithFacets / Get share it. Lock order: cmu, then index->mu, then col_mu.
  mutable std::shared_mutex cmu;
  // compiled predicate programs of the current collection state (see make_filter)
  std::mutex filter_cache_mu;
  std::vector<std::pair<std::string, std::shared_ptr<SharedFilter>>> filter_cache;
  uint64_t filter_cache_epoch = 0;

  const qh::Value* facet_value(const qh::Value& md, const std::string& path) const {
    // ExtractFacets dot-path walk (facets.go:405-421); nil values are dropped (:423)
    const qh::Value* v = &md;
    size_t start = 0;
    for (;;) {
      const size_t dot = path.find('.', start);
      const std::string part = path.substr(start, dot == std::string::npos ? std::string::npos : dot - start);
      if (v->type != qh::Value::Object) return nullptr;
      auto it = v->obj.find(part);
      if (it == v->obj.end()) return nullptr;
      v = it->second.get();
      if (dot == std::string::npos) break;
      start = dot + 1;
    }
    return v->type == qh::Value::Null ? nullptr : v;
  }

  DevCol& ensure(const std::string& field) {
    const std::string key = (family ? "f:" : "m:") + field;
    auto it = cols.find(key);
    if (it == cols.end()) it = cols.emplace(key, DevCol{(int)cols.size(), 0, {}}).first;
    DevCol& dc = it->second;
    if (dc.epoch == epoch) return dc;
    const size_t n = metadata.size();
    std::vector<qh::CellRef> cells(n);
    for (size_t r = 0; r < n; ++r) {
      const qh::Value* md = metadata[r].get();
      if (!md) { cells[r] = {nullptr, true}; continue; }
      if (family == 0) {
        auto f = md->obj.find(field);
        cells[r] = {f == md->obj.end() ? nullptr : f->second.get(), false};
      } else {
        // rows whose facet list is empty never match (MatchesAllFilters, facets.go:437-439)
        bool any = false;
        for (const std::string& ff : facet_fields)
          if (facet_value(*md, ff)) { any = true; break; }
        const bool is_facet = std::find(facet_fields.begin(), facet_fields.end(), field) != facet_fields.end();
        if (!any) cells[r] = {nullptr, true};
        else cells[r] = {is_facet ? facet_value(*md, field) : nullptr, false};
      }
    }
    qh::encode_column(cells, &dc.col);
    last_rc = qg_facets_set_column(index->h, dc.index, dc.col.kind.data(), dc.col.num.data(), dc.col.scode.data(),
                                   dc.col.fcode.data(), (int64_t)n);
    if (!last_rc && dc.col.has_array_rows)
      last_rc = qg_facets_set_array_column(index->h, dc.index, dc.col.arr_off.data(), dc.col.arr_code.data(), (int64_t)n,
                                           (int64_t)dc.col.arr_code.size());
    dc.epoch = epoch;
    return dc;
  }
  int last_rc = 0;
  int field_index(const std::string& name_) override { return ensure(name_).index; }
  const qh::Column& column(const std::string& name_) override { return ensure(name_).col; }
};

namespace {

int parse_operand(const char* text, qh::ValuePtr* out) {
  if (!text) { *out = std::make_shared<qh::Value>(); return 0; }
  std::string err;
  *out = qh::parse_json(text, true, &err);
  if (!*out) return fail(QG_ERR_INVALID, "filter value is not valid JSON: " + err);
  return 0;
}

int build_program(qh_collection* c, int which, const qh_filter* filters, const qh_facet_filter* ffilters, int n,
                  qh::Program* prog) {
  std::lock_guard<std::mutex> col_lock(c->col_mu);
  c->family = which;
  c->last_rc = 0;
  std::string err;
  int rc = 0;
  if (which == 0) {
    std::vector<qh::CoreFilter> fs((size_t)n);
    for (int i = 0; i < n; ++i) {
      fs[(size_t)i].field = filters[i].field ? filters[i].field : "";
      fs[(size_t)i].op = filters[i].op ? filters[i].op : "";
      if (int prc = parse_operand(filters[i].value_json, &fs[(size_t)i].value)) return prc;
    }
    rc = qh::compile_core_filters(fs, *c, prog, &err);
  } else {
    std::vector<qh::FacetFilter> fs((size_t)n);
    for (int i = 0; i < n; ++i) {
      qh::FacetFilter& f = fs[(size_t)i];
      const qh_facet_filter& s = ffilters[i];
      if (s.type < 0 || s.type > 3) return fail(QG_ERR_INVALID, "unknown facet filter type");
      f.type = (qh::FacetFilter::Type)s.type;
      f.field = s.field ? s.field : "";
      f.include_min = s.include_min != 0;
      f.include_max = s.include_max != 0;
      f.should_exist = s.should_exist != 0;
      if (s.type == 0) {
        if (int prc = parse_operand(s.value_json, &f.value)) return prc;
      } else if (s.type == 1) {
        if (int prc = parse_operand(s.min_json, &f.min)) return prc;
        if (int prc = parse_operand(s.max_json, &f.max)) return prc;
      } else if (s.type == 2) {
        qh::ValuePtr arr;
        if (int prc = parse_operand(s.value_json, &arr)) return prc;
        if (arr->type != qh::Value::Array) return fail(QG_ERR_INVALID, "set filter values must be a JSON array");
        f.values = arr->arr;
      }
    }
    rc = qh::compile_facet_filters(fs, *c, prog, &err);
  }
  if (rc) return fail(rc, err);
  if (c->last_rc) return gpu_fail(c->last_rc);
  return 0;
}

// A compiled predicate program on the device. Handles are shared through the collection's small cache
// (filter_cache below): the library keeps the evaluated mask — and, for batched searches, the dense view of
// the passing rows — inside the handle, so a predicate that comes back (the same category filter on every
// request) is evaluated and gathered once per collection state, not once per search.
struct FilterHandle {
  std::shared_ptr<SharedFilter> h;
  qg_filter* f = nullptr;
};

int make_filter(qh_collection* c, const qh::Program& prog, FilterHandle* out, int64_t* matches) {
  // cache key: the program's bytes; valid for the collection state (epoch) it was compiled in — column
  // indices and dictionary codes inside the clauses belong to that state
  std::string key;
  auto add = [&](const void* p, size_t n) { key.append(reinterpret_cast<const char*>(p), n); key.push_back('|'); };
  add(prog.preds.data(), prog.preds.size() * sizeof(qg_pred));
  add(prog.clauses.data(), prog.clauses.size() * sizeof(qg_clause));
  add(prog.iset.data(), prog.iset.size() * 4);
  add(prog.fset.data(), prog.fset.size() * 8);
  {
    std::lock_guard<std::mutex> lk(c->filter_cache_mu);
    if (c->filter_cache_epoch != c->epoch) {
      c->filter_cache.clear();
      c->filter_cache_epoch = c->epoch;
    }
    for (auto& e : c->filter_cache)
      if (e.first == key) out->h = e.second;
  }
  if (!out->h) {
    auto sh = std::make_shared<SharedFilter>();
    if (int rc = qg_filter_compile(c->index->h, prog.preds.data(), (int)prog.preds.size(), prog.clauses.data(),
                                   (int)prog.clauses.size(), prog.iset.data(), (int)prog.iset.size(), prog.fset.data(),
                                   (int)prog.fset.size(), &sh->f))
      return gpu_fail(rc);
    out->h = sh;
    std::lock_guard<std::mutex> lk(c->filter_cache_mu);
    if (c->filter_cache_epoch == c->epoch) {
      if (c->filter_cache.size() >= 4) c->filter_cache.erase(c->filter_cache.begin());  // oldest out
      c->filter_cache.emplace_back(key, sh);
    }
  }
  out->f = out->h->f;
  if (matches) {
    if (int rc = qg_filter_eval(c->index->h, out->f, nullptr, matches)) return gpu_fail(rc);
  }
  return 0;
}

}  // namespace

// ======================================================================================================
extern "C" {

const char* qh_last_error(void) { return g_err.c_str(); }

int qh_results_queries(const qh_results* r) { return r ? (int)r->lists.size() : 0; }
int qh_results_count(const qh_results* r, int q) {
  return (r && q >= 0 && q < (int)r->lists.size()) ? (int)r->lists[(size_t)q].size() : 0;
}
const char* qh_results_id(const qh_results* r, int q, int j) { return r->lists[(size_t)q][(size_t)j].id.c_str(); }
float qh_results_distance(const qh_results* r, int q, int j) { return r->lists[(size_t)q][(size_t)j].distance; }
float qh_results_score(const qh_results* r, int q, int j) { return 1.0f - r->lists[(size_t)q][(size_t)j].distance; }
const float* qh_results_vector(const qh_results* r, int q, int j, int* out_len) {
  const Hit& h = r->lists[(size_t)q][(size_t)j];
  if (out_len) *out_len = (int)h.vector.size();
  return h.vector.empty() ? nullptr : h.vector.data();
}
const char* qh_results_metadata(const qh_results* r, int q, int j) {
  const Hit& h = r->lists[(size_t)q][(size_t)j];
  return h.has_metadata ? h.metadata.c_str() : nullptr;
}
int qh_results_free(qh_results* r) {
  delete r;
  return 0;
}

int qh_index_create(qh_index** out, int dim, const char* distance, int arith, int device) {
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  qg_config cfg{};
  cfg.device = device;
  cfg.arith = arith;
  qg_index* h = nullptr;
  if (int rc = qg_index_create(&h, dim, metric_of(distance), &cfg)) return gpu_fail(rc);
  qh_index* idx = new qh_index();
  idx->h = h;
  idx->dim = dim;
  if (const char* e = std::getenv("QH_INSERT_BUFFER")) idx->pending_cap = std::max(0, std::atoi(e));
  *out = idx;
  return 0;
}

int qh_index_destroy(qh_index* idx) {
  if (!idx) return 0;
  qg_index_destroy(idx->h);
  delete idx;
  return 0;
}

int qh_index_insert(qh_index* idx, const char* id, const float* vec, int dim) {
  if (!idx || !id || !vec) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(idx->mu);
  const char* ids[1] = {id};
  return index_insert_locked(idx, ids, vec, 1, dim);
}

int qh_index_insert_batch(qh_index* idx, const char* const* ids, const float* vecs, int64_t n, int dim) {
  if (!idx || (n > 0 && (!ids || !vecs))) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(idx->mu);
  return index_insert_locked(idx, ids, vecs, n, dim);
}

int qh_index_delete(qh_index* idx, const char* id) {
  if (!idx || !id) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(idx->mu);
  auto it = idx->rows.find(id);
  if (it == idx->rows.end()) return fail(QG_ERR_INVALID, std::string("vector with ID ") + id + " not found");
  const int64_t row = it->second;
  if (int rc = flush_locked(idx)) return rc;
  if (int rc = qg_index_tombstone(idx->h, &row, 1)) return gpu_fail(rc);
  idx->rows.erase(it);
  return 0;
}

// HybridIndex.DeleteBatch (hybrid_index.go:293-375): every id must exist, else nothing is deleted.
int qh_index_delete_batch(qh_index* idx, const char* const* ids, int64_t n) {
  if (!idx || (n > 0 && !ids)) return fail(QG_ERR_INVALID, "null argument");
  if (n <= 0) return 0;
  std::unique_lock<std::shared_mutex> lk(idx->mu);
  std::string missing;
  std::vector<int64_t> rows;
  rows.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    auto it = idx->rows.find(ids[i] ? ids[i] : "");
    if (it == idx->rows.end()) missing += (missing.empty() ? "" : " ") + std::string(ids[i] ? ids[i] : "");
    else rows.push_back(it->second);
  }
  if (!missing.empty()) return fail(QG_ERR_INVALID, "some vectors not found: [" + missing + "]");  // %v of []string
  if (int rc = flush_locked(idx)) return rc;
  if (int rc = qg_index_tombstone(idx->h, rows.data(), (int64_t)rows.size())) return gpu_fail(rc);  // one launch
  for (int64_t i = 0; i < n; ++i) idx->rows.erase(ids[i] ? ids[i] : "");
  return 0;
}

namespace {

// qg_index_compact + the id <-> row tables renumbered with its map (caller holds idx->mu exclusively).
int index_compact_locked(qh_index* idx, std::vector<int64_t>* map_out, int64_t* out_removed) {
  if (int rc = flush_locked(idx)) return rc;
  const int64_t n_old = qg_index_rows(idx->h);
  std::vector<int64_t> map((size_t)n_old);
  int64_t n_new = 0;
  if (int rc = qg_index_compact(idx->h, map.data(), &n_new)) return gpu_fail(rc);
  std::vector<std::string> ids((size_t)n_new);
  for (int64_t r = 0; r < n_old; ++r)
    if (map[(size_t)r] >= 0) ids[(size_t)map[(size_t)r]] = std::move(idx->ids[(size_t)r]);
  idx->ids.swap(ids);
  for (auto& kv : idx->rows) kv.second = map[(size_t)kv.second];
  if (out_removed) *out_removed = n_old - n_new;
  if (map_out) map_out->swap(map);
  return 0;
}

}  // namespace

int qh_index_compact(qh_index* idx, int64_t* out_removed) {
  if (!idx) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> lk(idx->mu);
  return index_compact_locked(idx, nullptr, out_removed);
}

int qh_collection_compact(qh_collection* c, int64_t* out_removed) {
  if (!c) return fail(QG_ERR_INVALID, "null argument");
  // lock order as in a search: the index lock, then the column mutex
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  std::unique_lock<std::shared_mutex> lk(c->index->mu);
  std::lock_guard<std::mutex> col_lock(c->col_mu);
  std::vector<int64_t> map;
  if (int rc = index_compact_locked(c->index, &map, out_removed)) return rc;
  std::vector<qh::ValuePtr> md(c->index->ids.size());
  std::vector<std::string> raw(c->index->ids.size());
  for (size_t r = 0; r < map.size() && r < c->metadata.size(); ++r)
    if (map[r] >= 0) {
      md[(size_t)map[r]] = std::move(c->metadata[r]);
      if (r < c->metadata_raw.size()) raw[(size_t)map[r]] = std::move(c->metadata_raw[r]);
    }
  c->metadata.swap(md);
  c->metadata_raw.swap(raw);
  c->epoch++;  // host-side encoded columns follow the new numbering on their next use
  return 0;
}

int64_t qh_index_size(const qh_index* idx) {
  if (!idx) return 0;
  std::shared_lock<std::shared_mutex> lk(idx->mu);
  return (int64_t)idx->rows.size();
}

int qh_index_search(qh_index* idx, const float* query, int dim, int k, qh_results** out) {
  return qh_index_batch_search(idx, query, 1, dim, k, nullptr, 0, 0.f, "", out);
}

int qh_index_batch_search(qh_index* idx, const float* queries, int nq, int dim, int k, const float* negatives,
                          int neg_dim, float negative_weight, const char* force_strategy, qh_results** out) {
  if (!idx || !out) return fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (nq <= 0) return fail(QG_ERR_INVALID, "no queries provided");  // hybrid_index.go:678-680
  const std::string strategy = force_strategy ? force_strategy : "";
  if (strategy == "hnsw")
    return fail(QG_ERR_UNSUPPORTED, "the GPU index serves the exact strategy; use the HNSW adapter for graph search");
  if (!strategy.empty() && strategy != "exact") return fail(QG_ERR_INVALID, "invalid search strategy: " + strategy);
  int flush_rc = 0;
  std::shared_lock<std::shared_mutex> lk = lock_flushed(idx, &flush_rc);
  if (flush_rc) return flush_rc;
  // SearchWithRequest order (hybrid_index.go:392-402): query dim, negative dim, k
  // vectorDim is 0 while the index holds no vector (set at the first Insert, reset when the last one is deleted,
  // hybrid_index.go:118-119, 282-284): an empty index takes any query and answers with no results
  const bool dim_known = !idx->rows.empty();
  if (dim_known && dim != idx->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  const bool neg_given = negatives != nullptr && neg_dim > 0;
  if (neg_given && dim_known && neg_dim != idx->dim)
    return fail(QG_ERR_DIM, "negative example dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(neg_dim));
  if (k <= 0) return fail(QG_ERR_K, "k must be positive");
  std::unique_ptr<qh_results> res(new qh_results());
  if (int rc = index_search_locked(idx, queries, nq, dim, k, neg_given ? negatives : nullptr, negative_weight, nullptr,
                                   0, res.get()))
    return rc;
  *out = res.release();
  return 0;
}

// ---- collection ------------------------------------------------------------------------------------
int qh_collection_create(qh_collection** out, const char* name, int dim, const char* distance, int device) {
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  qh_index* idx = nullptr;
  if (int rc = qh_index_create(&idx, dim, distance, 0, device)) return rc;
  qh_collection* c = new qh_collection();
  c->name = name ? name : "";
  c->dim = dim;
  c->index = idx;
  *out = c;
  return 0;
}

int qh_collection_destroy(qh_collection* c) {
  if (!c) return 0;
  c->filter_cache.clear();  // the handles belong to the index: gone before it
  qh_index_destroy(c->index);
  delete c;
  return 0;
}

int qh_collection_add_batch(qh_collection* c, const char* const* ids, const float* vecs, int64_t n, int dim,
                            const char* const* metadata_json) {
  if (!c || (n > 0 && (!ids || !vecs))) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  std::vector<qh::ValuePtr> parsed((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (!ids[i] || !ids[i][0]) return fail(QG_ERR_INVALID, "vector ID cannot be empty");  // collection.go:143-148
    if (dim != c->dim)
      return fail(QG_ERR_DIM, "invalid vector dimension: expected " + std::to_string(c->dim) + ", got " +
                                  std::to_string(dim));
    const char* md = metadata_json ? metadata_json[i] : nullptr;
    if (md && md[0]) {
      std::string err;
      qh::ValuePtr v = qh::parse_json(md, false, &err);
      // json.Unmarshal into map[string]interface{}: objects and null are accepted (collection.go:160-168)
      if (!v || (v->type != qh::Value::Object && v->type != qh::Value::Null))
        return fail(QG_ERR_INVALID, "invalid metadata format: " + (v ? std::string("not a JSON object") : err));
      if (v->type == qh::Value::Object) parsed[(size_t)i] = v;  // a null document decodes to a nil map: no fields
    }
    if (c->index->rows.count(ids[i])) return fail(QG_ERR_INVALID, std::string("vector with the same ID already exists: ") + ids[i]);
  }
  {
    std::unique_lock<std::shared_mutex> lk(c->index->mu);
    if (int rc = index_insert_locked(c->index, ids, vecs, n, dim)) return rc;
  }
  for (int64_t i = 0; i < n; ++i) {
    c->metadata.push_back(parsed[(size_t)i]);
    const char* md = metadata_json ? metadata_json[i] : nullptr;
    c->metadata_raw.push_back(md ? md : "");
  }
  c->epoch++;
  return 0;
}

int qh_collection_add(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json) {
  const char* ids[1] = {id};
  const char* mds[1] = {metadata_json};
  return qh_collection_add_batch(c, ids, vec, 1, dim, mds);
}

int qh_collection_delete(qh_collection* c, const char* id) {
  if (!c || !id) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  if (!c->index->rows.count(id)) return fail(QG_ERR_INVALID, "vector not found");  // ErrVectorNotFound
  return qh_index_delete(c->index, id);
}

// Collection.DeleteBatch (collection.go:375-414): the first missing id is reported, nothing is deleted.
int qh_collection_delete_batch(qh_collection* c, const char* const* ids, int64_t n) {
  if (!c || (n > 0 && !ids)) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  for (int64_t i = 0; i < n; ++i)
    if (!ids[i] || !c->index->rows.count(ids[i]))
      return fail(QG_ERR_INVALID, std::string("vector not found: ") + (ids[i] ? ids[i] : ""));
  return qh_index_delete_batch(c->index, ids, n);
}

// Collection.Update (collection.go:417-466): a new vector is Delete + Insert under the same id (the row moves
// to the end, its metadata with it); new metadata replaces the old document.
static int collection_update_locked(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json);
int qh_collection_update(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json) {
  if (!c || !id) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  return collection_update_locked(c, id, vec, dim, metadata_json);
}
// (caller holds c->cmu exclusively)
static int collection_update_locked(qh_collection* c, const char* id, const float* vec, int dim, const char* metadata_json) {
  auto it = c->index->rows.find(id);
  if (it == c->index->rows.end()) return fail(QG_ERR_INVALID, "vector not found");
  if (vec && dim != c->dim)
    return fail(QG_ERR_DIM, "invalid vector dimension: expected " + std::to_string(c->dim) + ", got " +
                                std::to_string(dim));
  const bool has_md = metadata_json && metadata_json[0];
  qh::ValuePtr parsed;
  if (has_md) {
    std::string err;
    qh::ValuePtr v = qh::parse_json(metadata_json, false, &err);
    if (!v || (v->type != qh::Value::Object && v->type != qh::Value::Null))
      return fail(QG_ERR_INVALID, "invalid metadata format: " + (v ? std::string("not a JSON object") : err));
    if (v->type == qh::Value::Object) parsed = v;
  }
  int64_t row = it->second;
  if (vec) {
    if (int rc = qh_index_delete(c->index, id)) return rc;
    {
      std::unique_lock<std::shared_mutex> lk(c->index->mu);
      const char* ids1[1] = {id};
      if (int rc = index_insert_locked(c->index, ids1, vec, 1, dim)) return rc;
    }
  }
  std::lock_guard<std::mutex> col_lock(c->col_mu);
  if (vec) {
    qh::ValuePtr moved = (size_t)row < c->metadata.size() ? c->metadata[(size_t)row] : nullptr;
    std::string moved_raw = (size_t)row < c->metadata_raw.size() ? c->metadata_raw[(size_t)row] : std::string();
    if ((size_t)row < c->metadata.size()) c->metadata[(size_t)row].reset();
    if ((size_t)row < c->metadata_raw.size()) c->metadata_raw[(size_t)row].clear();
    c->metadata.push_back(moved);
    c->metadata_raw.push_back(moved_raw);
    row = (int64_t)c->metadata.size() - 1;
  }
  if (has_md) {
    c->metadata[(size_t)row] = parsed;
    c->metadata_raw[(size_t)row] = metadata_json;
  }
  c->epoch++;
  return 0;
}

// Collection.UpdateBatch (collection.go:469-529): everything is validated first; then every vector is deleted
// and re-inserted under its id — here as one tombstone launch and one upload for the whole batch.
int qh_collection_update_batch(qh_collection* c, const char* const* ids, const float* vecs, int64_t n, int dim,
                               const char* const* metadata_json) {
  if (!c || (n > 0 && (!ids || !vecs))) return fail(QG_ERR_INVALID, "null argument");
  if (n <= 0) return fail(QG_ERR_INVALID, "no vectors provided for batch update");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  std::vector<qh::ValuePtr> parsed((size_t)n);
  std::vector<char> has_md((size_t)n, 0);
  std::unordered_map<std::string, int> seen;
  bool duplicates = false;
  for (int64_t i = 0; i < n; ++i) {
    if (!ids[i] || !ids[i][0]) return fail(QG_ERR_INVALID, "vector ID cannot be empty");
    const std::string id = ids[i];
    if (!c->index->rows.count(id)) return fail(QG_ERR_INVALID, "vector not found: " + id);
    if (dim != c->dim)
      return fail(QG_ERR_DIM, "invalid vector dimension for vector " + id + ": expected " + std::to_string(c->dim) +
                                  ", got " + std::to_string(dim));
    const char* md = metadata_json ? metadata_json[i] : nullptr;
    if (md && md[0]) {
      std::string err;
      qh::ValuePtr v = qh::parse_json(md, false, &err);
      if (!v || (v->type != qh::Value::Object && v->type != qh::Value::Null))
        return fail(QG_ERR_INVALID, "invalid metadata format for vector " + id + ": " +
                                        (v ? std::string("not a JSON object") : err));
      if (v->type == qh::Value::Object) parsed[(size_t)i] = v;
      has_md[(size_t)i] = 1;
    }
    duplicates = duplicates || !seen.emplace(id, 1).second;
  }
  if (duplicates) {  // the reference applies them one after the other; so do we
    for (int64_t i = 0; i < n; ++i)
      if (int rc = collection_update_locked(c, ids[i], vecs + (size_t)i * dim, dim, metadata_json ? metadata_json[i] : nullptr))
        return rc;
    return 0;
  }
  std::vector<int64_t> old_rows((size_t)n);
  for (int64_t i = 0; i < n; ++i) old_rows[(size_t)i] = c->index->rows[ids[i]];
  if (int rc = qh_index_delete_batch(c->index, ids, n)) return rc;
  {
    std::unique_lock<std::shared_mutex> lk(c->index->mu);
    if (int rc = index_insert_locked(c->index, ids, vecs, n, dim)) return rc;
  }
  std::lock_guard<std::mutex> col_lock(c->col_mu);
  for (int64_t i = 0; i < n; ++i) {
    const size_t r = (size_t)old_rows[(size_t)i];
    qh::ValuePtr md = has_md[(size_t)i] ? parsed[(size_t)i] : (r < c->metadata.size() ? c->metadata[r] : nullptr);
    std::string raw = has_md[(size_t)i] ? std::string(metadata_json[i])
                                         : (r < c->metadata_raw.size() ? c->metadata_raw[r] : std::string());
    if (r < c->metadata.size()) c->metadata[r].reset();
    if (r < c->metadata_raw.size()) c->metadata_raw[r].clear();
    c->metadata.push_back(md);
    c->metadata_raw.push_back(raw);
  }
  c->epoch++;
  return 0;
}

int64_t qh_collection_count(const qh_collection* c) { return c ? qh_index_size(c->index) : 0; }
int64_t qh_collection_rows(const qh_collection* c) {
  if (!c) return 0;
  std::shared_lock<std::shared_mutex> clk(c->cmu);
  return (int64_t)c->metadata.size();
}
const char* qh_collection_row_id(const qh_collection* c, int64_t row) {
  if (!c) return "";
  std::shared_lock<std::shared_mutex> clk(c->cmu);
  if (row < 0 || row >= (int64_t)c->index->ids.size()) return "";
  return c->index->ids[(size_t)row].c_str();
}

int qh_collection_set_facet_fields(qh_collection* c, const char* const* fields, int n) {
  if (!c) return fail(QG_ERR_INVALID, "null argument");
  std::unique_lock<std::shared_mutex> clk(c->cmu);
  c->facet_fields.clear();
  for (int i = 0; i < n; ++i) c->facet_fields.push_back(fields[i] ? fields[i] : "");
  c->epoch++;  // SetFacetFields re-indexes every row (collection.go:1111-1130)
  return 0;
}

static int collection_search(qh_collection* c, int which, const float* query, int dim, int k, const qh_filter* filters,
                             const qh_facet_filter* ffilters, int n_filters, qh_results** out) {
  if (!c || !out) return fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (which == 0) {
    // Collection.Search (collection.go:650-665): dimension, then top_k
    if (dim != c->dim)
      return fail(QG_ERR_DIM, "invalid vector dimension: expected " + std::to_string(c->dim) + ", got " +
                                  std::to_string(dim));
    if (k <= 0) return fail(QG_ERR_K, "top_k must be greater than 0");
  } else {
    // SearchWithFacets (collection.go:1146-1151): k, then dimension
    if (k <= 0) return fail(QG_ERR_K, "k must be positive");
    if (dim != c->dim)
      return fail(QG_ERR_DIM, "query vector dimension mismatch, expected " + std::to_string(c->dim) + ", got " +
                                  std::to_string(dim));
  }
  std::unique_ptr<qh_results> res(new qh_results());
  res->lists.assign(1, {});
  std::shared_lock<std::shared_mutex> clk(c->cmu);
  int flush_rc = 0;
  std::shared_lock<std::shared_mutex> lk = lock_flushed(c->index, &flush_rc);
  if (flush_rc) return flush_rc;
  if (c->index->rows.empty()) {  // empty index: no results, no error (collection.go:666-677, 1179-1182)
    *out = res.release();
    return 0;
  }
  if (n_filters <= 0) {
    if (int rc = index_search_locked(c->index, query, 1, dim, k, nullptr, 0.f, nullptr, 0, res.get())) return rc;
    *out = res.release();
    return 0;
  }
  qh::Program prog;
  if (int rc = build_program(c, which, filters, ffilters, n_filters, &prog)) return rc;
  FilterHandle fh;
  int64_t matches = 0;
  if (int rc = make_filter(c, prog, &fh, &matches)) return rc;
  if (int rc = index_search_locked(c->index, query, 1, dim, k, nullptr, 0.f, fh.f, matches, res.get())) return rc;
  *out = res.release();
  return 0;
}

int qh_collection_search(qh_collection* c, const float* query, int dim, int k, const qh_filter* filters, int n_filters,
                         qh_results** out) {
  return collection_search(c, 0, query, dim, k, filters, nullptr, n_filters, out);
}

// Collection.Search(types.SearchRequest) (collection.go:637-807): the filtered search above plus the
// SearchOptions decoration of every result (collection.go:758-779). Options.ExactSearch and NamespaceID are
// carried by the reference's request but not consulted by Collection.Search (only DB.BatchSearch groups by
// them, db.go:856-862); this index is always exact.
int qh_collection_search_request(qh_collection* c, const float* query, int dim, int k, const qh_filter* filters,
                                 int n_filters, const qh_search_options* opt, qh_results** out) {
  if (int rc = collection_search(c, 0, query, dim, k, filters, nullptr, n_filters, out)) return rc;
  if (!opt || (!opt->include_vectors && !opt->include_metadata)) return 0;
  qh_results* res = *out;
  std::shared_lock<std::shared_mutex> clk(c->cmu);
  int flush_rc = 0;
  std::shared_lock<std::shared_mutex> lk = lock_flushed(c->index, &flush_rc);
  if (flush_rc) return flush_rc;
  std::vector<Hit>& list = res->lists[0];
  std::vector<int64_t> rows;
  std::vector<size_t> which;
  for (size_t j = 0; j < list.size(); ++j) {
    auto it = c->index->rows.find(list[j].id);
    if (it == c->index->rows.end()) continue;  // deleted since the search: the reference's map lookup would miss too
    if (opt->include_metadata && (size_t)it->second < c->metadata_raw.size() && !c->metadata_raw[(size_t)it->second].empty()) {
      list[j].metadata = c->metadata_raw[(size_t)it->second];
      list[j].has_metadata = true;
    }
    if (opt->include_vectors) {
      rows.push_back(it->second);
      which.push_back(j);
    }
  }
  if (!rows.empty()) {
    std::vector<float> buf(rows.size() * (size_t)c->dim);
    if (int rc = qg_index_fetch(c->index->h, rows.data(), (int64_t)rows.size(), buf.data())) return gpu_fail(rc);
    for (size_t t = 0; t < rows.size(); ++t)
      list[which[t]].vector.assign(buf.begin() + (long)(t * (size_t)c->dim), buf.begin() + (long)((t + 1) * (size_t)c->dim));
  }
  return 0;
}

// persistence.Collection.Search / SearchWithFacets (pkg/persistence/collection.go:226-261, 327-378): the same
// prefilter path under that type's argument checks — "query vector is nil", "query vector dimension mismatch:
// got %d, expected %d" (got first), limit <= 0 returns every row, vectors without facet values never pass a
// filter. The reference orders by Distance only (an O(n^2) selection sort over Go's map order, so ties come in
// any order); here ties are ordered by row.
int qh_collection_persistence_search(qh_collection* c, const float* query, int dim, int limit,
                                     const qh_facet_filter* filters, int n_filters, qh_results** out) {
  if (!c || !out) return fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (!query) return fail(QG_ERR_INVALID, "query vector is nil");
  if (dim != c->dim)
    return fail(QG_ERR_DIM, "query vector dimension mismatch: got " + std::to_string(dim) + ", expected " +
                                std::to_string(c->dim));
  int64_t count = qh_collection_count(c);
  if (count == 0) {
    std::unique_ptr<qh_results> res(new qh_results());
    res->lists.assign(1, {});
    *out = res.release();
    return 0;
  }
  const int k = (limit > 0 && (int64_t)limit < count) ? limit : (int)std::min<int64_t>(count, 0x7fffffff);
  if (n_filters <= 0) return collection_search(c, 0, query, dim, k, nullptr, nullptr, 0, out);
  return collection_search(c, 1, query, dim, k, nullptr, filters, n_filters, out);
}

int qh_collection_search_with_facets(qh_collection* c, const float* query, int dim, int k,
                                     const qh_facet_filter* filters, int n_filters, qh_results** out) {
  return collection_search(c, 1, query, dim, k, nullptr, filters, n_filters, out);
}

int qh_collection_filter_mask(qh_collection* c, int which, const qh_filter* filters, const qh_facet_filter* ffilters,
                              int n_filters, uint8_t* mask_out, int64_t n_rows) {
  if (!c || !mask_out) return fail(QG_ERR_INVALID, "null argument");
  std::shared_lock<std::shared_mutex> clk(c->cmu);
  const int64_t rows = (int64_t)c->metadata.size();
  if (n_rows < rows) return fail(QG_ERR_INVALID, "mask buffer too small");
  int flush_rc = 0;
  std::shared_lock<std::shared_mutex> lk = lock_flushed(c->index, &flush_rc);
  if (flush_rc) return flush_rc;
  qh::Program prog;
  if (int rc = build_program(c, which, filters, ffilters, n_filters, &prog)) return rc;
  FilterHandle fh;
  if (int rc = make_filter(c, prog, &fh, nullptr)) return rc;
  std::vector<uint64_t> words((size_t)((rows + 63) / 64) + 1);
  int64_t matches = 0;
  if (int rc = qg_filter_eval(c->index->h, fh.f, words.data(), &matches)) return gpu_fail(rc);
  for (int64_t r = 0; r < rows; ++r) mask_out[r] = (uint8_t)((words[(size_t)(r >> 6)] >> (r & 63)) & 1ull);
  return 0;
}

int qh_debug_sprint_v(const char* value_json, int typed_literals, char* buf, int buf_len) {
  std::string err;
  qh::ValuePtr v = qh::parse_json(value_json ? value_json : "", typed_literals != 0, &err);
  if (!v) {
    g_err = err;
    return -1;
  }
  const std::string s = qh::sprint_v(*v);
  if (buf && buf_len > 0) {
    const size_t n = std::min<size_t>(s.size(), (size_t)buf_len - 1);
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int)s.size();
}

int qh_debug_equal_fold(const char* a, const char* b) {
  return qh::fold_key(a ? a : "") == qh::fold_key(b ? b : "") ? 1 : 0;
}

}  // extern "C"

// ---- internals shared with hnsw_walk.cpp -------------------------------------------------------------
extern "C" {
// The HNSW walks read device rows and the id table for as long as they run: they hold the index's shared lock
// (taken once nothing is pending, like every other reader) through this guard, so an Insert that grows the
// arrays or a compaction cannot free them under kernels in flight. *guard is released with
// qh_internal_index_unlock; the two lookups below must only be called while it is held.
int qh_internal_index_lock(qh_index* idx, qg_index** h, int* dim, void** guard) {
  *guard = nullptr;
  int flush_rc = 0;
  std::shared_lock<std::shared_mutex> lk = lock_flushed(idx, &flush_rc);
  if (flush_rc) return flush_rc;
  *h = idx->h;
  *dim = idx->dim;
  *guard = new std::shared_lock<std::shared_mutex>(std::move(lk));
  return 0;
}
void qh_internal_index_unlock(void* guard) { delete static_cast<std::shared_lock<std::shared_mutex>*>(guard); }
const char* qh_internal_row_id(qh_index* idx, int64_t row) {
  return (row >= 0 && row < (int64_t)idx->ids.size()) ? idx->ids[(size_t)row].c_str() : "";
}
int64_t qh_internal_id_row(qh_index* idx, const char* id) {
  auto it = idx->rows.find(id ? id : "");
  return it == idx->rows.end() ? -1 : it->second;
}
int qh_internal_fail(int code, const char* msg) { return fail(code, msg ? msg : ""); }
qh_results* qh_internal_results_new(int nq) {
  qh_results* r = new qh_results();
  r->lists.assign((size_t)nq, {});
  return r;
}
void qh_internal_results_push(qh_results* r, int q, const char* id, float dist) {
  r->lists[(size_t)q].push_back(Hit{id, dist, {}, {}, false});
}
}
