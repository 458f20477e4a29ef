// hnsw_walk.cpp — hnsw.Search (pkg/hnsw/hnsw.go:602-713) over the reference's graph with the
// per-neighbour distance calls of searchLayer (hnsw.go:536-563) batched on the GPU: all queries
// of a batch advance in lock step, one expansion step (<= MaxM0 neighbours per query) per
// qg_batch_distance_queries call. Heaps, visit order, stop and admit rules are the reference's,
// so with bit-identical distances the walk is step-identical (tests/test_gpu_hnsw.py).
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../include/quiver_gpu.h"
#include "../../include/quiver_host.h"

namespace qh_hnsw {

struct Res {
  uint32_t idx;
  float dist;
};

// hnsw.go:101-144 (min-heap) and :153-196 (max-heap): same sift rules, so equal distances leave
// the heaps in the same shape as the reference.
struct MinHeap {
  std::vector<Res> a;
  void push(Res x) {
    a.push_back(x);
    int j = (int)a.size() - 1;
    for (;;) {
      int i = (j - 1) / 2;
      if (i == j || a[j].dist >= a[i].dist) break;
      std::swap(a[i], a[j]);
      j = i;
    }
  }
  Res pop() {
    int n = (int)a.size() - 1;
    std::swap(a[0], a[n]);
    int i = 0;
    for (;;) {
      int j1 = 2 * i + 1;
      if (j1 >= n || j1 < 0) break;
      int j = j1, j2 = j1 + 1;
      if (j2 < n && a[j2].dist < a[j1].dist) j = j2;
      if (a[i].dist <= a[j].dist) break;
      std::swap(a[i], a[j]);
      i = j;
    }
    Res r = a[n];
    a.pop_back();
    return r;
  }
};
struct MaxHeap {
  std::vector<Res> a;
  void push(Res x) {
    a.push_back(x);
    int j = (int)a.size() - 1;
    for (;;) {
      int i = (j - 1) / 2;
      if (i == j || a[j].dist <= a[i].dist) break;
      std::swap(a[i], a[j]);
      j = i;
    }
  }
  Res pop() {
    int n = (int)a.size() - 1;
    std::swap(a[0], a[n]);
    int i = 0;
    for (;;) {
      int j1 = 2 * i + 1;
      if (j1 >= n || j1 < 0) break;
      int j = j1, j2 = j1 + 1;
      if (j2 < n && a[j2].dist > a[j1].dist) j = j2;
      if (a[i].dist >= a[j].dist) break;
      std::swap(a[i], a[j]);
      i = j;
    }
    Res r = a[n];
    a.pop_back();
    return r;
  }
};

enum Stage { ENTRY_DISTANCE, LAYER_START, EXPAND, DONE };

// The walks of a batch are independent between two distance rounds: their heap work (the part the
// GPU does not take over) is spread over the host cores by a small persistent pool.
class WalkPool {
 public:
  explicit WalkPool(int n_threads) {
    for (int t = 1; t < n_threads; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~WalkPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
  }
  // fn(i) for every i in [0, n), in chunks claimed from a shared counter; returns when all are done.
  void run(int n, const std::function<void(int)>& fn) {
    if (workers_.empty() || n < 256) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      busy_ = (int)workers_.size();
      ++gen_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return busy_ == 0; });
    fn_ = nullptr;
  }

 private:
  void work() {
    const int chunk = 64;
    for (;;) {
      const int i0 = next_.fetch_add(chunk);
      if (i0 >= n_) break;
      const int i1 = std::min(n_, i0 + chunk);
      for (int i = i0; i < i1; ++i) (*fn_)(i);
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--busy_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, busy_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

struct Walk {
  Stage stage = ENTRY_DISTANCE;
  int level = 0, ef = 1;
  uint32_t entry = 0;
  MinHeap cand;
  MaxHeap res;
  std::vector<uint64_t> visited;   // bitset over nodes
  std::vector<uint32_t> touched;   // words to clear between layers
  std::vector<uint32_t> pending;   // rows whose distances were requested this round
  std::vector<Res> layer_result;   // ascending
  int64_t evals = 0;

  bool test_and_set(uint32_t id) {
    uint64_t& w = visited[id >> 6];
    const uint64_t bit = 1ull << (id & 63);
    if (w & bit) return true;
    if (w == 0) touched.push_back(id >> 6);
    w |= bit;
    return false;
  }
  void clear_visited() {
    for (uint32_t wi : touched) visited[wi] = 0;
    touched.clear();
  }
};

}  // namespace qh_hnsw

// qh_index internals needed here (defined in host.cpp)
extern "C" int qh_internal_index_lock(qh_index* idx, qg_index** h, int* dim, void** guard);
extern "C" void qh_internal_index_unlock(void* guard);
namespace {
// shared lock of the index for the duration of a walk (host.cpp: qh_internal_index_lock)
struct IndexReadGuard {
  void* g = nullptr;
  ~IndexReadGuard() { if (g) qh_internal_index_unlock(g); }
};
}  // namespace
extern "C" const char* qh_internal_row_id(qh_index* idx, int64_t row);
extern "C" int64_t qh_internal_id_row(qh_index* idx, const char* id);
extern "C" int qh_internal_fail(int code, const char* msg);
extern "C" qh_results* qh_internal_results_new(int nq);
extern "C" void qh_internal_results_push(qh_results* r, int q, const char* id, float dist);

// the lock-step walk; the caller holds the index's shared lock (IndexReadGuard) and passes the device handle
static int hnsw_search_batch_held(qh_index* idx, qg_index* h, int idim, const qh_hnsw_graph* g, const float* queries,
                                  int nq, int dim, int k, qh_results** out, int64_t* out_evals, int64_t* out_steps);

extern "C" int qh_hnsw_search_batch(qh_index* idx, const qh_hnsw_graph* g, const float* queries, int nq, int dim,
                                    int k, qh_results** out, int64_t* out_evals, int64_t* out_steps) {
  if (!idx || !g || !out) return qh_internal_fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  qg_index* h = nullptr;
  int idim = 0;
  IndexReadGuard guard;
  if (int rc = qh_internal_index_lock(idx, &h, &idim, &guard.g)) return rc;  // also uploads write-combined Inserts
  return hnsw_search_batch_held(idx, h, idim, g, queries, nq, dim, k, out, out_evals, out_steps);
}

static int hnsw_search_batch_held(qh_index* idx, qg_index* h, int idim, const qh_hnsw_graph* g, const float* queries,
                                  int nq, int dim, int k, qh_results** out, int64_t* out_evals, int64_t* out_steps) {
  using namespace qh_hnsw;
  *out = nullptr;
  if (nq <= 0 || !queries) return qh_internal_fail(QG_ERR_INVALID, "no queries provided");
  if (dim != idim) {
    const std::string msg = "query dimension mismatch: expected " + std::to_string(idim) + ", got " + std::to_string(dim);
    return qh_internal_fail(QG_ERR_DIM, msg.c_str());
  }
  std::unique_ptr<qh_results, int (*)(qh_results*)> res(qh_internal_results_new(nq), qh_results_free);
  const int64_t n = g->n_nodes;
  if (n == 0) {  // hnsw.go:606-608: empty graph => no results, no error
    *out = res.release();
    return 0;
  }
  if (k <= 0) return qh_internal_fail(QG_ERR_K, "k must be positive");
  const int kk = (int)std::min<int64_t>(k, n);
  // entry point validation (hnsw.go:620-634)
  uint32_t entry = (uint32_t)g->entry_point;
  if ((int64_t)entry >= n || g->level[entry] < 0) {
    entry = 0xFFFFFFFFu;
    for (int64_t i = 0; i < n; ++i)
      if (g->level[i] >= 0) { entry = (uint32_t)i; break; }
    if (entry == 0xFFFFFFFFu) {
      *out = res.release();
      return 0;
    }
  }
  const int ef0 = std::max(g->ef_search, kk);  // hnsw.go:660-663
  const int m = g->max_m0;                     // rows per query per round

  qg_queries* qs = nullptr;
  if (int rc = qg_queries_upload(h, queries, nq, dim, &qs)) return qh_internal_fail(rc, qg_last_error());
  std::vector<Walk> walks((size_t)nq);
  for (Walk& w : walks) {
    w.visited.assign((size_t)((n + 63) / 64), 0);
    w.entry = entry;
    w.level = g->current_level;
  }
  std::vector<uint32_t> rows((size_t)nq * m);
  std::vector<float> dist((size_t)nq * m);
  int64_t steps = 0;

  auto connections = [&](uint32_t node, int level, const uint32_t** lst, int* cnt) {
    if (level == 0) {
      *lst = g->adj0 + (size_t)node * g->max_m0;
      *cnt = g->max_m0;
    } else {
      *lst = g->upper_adj + g->upper_off[node] + (size_t)(level - 1) * g->m;
      *cnt = g->m;
    }
  };
  // Advance a walk until it needs distances (fills w.pending) or is done.
  auto advance = [&](Walk& w) {
    w.pending.clear();
    for (;;) {
      if (w.stage == DONE) return;
      if (w.stage == ENTRY_DISTANCE) {  // entryDistance, hnsw.go:637 (its value is not used further)
        w.pending.push_back(w.entry);
        return;
      }
      if (w.stage == LAYER_START) {     // searchLayer prologue, hnsw.go:483-505
        w.ef = w.level > 0 ? 1 : ef0;
        w.clear_visited();
        w.test_and_set(w.entry);
        w.cand.a.clear();
        w.res.a.clear();
        w.pending.push_back(w.entry);
        return;
      }
      // EXPAND: pop candidates (hnsw.go:508-563) until one has unvisited neighbours
      bool finished = true;
      while (!w.cand.a.empty()) {
        const Res cur = w.cand.pop();
        if ((int)w.res.a.size() >= w.ef && cur.dist > w.res.a[0].dist) break;
        if (g->level[cur.idx] < 0 || w.level > g->level[cur.idx]) continue;
        const uint32_t* lst;
        int cnt;
        connections(cur.idx, w.level, &lst, &cnt);
        for (int c = 0; c < cnt; ++c) {
          const uint32_t id = lst[c];
          if (id == 0xFFFFFFFFu) break;  // end of the list
          if ((int64_t)id >= n || g->level[id] < 0) continue;
          if (!w.test_and_set(id)) w.pending.push_back(id);
        }
        if (!w.pending.empty()) { finished = false; break; }
      }
      if (!finished) return;
      // layer finished: drain the max-heap into an ascending list (hnsw.go:567-577)
      const int rn = (int)w.res.a.size();
      w.layer_result.resize((size_t)rn);
      for (int i = rn - 1; i >= 0; --i) w.layer_result[(size_t)i] = w.res.pop();
      if (w.level > 0) {
        if (rn > 0) w.entry = w.layer_result[0].idx;  // hnsw.go:650-656
        w.level--;
        w.stage = LAYER_START;
        continue;
      }
      w.stage = DONE;
      return;
    }
  };
  // Feed the distances of the pending rows back, in connection order (hnsw.go:547-560).
  auto consume = [&](Walk& w, const float* d) {
    w.evals += (int64_t)w.pending.size();
    if (w.stage == ENTRY_DISTANCE) {
      w.stage = LAYER_START;
      if (w.level <= 0) w.level = 0;
      return;
    }
    if (w.stage == LAYER_START) {
      const Res e{w.entry, d[0]};
      w.cand.push(e);
      w.res.push(e);
      w.stage = EXPAND;
      return;
    }
    for (size_t j = 0; j < w.pending.size(); ++j) {
      const float cd = d[j];
      if ((int)w.res.a.size() < w.ef || cd < w.res.a[0].dist) {
        const Res r{w.pending[j], cd};
        w.cand.push(r);
        w.res.push(r);
        if ((int)w.res.a.size() > w.ef) w.res.pop();
      }
    }
  };

  int rc = 0;
  const int n_threads = (int)std::max(1u, std::min({std::thread::hardware_concurrency(), 32u, (unsigned)(nq / 256 + 1)}));
  WalkPool pool(n_threads);
  std::atomic<int> any{0};
  const std::function<void(int)> do_advance = [&](int i) {
    Walk& w = walks[(size_t)i];
    uint32_t* slot = rows.data() + (size_t)i * m;
    advance(w);
    const size_t np = w.stage == DONE ? 0 : w.pending.size();
    std::copy(w.pending.begin(), w.pending.begin() + (long)np, slot);
    std::fill(slot + np, slot + m, 0xFFFFFFFFu);
    if (w.stage != DONE) any.store(1, std::memory_order_relaxed);
  };
  const std::function<void(int)> do_consume = [&](int i) {
    Walk& w = walks[(size_t)i];
    if (w.stage != DONE && !w.pending.empty()) consume(w, dist.data() + (size_t)i * m);
  };
  for (;;) {
    any.store(0);
    pool.run(nq, do_advance);
    if (!any.load()) break;
    ++steps;
    if ((rc = qg_batch_distance_queries(h, qs, rows.data(), m, dist.data()))) break;
    pool.run(nq, do_consume);
  }
  qg_queries_destroy(qs);
  if (rc) return qh_internal_fail(rc, qg_last_error());

  // truncate to k (hnsw.go:670-672); under-filled queries get the exact pass (hnsw.go:676-710)
  std::vector<int> underfilled;
  for (int i = 0; i < nq; ++i) {
    Walk& w = walks[(size_t)i];
    const int cnt = std::min<int>((int)w.layer_result.size(), kk);
    if (cnt < kk) { underfilled.push_back(i); continue; }
    for (int j = 0; j < cnt; ++j)
      qh_internal_results_push(res.get(), i, qh_internal_row_id(idx, w.layer_result[(size_t)j].idx), w.layer_result[(size_t)j].dist);
  }
  if (!underfilled.empty()) {
    const int nu = (int)underfilled.size();
    std::vector<float> uq((size_t)nu * dim), ud((size_t)nu * kk);
    std::vector<int64_t> ur((size_t)nu * kk);
    std::vector<int> uc((size_t)nu);
    for (int u = 0; u < nu; ++u)
      std::memcpy(uq.data() + (size_t)u * dim, queries + (size_t)underfilled[(size_t)u] * dim, (size_t)dim * 4);
    if (int erc = qg_search_batch(h, uq.data(), nu, dim, kk, nullptr, nullptr, ud.data(), nullptr, ur.data(), uc.data()))
      return qh_internal_fail(erc, qg_last_error());
    for (int u = 0; u < nu; ++u) {
      // sort by (Distance, VectorID) like the reference's supplement (hnsw.go:699-704)
      std::vector<std::pair<float, std::string>> lst;
      for (int j = 0; j < uc[(size_t)u]; ++j)
        lst.emplace_back(ud[(size_t)u * kk + j], qh_internal_row_id(idx, ur[(size_t)u * kk + j]));
      std::stable_sort(lst.begin(), lst.end(), [](const auto& a, const auto& b) {
        if (a.first == b.first) return a.second < b.second;
        return a.first < b.first;
      });
      for (const auto& e : lst) qh_internal_results_push(res.get(), underfilled[(size_t)u], e.second.c_str(), e.first);
    }
  }
  if (out_evals)
    for (int i = 0; i < nq; ++i) out_evals[i] = walks[(size_t)i].evals;
  if (out_steps) *out_steps = steps;
  *out = res.release();
  return 0;
}


// ================================================================================================
// The same search with the WHOLE walk on the device (qg_hnsw_search_batch: one persistent kernel, a
// warp per query, csrc/hnsw.cu). The host keeps what needs string ids: the under-fill exact pass
// ordered by (Distance, VectorID) (hnsw.go:676-710) and the id lookup of the results.
// ================================================================================================
struct qh_hnsw_dev {
  qh_index* owner = nullptr;
  qg_hnsw* dev = nullptr;
  qh_hnsw_graph host{};          // the caller's arrays (kept alive by the caller): the lock-step fallback uses them
  int64_t n_nodes = 0;
  int ef_search = 0;
  bool has_entry = false;
};

extern "C" int qh_hnsw_upload(qh_index* idx, const qh_hnsw_graph* g, qh_hnsw_dev** out) {
  if (!idx || !g || !out) return qh_internal_fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  qg_index* h = nullptr;
  int idim = 0;
  IndexReadGuard guard;
  if (int rc = qh_internal_index_lock(idx, &h, &idim, &guard.g)) return rc;
  std::unique_ptr<qh_hnsw_dev> d(new qh_hnsw_dev());
  d->owner = idx;
  d->host = *g;
  d->n_nodes = g->n_nodes;
  d->ef_search = g->ef_search;
  // entry point validation (hnsw.go:620-634): a stale entry is replaced by the first live node
  int64_t entry = g->entry_point;
  if (g->n_nodes > 0 && (entry < 0 || entry >= g->n_nodes || g->level[entry] < 0)) {
    entry = -1;
    for (int64_t i = 0; i < g->n_nodes; ++i)
      if (g->level[i] >= 0) { entry = i; break; }
  }
  d->has_entry = g->n_nodes > 0 && entry >= 0;
  if (d->has_entry) {
    if (int rc = qg_hnsw_upload(h, g->n_nodes, g->m, g->max_m0, (int)entry, g->current_level, g->level, g->adj0,
                                g->upper_off, g->upper_adj, &d->dev))
      return qh_internal_fail(rc, qg_last_error());
  }
  *out = d.release();
  return 0;
}

extern "C" int qh_hnsw_dev_free(qh_hnsw_dev* d) {
  if (!d) return 0;
  if (d->dev) qg_hnsw_destroy(d->dev);
  delete d;
  return 0;
}

extern "C" int qh_hnsw_search_device(qh_index* idx, qh_hnsw_dev* d, const float* queries, int nq, int dim, int k,
                                     qh_results** out, int64_t* out_evals, int* out_fallbacks) {
  if (!idx || !d || !out) return qh_internal_fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (out_fallbacks) *out_fallbacks = 0;
  if (d->owner != idx) return qh_internal_fail(QG_ERR_INVALID, "graph does not belong to this index");
  qg_index* h = nullptr;
  int idim = 0;
  IndexReadGuard guard;
  if (int rc = qh_internal_index_lock(idx, &h, &idim, &guard.g)) return rc;
  if (nq <= 0 || !queries) return qh_internal_fail(QG_ERR_INVALID, "no queries provided");
  if (dim != idim) {
    const std::string msg = "query dimension mismatch: expected " + std::to_string(idim) + ", got " + std::to_string(dim);
    return qh_internal_fail(QG_ERR_DIM, msg.c_str());
  }
  std::unique_ptr<qh_results, int (*)(qh_results*)> res(qh_internal_results_new(nq), qh_results_free);
  if (d->n_nodes == 0 || !d->has_entry) {  // hnsw.go:606-608, 628-631: nothing to search
    *out = res.release();
    return 0;
  }
  if (k <= 0) return qh_internal_fail(QG_ERR_K, "k must be positive");
  const int kk = (int)std::min<int64_t>(k, d->n_nodes);
  std::vector<uint32_t> ridx((size_t)nq * kk);
  std::vector<float> rdist((size_t)nq * kk);
  std::vector<int> rcnt((size_t)nq);
  std::vector<int64_t> evals((size_t)nq, 0);
  if (int rc = qg_hnsw_search_batch(h, d->dev, queries, nq, dim, kk, d->ef_search, ridx.data(), rdist.data(), rcnt.data(),
                                    evals.data()))
    return qh_internal_fail(rc, qg_last_error());
  // queries whose candidate heap outgrew the kernel's shared-memory slice: the lock-step host walk
  std::vector<int> redo;
  for (int i = 0; i < nq; ++i)
    if (rcnt[(size_t)i] < 0) redo.push_back(i);
  std::vector<std::vector<std::pair<std::string, float>>> redone(redo.size());
  if (!redo.empty()) {
    std::vector<float> rq(redo.size() * (size_t)dim);
    for (size_t u = 0; u < redo.size(); ++u)
      std::memcpy(rq.data() + u * dim, queries + (size_t)redo[u] * dim, (size_t)dim * 4);
    qh_results* rr = nullptr;
    std::vector<int64_t> rev(redo.size());
    if (int rc = hnsw_search_batch_held(idx, h, idim, &d->host, rq.data(), (int)redo.size(), dim, k, &rr, rev.data(), nullptr)) return rc;
    for (size_t u = 0; u < redo.size(); ++u) {
      const int cnt = qh_results_count(rr, (int)u);
      for (int j = 0; j < cnt; ++j) redone[u].emplace_back(qh_results_id(rr, (int)u, j), qh_results_distance(rr, (int)u, j));
      evals[(size_t)redo[u]] = rev[u];
    }
    qh_results_free(rr);
    if (out_fallbacks) *out_fallbacks = (int)redo.size();
  }
  // under-filled walks get the exact pass (hnsw.go:676-710), all of them in one batched search
  std::vector<int> underfilled;
  for (int i = 0; i < nq; ++i)
    if (rcnt[(size_t)i] >= 0 && rcnt[(size_t)i] < kk) underfilled.push_back(i);
  std::vector<float> ud;
  std::vector<int64_t> ur;
  std::vector<int> uc;
  if (!underfilled.empty()) {
    const int nu = (int)underfilled.size();
    std::vector<float> uq((size_t)nu * dim);
    ud.resize((size_t)nu * kk);
    ur.resize((size_t)nu * kk);
    uc.resize((size_t)nu);
    for (int u = 0; u < nu; ++u)
      std::memcpy(uq.data() + (size_t)u * dim, queries + (size_t)underfilled[(size_t)u] * dim, (size_t)dim * 4);
    if (int erc = qg_search_batch(h, uq.data(), nu, dim, kk, nullptr, nullptr, ud.data(), nullptr, ur.data(), uc.data()))
      return qh_internal_fail(erc, qg_last_error());
  }
  size_t next_redo = 0, next_under = 0;
  for (int i = 0; i < nq; ++i) {
    if (rcnt[(size_t)i] < 0) {
      for (const auto& e : redone[next_redo]) qh_internal_results_push(res.get(), i, e.first.c_str(), e.second);
      ++next_redo;
    } else if (rcnt[(size_t)i] < kk) {
      const size_t u = next_under++;
      std::vector<std::pair<float, std::string>> lst;  // (Distance, VectorID) order, hnsw.go:699-704
      for (int j = 0; j < uc[u]; ++j) lst.emplace_back(ud[u * kk + j], qh_internal_row_id(idx, ur[u * kk + j]));
      std::stable_sort(lst.begin(), lst.end(), [](const auto& a, const auto& b) {
        if (a.first == b.first) return a.second < b.second;
        return a.first < b.first;
      });
      for (const auto& e : lst) qh_internal_results_push(res.get(), i, e.second.c_str(), e.first);
    } else {
      for (int j = 0; j < kk; ++j)
        qh_internal_results_push(res.get(), i, qh_internal_row_id(idx, ridx[(size_t)i * kk + j]), rdist[(size_t)i * kk + j]);
    }
  }
  if (out_evals) std::memcpy(out_evals, evals.data(), (size_t)nq * 8);
  *out = res.release();
  return 0;
}


extern "C" int64_t qh_index_size(const qh_index* idx);

// HNSWAdapter.SearchWithNegativeExample (pkg/hnsw/adapter.go:345-437).
extern "C" int qh_hnsw_search_negative(qh_index* idx, qh_hnsw_dev* d, const float* query, int dim, const float* negative,
                                       int neg_dim, float negative_weight, int k, qh_results** out) {
  if (!idx || !d || !out) return qh_internal_fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (k <= 0) return qh_internal_fail(QG_ERR_K, "k must be positive");  // adapter.go:347-349
  int64_t retrieve = std::max(2 * (int64_t)k, (int64_t)30);               // :353-356
  const int64_t size = qh_index_size(idx);
  if (retrieve > size) retrieve = size;
  qh_results* initial = nullptr;
  if (retrieve <= 0) {  // empty index: Search returns no results
    *out = qh_internal_results_new(1);
    return 0;
  }
  if (int rc = qh_hnsw_search_device(idx, d, query, 1, dim, (int)retrieve, &initial, nullptr, nullptr)) {
    const std::string msg = std::string("initial search failed: ") + qh_last_error();  // :360-362
    return qh_internal_fail(rc, msg.c_str());
  }
  std::unique_ptr<qh_results, int (*)(qh_results*)> init(initial, qh_results_free);
  const int n0 = qh_results_count(initial, 0);
  std::unique_ptr<qh_results, int (*)(qh_results*)> res(qh_internal_results_new(1), qh_results_free);
  if (!negative || neg_dim == 0 || negative_weight <= 0.f || n0 <= k) {  // :366-372
    for (int j = 0; j < std::min(n0, k); ++j)
      qh_internal_results_push(res.get(), 0, qh_results_id(initial, 0, j), qh_results_distance(initial, 0, j));
    *out = res.release();
    return 0;
  }
  if (negative_weight > 1.0f) negative_weight = 1.0f;  // :375-377
  struct Ext { std::string id; float dist; };
  std::vector<Ext> ext;
  if (neg_dim == dim) {  // a dimension mismatch makes DistanceFunc fail for every candidate: all are skipped (:404-406)
    qg_index* h = nullptr;
    int idim = 0;
    IndexReadGuard guard;
    if (int rc = qh_internal_index_lock(idx, &h, &idim, &guard.g)) return rc;
    std::vector<uint32_t> rows((size_t)n0);
    for (int j = 0; j < n0; ++j) {
      const int64_t r = qh_internal_id_row(idx, qh_results_id(initial, 0, j));
      rows[(size_t)j] = r < 0 ? 0xFFFFFFFFu : (uint32_t)r;
    }
    std::vector<float> nd((size_t)n0);
    if (int rc = qg_batch_distance(h, negative, dim, rows.data(), n0, nd.data())) return qh_internal_fail(rc, qg_last_error());
    for (int j = 0; j < n0; ++j) {
      if (rows[(size_t)j] == 0xFFFFFFFFu) continue;  // id no longer in the index (:386-390)
      // Distance - (negativeWeight * negDistance) in float32 (:415); this file is compiled with -ffp-contract=off
      const float prod = negative_weight * nd[(size_t)j];
      ext.push_back(Ext{qh_results_id(initial, 0, j), qh_results_distance(initial, 0, j) - prod});
    }
  }
  std::stable_sort(ext.begin(), ext.end(), [](const Ext& a, const Ext& b) {  // :418-423
    if (a.dist == b.dist) return a.id < b.id;
    return a.dist < b.dist;
  });
  for (int j = 0; j < std::min<int>(k, (int)ext.size()); ++j) qh_internal_results_push(res.get(), 0, ext[(size_t)j].id.c_str(), ext[(size_t)j].dist);
  *out = res.release();
  return 0;
}
