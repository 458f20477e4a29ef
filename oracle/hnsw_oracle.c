/*
 * oracle/hnsw_oracle.c — CPU restatement of Quiver's HNSW (pkg/hnsw/hnsw.go). TEST INFRASTRUCTURE.
 *
 * Restated line by line: the two binary heaps with the reference's own sift rules
 * (hnsw.go:100-196 — equal distances are resolved by heap shape, so the rules matter),
 * searchLayer (:471-580: visited marked BEFORE the distance call, stop on strict `>`, admit on
 * strict `<`), selectNeighbors (:583-599: plain closest-k, ties by smaller index — not the
 * paper's diversity heuristic), Insert/connectNode (:266-468, including the prune of an
 * over-full neighbour list) and Search (:602-713: ef=1 descent, base layer ef = max(EfSearch, k),
 * truncate to k; the under-fill exact pass is restated in the Python wrapper).
 * One deliberate difference: randomLevel (:716-738) draws from Go's math/rand, whose stream is
 * not reproduced; levels come from the counter-based hash of oracle/synth.h with the same law
 * (p = 0.25 per level, at most min(MaxLevel, 10) promotions).
 * Distances are the reference's (quiver_oracle.c: qo_distance).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#include "synth.h"

float qo_distance(int metric, int arith, const float* a, const float* b, int d);

typedef struct {
  uint32_t idx;
  float dist;
} qh_res;

typedef struct {
  qh_res* a;
  int n, cap;
} qh_heap;

static void heap_reserve(qh_heap* h, int need) {
  if (need > h->cap) {
    h->cap = need * 2 + 16;
    h->a = (qh_res*)realloc(h->a, (size_t)h->cap * sizeof(qh_res));
  }
}

/* min-heap: hnsw.go:101-144 */
static void min_up(qh_res* a, int j) {
  for (;;) {
    int i = (j - 1) / 2;
    if (i == j || a[j].dist >= a[i].dist) break;
    qh_res t = a[i]; a[i] = a[j]; a[j] = t;
    j = i;
  }
}
static void min_down(qh_res* a, int i0, int n) {
  int i = i0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= n || j1 < 0) break;
    int j = j1, j2 = j1 + 1;
    if (j2 < n && a[j2].dist < a[j1].dist) j = j2;
    if (a[i].dist <= a[j].dist) break;
    qh_res t = a[i]; a[i] = a[j]; a[j] = t;
    i = j;
  }
}
static void min_push(qh_heap* h, qh_res x) {
  heap_reserve(h, h->n + 1);
  h->a[h->n++] = x;
  min_up(h->a, h->n - 1);
}
static qh_res min_pop(qh_heap* h) {
  int n = h->n - 1;
  qh_res t = h->a[0]; h->a[0] = h->a[n]; h->a[n] = t;
  min_down(h->a, 0, n);
  h->n = n;
  return h->a[n];
}
/* max-heap: hnsw.go:153-196 */
static void max_up(qh_res* a, int j) {
  for (;;) {
    int i = (j - 1) / 2;
    if (i == j || a[j].dist <= a[i].dist) break;
    qh_res t = a[i]; a[i] = a[j]; a[j] = t;
    j = i;
  }
}
static void max_down(qh_res* a, int i0, int n) {
  int i = i0;
  for (;;) {
    int j1 = 2 * i + 1;
    if (j1 >= n || j1 < 0) break;
    int j = j1, j2 = j1 + 1;
    if (j2 < n && a[j2].dist > a[j1].dist) j = j2;
    if (a[i].dist >= a[j].dist) break;
    qh_res t = a[i]; a[i] = a[j]; a[j] = t;
    i = j;
  }
}
static void max_push(qh_heap* h, qh_res x) {
  heap_reserve(h, h->n + 1);
  h->a[h->n++] = x;
  max_up(h->a, h->n - 1);
}
static qh_res max_pop(qh_heap* h) {
  int n = h->n - 1;
  qh_res t = h->a[0]; h->a[0] = h->a[n]; h->a[n] = t;
  max_down(h->a, 0, n);
  h->n = n;
  return h->a[n];
}

typedef struct {
  int64_t n, cap_nodes;
  int d, metric, arith, M, max_m0, ef_construction, ef_search, max_level;
  const float* vec;     /* borrowed: [n x d] */
  int32_t* level;       /* [n] */
  uint32_t** conn;      /* conn[i][l] pointer table flattened: conn[i] -> array of (level+1) lists */
  int32_t** conn_n;     /* conn_n[i][l] list length */
  uint32_t entry;
  int cur_level;
  uint8_t* visited;
  qh_heap cand, res;
  int64_t dist_evals;
  uint64_t trace_hash;  /* order-sensitive hash of the rows whose distance was computed */
} qo_hnsw;

static uint32_t* list_ptr(const qo_hnsw* h, int64_t i, int l) {
  /* lists of node i are stored back to back: level 0 (max_m0+1 slots), then M+1 slots per level */
  uint32_t* base = h->conn[i];
  return l == 0 ? base : base + (h->max_m0 + 1) + (size_t)(l - 1) * (h->M + 1);
}

static float hdist(qo_hnsw* h, const float* q, uint32_t row) {
  h->dist_evals++;
  h->trace_hash = qo_mix64(h->trace_hash ^ (uint64_t)row);
  return qo_distance(h->metric, h->arith, q, h->vec + (size_t)row * h->d, h->d);
}

/* hnsw.go:471-580. Results ascending into out (at most ef); returns the count. */
static int search_layer(qo_hnsw* h, const float* q, uint32_t entry, int ef, int level, qh_res* out) {
  if (h->n == 0) return 0;
  memset(h->visited, 0, (size_t)h->n);
  h->visited[entry] = 1;
  float d0 = hdist(h, q, entry);
  h->cand.n = 0;
  h->res.n = 0;
  qh_res e = {entry, d0};
  min_push(&h->cand, e);
  max_push(&h->res, e);
  while (h->cand.n > 0) {
    qh_res cur = min_pop(&h->cand);
    if (h->res.n >= ef && cur.dist > h->res.a[0].dist) break;
    if (level > h->level[cur.idx]) continue; /* level >= len(Connections) */
    const uint32_t* lst = list_ptr(h, cur.idx, level);
    const int ln = h->conn_n[cur.idx][level];
    for (int c = 0; c < ln; ++c) {
      const uint32_t id = lst[c];
      if ((int64_t)id >= h->n) continue;
      if (!h->visited[id]) {
        h->visited[id] = 1;
        const float cd = hdist(h, q, id);
        if (h->res.n < ef || cd < h->res.a[0].dist) {
          qh_res r = {id, cd};
          min_push(&h->cand, r);
          max_push(&h->res, r);
          if (h->res.n > ef) max_pop(&h->res);
        }
      }
    }
  }
  const int n = h->res.n;
  for (int i = n - 1; i >= 0; --i) out[i] = max_pop(&h->res);
  return n;
}

static int res_cmp(const void* pa, const void* pb) {
  const qh_res* a = (const qh_res*)pa;
  const qh_res* b = (const qh_res*)pb;
  if (a->dist == b->dist) return (a->idx > b->idx) - (a->idx < b->idx);
  return a->dist < b->dist ? -1 : 1;
}
/* hnsw.go:583-599 */
static int select_neighbors(qh_res* c, int n, int k) {
  if (k <= 0 || n == 0) return 0;
  qsort(c, (size_t)n, sizeof(qh_res), res_cmp);
  return n > k ? k : n;
}

static int random_level(const qo_hnsw* h, uint64_t seed, uint64_t node) {
  int level = 0;
  const int attempts = h->max_level < 10 ? h->max_level : 10;
  for (int i = 0; i < attempts; ++i) {
    if (qo_uniform(seed, node, (uint32_t)i, 7) < 0.25f) level++;
    else break;
  }
  if (level >= h->max_level) level = h->max_level - 1;
  return level;
}

qo_hnsw* qo_hnsw_build(const float* vec, int64_t n, int d, int metric, int arith, int M, int max_m0,
                       int ef_construction, int ef_search, int max_level, uint64_t seed) {
  qo_hnsw* h = (qo_hnsw*)calloc(1, sizeof(qo_hnsw));
  h->d = d; h->metric = metric; h->arith = arith; h->M = M; h->max_m0 = max_m0;
  h->ef_construction = ef_construction; h->ef_search = ef_search; h->max_level = max_level;
  h->vec = vec;
  h->level = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  h->conn = (uint32_t**)calloc((size_t)n + 1, sizeof(uint32_t*));
  h->conn_n = (int32_t**)calloc((size_t)n + 1, sizeof(int32_t*));
  h->visited = (uint8_t*)calloc((size_t)n + 1, 1);
  qh_res* buf = (qh_res*)malloc(sizeof(qh_res) * (size_t)(ef_construction + max_m0 + M + 8));
  qh_res* nd = (qh_res*)malloc(sizeof(qh_res) * (size_t)(max_m0 + M + 8));
  const char* std_env = getenv("QO_HNSW_STANDARD");
  const int standard = std_env != NULL && std_env[0] == '1';
  for (int64_t i = 0; i < n; ++i) {
    /* Insert, hnsw.go:266-334 */
    int level = random_level(h, seed, (uint64_t)i);
    const int old_level = h->cur_level;
    h->level[i] = level;
    h->conn[i] = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)((max_m0 + 1) + level * (M + 1)));
    h->conn_n[i] = (int32_t*)calloc((size_t)level + 1, sizeof(int32_t));
    h->n = i + 1;
    if (i == 0) {
      h->entry = 0;
      h->cur_level = level;
      continue;
    }
    /* connectNode, hnsw.go:337-468 */
    const float* v = vec + (size_t)i * d;
    if (level >= h->max_level) level = h->max_level - 1;
    uint32_t ep = h->entry;
    for (int lc = old_level; lc > level; --lc) {
      if (lc > h->level[ep]) continue;
      int m = search_layer(h, v, ep, 1, lc, buf);
      if (m > 0) ep = buf[0].idx;
    }
    for (int lc = (level < old_level ? level : old_level); lc >= 0; --lc) {
      int m = search_layer(h, v, ep, h->ef_construction, lc, buf);
      if (m == 0) continue;
      const int maxc = lc == 0 ? max_m0 : M;
      const int ns = select_neighbors(buf, m, maxc < m ? maxc : m);
      uint32_t* mine = list_ptr(h, i, lc);
      for (int s = 0; s < ns; ++s) mine[h->conn_n[i][lc]++] = buf[s].idx;
      for (int s = 0; s < ns; ++s) {
        const uint32_t nb = buf[s].idx;
        if (lc > h->level[nb]) continue;
        uint32_t* nl = list_ptr(h, nb, lc);
        nl[h->conn_n[nb][lc]++] = (uint32_t)i;
        if (h->conn_n[nb][lc] > maxc) {
          const int cn = h->conn_n[nb][lc];
          for (int c = 0; c < cn; ++c) {
            nd[c].idx = nl[c];
            nd[c].dist = qo_distance(metric, arith, vec + (size_t)nb * d, vec + (size_t)nl[c] * d, d);
          }
          const int keep = select_neighbors(nd, cn, maxc);
          for (int c = 0; c < keep; ++c) nl[c] = nd[c].idx;
          h->conn_n[nb][lc] = keep;
        }
      }
      /* hnsw.go sets the next layer's entry point to the node being inserted (which has no links
       * there yet) - kept faithfully. QO_HNSW_STANDARD=1 descends from the closest neighbour found
       * instead (the textbook algorithm): measurement aid only, never used by a parity test. */
      if (ns > 0) ep = standard ? buf[0].idx : (uint32_t)i;
    }
    if (level > old_level) {
      h->entry = (uint32_t)i;
      h->cur_level = level;
    }
  }
  free(buf);
  free(nd);
  h->dist_evals = 0;
  h->trace_hash = 0;
  return h;
}

void qo_hnsw_free(qo_hnsw* h) {
  if (!h) return;
  for (int64_t i = 0; i < h->n; ++i) { free(h->conn[i]); free(h->conn_n[i]); }
  free(h->conn); free(h->conn_n); free(h->level); free(h->visited); free(h->cand.a); free(h->res.a);
  free(h);
}

/* hnsw.go:602-672 (without the under-fill pass). Returns the result count; evals / trace are
 * accumulated over the call. */
int qo_hnsw_search(qo_hnsw* h, const float* q, int k, float* out_dist, uint32_t* out_idx, int64_t* evals,
                   uint64_t* trace) {
  h->dist_evals = 0;
  h->trace_hash = 0;
  if (h->n == 0) return 0;
  if (k <= 0) return -1;
  if (k > h->n) k = (int)h->n;
  uint32_t ep = h->entry;
  (void)hdist(h, q, ep); /* entryDistance, hnsw.go:637 */
  int ef = h->ef_search < k ? k : h->ef_search;
  qh_res* buf = (qh_res*)malloc(sizeof(qh_res) * (size_t)(ef + 8));
  for (int level = h->cur_level; level > 0; --level) {
    int m = search_layer(h, q, ep, 1, level, buf);
    if (m == 0) continue;
    ep = buf[0].idx;
  }
  int m = search_layer(h, q, ep, ef, 0, buf);
  if (m > k) m = k;
  for (int i = 0; i < m; ++i) { out_dist[i] = buf[i].dist; out_idx[i] = buf[i].idx; }
  free(buf);
  if (evals) *evals = h->dist_evals;
  if (trace) *trace = h->trace_hash;
  return m;
}

/* Flat views for the GPU-batched walk: level[n]; adj0[n x max_m0] (0xFFFFFFFF padded, list order);
 * upper_off[n+1] (uint32 units), upper_adj = per node level[i] blocks of M entries. */
void qo_hnsw_info(const qo_hnsw* h, int64_t* n, int* entry, int* cur_level, int64_t* upper_len) {
  *n = h->n; *entry = (int)h->entry; *cur_level = h->cur_level;
  int64_t u = 0;
  for (int64_t i = 0; i < h->n; ++i) u += (int64_t)h->level[i] * h->M;
  *upper_len = u;
}
void qo_hnsw_export(const qo_hnsw* h, int32_t* level, uint32_t* adj0, int64_t* upper_off, uint32_t* upper_adj) {
  int64_t u = 0;
  for (int64_t i = 0; i < h->n; ++i) {
    level[i] = h->level[i];
    for (int c = 0; c < h->max_m0; ++c)
      adj0[i * h->max_m0 + c] = c < h->conn_n[i][0] ? list_ptr(h, i, 0)[c] : 0xFFFFFFFFu;
    upper_off[i] = u;
    for (int l = 1; l <= h->level[i]; ++l)
      for (int c = 0; c < h->M; ++c) upper_adj[u++] = c < h->conn_n[i][l] ? list_ptr(h, i, l)[c] : 0xFFFFFFFFu;
  }
  upper_off[h->n] = u;
}


/* A graph from its flat arrays (the layout qo_hnsw_export writes): lets a stored graph be walked on the host
 * without rebuilding it (tests/bench_hnsw_c5.py at 1M nodes). */
qo_hnsw* qo_hnsw_import(const float* vec, int64_t n, int d, int metric, int arith, int M, int max_m0, int ef_search,
                        int entry, int cur_level, const int32_t* level, const uint32_t* adj0, const int64_t* upper_off,
                        const uint32_t* upper_adj) {
  qo_hnsw* h = (qo_hnsw*)calloc(1, sizeof(qo_hnsw));
  h->d = d; h->metric = metric; h->arith = arith; h->M = M; h->max_m0 = max_m0;
  h->ef_construction = 0; h->ef_search = ef_search; h->max_level = 16;
  h->vec = vec;
  h->n = n;
  h->entry = (uint32_t)entry;
  h->cur_level = cur_level;
  h->level = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  h->conn = (uint32_t**)calloc((size_t)n + 1, sizeof(uint32_t*));
  h->conn_n = (int32_t**)calloc((size_t)n + 1, sizeof(int32_t*));
  h->visited = (uint8_t*)calloc((size_t)n + 1, 1);
  for (int64_t i = 0; i < n; ++i) {
    const int lv = level[i] < 0 ? 0 : level[i];
    h->level[i] = level[i];
    h->conn[i] = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)((max_m0 + 1) + lv * (M + 1)));
    h->conn_n[i] = (int32_t*)calloc((size_t)lv + 1, sizeof(int32_t));
    uint32_t* l0 = list_ptr(h, i, 0);
    for (int c = 0; c < max_m0 && adj0[i * max_m0 + c] != 0xFFFFFFFFu; ++c) l0[h->conn_n[i][0]++] = adj0[i * max_m0 + c];
    for (int l = 1; l <= lv; ++l) {
      uint32_t* ll = list_ptr(h, i, l);
      const uint32_t* src = upper_adj + upper_off[i] + (size_t)(l - 1) * M;
      for (int c = 0; c < M && src[c] != 0xFFFFFFFFu; ++c) ll[h->conn_n[i][l]++] = src[c];
    }
  }
  return h;
}

/* hnsw.Search for a batch on `threads` host threads (one goroutine per query in the reference's BatchSearch):
 * every thread walks with its own visited array and heaps over the shared, read-only graph. */
typedef struct {
  const qo_hnsw* h;
  const float* queries;
  int64_t nq;
  int k, tid, threads;
  float* out_dist;
  uint32_t* out_idx;
  int32_t* out_cnt;
  int64_t* out_evals;
} qo_walk_job;

static void* walk_thread(void* arg) {
  qo_walk_job* j = (qo_walk_job*)arg;
  qo_hnsw local = *j->h; /* shallow copy: graph arrays shared, scratch private */
  local.visited = (uint8_t*)calloc((size_t)local.n + 1, 1);
  memset(&local.cand, 0, sizeof(local.cand));
  memset(&local.res, 0, sizeof(local.res));
  for (int64_t i = j->tid; i < j->nq; i += j->threads) {
    int64_t ev = 0;
    const int m = qo_hnsw_search(&local, j->queries + (size_t)i * local.d, j->k, j->out_dist + (size_t)i * j->k,
                                 j->out_idx + (size_t)i * j->k, &ev, NULL);
    j->out_cnt[i] = m;
    if (j->out_evals) j->out_evals[i] = ev;
  }
  free(local.visited);
  free(local.cand.a);
  free(local.res.a);
  return NULL;
}

void qo_hnsw_search_batch_mt(const qo_hnsw* h, const float* queries, int64_t nq, int k, int threads, float* out_dist,
                             uint32_t* out_idx, int32_t* out_cnt, int64_t* out_evals) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  qo_walk_job jobs[256];
  for (int t = 0; t < threads; ++t) {
    qo_walk_job jb = {h, queries, nq, k, t, threads, out_dist, out_idx, out_cnt, out_evals};
    jobs[t] = jb;
    pthread_create(&th[t], NULL, walk_thread, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
