/*
 * oracle/quiver_oracle.c — CPU restatement of Quiver's exact-search hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this. The product path
 * (quiver_b200/) never links or calls it.
 *
 * The reference (TFMV/quiver) is pure Go and there is no Go toolchain in the build
 * image, so the reference itself cannot be compiled here (oracle/_ref is therefore
 * absent). This file restates the algorithm line by line; every
 * function cites the reference file:line it follows. It is pinned against every
 * known-answer vector the reference's own tests hold for this path
 * (the JSON files under tests/golden, extracted from the Go test sources by
 * tests/golden/make_golden.py; checked in tests/test_oracle_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math  (Go on amd64 never fuses
 * multiply-add; see oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "synth.h"

enum { QO_COSINE = 0, QO_L2 = 1, QO_DOT = 2, QO_SQL2 = 3, QO_L1 = 4 };
enum { QO_ARITH_VECTORTYPES = 0, QO_ARITH_HNSW_F32 = 1 };

/* pkg/vectortypes/distances.go:12-40 */
float qo_cosine(const float* a, const float* b, int d) {
  double dot = 0, ma = 0, mb = 0;
  for (int i = 0; i < d; ++i) {
    dot += (double)a[i] * (double)b[i];
    ma += (double)a[i] * (double)a[i];
    mb += (double)b[i] * (double)b[i];
  }
  if (ma == 0 || mb == 0) return 1.0f;
  double sim = dot / (sqrt(ma) * sqrt(mb));
  if (sim > 1.0) sim = 1.0;
  else if (sim < -1.0) sim = -1.0;
  return (float)(1.0 - sim);
}

/* pkg/vectortypes/distances.go:43-55 (the subtraction is float32, :50) */
float qo_euclidean(const float* a, const float* b, int d) {
  double sum = 0;
  for (int i = 0; i < d; ++i) {
    float df = a[i] - b[i];
    double diff = (double)df;
    sum += diff * diff;
  }
  return (float)sqrt(sum);
}

/* pkg/vectortypes/distances.go:60-72 (all float32) */
float qo_sq_euclidean(const float* a, const float* b, int d) {
  float sum = 0;
  for (int i = 0; i < d; ++i) {
    float diff = a[i] - b[i];
    sum += diff * diff;
  }
  return sum;
}

/* pkg/vectortypes/distances.go:77-90 */
float qo_dot(const float* a, const float* b, int d) {
  double dot = 0;
  for (int i = 0; i < d; ++i) dot += (double)a[i] * (double)b[i];
  return (float)(1.0 - dot);
}

/* pkg/vectortypes/distances.go:93-104 */
float qo_manhattan(const float* a, const float* b, int d) {
  double sum = 0;
  for (int i = 0; i < d; ++i) {
    float df = a[i] - b[i];
    sum += fabs((double)df);
  }
  return (float)sum;
}

/* pkg/hnsw/adapter.go:105-136 (all float32; sqrt through float64 then rounded) */
float qo_hnsw_cosine(const float* a, const float* b, int d) {
  float dot = 0, na = 0, nb = 0;
  for (int i = 0; i < d; ++i) {
    dot += a[i] * b[i];
    na += a[i] * a[i];
    nb += b[i] * b[i];
  }
  if (na == 0 || nb == 0) return 1.0f;
  float sa = (float)sqrt((double)na), sb = (float)sqrt((double)nb);
  float den = sa * sb;
  float sim = dot / den;
  if (sim > 1.0f) sim = 1.0f;
  else if (sim < -1.0f) sim = -1.0f;
  return 1.0f - sim;
}

/* pkg/hnsw/adapter.go:139-151 */
float qo_hnsw_euclidean(const float* a, const float* b, int d) {
  float sum = 0;
  for (int i = 0; i < d; ++i) {
    float diff = a[i] - b[i];
    sum += diff * diff;
  }
  return (float)sqrt((double)sum);
}

/* pkg/hnsw/adapter.go:154-166 */
float qo_hnsw_dot(const float* a, const float* b, int d) {
  float dot = 0;
  for (int i = 0; i < d; ++i) dot += a[i] * b[i];
  return 1.0f - dot;
}

/* metric/arith dispatch: types.go:36-65 maps the enum to the vectortypes functions;
 * db.go:181-188 wires the hnsw float32 variants after a reload. */
float qo_distance(int metric, int arith, const float* a, const float* b, int d) {
  if (arith == QO_ARITH_HNSW_F32) {
    switch (metric) {
      case QO_COSINE: return qo_hnsw_cosine(a, b, d);
      case QO_L2: return qo_hnsw_euclidean(a, b, d);
      case QO_DOT: return qo_hnsw_dot(a, b, d);
      default: break;
    }
  }
  switch (metric) {
    case QO_COSINE: return qo_cosine(a, b, d);
    case QO_L2: return qo_euclidean(a, b, d);
    case QO_DOT: return qo_dot(a, b, d);
    case QO_SQL2: return qo_sq_euclidean(a, b, d);
    case QO_L1: return qo_manhattan(a, b, d);
    default: return qo_cosine(a, b, d); /* unknown => cosine, types.go:46-47 */
  }
}

void qo_distances(int metric, int arith, const float* corpus, int64_t n, int d, const float* q,
                  float* out) {
  for (int64_t r = 0; r < n; ++r) out[r] = qo_distance(metric, arith, q, corpus + r * (int64_t)d, d);
}

typedef struct {
  float dist;
  int64_t row;
} qo_hit;

/* exact.go:75 orders on Distance only and Go's sort is unstable over a random map order,
 * so ties are arbitrary in the reference. The oracle fixes them by row index so that
 * the comparison with the device path (same rule) can be exact. */
static int qo_hit_cmp(const void* pa, const void* pb) {
  const qo_hit* a = (const qo_hit*)pa;
  const qo_hit* b = (const qo_hit*)pb;
  if (a->dist < b->dist) return -1;
  if (a->dist > b->dist) return 1;
  return (a->row > b->row) - (a->row < b->row);
}

/* pkg/hybrid/exact.go:92-133 — scan every live row, sort ALL results, slice to k.
 * `live` (nullable) has one byte per row: 0 = deleted (exact.go:61-70) or filtered out.
 * Returns the number of results written, -1 for k <= 0 (error "k must be positive",
 * exact.go:104-106), 0 for an empty index regardless of k (exact.go:96-98). */
int64_t qo_exact_search(const float* corpus, int64_t n, int d, int metric, int arith,
                        const uint8_t* live, const float* q, int64_t k, float* out_dist,
                        int64_t* out_row) {
  int64_t n_live = 0;
  if (live) {
    for (int64_t r = 0; r < n; ++r) n_live += live[r] != 0;
  } else {
    n_live = n;
  }
  if (n_live == 0) return 0;
  if (k <= 0) return -1;
  if (k > n_live) k = n_live;
  qo_hit* hits = (qo_hit*)malloc((size_t)n_live * sizeof(qo_hit));
  if (!hits) return -2;
  int64_t m = 0;
  for (int64_t r = 0; r < n; ++r) {
    if (live && !live[r]) continue;
    hits[m].dist = qo_distance(metric, arith, q, corpus + r * (int64_t)d, d);
    hits[m].row = r;
    ++m;
  }
  qsort(hits, (size_t)m, sizeof(qo_hit), qo_hit_cmp);
  for (int64_t i = 0; i < k; ++i) {
    out_dist[i] = hits[i].dist;
    out_row[i] = hits[i].row;
  }
  free(hits);
  return k;
}

/* pkg/hybrid/hybrid_index.go:703-795 — BatchSearch runs one goroutine per query; each
 * goroutine is exactly qo_exact_search. `threads` workers pull queries from a shared
 * counter (the Go scheduler multiplexes goroutines over GOMAXPROCS threads). */
typedef struct {
  const float* corpus;
  int64_t n;
  int d, metric, arith;
  const uint8_t* live;
  const float* queries;
  int64_t nq, k;
  float* out_dist;
  int64_t* out_row;
  int64_t* out_count;
  volatile int64_t* next;
} qo_batch_job;

static void* qo_batch_worker(void* p) {
  qo_batch_job* j = (qo_batch_job*)p;
  for (;;) {
    int64_t i = __sync_fetch_and_add(j->next, 1);
    if (i >= j->nq) break;
    j->out_count[i] = qo_exact_search(j->corpus, j->n, j->d, j->metric, j->arith, j->live,
                                      j->queries + i * (int64_t)j->d, j->k,
                                      j->out_dist + i * j->k, j->out_row + i * j->k);
  }
  return NULL;
}

int qo_exact_search_batch(const float* corpus, int64_t n, int d, int metric, int arith,
                          const uint8_t* live, const float* queries, int64_t nq, int64_t k,
                          int threads, float* out_dist, int64_t* out_row, int64_t* out_count) {
  if (k <= 0 || nq <= 0) return -1;
  for (int64_t i = 0; i < nq * k; ++i) {
    out_dist[i] = INFINITY;
    out_row[i] = -1;
  }
  volatile int64_t next = 0;
  qo_batch_job job = {corpus, n, d, metric, arith, live, queries, nq, k, out_dist, out_row, out_count, &next};
  if (threads <= 1) {
    qo_batch_worker(&job);
    return 0;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, qo_batch_worker, &job);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  return 0;
}

/* Synthetic matrices (see synth.h). Rows [row0, row0+n). */
void qo_synth_fill(int kind, uint64_t seed, int64_t row0, int64_t n, int dim, float* out) {
  for (int64_t r = 0; r < n; ++r) qo_synth_row(kind, seed, (uint64_t)(row0 + r), dim, out + r * (int64_t)dim);
}

typedef struct {
  int kind;
  uint64_t seed;
  int64_t row0, n;
  int dim;
  float* out;
  int t, nt;
} qo_fill_job;

static void* qo_fill_worker(void* p) {
  qo_fill_job* j = (qo_fill_job*)p;
  int64_t lo = j->n * j->t / j->nt, hi = j->n * (j->t + 1) / j->nt;
  qo_synth_fill(j->kind, j->seed, j->row0 + lo, hi - lo, j->dim, j->out + lo * (int64_t)j->dim);
  return NULL;
}

void qo_synth_fill_mt(int kind, uint64_t seed, int64_t row0, int64_t n, int dim, float* out, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  pthread_t th[64];
  qo_fill_job jobs[64];
  for (int t = 0; t < threads; ++t) {
    qo_fill_job j = {kind, seed, row0, n, dim, out, t, threads};
    jobs[t] = j;
    pthread_create(&th[t], NULL, qo_fill_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
