// hnsw.cu — hnsw.Search (reference pkg/hnsw/hnsw.go:602-713) as one persistent kernel: a warp per query.
//
// The round-1 walk kept the heaps on the host and sent one distance batch per expansion step and lock
// step to the device: one launch + two PCIe trips per step, and it lost to a single CPU thread on short
// walks. Here the whole walk lives on the device: the candidate min-heap and the result max-heap of a
// query sit in the warp's slice of shared memory and are driven by lane 0 with the reference's own sift
// rules (hnsw.go:101-196: heaps compare on Distance only, so equal distances are resolved by heap shape —
// the rules must be the same for the walk to be step-identical); the visited set is a bitset in global
// memory (marked BEFORE the distance is computed, hnsw.go:542-543); the <= 32 unvisited neighbours of an
// expansion step (hnsw.go:536-563) are taken by one lane each, which evaluates the whole distance in the
// reference's arithmetic and operation order (exact_distance_lane: one sequential chain per lane, 32
// chains side by side); lane 0 then admits them in connection order (strict `<`, hnsw.go:553; stop on
// strict `>`, :513). Thousands of warps are in flight, so the per-query latency chain (heap sifts,
// dependent gathers) is hidden by other queries — throughput comes from query parallelism, exactly as
// SURVEY Appendix C prescribes. A query whose candidate heap outgrows its shared-memory slice is
// reported with count -1 and repeated by the host walk (never seen at efSearch = 128).
#include <algorithm>

#include "exact.cuh"
#include "hnsw.cuh"

namespace qg {

// ---- one (query, row) distance by ONE thread, reference arithmetic and order ----------------------
// q: the query (shared memory, all lanes read the same element: broadcast), x: the stored row.
// Same value as exact_distance_warp(metric, arith, q, x, d): the terms are formed with the same
// roundings and added in index order.
__device__ __forceinline__ float exact_distance_lane(int metric, int arith, const float* __restrict__ q,
                                                     const float* __restrict__ x, int d, int dp) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int n4 = dp >> 2;  // rows are zero padded to a multiple of 4 floats; zero terms do not change a sum
  if (arith == ARITH_HNSW_F32 && (metric == METRIC_COSINE || metric == METRIC_L2 || metric == METRIC_DOT)) {
    // pkg/hnsw/adapter.go:105-167 — everything float32, sequential, no fused multiply-add
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int i = 0; i < n4; ++i) {
      const float4 v = __ldg(x4 + i);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i * 4 + j >= d) break;
        const float a = q[i * 4 + j], b = xs[j];
        if (metric == METRIC_L2) {
          const float df = __fsub_rn(a, b);
          s0 = __fadd_rn(s0, __fmul_rn(df, df));
        } else {
          s0 = __fadd_rn(s0, __fmul_rn(a, b));
          if (metric == METRIC_COSINE) {
            s1 = __fadd_rn(s1, __fmul_rn(a, a));
            s2 = __fadd_rn(s2, __fmul_rn(b, b));
          }
        }
      }
    }
    if (metric == METRIC_L2) return (float)sqrt((double)s0);
    if (metric == METRIC_DOT) return __fsub_rn(1.0f, s0);
    if (s1 == 0.f || s2 == 0.f) return 1.0f;
    const float sa = (float)sqrt((double)s1), sb = (float)sqrt((double)s2);
    float sim = __fdiv_rn(s0, __fmul_rn(sa, sb));
    if (sim > 1.0f) sim = 1.0f;
    else if (sim < -1.0f) sim = -1.0f;
    return __fsub_rn(1.0f, sim);
  }
  if (metric == METRIC_SQL2) {  // distances.go:60-72
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < n4; ++i) {
      const float4 v = __ldg(x4 + i);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i * 4 + j >= d) break;
        const float df = __fsub_rn(q[i * 4 + j], xs[j]);
        s = __fadd_rn(s, __fmul_rn(df, df));
      }
    }
    return s;
  }
  // float64 accumulators (distances.go:17-22, 48-52, 82-85, 99-101); a float32 product is exact in float64
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 4
  for (int i = 0; i < n4; ++i) {
    const float4 v = __ldg(x4 + i);
    const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i * 4 + j >= d) break;
      const float a = q[i * 4 + j], b = xs[j];
      if (metric == METRIC_L2) {
        const double df = (double)__fsub_rn(a, b);
        s0 = __dadd_rn(s0, __dmul_rn(df, df));
      } else if (metric == METRIC_L1) {
        s0 = __dadd_rn(s0, fabs((double)__fsub_rn(a, b)));
      } else {
        s0 = __dadd_rn(s0, __dmul_rn((double)a, (double)b));
        if (metric == METRIC_COSINE) {
          s1 = __dadd_rn(s1, __dmul_rn((double)a, (double)a));
          s2 = __dadd_rn(s2, __dmul_rn((double)b, (double)b));
        }
      }
    }
  }
  if (metric == METRIC_L2) return (float)sqrt(s0);
  if (metric == METRIC_L1) return (float)s0;
  if (metric == METRIC_DOT) return (float)(1.0 - s0);
  if (s1 == 0.0 || s2 == 0.0) return 1.0f;
  double sim = s0 / (sqrt(s1) * sqrt(s2));
  if (sim > 1.0) sim = 1.0;
  else if (sim < -1.0) sim = -1.0;
  return (float)(1.0 - sim);
}

// ---- the reference's binary heaps over (idx, dist) pairs in shared memory --------------------------
struct HRes {
  uint32_t idx;
  float dist;
};
// hnsw.go:101-144
__device__ __forceinline__ void min_push(HRes* a, int& n, HRes x) {
  a[n] = x;
  int j = n++;
  for (;;) {
    const int i = (j - 1) / 2;
    if (i == j || a[j].dist >= a[i].dist) break;
    const HRes t = a[i];
    a[i] = a[j];
    a[j] = t;
    j = i;
  }
}
__device__ __forceinline__ HRes min_pop(HRes* a, int& n) {
  const int m = n - 1;
  HRes t = a[0];
  a[0] = a[m];
  a[m] = t;
  int i = 0;
  for (;;) {
    const int j1 = 2 * i + 1;
    if (j1 >= m || j1 < 0) break;
    int j = j1;
    const int j2 = j1 + 1;
    if (j2 < m && a[j2].dist < a[j1].dist) j = j2;
    if (a[i].dist <= a[j].dist) break;
    t = a[i];
    a[i] = a[j];
    a[j] = t;
    i = j;
  }
  n = m;
  return a[m];
}
// hnsw.go:153-196
__device__ __forceinline__ void max_push(HRes* a, int& n, HRes x) {
  a[n] = x;
  int j = n++;
  for (;;) {
    const int i = (j - 1) / 2;
    if (i == j || a[j].dist <= a[i].dist) break;
    const HRes t = a[i];
    a[i] = a[j];
    a[j] = t;
    j = i;
  }
}
__device__ __forceinline__ HRes max_pop(HRes* a, int& n) {
  const int m = n - 1;
  HRes t = a[0];
  a[0] = a[m];
  a[m] = t;
  int i = 0;
  for (;;) {
    const int j1 = 2 * i + 1;
    if (j1 >= m || j1 < 0) break;
    int j = j1;
    const int j2 = j1 + 1;
    if (j2 < m && a[j2].dist > a[j1].dist) j = j2;
    if (a[i].dist >= a[j].dist) break;
    t = a[i];
    a[i] = a[j];
    a[j] = t;
    i = j;
  }
  n = m;
  return a[m];
}

struct HnswKParams {
  HnswDevGraph g;
  const float* vec;
  int dp, d, metric, arith;
  const float* queries;  // [nq x d] (unpadded)
  int nq, kk, ef0, cand_cap;
  uint32_t* visited;     // [slots x n_words]
  uint32_t* touched;     // [slots x HNSW_TOUCH_CAP]
  long long n_words;
  int* next_query;       // work counter
  uint32_t* out_idx;     // [nq x kk]
  float* out_dist;       // [nq x kk]
  int* out_count;        // [nq]
  long long* out_evals;  // [nq] or nullptr
};

// searchLayer (hnsw.go:471-580) for the warp's query. Results stay in the max-heap `res` (n_res entries);
// returns false when the candidate heap overflowed.
__device__ __forceinline__ bool search_layer(const HnswKParams& p, const float* q_s, uint32_t entry, int ef, int level,
                                             HRes* cand, HRes* res, int& n_res, uint32_t* pend_idx, float* pend_d,
                                             uint32_t* vis, uint32_t* touched, int& n_touched, bool& touched_overflow,
                                             long long& evals) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  // visited = make([]bool, ...) (hnsw.go:483): clear what the previous layer / query marked
  if (touched_overflow) {
    for (long long i = lane; i < p.n_words; i += 32) vis[i] = 0u;
  } else {
    for (int i = lane; i < n_touched; i += 32) vis[touched[i]] = 0u;
  }
  __syncwarp();
  n_touched = 0;
  touched_overflow = false;
  int n_cand = 0;
  n_res = 0;
  if (lane == 0) {
    vis[entry >> 5] = 1u << (entry & 31);
    touched[0] = entry >> 5;
  }
  n_touched = 1;
  // entry distance (hnsw.go:488-505)
  float d0 = 0.f;
  if (lane == 0) {
    d0 = exact_distance_lane(p.metric, p.arith, q_s, p.vec + (size_t)entry * p.dp, p.d, p.dp);
    const HRes e{entry, d0};
    cand[0] = e;
    res[0] = e;
  }
  evals += 1;
  n_cand = 1;
  n_res = 1;
  __syncwarp();
  bool ok = true;
  for (;;) {
    // ---- lane 0: pop the closest candidate (hnsw.go:508-520) ----
    int action = 0;  // 0 = layer finished, 1 = expand cur, 2 = skip cur
    uint32_t cur = 0;
    if (lane == 0) {
      if (n_cand > 0) {
        const HRes c = min_pop(cand, n_cand);
        if (!(n_res >= ef && c.dist > res[0].dist)) {
          cur = c.idx;
          const int lv = p.g.level[c.idx];
          action = (lv < 0 || level > lv) ? 2 : 1;
        }
      }
    }
    action = __shfl_sync(0xffffffffu, action, 0);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    n_cand = __shfl_sync(0xffffffffu, n_cand, 0);
    if (action == 0) break;
    if (action == 2) continue;
    // ---- the connection list of cur at this level, 32 entries at a time, in list order ----
    const uint32_t* lst;
    int cnt;
    if (level == 0) {
      lst = p.g.adj0 + (size_t)cur * p.g.max_m0;
      cnt = p.g.max_m0;
    } else {
      lst = p.g.upper_adj + p.g.upper_off[cur] + (size_t)(level - 1) * p.g.m;
      cnt = p.g.m;
    }
    bool list_done = false;
    for (int c0 = 0; c0 < cnt && !list_done; c0 += 32) {
      const int c = c0 + lane;
      uint32_t id = 0xFFFFFFFFu;
      if (c < cnt) id = __ldg(lst + c);
      const unsigned none = __ballot_sync(0xffffffffu, c < cnt && id == 0xFFFFFFFFu);
      const unsigned upto = none ? ((1u << (__ffs((int)none) - 1)) - 1u) : 0xffffffffu;  // entries before the end mark
      if (none) list_done = true;
      bool want = c < cnt && ((upto >> lane) & 1u) && (long long)id < p.g.n_nodes;
      if (want) want = __ldg(p.g.level + id) >= 0;
      // a list never holds a node twice, but if it did the first occurrence marks it (sequential semantics)
      const unsigned same = __match_any_sync(0xffffffffu, want ? id : (0xFFFFFF00u | (unsigned)lane));
      if (want && (same & lt) != 0u) want = false;
      bool fresh = false, first_in_word = false;
      if (want) {
        const uint32_t bit = 1u << (id & 31);
        const uint32_t old = atomicOr(vis + (id >> 5), bit);
        fresh = (old & bit) == 0u;
        first_in_word = old == 0u;  // exactly one lane sees the word empty: it records the word for the next clear
      }
      const unsigned nw = __ballot_sync(0xffffffffu, first_in_word);
      if (nw != 0u) {
        if (first_in_word) {
          const int pos = n_touched + __popc(nw & lt);
          if (pos < HNSW_TOUCH_CAP) touched[pos] = id >> 5;
        }
        n_touched += __popc(nw);
        if (n_touched > HNSW_TOUCH_CAP) touched_overflow = true;  // the next clear wipes the whole bitset
      }
      const unsigned fm = __ballot_sync(0xffffffffu, fresh);
      const int np = __popc(fm);
      if (np > 0) {
        float dist = 0.f;
        if (fresh) dist = exact_distance_lane(p.metric, p.arith, q_s, p.vec + (size_t)id * p.dp, p.d, p.dp);
        if (fresh) {
          const int pos = __popc(fm & lt);
          pend_idx[pos] = id;
          pend_d[pos] = dist;
        }
        evals += np;
        __syncwarp();
        // ---- lane 0 admits them in connection order (hnsw.go:547-560) ----
        if (lane == 0) {
          for (int j = 0; j < np; ++j) {
            const float cd = pend_d[j];
            if (n_res < ef || cd < res[0].dist) {
              if (n_cand >= p.cand_cap) {
                ok = false;
                break;
              }
              const HRes r{pend_idx[j], cd};
              min_push(cand, n_cand, r);
              max_push(res, n_res, r);
              if (n_res > ef) (void)max_pop(res, n_res);
            }
          }
        }
        n_cand = __shfl_sync(0xffffffffu, n_cand, 0);
        n_res = __shfl_sync(0xffffffffu, n_res, 0);
        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
        if (!ok) return false;
      }
    }
  }
  n_res = __shfl_sync(0xffffffffu, n_res, 0);
  return ok;
}


constexpr int HNSW_WARPS = 16;

__global__ void __launch_bounds__(HNSW_WARPS * 32, 1) hnsw_search_kernel(const HnswKParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-warp slice: query | result heap (ef0 + 1) | candidate heap | pending ids / distances
  const size_t q_bytes = ((size_t)p.dp * 4 + 15) & ~(size_t)15;
  const size_t res_bytes = ((size_t)(p.ef0 + 2) * 8 + 15) & ~(size_t)15;
  const size_t per_warp = q_bytes + res_bytes + (size_t)p.cand_cap * 8 + 256;
  unsigned char* base = smem_raw + (size_t)warp * per_warp;
  float* q_s = reinterpret_cast<float*>(base);
  HRes* res = reinterpret_cast<HRes*>(base + q_bytes);
  HRes* cand = reinterpret_cast<HRes*>(base + q_bytes + res_bytes);
  uint32_t* pend_idx = reinterpret_cast<uint32_t*>(base + q_bytes + res_bytes + (size_t)p.cand_cap * 8);
  float* pend_d = reinterpret_cast<float*>(pend_idx + 32);
  const size_t slot = (size_t)blockIdx.x * HNSW_WARPS + warp;
  uint32_t* vis = p.visited + slot * (size_t)p.n_words;
  uint32_t* touched = p.touched + slot * (size_t)HNSW_TOUCH_CAP;
  int n_touched = 0;
  bool touched_overflow = false;

  for (;;) {
    int qi = 0;
    if (lane == 0) qi = atomicAdd(p.next_query, 1);
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi >= p.nq) break;
    for (int i = lane; i < p.dp; i += 32) q_s[i] = i < p.d ? __ldg(p.queries + (size_t)qi * p.d + i) : 0.f;
    __syncwarp();
    long long evals = 1;  // entryDistance (hnsw.go:637): evaluated by the reference, its value unused
    uint32_t ep = (uint32_t)p.g.entry_point;
    int n_res = 0;
    bool ok = true;
    // descent with ef = 1 (hnsw.go:640-657)
    for (int level = p.g.current_level; level > 0 && ok; --level) {
      ok = search_layer(p, q_s, ep, 1, level, cand, res, n_res, pend_idx, pend_d, vis, touched, n_touched,
                        touched_overflow, evals);
      if (ok && n_res > 0) {
        // results[0] of the ascending list = the smallest of the heap; with ef = 1 the heap holds one entry
        uint32_t best = 0;
        if (lane == 0) {
          int m = n_res;
          HRes last = res[0];
          while (m > 0) last = max_pop(res, m);
          best = last.idx;
        }
        ep = __shfl_sync(0xffffffffu, best, 0);
      }
    }
    // base layer with ef = max(EfSearch, k) (hnsw.go:660-668)
    if (ok)
      ok = search_layer(p, q_s, ep, p.ef0, 0, cand, res, n_res, pend_idx, pend_d, vis, touched, n_touched,
                        touched_overflow, evals);
    if (lane == 0) {
      uint32_t* oi = p.out_idx + (size_t)qi * p.kk;
      float* od = p.out_dist + (size_t)qi * p.kk;
      if (!ok) {
        p.out_count[qi] = -1;
      } else {
        // ascending list = the heap popped backwards (hnsw.go:567-577), truncated to k (:670-672)
        int m = n_res;
        const int total = n_res;
        for (int i = total - 1; i >= 0; --i) {
          const HRes r = max_pop(res, m);
          if (i < p.kk) {
            oi[i] = r.idx;
            od[i] = r.dist;
          }
        }
        for (int i = total; i < p.kk; ++i) {
          oi[i] = 0xFFFFFFFFu;
          od[i] = __int_as_float(0x7f800000);
        }
        p.out_count[qi] = total < p.kk ? total : p.kk;
      }
      if (p.out_evals) p.out_evals[qi] = evals;
    }
    __syncwarp();
  }
  // leave the bitset clean for the next launch
  if (touched_overflow) {
    for (long long i = lane; i < p.n_words; i += 32) vis[i] = 0u;
  } else {
    for (int i = lane; i < n_touched; i += 32) vis[touched[i]] = 0u;
  }
}

size_t hnsw_workspace_bytes(long long n_nodes, int sm_count) {
  const size_t slots = (size_t)sm_count * HNSW_WARPS;
  const size_t n_words = (size_t)((n_nodes + 31) / 32);
  return slots * n_words * 4 + slots * (size_t)HNSW_TOUCH_CAP * 4 + 256;
}

int launch_hnsw_search(const HnswDevGraph& g, const float* vec, int dp, int d, int metric, int arith,
                       const float* d_queries, int nq, int kk, int ef0, void* workspace, int sm_count, uint32_t* out_idx,
                       float* out_dist, int* out_count, long long* out_evals, cudaStream_t st) {
  if (nq <= 0) return 0;
  HnswKParams p{};
  p.g = g;
  p.vec = vec;
  p.dp = dp;
  p.d = d;
  p.metric = metric;
  p.arith = arith;
  p.queries = d_queries;
  p.nq = nq;
  p.kk = kk;
  p.ef0 = ef0;
  const size_t slots = (size_t)sm_count * HNSW_WARPS;
  p.n_words = (g.n_nodes + 31) / 32;
  p.visited = static_cast<uint32_t*>(workspace);
  p.touched = p.visited + slots * (size_t)p.n_words;
  p.next_query = reinterpret_cast<int*>(p.touched + slots * (size_t)HNSW_TOUCH_CAP);
  p.out_idx = out_idx;
  p.out_dist = out_dist;
  p.out_count = out_count;
  p.out_evals = out_evals;
  // shared memory per warp: query, result heap, pending slots, and the rest of ~13.5 KB for the candidate heap
  const size_t q_bytes = ((size_t)dp * 4 + 15) & ~(size_t)15;
  const size_t res_bytes = ((size_t)(ef0 + 2) * 8 + 15) & ~(size_t)15;
  const size_t budget = (size_t)216 * 1024 / HNSW_WARPS;
  if (q_bytes + res_bytes + 256 + 64 * 8 > budget)
    return fail(QG_ERR_UNSUPPORTED, "hnsw search: dimension / efSearch too large for the kernel's shared-memory slice");
  p.cand_cap = (int)((budget - q_bytes - res_bytes - 256) / 8);
  const size_t smem = (q_bytes + res_bytes + (size_t)p.cand_cap * 8 + 256) * HNSW_WARPS;
  static bool attr_done[64] = {};
  int dev = 0;
  QG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    QG_CUDA_OK(cudaFuncSetAttribute(hnsw_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done[dev] = true;
  }
  QG_CUDA_OK(cudaMemsetAsync(p.next_query, 0, 4, st));
  const int grid = (int)std::min<long long>(sm_count, (nq + HNSW_WARPS - 1) / HNSW_WARPS);
  hnsw_search_kernel<<<grid, HNSW_WARPS * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qg
