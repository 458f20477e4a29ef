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

#include <cub/device/device_radix_sort.cuh>

#include "exact.cuh"
#include "hnsw.cuh"

namespace qg {

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


constexpr int HNSW_WARPS = 16;  // the builder's kernels
#ifndef QG_HNSW_SEARCH_WARPS
#define QG_HNSW_SEARCH_WARPS 16
#endif
// walks in flight per SM. Measured at 1M x 128, efSearch 128 (10 000 queries): 16 warps 239-253 k queries/s,
// 24 warps 255 k, 32 warps (64 registers, smaller candidate heaps) 240 k: occupancy is not what bounds the walk.
constexpr int HNSW_SEARCH_WARPS = QG_HNSW_SEARCH_WARPS;
constexpr int HNSW_MAX_WARPS = HNSW_WARPS > HNSW_SEARCH_WARPS ? HNSW_WARPS : HNSW_SEARCH_WARPS;

__global__ void __launch_bounds__(HNSW_SEARCH_WARPS * 32, 1) hnsw_search_kernel(const HnswKParams p) {
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
  const size_t slot = (size_t)blockIdx.x * HNSW_SEARCH_WARPS + warp;
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

// ================================================================================================
// Graph construction on the device (reference hnsw.Insert / connectNode,
// pkg/hnsw/hnsw.go:266-468, selectNeighbors :583-599). The reference inserts one node at a time, each
// insert walking the graph its predecessors built; a graph IDENTICAL to the reference's therefore admits
// one insert in flight. Here nodes are inserted in batches: every node of a batch searches the graph of
// the already committed nodes (a warp per node, the same search_layer as above with ef = efConstruction),
// takes its closest M / MaxM0 results as neighbours (the reference's plain closest-k rule, ties by smaller
// index, not the paper's diversity heuristic), and the reverse links are applied afterwards — sorted by
// (level, neighbour, new node), one warp per neighbour list, pruned to the closest M / MaxM0 when a list
// overflows (the union's closest-k equals the reference's prune-after-every-append). Nodes of one batch
// do not see each other, so the graph differs from the sequential one: the bar is recall at equal
// efSearch against the host-built graph, not step identity (tests/test_gpu_hnsw_build.py). The descent
// follows the textbook rule (next layer starts at the closest node found); the reference's own rule
// (start at the node being inserted, which has no links there yet) fragments layer 0 — reproduced
// faithfully by the oracle's host builder, not here.
// ================================================================================================
struct HnswBuildKParams {
  HnswDevGraph g;          // level[] filled for ALL nodes; lists of uncommitted nodes are empty (0xFFFFFFFF)
  uint32_t* adj0_w;        // writable views of the lists
  uint32_t* upper_w;
  const float* vec;
  int dp, d, metric, arith;
  int ef_c, cand_cap;
  long long first;         // nodes [first, first + count) are inserted by this launch
  int count;
  uint32_t* visited;
  uint32_t* touched;
  long long n_words;
  int* next;
  unsigned long long* rev_key;  // (level << 60) | (neighbour << 30) | new node
  float* rev_dist;
  unsigned int* rev_count;
  unsigned int rev_cap;
};

__global__ void __launch_bounds__(HNSW_WARPS * 32, 1) hnsw_insert_kernel(const HnswBuildKParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t q_bytes = ((size_t)p.dp * 4 + 15) & ~(size_t)15;
  const size_t res_bytes = ((size_t)(p.ef_c + 2) * 8 + 15) & ~(size_t)15;
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
  // search_layer reads its parameters from an HnswKParams
  HnswKParams sp{};
  sp.g = p.g;
  sp.vec = p.vec;
  sp.dp = p.dp;
  sp.d = p.d;
  sp.metric = p.metric;
  sp.arith = p.arith;
  sp.cand_cap = p.cand_cap;
  sp.n_words = p.n_words;

  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(p.next, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= p.count) break;
    const uint32_t node = (uint32_t)(p.first + t);
    const float* row = p.vec + (size_t)node * p.dp;
    for (int i = lane; i < p.dp; i += 32) q_s[i] = __ldg(row + i);
    __syncwarp();
    int level = __ldg(p.g.level + node);
    long long evals = 0;
    uint32_t ep = (uint32_t)p.g.entry_point;
    int n_res = 0;
    // ef = 1 descent through the layers above the node's level (hnsw.go:357-372)
    for (int lc = p.g.current_level; lc > level; --lc) {
      search_layer(sp, q_s, ep, 1, lc, cand, res, n_res, pend_idx, pend_d, vis, touched, n_touched, touched_overflow, evals);
      if (n_res > 0) {
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
    // the node's own layers, top down (hnsw.go:375-460)
    for (int lc = level < p.g.current_level ? level : p.g.current_level; lc >= 0; --lc) {
      search_layer(sp, q_s, ep, p.ef_c, lc, cand, res, n_res, pend_idx, pend_d, vis, touched, n_touched, touched_overflow,
                   evals);  // (an overflowing candidate heap only truncates the search: the results so far are used)
      n_res = __shfl_sync(0xffffffffu, n_res, 0);
      if (n_res == 0) continue;
      const int total = n_res;
      HRes* out = cand;  // the candidate heap is dead: its slice takes the ascending result list
      if (lane == 0) {
        int m = total;
        for (int i = total - 1; i >= 0; --i) out[i] = max_pop(res, m);
        // selectNeighbors orders equal distances by index (hnsw.go:586-592)
        for (int i = 1; i < total; ++i) {
          const HRes x = out[i];
          int j = i - 1;
          while (j >= 0 && out[j].dist == x.dist && out[j].idx > x.idx) {
            out[j + 1] = out[j];
            --j;
          }
          out[j + 1] = x;
        }
      }
      __syncwarp();
      const int maxc = lc == 0 ? p.g.max_m0 : p.g.m;
      const int ns = total < maxc ? total : maxc;
      uint32_t* mine = lc == 0 ? p.adj0_w + (size_t)node * p.g.max_m0
                               : p.upper_w + p.g.upper_off[node] + (size_t)(lc - 1) * p.g.m;
      unsigned int rbase = 0;
      if (lane == 0) rbase = atomicAdd(p.rev_count, (unsigned int)ns);
      rbase = __shfl_sync(0xffffffffu, rbase, 0);
      for (int s = lane; s < ns; s += 32) {
        const HRes r = out[s];
        mine[s] = r.idx;
        const unsigned int pos = rbase + (unsigned int)s;
        if (pos < p.rev_cap) {
          p.rev_key[pos] = ((unsigned long long)lc << 60) | ((unsigned long long)r.idx << 30) | (unsigned long long)node;
          p.rev_dist[pos] = r.dist;
        }
      }
      ep = out[0].idx;  // next layer starts at the closest node found
      __syncwarp();
    }
  }
  if (touched_overflow) {
    for (long long i = lane; i < p.n_words; i += 32) vis[i] = 0u;
  } else {
    for (int i = lane; i < n_touched; i += 32) vis[touched[i]] = 0u;
  }
}

// ascending bitonic sort of 64 keys held two per lane: element e = lane in `a`, e = 32 + lane in `b`
__device__ __forceinline__ void warp_sort64(unsigned long long& a, unsigned long long& b) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j == 32) {  // partner is the other register of the same lane; k == 64: ascending everywhere
        if (a > b) {
          const unsigned long long t = a;
          a = b;
          b = t;
        }
      } else {
        const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, j);
        const unsigned long long pb = __shfl_xor_sync(0xffffffffu, b, j);
        const bool lower = (lane & j) == 0;                 // this element is the lower index of its pair
        const bool up_a = (lane & k) == 0;                  // direction of element `lane`
        const bool up_b = ((32 + lane) & k) == 0;           // direction of element `32 + lane`
        a = (lower == up_a) ? (a < pa ? a : pa) : (a > pa ? a : pa);
        b = (lower == up_b) ? (b < pb ? b : pb) : (b > pb ? b : pb);
      }
    }
  }
}

// Reverse links of a batch: records sorted by key = (level, neighbour, new node). One warp per record looks
// whether it leads a (level, neighbour) group and, if so, merges the whole group into that list.
__global__ void __launch_bounds__(256) hnsw_link_kernel(const HnswBuildKParams p, const unsigned long long* __restrict__ keys,
                                                        const unsigned int* __restrict__ perm, unsigned int n_rec) {
  const int lane = threadIdx.x & 31;
  const unsigned int r0 = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (r0 >= n_rec) return;
  const unsigned long long k0 = keys[r0];
  const unsigned long long grp = k0 >> 30;
  if (r0 > 0 && (keys[r0 - 1] >> 30) == grp) return;  // not the first record of its list
  const int lc = (int)(k0 >> 60);
  const uint32_t nb = (uint32_t)((k0 >> 30) & 0x3FFFFFFFu);
  if (lc > __ldg(p.g.level + nb)) return;  // the neighbour has no list at this level (hnsw.go:419-421)
  const int cap = lc == 0 ? p.g.max_m0 : p.g.m;
  uint32_t* lst = lc == 0 ? p.adj0_w + (size_t)nb * p.g.max_m0 : p.upper_w + p.g.upper_off[nb] + (size_t)(lc - 1) * p.g.m;
  // current list (<= cap <= 32 entries, 0xFFFFFFFF terminated): one entry per lane
  uint32_t id = lane < cap ? lst[lane] : 0xFFFFFFFFu;
  const unsigned have = __ballot_sync(0xffffffffu, id != 0xFFFFFFFFu);
  int n_old = have == 0xffffffffu ? 32 : __ffs((int)~have) - 1;
  if (lane >= n_old) id = 0xFFFFFFFFu;
  unsigned int g = 0;  // group size
  while (r0 + g < n_rec && (keys[r0 + g] >> 30) == grp) ++g;  // (uniform: every lane walks the same records)
  if (n_old + (int)g <= cap) {  // room for all: append in new-node order
    for (unsigned int j = lane; j < g; j += 32) lst[n_old + (int)j] = (uint32_t)(keys[r0 + j] & 0x3FFFFFFFu);
    return;
  }
  // overflow: keep the closest `cap` of (old members, new nodes) by (distance to nb, index) — selectNeighbors.
  // Distances of the old members are evaluated here, as the reference does when it prunes (hnsw.go:433-446).
  const float* vn = p.vec + (size_t)nb * p.dp;
  unsigned long long a = ~0ull;
  if (id != 0xFFFFFFFFu) {
    const float dist = exact_distance_lane(p.metric, p.arith, vn, p.vec + (size_t)id * p.dp, p.d, p.dp);
    a = ((unsigned long long)f32_to_ordered(dist) << 32) | id;
  }
  for (unsigned int c0 = 0; c0 < g; c0 += 32) {
    unsigned long long b = ~0ull;
    if (c0 + lane < g) {
      const unsigned int r = r0 + c0 + lane;
      b = ((unsigned long long)f32_to_ordered(p.rev_dist[perm[r]]) << 32) | (keys[r] & 0x3FFFFFFFu);
    }
    warp_sort64(a, b);
    if (lane >= cap) a = ~0ull;  // only the closest `cap` survive into the next round
  }
  if (lane < cap) lst[lane] = a == ~0ull ? 0xFFFFFFFFu : (uint32_t)a;
}

size_t hnsw_workspace_bytes(long long n_nodes, int sm_count) {
  const size_t slots = (size_t)sm_count * HNSW_MAX_WARPS;
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
  const size_t slots = (size_t)sm_count * HNSW_SEARCH_WARPS;
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
  const size_t budget = (size_t)216 * 1024 / HNSW_SEARCH_WARPS;
  if (q_bytes + res_bytes + 256 + 64 * 8 > budget)
    return fail(QG_ERR_UNSUPPORTED, "hnsw search: dimension / efSearch too large for the kernel's shared-memory slice");
  p.cand_cap = (int)((budget - q_bytes - res_bytes - 256) / 8);
  const size_t smem = (q_bytes + res_bytes + (size_t)p.cand_cap * 8 + 256) * HNSW_SEARCH_WARPS;
  static bool attr_done[64] = {};
  int dev = 0;
  QG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    QG_CUDA_OK(cudaFuncSetAttribute(hnsw_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done[dev] = true;
  }
  QG_CUDA_OK(cudaMemsetAsync(p.next_query, 0, 4, st));
  const int grid = (int)std::min<long long>(sm_count, (nq + HNSW_SEARCH_WARPS - 1) / HNSW_SEARCH_WARPS);
  hnsw_search_kernel<<<grid, HNSW_SEARCH_WARPS * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}


size_t hnsw_build_smem(int dp, int ef_c, int* cand_cap) {
  const size_t q_bytes = ((size_t)dp * 4 + 15) & ~(size_t)15;
  const size_t res_bytes = ((size_t)(ef_c + 2) * 8 + 15) & ~(size_t)15;
  const size_t budget = (size_t)216 * 1024 / HNSW_WARPS;
  if (q_bytes + res_bytes + 256 + (size_t)(ef_c + 8) * 8 > budget) return 0;
  *cand_cap = (int)((budget - q_bytes - res_bytes - 256) / 8);
  return (q_bytes + res_bytes + (size_t)*cand_cap * 8 + 256) * HNSW_WARPS;
}

int launch_hnsw_insert_batch(const HnswDevGraph& g, uint32_t* adj0_w, uint32_t* upper_w, const float* vec, int dp, int d,
                             int metric, int arith, int ef_c, long long first, int count, void* workspace, int sm_count,
                             unsigned long long* rev_key, float* rev_dist, unsigned int* rev_count, unsigned int rev_cap,
                             cudaStream_t st) {
  if (count <= 0) return 0;
  HnswBuildKParams p{};
  p.g = g;
  p.adj0_w = adj0_w;
  p.upper_w = upper_w;
  p.vec = vec;
  p.dp = dp;
  p.d = d;
  p.metric = metric;
  p.arith = arith;
  p.ef_c = ef_c;
  int cand_cap = 0;
  const size_t smem = hnsw_build_smem(dp, ef_c, &cand_cap);
  if (smem == 0) return fail(QG_ERR_UNSUPPORTED, "hnsw build: dimension / efConstruction too large for the kernel's shared-memory slice");
  p.cand_cap = cand_cap;
  p.first = first;
  p.count = count;
  const size_t slots = (size_t)sm_count * HNSW_WARPS;
  p.n_words = (g.n_nodes + 31) / 32;
  p.visited = static_cast<uint32_t*>(workspace);
  p.touched = p.visited + slots * (size_t)p.n_words;
  p.next = reinterpret_cast<int*>(p.touched + slots * (size_t)HNSW_TOUCH_CAP);
  p.rev_key = rev_key;
  p.rev_dist = rev_dist;
  p.rev_count = rev_count;
  p.rev_cap = rev_cap;
  static bool attr_done[64] = {};
  int dev = 0;
  QG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    QG_CUDA_OK(cudaFuncSetAttribute(hnsw_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done[dev] = true;
  }
  QG_CUDA_OK(cudaMemsetAsync(p.next, 0, 4, st));
  QG_CUDA_OK(cudaMemsetAsync(rev_count, 0, 4, st));
  const int grid = (int)std::min<long long>(sm_count, (count + HNSW_WARPS - 1) / HNSW_WARPS);
  hnsw_insert_kernel<<<grid, HNSW_WARPS * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_hnsw_link_batch(const HnswDevGraph& g, uint32_t* adj0_w, uint32_t* upper_w, const float* vec, int dp, int d,
                           int metric, int arith, const unsigned long long* sorted_keys, const unsigned int* perm,
                           const float* rev_dist, unsigned int n_rec, cudaStream_t st) {
  if (n_rec == 0) return 0;
  HnswBuildKParams p{};
  p.g = g;
  p.adj0_w = adj0_w;
  p.upper_w = upper_w;
  p.vec = vec;
  p.dp = dp;
  p.d = d;
  p.metric = metric;
  p.arith = arith;
  p.rev_dist = const_cast<float*>(rev_dist);
  hnsw_link_kernel<<<(n_rec + 7) / 8, 256, 0, st>>>(p, sorted_keys, perm, n_rec);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}


__global__ void iota_u32_kernel(unsigned int* p, unsigned int n) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i;
}

int hnsw_sort_records(const unsigned long long* keys, unsigned long long* keys_out, unsigned int* perm_in,
                      unsigned int* perm_out, unsigned int n, void** temp, size_t* temp_bytes, cudaStream_t st) {
  if (n == 0) return 0;
  iota_u32_kernel<<<(int)std::min<unsigned int>((n + 255) / 256, 1184u), 256, 0, st>>>(perm_in, n);
  QG_CUDA_OK(cudaGetLastError());
  size_t need = 0;
  QG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_out, perm_in, perm_out, (int)n, 0, 64, st));
  if (need > *temp_bytes) {
    QG_CUDA_OK(cudaStreamSynchronize(st));
    if (*temp) cudaFree(*temp);
    *temp = nullptr;
    *temp_bytes = 0;
    QG_CUDA_OK(cudaMalloc(temp, need));
    *temp_bytes = need;
  }
  size_t tb = *temp_bytes;
  QG_CUDA_OK(cub::DeviceRadixSort::SortPairs(*temp, tb, keys, keys_out, perm_in, perm_out, (int)n, 0, 64, st));
  return 0;
}

}  // namespace qg
