// finalize.cu — merge of the per-CTA candidate lists, exact re-rank and exactness certificate.
//
// Replaces `sort.Sort(&results); results = results[:k]` (reference pkg/hybrid/exact.go:124-129)
// for the candidates that survive the fp32 scan, and produces the float32 distances the
// reference would return by recomputing them in its arithmetic (exact.cuh). One CTA per query.
//
// Exactness: the scan ranks rows by an fp32 score whose error is bounded (gamma). After the
// re-rank the kernel derives T, the largest scan score a row could have and still belong to
// the exact top-k, re-ranks any listed candidate with score <= T that the first cut missed,
// and certifies the result only if every scan CTA whose list was full dropped nothing with
// score <= T. An uncertified query is reported with count -1 and redone by the exhaustive
// path, so recall is 1.0 by construction, not by a margin heuristic.
#include <cstdlib>

#include "exact.cuh"
#include "finalize.cuh"
#include "scan.cuh"
#include "select.cuh"

namespace qg {

constexpr int FIN_THREADS = 1024;
constexpr int FIN_WARPS = FIN_THREADS / 32;

__host__ __device__ constexpr int fin_slots(int kp) { return kp <= 512 ? 2048 : 4096; }

__host__ __device__ inline size_t finalize_smem(int kp) {
  return (size_t)fin_slots(kp) * 8 * 2 + (size_t)FIN_WARPS * EXACT_SCRATCH_BYTES + 64;
}

// Largest fp32 scan score a row with exact (reference-arithmetic) distance <= E can have.
__device__ __forceinline__ float score_upper_bound(int metric, int mode, int cosine, float E, double gamma,
                                                   double qnorm2, double max_norm2) {
  double t;
  if (mode == MODE_L2) {
    if (metric == METRIC_SQL2) {
      t = (double)E * (1.0 + 2.0 * gamma + 4e-7);
    } else {  // E = fl32(sqrt(sum)); the score is the fp32 sum of squares
      const double e = (double)E * (1.0 + 1.2e-7);
      t = e * e * (1.0 + gamma + 2e-7);
    }
    t += 1e-37;
  } else if (mode == MODE_L1) {
    t = (double)E * (1.0 + gamma + 4e-7) + 1e-37;
  } else if (cosine) {
    t = (double)E + gamma + 2e-6;
  } else {
    const double b = gamma * sqrt(qnorm2 * max_norm2);
    t = (double)E + b + (fabs((double)E) + 1.0 + b) * 4e-7;
  }
  float tf = (float)t;
  if ((double)tf < t) tf = nextafterf(tf, __int_as_float(0x7f800000));
  return tf;
}

// Best kp keys of the nb sorted lists into pool[0..*cnt) (sorted). Keys are read in rank-major
// rounds of one key per thread, the next round's key already in flight while the current one is
// filtered. pool has `slots` >= kp + FIN_THREADS entries. All threads of the block call.
__device__ __forceinline__ void generic_select(const uint64_t* part, int total, int nb, int kp, uint64_t* pool,
                                               int slots, int* s_cnt, float* s_tau) {
  const int tid = threadIdx.x;
  const int highwater = slots - FIN_THREADS;
  auto load_key = [&](int i) -> uint64_t {
    const uint64_t key = i < total ? __ldcg(part + (size_t)(i % nb) * kp + (i / nb)) : KEY_NONE;
    return key_row(key) == XCHG_ROW ? KEY_NONE : key;  // the threshold marker of a cut list is not a candidate
  };
  uint64_t next_key = load_key(tid);
  __syncthreads();
  if (tid == 0) {
    *s_cnt = 0;
    *s_tau = __int_as_float(0x7f800000);
  }
  __syncthreads();
  PoolRef pr{pool, s_cnt, s_tau};
  for (int base = 0; base < total; base += FIN_THREADS) {
    const uint64_t key = next_key;
    next_key = load_key(base + FIN_THREADS + tid);
    const bool pass = (key != KEY_NONE) && (key_score(key) <= *s_tau);
    warp_append(pr, pass, key);
    __syncthreads();
    if (*s_cnt >= highwater || (base == 0 && *s_cnt > kp)) block_prune(pr, kp);
  }
  block_prune(pr, kp);
}

// One CTA (1024 threads) per query. The nb*kp scan keys are read in rank-major rounds of one key
// per thread, the next round's key already in flight while the current one is filtered, so the
// whole selection costs a handful of L2 round trips instead of one per round.
__global__ void __launch_bounds__(FIN_THREADS, 1) finalize_kernel(const FinalizeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int q = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kp = p.kp, nb = p.nb, k = p.k;
  const int slots = fin_slots(kp);
  uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* ex = pool + slots;
  double* scratch = reinterpret_cast<double*>(ex + slots) + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  __shared__ int s_cnt;
  __shared__ float s_tau;
  __shared__ int s_extra;
  __shared__ int s_bad;
  __shared__ double s_qn2;

  const uint64_t* part = p.partial + (size_t)q * nb * kp;
  const float* qv = p.queries + (size_t)q * p.dp;
  const int total = nb * kp;
  // rank-major order: index i -> list i % nb, rank i / nb (so the threshold tightens early)
  if (tid == 0) {
    s_extra = 0;
    s_bad = 0;
  }
  if (warp == 1) {  // |q|^2 for the dot-product bound
    double s = 0.0;
    for (int i = lane; i < p.d; i += 32) s += (double)qv[i] * (double)qv[i];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qn2 = s;
  }
  __syncthreads();

  // ---- 1. best kp scan keys among nb*kp ----------------------------------------------------------
  generic_select(part, total, nb, kp, pool, slots, &s_cnt, &s_tau);
  const int ncand = s_cnt;
  const uint64_t last_key = ncand > 0 ? pool[ncand - 1] : 0ull;

  // ---- 2. exact re-rank ----------------------------------------------------------------------------
  for (int c = warp; c < ncand; c += FIN_WARPS) {
    const uint32_t row = key_row(pool[c]);
    const float dist = exact_distance_warp(p.metric, p.arith, qv, p.vec + (size_t)row * p.dp, p.d, scratch);
    if (lane == 0) ex[c] = make_key(dist, row);
  }
  __syncthreads();
  int nex = ncand;
  {
    const int n2 = next_pow2(nex);
    for (int i = nex + tid; i < n2; i += FIN_THREADS) ex[i] = KEY_NONE;
    block_bitonic_sort(ex, n2);
  }

  // ---- 3. certificate and widening -----------------------------------------------------------------
  if (ncand >= k && ncand > 0) {
    const float E = key_score(ex[k - 1]);
    const float T = score_upper_bound(p.metric, p.mode, p.cosine, E, (double)p.gamma, s_qn2,
                                      (double)(p.max_norm2 ? *p.max_norm2 : 0.f));
    // every list is sorted: a list can only hold something <= T if its rank-0 key is <= T, and its
    // last key decides whether the scan CTA may have dropped such rows.
    for (int b = tid; b < nb; b += FIN_THREADS) {
      const uint64_t* lst = part + (size_t)b * kp;
      const uint64_t lastk = __ldcg(lst + kp - 1);
      // a full list — or one cut by an exchanged threshold, whose marker sits here — may have dropped rows <= T
      if (lastk != KEY_NONE && key_score(lastk) <= T) s_bad = 1;
      for (int j = 0; j < kp; ++j) {
        const uint64_t key = __ldcg(lst + j);
        if (key_row(key) == XCHG_ROW || key_score(key) > T) break;  // end of list, or its threshold marker
        if (key > last_key) {
          const int pos = atomicAdd(&s_extra, 1);
          if (ncand + pos < slots) pool[ncand + pos] = key;
        }
      }
    }
    __syncthreads();
    int extra = s_extra;
    if (ncand + extra > slots) {
      extra = slots - ncand;
      if (tid == 0) s_bad = 1;
    }
    if (extra > 0) {
      for (int c = warp; c < extra; c += FIN_WARPS) {
        const uint32_t row = key_row(pool[ncand + c]);
        const float dist = exact_distance_warp(p.metric, p.arith, qv, p.vec + (size_t)row * p.dp, p.d, scratch);
        if (lane == 0) ex[ncand + c] = make_key(dist, row);
      }
      __syncthreads();
      nex = ncand + extra;
      const int n2 = next_pow2(nex);
      for (int i = nex + tid; i < n2; i += FIN_THREADS) ex[i] = KEY_NONE;
      block_bitonic_sort(ex, n2);
    }
  } else {
    // fewer candidates than k: nothing may have been dropped anywhere
    for (int b = tid; b < nb; b += FIN_THREADS)
      if (__ldcg(part + (size_t)b * kp + kp - 1) != KEY_NONE) s_bad = 1;
  }
  __syncthreads();

  // ---- 4. results ----------------------------------------------------------------------------------
  const int kk = nex < k ? nex : k;
  const bool bad = s_bad != 0;
  if (p.out_keys != nullptr) {
    for (int j = tid; j < k; j += FIN_THREADS) {
      uint64_t o = KEY_NONE;
      if (j < kk && !bad) o = (ex[j] & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)(p.row_base + key_row(ex[j]));
      p.out_keys[(size_t)q * k + j] = o;
    }
    if (tid == 0 && p.out_count) p.out_count[q] = bad ? -1 : kk;
    return;
  }
  for (int j = tid; j < k; j += FIN_THREADS) {
    const bool ok = j < kk;
    p.out_dist[(size_t)q * k + j] = ok ? key_score(ex[j]) : __int_as_float(0x7f800000);
    p.out_row[(size_t)q * k + j] = ok ? (long long)key_row(ex[j]) + p.row_base : -1ll;
  }
  if (p.out_negdist != nullptr) {
    const float* nv = p.negatives + (size_t)q * p.dp;
    for (int j = warp; j < k; j += FIN_WARPS) {
      float nd = __int_as_float(0x7f800000);
      if (j < kk) {
        // hybrid_index.go:544  negDistance: idx.distFunc(vector, negExample)
        nd = exact_distance_warp(p.metric, p.arith, p.vec + (size_t)key_row(ex[j]) * p.dp, nv, p.d, scratch);
      }
      if (lane == 0) p.out_negdist[(size_t)q * k + j] = nd;
    }
  }
  if (tid == 0) p.out_count[q] = bad ? -1 : kk;
}

// ------------------------------------------------------------------------------------------------
// Low-latency variant for kp <= 128 (k <= 112): every thread keeps its share of the nb*kp scan
// keys in registers (one L2 round trip), the threshold, the selection and both sorts are done by
// all-pairs rank counting (a handful of barriers instead of ~100 bitonic steps), and the
// certificate re-uses the register-resident keys.
// ------------------------------------------------------------------------------------------------
constexpr int FF_MAXR = 19;        // keys per thread: 148 lists * 128 keys / 1024 threads
constexpr int FF_SAMPLE = 256;     // keys used to derive the selection threshold
constexpr int FF_CAP = 1024;       // candidate capacity

__host__ __device__ inline size_t finalize_fast_smem() {
  return (size_t)FF_SAMPLE * 8 + (size_t)FF_CAP * 4 + (size_t)FF_CAP * 8 * 3 +
         (size_t)FIN_WARPS * EXACT_SCRATCH_BYTES + 64;
}

// dst[rank of src[i]] = src[i] for n <= FF_CAP distinct keys; all FIN_THREADS threads call.
__device__ __forceinline__ void block_rank_sort(const uint64_t* src, int n, uint64_t* dst, int* rank) {
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += FIN_THREADS) rank[i] = 0;
  __syncthreads();
  if (n > 0) {
    const int nparts = FIN_THREADS / n > 0 ? FIN_THREADS / n : 1;
    const int chunk = (n + nparts - 1) / nparts;
    const int a = tid % n, part = tid / n;
    if (part < nparts) {
      const uint64_t mine = src[a];
      const int lo = part * chunk, hi = min(n, lo + chunk);
      int c = 0;
      for (int j = lo; j < hi; ++j) c += src[j] < mine;
      if (c) atomicAdd(&rank[a], c);
    }
  }
  __syncthreads();
  if (tid < n) dst[rank[tid]] = src[tid];
  __syncthreads();
}

template <int FF_R>
__global__ void __launch_bounds__(FIN_THREADS, 1) finalize_fast_kernel(const FinalizeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int q = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kp = p.kp, nb = p.nb, k = p.k;

  uint64_t* samp = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* cand = samp + FF_SAMPLE;
  uint64_t* sel = cand + FF_CAP;
  uint64_t* ex = sel + FF_CAP;
  int* rank = reinterpret_cast<int*>(ex + FF_CAP);
  double* scratch = reinterpret_cast<double*>(rank + FF_CAP) + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  __shared__ int s_ncand, s_extra, s_bad, s_gcnt;
  __shared__ float s_gtau;
  __shared__ unsigned long long s_taukey;
  __shared__ double s_qn2;

  const uint64_t* part = p.partial + (size_t)q * nb * kp;
  const float* qv = p.queries + (size_t)q * p.dp;
  const int total = nb * kp;

  // ---- 0. all keys into registers, rank-major (index i -> list i % nb, rank i / nb) -------------
  uint64_t keys[FF_R];
#pragma unroll
  for (int r = 0; r < FF_R; ++r) {
    const int i = r * FIN_THREADS + tid;
    keys[r] = i < total ? __ldcg(part + (size_t)(i % nb) * kp + (i / nb)) : KEY_NONE;
  }
  if (tid == 0) {
    s_ncand = 0;
    s_extra = 0;
    s_bad = 0;
    s_taukey = KEY_NONE;
  }
  if (warp == 1) {  // |q|^2 for the dot-product bound
    double s = 0.0;
    for (int i = lane; i < p.d; i += 32) s += (double)qv[i] * (double)qv[i];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qn2 = s;
  }
  if (tid < FF_SAMPLE) {
    samp[tid] = key_row(keys[0]) == XCHG_ROW ? KEY_NONE : keys[0];
    rank[tid] = 0;
  }
  __syncthreads();

  // ---- 1. threshold: the kp-th smallest of the first 256 keys bounds the global kp-th -------------
  if (total >= FF_SAMPLE) {
    const int a = tid & (FF_SAMPLE - 1), part_i = tid >> 8;  // 4 parts of 64 comparisons
    const uint64_t mine = samp[a];
    int c = 0;
    for (int j = part_i * 64; j < part_i * 64 + 64; ++j) c += samp[j] < mine;
    if (c) atomicAdd(&rank[a], c);
    __syncthreads();
    if (tid < FF_SAMPLE && rank[tid] == kp - 1 && samp[tid] != KEY_NONE) s_taukey = samp[tid];
    __syncthreads();
  }
  const uint64_t taukey = s_taukey;

  // ---- 2. candidates = keys <= threshold ---------------------------------------------------------
#pragma unroll
  for (int r = 0; r < FF_R; ++r) {
    if (key_row(keys[r]) != XCHG_ROW && keys[r] <= taukey) {  // neither padding nor a threshold marker
      const int pos = atomicAdd(&s_ncand, 1);
      if (pos < FF_CAP) cand[pos] = keys[r];
    }
  }
  __syncthreads();
  int ncand = s_ncand;
  int nsel;
  if (ncand > FF_CAP) {
    // very uneven lists (the sample held fewer than kp keys): general pool selection instead.
    // cand and sel are contiguous: 2 * FF_CAP slots.
    generic_select(part, total, nb, kp, cand, 2 * FF_CAP, &s_gcnt, &s_gtau);
    nsel = s_gcnt;
    __syncthreads();
    uint64_t tmp = tid < nsel ? cand[tid] : 0ull;
    __syncthreads();
    if (tid < nsel) sel[tid] = tmp;
    __syncthreads();
  } else {
    // ---- 3. sort candidates by scan key, keep the best kp --------------------------------------------
    block_rank_sort(cand, ncand, sel, rank);
    nsel = ncand < kp ? ncand : kp;
  }
  const uint64_t last_key = nsel > 0 ? sel[nsel - 1] : 0ull;

  // ---- 4. exact re-rank --------------------------------------------------------------------------------
  for (int c = warp; c < nsel; c += FIN_WARPS) {
    const uint32_t row = key_row(sel[c]);
    const float dist = exact_distance_warp(p.metric, p.arith, qv, p.vec + (size_t)row * p.dp, p.d, scratch);
    if (lane == 0) cand[c] = make_key(dist, row);
  }
  __syncthreads();
  int nex = nsel;
  block_rank_sort(cand, nex, ex, rank);

  // ---- 5. certificate and widening --------------------------------------------------------------------
  if (nsel >= k && nsel > 0) {
    const float E = key_score(ex[k - 1]);
    const float T = score_upper_bound(p.metric, p.mode, p.cosine, E, (double)p.gamma, s_qn2,
                                      (double)(p.max_norm2 ? *p.max_norm2 : 0.f));
#pragma unroll
    for (int r = 0; r < FF_R; ++r) {
      const uint64_t key = keys[r];
      if (key != KEY_NONE && key_score(key) <= T) {
        const int i = r * FIN_THREADS + tid;
        if (i / nb == kp - 1) s_bad = 1;  // a full (or threshold-cut) list may have dropped rows with score <= T
        if (key > last_key && key_row(key) != XCHG_ROW) {
          const int pos = atomicAdd(&s_extra, 1);
          if (nsel + pos < FF_CAP) sel[nsel + pos] = key;
        }
      }
    }
    __syncthreads();
    int extra = s_extra;
    if (nsel + extra > FF_CAP) {
      extra = FF_CAP - nsel;
      if (tid == 0) s_bad = 1;
    }
    if (extra > 0) {
      for (int c = tid; c < nsel; c += FIN_THREADS) cand[c] = ex[c];
      for (int c = warp; c < extra; c += FIN_WARPS) {
        const uint32_t row = key_row(sel[nsel + c]);
        const float dist = exact_distance_warp(p.metric, p.arith, qv, p.vec + (size_t)row * p.dp, p.d, scratch);
        if (lane == 0) cand[nsel + c] = make_key(dist, row);
      }
      __syncthreads();
      nex = nsel + extra;
      block_rank_sort(cand, nex, ex, rank);
    }
  } else {
    // fewer candidates than k: nothing may have been dropped anywhere
#pragma unroll
    for (int r = 0; r < FF_R; ++r) {
      const int i = r * FIN_THREADS + tid;
      if (keys[r] != KEY_NONE && i / nb == kp - 1) s_bad = 1;
    }
  }
  __syncthreads();

  // ---- 6. results -------------------------------------------------------------------------------------
  const int kk = nex < k ? nex : k;
  const bool bad = s_bad != 0;
  if (p.out_keys != nullptr) {
    for (int j = tid; j < k; j += FIN_THREADS) {
      uint64_t o = KEY_NONE;
      if (j < kk && !bad) o = (ex[j] & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)(p.row_base + key_row(ex[j]));
      p.out_keys[(size_t)q * k + j] = o;
    }
    if (tid == 0 && p.out_count) p.out_count[q] = bad ? -1 : kk;
    return;
  }
  for (int j = tid; j < k; j += FIN_THREADS) {
    const bool ok = j < kk;
    p.out_dist[(size_t)q * k + j] = ok ? key_score(ex[j]) : __int_as_float(0x7f800000);
    p.out_row[(size_t)q * k + j] = ok ? (long long)key_row(ex[j]) + p.row_base : -1ll;
  }
  if (p.out_negdist != nullptr) {
    const float* nv = p.negatives + (size_t)q * p.dp;
    for (int j = warp; j < k; j += FIN_WARPS) {
      float nd = __int_as_float(0x7f800000);
      if (j < kk) {
        // hybrid_index.go:544  negDistance: idx.distFunc(vector, negExample)
        nd = exact_distance_warp(p.metric, p.arith, p.vec + (size_t)key_row(ex[j]) * p.dp, nv, p.d, scratch);
      }
      if (lane == 0) p.out_negdist[(size_t)q * k + j] = nd;
    }
  }
  if (tid == 0) p.out_count[q] = bad ? -1 : kk;
}

__global__ void finalize_cand_kernel(const FinalizeCandParams cp);
__global__ void merge_shards_kernel(const uint64_t* __restrict__ keys, int world, int nq, int k, float* out_dist,
                                    long long* out_row, int* out_count);

int finalize_set_attributes() {
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
  // same L1 / shared-memory split as the scan kernels, so that back-to-back launches of a pass do
  // not make the SMs reconfigure their carve-out
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_cand_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
  QG_CUDA_OK(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)finalize_smem(1024)));
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_fast_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)finalize_fast_smem()));
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_fast_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)finalize_fast_smem()));
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_fast_kernel<FF_MAXR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)finalize_fast_smem()));
  return 0;
}

int launch_finalize(const FinalizeParams& p, int nq, cudaStream_t st) {
  if (nq <= 0) return 0;
  const long long per_thread = ((long long)p.nb * p.kp + FIN_THREADS - 1) / FIN_THREADS;
  if (p.kp <= 128 && per_thread <= 5)
    finalize_fast_kernel<5><<<nq, FIN_THREADS, finalize_fast_smem(), st>>>(p);
  else if (p.kp <= 128 && per_thread <= 10)
    finalize_fast_kernel<10><<<nq, FIN_THREADS, finalize_fast_smem(), st>>>(p);
  else if (p.kp <= 128 && per_thread <= FF_MAXR)
    finalize_fast_kernel<FF_MAXR><<<nq, FIN_THREADS, finalize_fast_smem(), st>>>(p);
  else
    finalize_kernel<<<nq, FIN_THREADS, finalize_smem(p.kp), st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tensor-core regime: candidates admitted by `score <= tau` (tc_scan.cu), one unordered list per
// query. Sort by scan key, re-rank the best kp exactly, derive T (largest tf32 scan score a row of
// the exact top-k can have), re-rank every further candidate with score <= T, and certify only if
// T <= tau (so every such row was admitted) and the list did not overflow.
// ------------------------------------------------------------------------------------------------
constexpr int FC_THREADS = 512;
constexpr int FC_WARPS = FC_THREADS / 32;
constexpr int FC_RANK_MAX = 1024;  // up to this many candidates are ordered by all-pairs rank counting

__host__ __device__ inline size_t finalize_cand_smem(int cap) {
  return (size_t)cap * 8 * 2 + (size_t)FC_WARPS * EXACT_SCRATCH_BYTES + 64;
}

// Largest tensor-core scan score of a row whose exact (reference-arithmetic) distance is <= E.
// L2 scores are |x|^2 - 2 q.x (the |q|^2 term is dropped by the scan).
__device__ __forceinline__ float tc_score_upper_bound(int metric, int mode, int cosine, float E, double g,
                                                      double gamma_seq, double qn2, double max_norm2) {
  const double qx = sqrt(qn2 * max_norm2);
  double t;
  if (mode == MODE_L2) {
    double d2;
    if (metric == METRIC_SQL2) {
      d2 = (double)E * (1.0 + 2.0 * gamma_seq + 4e-7);
    } else {
      const double e = (double)E * (1.0 + 1.2e-7);
      d2 = e * e * (1.0 + 3e-7);
    }
    t = d2 - qn2 + 2.0 * g * qx + 2.4e-7 * (max_norm2 + qn2 + 2.0 * qx) + 1e-30;
  } else if (cosine) {
    t = (double)E + g + 2e-6;
  } else {
    t = (double)E + g * qx + (fabs((double)E) + 1.0 + qx) * 4e-7;
  }
  float tf = (float)t;
  if ((double)tf < t) tf = nextafterf(tf, __int_as_float(0x7f800000));
  return tf;
}

// Smallest 32-bit score image P such that at least `want` keys have an image <= P: radix select, four
// bits per round (eight block-wide rounds whatever n is, unlike all-pairs ranking). The keys stay in
// registers: thread t owns key t, t + FC_THREADS, ... (KEY_NONE beyond n). Each round histograms the
// next digit of the keys that still match the prefix (warp-aggregated with match.any, one shared
// atomic per distinct digit and warp), then every thread walks the 16 counters itself. s_hist[48] is
// scratch: three rotating histograms — the one cleared after a round's barrier was last read before
// that barrier and is next written after the following one. All threads of the block call.
template <int PER>
__device__ __forceinline__ uint32_t fc_select_pivot(const uint64_t (&mine)[PER], int n, int want, int* s_hist) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0;
  int below = 0;  // keys below the prefix range
  if (threadIdx.x < 48) s_hist[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < 8; ++r) {
    const int shift = 28 - 4 * r;
    int* h = s_hist + (r % 3) * 16;
#pragma unroll
    for (int s = 0; s < PER; ++s) {
      if (s * FC_THREADS >= n) break;  // block-uniform: slots beyond n hold no key in any thread
      const uint32_t img = (uint32_t)(mine[s] >> 32);
      // r == 0: every key matches; later: the bits above the digit must equal the prefix
      const bool active = mine[s] != KEY_NONE && (r == 0 || (img >> (shift + 4)) == (prefix >> (shift + 4)));
      const int digit = active ? (int)((img >> shift) & 15u) : 16;
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (active && lane == __ffs(peers) - 1) atomicAdd(h + digit, __popc(peers));
    }
    __syncthreads();
    // lanes 0..15 take one counter each; inclusive warp scan; the first digit whose running count
    // reaches `want` is the next digit of the pivot
    const int c = lane < 16 ? h[lane] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    const unsigned reach = __ballot_sync(0xffffffffu, lane < 16 && below + inc >= want);
    const int pick = reach ? __ffs(reach) - 1 : 15;
    const int before = __shfl_sync(0xffffffffu, inc - c, pick);
    prefix |= (uint32_t)pick << shift;
    below += before;
    if (threadIdx.x < 16) s_hist[((r + 2) % 3) * 16 + threadIdx.x] = 0;
  }
  __syncthreads();
  return prefix;
}

__global__ void __launch_bounds__(FC_THREADS, 2) finalize_cand_kernel(const FinalizeCandParams cp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FinalizeParams& p = cp.base;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cap = cp.cap, k = p.k;
  uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);  // selected scan keys (unordered), later the output order
  uint64_t* ex = sel + cap;                               // exact keys
  double* scratch = reinterpret_cast<double*>(ex + cap) + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  __shared__ double s_qn2;
  __shared__ int s_extra;
  __shared__ int s_sel[48];
  __shared__ float s_E;

  long long ts[8];
  int nts = 0;
  ts[nts++] = clock64();
  pdl_launch_dependents();
  pdl_wait();  // candidate lists, counts and thresholds come from the scan / threshold kernels just before
  const int n_raw = cp.cand_cnt[q];
  const bool overflow = n_raw > cap;
  const int n = overflow ? cap : n_raw;
  const float tau = cp.tau[q];
  const bool all_admitted = tau == __int_as_float(0x7f800000);
  const float* qv = p.queries + (size_t)q * p.dp;

  if (warp == 0) {
    double s = 0.0;
    for (int i = lane; i < p.d; i += 32) s += (double)qv[i] * (double)qv[i];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qn2 = s;
  }
  if (tid == 0) s_extra = 0;
  const uint64_t* src = cp.cand + (size_t)q * cap;
  constexpr int PER = 4;  // keys per thread: cap <= PER * FC_THREADS
  uint64_t mine[PER];
#pragma unroll
  for (int s = 0; s < PER; ++s) {
    const int i = tid + s * FC_THREADS;
    mine[s] = i < n ? __ldcg(src + i) : KEY_NONE;
  }
  // the best scan keys = every key whose score image is <= the pivot: at least min(kp, n) keys
  // (ties on the score image may add a few)
  const int want = n < cp.kp ? n : cp.kp;
  long long t_keys = 0, t_pivot = 0;
  if (cp.dbg != nullptr) {
    t_keys = (long long)(mine[0] & 1) + clock64();  // (depends on the first key: stamped once it has arrived)
  }
  const uint32_t pivot = want > 0 ? fc_select_pivot<PER>(mine, n, want, s_sel) : 0u;
  if (cp.dbg != nullptr) t_pivot = clock64();
  if (tid == 0) s_sel[0] = 0;
  __syncthreads();
#pragma unroll
  for (int s = 0; s < PER; ++s) {
    if (want > 0 && mine[s] != KEY_NONE && (uint32_t)(mine[s] >> 32) <= pivot) {
      const int pos = atomicAdd(&s_sel[0], 1);
      sel[pos] = mine[s];
    }
  }
  __syncthreads();
  const int nsel = want > 0 ? s_sel[0] : 0;
  ts[nts++] = clock64();  // keys loaded and selected

  // ---- exact re-rank of the selected candidates (one warp per candidate) ----
  // the rows are scattered over the corpus: the warp's later candidates are pulled towards L2 / L1 first
  const bool lane_rerank = cp.lane_rerank != 0;
  if (lane_rerank) {
    // one LANE per candidate: every lane runs the reference's sequential sum for its own row (32 independent
    // chains per warp), a few warps do all the work — a twentieth of the instructions of the warp-per-candidate
    // form, whose single summing lane keeps 31 others waiting
    for (int c = tid; c < nsel; c += FC_THREADS) {
      const uint32_t rid = key_row(sel[c]);
      ex[c] = make_key(exact_distance_lane(p.metric, p.arith, qv, p.vec + (size_t)rid * p.dp, p.d, p.dp), rid);
    }
  }
  for (int c = warp + FC_WARPS; !lane_rerank && c < nsel; c += FC_WARPS) {
    const char* rowp = reinterpret_cast<const char*>(p.vec + (size_t)key_row(sel[c]) * p.dp);
    if (lane * 128 < p.dp * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + lane * 128));
  }
  for (int c = warp; !lane_rerank && c < nsel; c += 3 * FC_WARPS) {  // three candidates of the warp per pass
    const float* rows3[3];
    uint32_t rid[3];
    int nb = 0;
    for (int j = 0; j < 3 && c + j * FC_WARPS < nsel; ++j, ++nb) {
      rid[j] = key_row(sel[c + j * FC_WARPS]);
      rows3[j] = p.vec + (size_t)rid[j] * p.dp;
    }
    float dist3[3];
    exact_distance_warp3(p.metric, p.arith, qv, rows3, nb, p.d, scratch, dist3);
    if (lane == 0)
      for (int j = 0; j < nb; ++j) ex[c + j * FC_WARPS] = make_key(dist3[j], rid[j]);
  }
  __syncthreads();
  ts[nts++] = clock64();  // first re-rank done

  bool certified = !overflow;
  int nex = nsel;
  if (nsel >= k) {
    // E = k-th smallest exact distance among the re-ranked candidates (all-pairs rank, nsel is small)
    for (int i = tid; i < nsel; i += FC_THREADS) {
      const uint64_t me = ex[i];
      int r = 0;
      for (int j = 0; j < nsel; ++j) r += ex[j] < me;
      if (r == k - 1) s_E = key_score(me);
    }
    __syncthreads();
    float T = tc_score_upper_bound(p.metric, p.mode, p.cosine, s_E, cp.tc_gamma, (double)p.gamma, s_qn2,
                                   (double)(p.max_norm2 ? *p.max_norm2 : 0.f));
    if (cp.tc_norm_gamma > 0.0) {  // raw L2 scan: the norm term is accumulated by the tensor core as well
      const double t2 = (double)T + cp.tc_norm_gamma * (double)(p.max_norm2 ? *p.max_norm2 : 0.f);
      T = (float)t2;
      if ((double)T < t2) T = nextafterf(T, __int_as_float(0x7f800000));
    }
    if (!(T <= tau) && !all_admitted) certified = false;
    // every further candidate whose scan score is <= T may still belong to the exact top-k
#pragma unroll
    for (int s = 0; s < PER; ++s) {
      if (mine[s] != KEY_NONE && (uint32_t)(mine[s] >> 32) > pivot && key_score(mine[s]) <= T) {
        const int pos = atomicAdd(&s_extra, 1);
        if (nsel + pos < cap) ex[nsel + pos] = mine[s];  // scan key for now; replaced by the exact key below
      }
    }
    __syncthreads();
    int extra = s_extra;
    if (nsel + extra > cap) {
      extra = cap - nsel;
      certified = false;
    }
    if (lane_rerank) {
      for (int c = tid; c < extra; c += FC_THREADS) {
        const uint32_t rid = key_row(ex[nsel + c]);
        ex[nsel + c] = make_key(exact_distance_lane(p.metric, p.arith, qv, p.vec + (size_t)rid * p.dp, p.d, p.dp), rid);
      }
    }
    for (int c = warp; !lane_rerank && c < extra; c += 3 * FC_WARPS) {
      const float* rows3[3];
      uint32_t rid[3];
      int nb = 0;
      for (int j = 0; j < 3 && c + j * FC_WARPS < extra; ++j, ++nb) {
        rid[j] = key_row(ex[nsel + c + j * FC_WARPS]);
        rows3[j] = p.vec + (size_t)rid[j] * p.dp;
      }
      float dist3[3];
      exact_distance_warp3(p.metric, p.arith, qv, rows3, nb, p.d, scratch, dist3);
      __syncwarp();
      if (lane == 0)
        for (int j = 0; j < nb; ++j) ex[nsel + c + j * FC_WARPS] = make_key(dist3[j], rid[j]);
    }
    __syncthreads();
    nex = nsel + extra;
  } else {
    // fewer than k candidates: complete only if the scan admitted every row
    if (!all_admitted) certified = false;
  }
  ts[nts++] = clock64();  // certificate + extras done

  // ---- order the exact keys and emit the first k ----
  const int kk = nex < k ? nex : k;
  uint64_t* outk = sel;  // the scan keys are no longer needed
  if (nex <= FC_RANK_MAX) {
    for (int i = tid; i < nex; i += FC_THREADS) {
      const uint64_t me = ex[i];
      int r = 0;
      for (int j = 0; j < nex; ++j) r += ex[j] < me;
      if (r < k) outk[r] = me;
    }
    __syncthreads();
  } else {
    const int m2 = next_pow2(nex);
    for (int i = nex + tid; i < m2; i += FC_THREADS) ex[i] = KEY_NONE;
    block_bitonic_sort(ex, m2);
    outk = ex;
  }

  if (p.out_keys != nullptr) {
    for (int j = tid; j < k; j += FC_THREADS) {
      uint64_t o = KEY_NONE;
      if (j < kk && certified) o = (outk[j] & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)(p.row_base + key_row(outk[j]));
      p.out_keys[(size_t)q * k + j] = o;
    }
    if (tid == 0 && p.out_count) p.out_count[q] = certified ? kk : -1;
  } else {
    for (int j = tid; j < k; j += FC_THREADS) {
      const bool ok = j < kk;
      p.out_dist[(size_t)q * k + j] = ok ? key_score(outk[j]) : __int_as_float(0x7f800000);
      p.out_row[(size_t)q * k + j] = ok ? (long long)key_row(outk[j]) + p.row_base : -1ll;
    }
    if (p.out_negdist != nullptr) {
      const float* nv = p.negatives + (size_t)q * p.dp;
      for (int j = warp; j < k; j += FC_WARPS) {
        float nd = __int_as_float(0x7f800000);
        if (j < kk) nd = exact_distance_warp(p.metric, p.arith, p.vec + (size_t)key_row(outk[j]) * p.dp, nv, p.d, scratch);
        if (lane == 0) p.out_negdist[(size_t)q * k + j] = nd;
      }
    }
    if (tid == 0) p.out_count[q] = certified ? kk : -1;
  }
  // the next pass of the search reuses the candidate lists: its scan starts only after this kernel
  // has completed (dependent launch), so the counters can be handed back zeroed from here
  if (tid == 0 && cp.reset_cnt != nullptr) {
    cp.reset_cnt[q] = 0;
    if (q < cp.n_reset_work && cp.reset_work != nullptr) cp.reset_work[q] = 0;
  }
  if (cp.dbg != nullptr && q == 0 && tid == 0) {
    ts[nts++] = clock64();
    for (int i = 0; i < nts; ++i) cp.dbg[i] = (unsigned long long)(ts[i] - ts[0]);
    cp.dbg[8] = (unsigned long long)n;
    cp.dbg[9] = (unsigned long long)nex;
    cp.dbg[10] = (unsigned long long)(t_keys - ts[0]);
    cp.dbg[11] = (unsigned long long)(t_pivot - ts[0]);
  }
}

// ------------------------------------------------------------------------------------------------
// The same stage with ONE WARP per query (round 2). The CTA-per-query kernel above is a chain of
// block-wide barriers (eight radix rounds, rank counting, two re-rank phases): 14 us of latency per query at
// two resident CTAs per SM, 98 us for the 2 048 queries of a pass group — 14 % of the headline step. A
// warp needs no block barrier at all and sixteen to twenty-four queries are in flight per SM:
//   1. radix-256 select over the 32-bit score images (four rounds; the warp's 256 shared-memory counters),
//      keys re-read from global memory each round (they sit in L1 / L2: ~2 KB per query);
//   2. the selected candidates are re-ranked exactly, ONE LANE per candidate (exact_distance_lane: the
//      reference's sequential sum is a chain anyway — 32 chains run side by side);
//   3. E = k-th smallest exact distance (warp bitonic sort in shared memory), T from the same error bound,
//      every further candidate with score <= T is re-ranked too, certificate as above;
//   4. the exact keys are sorted and the first k emitted.
// Any selection of >= k candidates followed by steps 3-4 yields the exact top-k when the certificate holds,
// so the results equal the CTA kernel's bit for bit (tests: every tensor-core parity test runs through it).
// ------------------------------------------------------------------------------------------------
constexpr int FW_WARPS = 8;          // queries per CTA
constexpr int FW_SEL_MAX = 64;       // selected candidates per query (kp <= 32 plus ties on the score image)
constexpr int FW_EX_MAX = 512;       // exact keys per query: selected + extras (more extras => not certified).
// For k <= 16 (kp = 32) the scan admits ~256 candidates per query, so 512 holds every list the sampled
// threshold produces; larger k (kp = 64 / 128: ~750 admitted, most of them below T) stay on the CTA kernel.

// ascending bitonic sort of n2 (power of two <= FW_EX_MAX) keys in the warp's shared memory
__device__ __forceinline__ void warp_bitonic_sort_smem(uint64_t* a, int n2) {
  const int lane = threadIdx.x & 31;
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const uint64_t x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(FW_WARPS * 32) finalize_cand_warp_kernel(const FinalizeCandParams cp, int nq) {
  __shared__ int s_hist[FW_WARPS][256];
  __shared__ uint64_t s_ex[FW_WARPS][FW_EX_MAX];
  const FinalizeParams& p = cp.base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int q = blockIdx.x * FW_WARPS + warp;
  // (no early griddepcontrol.launch_dependents here: all CTAs of this small grid are resident at once, so the
  //  next scan's persistent CTAs — 61 K registers each — would be launched at once too and take every SM a
  //  finished CTA of this kernel frees, leaving the CTAs still queued nowhere to run: measured 3.36 ms per
  //  10 000-query step with the early trigger against 2.99 ms without.)
  pdl_wait();  // candidate lists, counts and thresholds come from the scan / threshold kernels just before
  if (q >= nq) return;
  int* hist = s_hist[warp];
  uint64_t* ex = s_ex[warp];
  const int cap = cp.cap, k = p.k;
  const int n_raw = cp.cand_cnt[q];
  const bool overflow = n_raw > cap;
  const int n = overflow ? cap : n_raw;
  const float tau = cp.tau[q];
  const bool all_admitted = tau == __int_as_float(0x7f800000);
  const float* qv = p.queries + (size_t)q * p.dp;
  const uint64_t* src = cp.cand + (size_t)q * cap;

  // |q|^2 in float64 (same reduction order as the CTA kernel: lane-strided partial sums, xor butterfly)
  double qn2 = 0.0;
  for (int i = lane; i < p.d; i += 32) qn2 += (double)qv[i] * (double)qv[i];
  for (int o = 16; o >= 1; o >>= 1) qn2 += __shfl_xor_sync(0xffffffffu, qn2, o);

  // ---- 1. pivot = smallest score image P with at least `want` keys <= P (radix 256, four rounds) ----
  const int want = n < cp.kp ? n : cp.kp;
  uint32_t prefix = 0;
  int below = 0;
  if (want > 0) {
    for (int r = 0; r < 4; ++r) {
      const int shift = 24 - 8 * r;
      for (int i = lane; i < 256; i += 32) hist[i] = 0;
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        const uint32_t img = (uint32_t)(__ldcg(src + i) >> 32);
        if (r == 0 || (img >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(img >> shift) & 255u], 1);
      }
      __syncwarp();
      // lane l owns counters 8l .. 8l+7: inclusive scan across lanes, then inside the lane
      int c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j] = hist[lane * 8 + j];
        sum += c[j];
      }
      int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
      }
      const int before_lane = inc - sum;  // keys in the digits of lower lanes
      const unsigned reach = __ballot_sync(0xffffffffu, below + inc >= want);
      const int pick_lane = reach ? __ffs((int)reach) - 1 : 31;
      int digit = 0, before = 0;
      if (lane == pick_lane) {
        int run = before_lane;
        digit = 7;
        before = run;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (below + run + c[j] >= want) {
            digit = j;
            before = run;
            break;
          }
          run += c[j];
          before = run;
        }
        digit += lane * 8;
      }
      digit = __shfl_sync(0xffffffffu, digit, pick_lane);
      before = __shfl_sync(0xffffffffu, before, pick_lane);
      prefix |= (uint32_t)digit << shift;
      below += before;
      __syncwarp();
    }
  }
  const uint32_t pivot = prefix;

  // ---- 2. selected candidates (image <= pivot, the first FW_SEL_MAX in list order) -> exact keys ----
  int nsel = 0;
  if (want > 0) {
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const uint64_t key = i < n ? __ldcg(src + i) : KEY_NONE;
      const bool sel = i < n && (uint32_t)(key >> 32) <= pivot;
      const unsigned m = __ballot_sync(0xffffffffu, sel);
      const int pos = nsel + __popc(m & lt);
      if (sel && pos < FW_SEL_MAX) ex[pos] = key;  // scan key for now
      nsel += __popc(m);
    }
  }
  const int nsel_all = nsel;  // how many images are <= pivot (those beyond FW_SEL_MAX are treated as extras)
  if (nsel > FW_SEL_MAX) nsel = FW_SEL_MAX;
  __syncwarp();
  for (int c0 = 0; c0 < nsel; c0 += 32) {
    const int c = c0 + lane;
    if (c < nsel) {
      const uint32_t rid = key_row(ex[c]);
      const float dist = exact_distance_lane(p.metric, p.arith, qv, p.vec + (size_t)rid * p.dp, p.d, p.dp);
      ex[c] = make_key(dist, rid);
    }
  }
  __syncwarp();

  // ---- 3. certificate ----
  bool certified = !overflow;
  int nex = nsel;
  if (nsel >= k) {
    // E = k-th smallest exact distance of the selected candidates
    int n2 = 32;
    while (n2 < nsel) n2 <<= 1;
    for (int i = nsel + lane; i < n2; i += 32) ex[i] = KEY_NONE;
    __syncwarp();
    warp_bitonic_sort_smem(ex, n2);
    const float E = key_score(ex[k - 1]);
    const double mx = (double)(p.max_norm2 ? *p.max_norm2 : 0.f);
    float T = tc_score_upper_bound(p.metric, p.mode, p.cosine, E, cp.tc_gamma, (double)p.gamma, qn2, mx);
    if (cp.tc_norm_gamma > 0.0) {  // raw L2 scan: the norm term is accumulated by the tensor core as well
      const double t2 = (double)T + cp.tc_norm_gamma * mx;
      T = (float)t2;
      if ((double)T < t2) T = nextafterf(T, __int_as_float(0x7f800000));
    }
    if (!(T <= tau) && !all_admitted) certified = false;
    // every further candidate whose scan score is <= T may still belong to the exact top-k
    int ordinal = 0;  // position among the keys with image <= pivot (list order, as in step 2)
    const int nex0 = nex;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const uint64_t key = i < n ? __ldcg(src + i) : KEY_NONE;
      const bool in_piv = i < n && (uint32_t)(key >> 32) <= pivot;
      const unsigned mp = __ballot_sync(0xffffffffu, in_piv);
      const int my_ord = ordinal + __popc(mp & lt);
      ordinal += __popc(mp);
      const bool extra = i < n && key_score(key) <= T && (in_piv ? my_ord >= FW_SEL_MAX : true);
      const unsigned me = __ballot_sync(0xffffffffu, extra);
      const int pos = nex + __popc(me & lt);
      if (extra && pos < FW_EX_MAX) ex[pos] = key;
      nex += __popc(me);
    }
    if (nex > FW_EX_MAX) {
      nex = FW_EX_MAX;
      certified = false;
    }
    __syncwarp();
    for (int c0 = nex0; c0 < nex; c0 += 32) {
      const int c = c0 + lane;
      if (c < nex) {
        const uint32_t rid = key_row(ex[c]);
        const float dist = exact_distance_lane(p.metric, p.arith, qv, p.vec + (size_t)rid * p.dp, p.d, p.dp);
        ex[c] = make_key(dist, rid);
      }
    }
    __syncwarp();
  } else {
    // fewer than k candidates: complete only if the scan admitted every row
    if (!all_admitted) certified = false;
    (void)nsel_all;
  }

  // ---- 4. order the exact keys and emit the first k ----
  pdl_launch_dependents();  // the next kernel's CTAs may take the SMs this kernel's tail leaves idle
  {
    int n2 = 32;
    while (n2 < nex) n2 <<= 1;
    for (int i = nex + lane; i < n2; i += 32) ex[i] = KEY_NONE;
    __syncwarp();
    warp_bitonic_sort_smem(ex, n2);
  }
  const int kk = nex < k ? nex : k;
  if (p.out_keys != nullptr) {
    for (int j = lane; j < k; j += 32) {
      uint64_t o = KEY_NONE;
      if (j < kk && certified) o = (ex[j] & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)(p.row_base + key_row(ex[j]));
      p.out_keys[(size_t)q * k + j] = o;
    }
    if (lane == 0 && p.out_count) p.out_count[q] = certified ? kk : -1;
  } else {
    for (int j = lane; j < k; j += 32) {
      const bool ok = j < kk;
      p.out_dist[(size_t)q * k + j] = ok ? key_score(ex[j]) : __int_as_float(0x7f800000);
      p.out_row[(size_t)q * k + j] = ok ? (long long)key_row(ex[j]) + p.row_base : -1ll;
      if (p.out_negdist != nullptr) {
        float nd = __int_as_float(0x7f800000);
        if (ok) nd = exact_distance_lane(p.metric, p.arith, p.vec + (size_t)key_row(ex[j]) * p.dp,
                                         p.negatives + (size_t)q * p.dp, p.d, p.dp);
        p.out_negdist[(size_t)q * k + j] = nd;
      }
    }
    if (lane == 0) p.out_count[q] = certified ? kk : -1;
  }
  if (lane == 0 && cp.reset_cnt != nullptr) {
    cp.reset_cnt[q] = 0;
    if (q < cp.n_reset_work && cp.reset_work != nullptr) cp.reset_work[q] = 0;
  }
}

int launch_finalize_cand(const FinalizeCandParams& p, int nq, cudaStream_t st) {
  if (nq <= 0) return 0;
  if (p.cap > 4 * FC_THREADS) return fail(1, "finalize_cand: candidate capacity must be at most 2048");
  static const bool warp_on = [] {
    // opt-in: 80 us instead of 99 us per 2 048 queries on its own, but the step is slower with it (the next
    // pass's threshold kernel no longer hides under the finalize)
    const char* e = std::getenv("QG_FINALIZE_WARP");
    return e != nullptr && std::atoi(e) != 0;
  }();
  // rows must be 16-byte aligned for the per-lane row loads (dp % 4 == 0)
  if (warp_on && p.dbg == nullptr && p.kp <= 32 && (p.base.dp & 3) == 0) {
    QG_CUDA_OK(launch_chained(finalize_cand_warp_kernel, dim3((nq + FW_WARPS - 1) / FW_WARPS), dim3(FW_WARPS * 32),
                              (size_t)0, st, p, nq));
    return 0;
  }
  static const bool lane_on = [] {
    const char* e = std::getenv("QG_FINALIZE_LANE");
    return e == nullptr || std::atoi(e) != 0;
  }();
  FinalizeCandParams pl = p;
  // Measured (1M x 128, 10 000 queries, us per 2 048 queries): L2 k=10 100 -> 181, k=100 297 -> 342 (two CTAs per
  // SM: the latency of a lane's 128-step chain counts, not the instructions saved); cosine, whose distance is
  // three sums, k=10 145 -> 171 but k=100 551 -> 321. So: cosine with 64 or more candidates to re-rank.
  pl.lane_rerank = (lane_on && (p.base.dp & 3) == 0 && p.base.metric == METRIC_COSINE && p.kp >= 64) ? 1 : 0;
  QG_CUDA_OK(launch_chained(finalize_cand_kernel, dim3(nq), dim3(FC_THREADS), finalize_cand_smem(p.cap), st, pl));
  return 0;
}

// ---- K8: k-way merge of the per-shard lists after the all-gather --------------------------------
// keys: [world][nq][k] exact keys (ordered float32 distance << 32 | global row). One CTA per query.
__global__ void __launch_bounds__(256) merge_shards_kernel(const uint64_t* __restrict__ keys, int world, int nq,
                                                           int k, float* out_dist, long long* out_row,
                                                           int* out_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
  const int q = blockIdx.x;
  const int total = world * k;
  int n2 = 32;
  while (n2 < total) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    uint64_t v = KEY_NONE;
    if (i < total) {
      const int w = i / k, j = i % k;
      v = keys[((size_t)w * nq + q) * k + j];
    }
    buf[i] = v;
  }
  block_bitonic_sort(buf, n2);
  int cnt = 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t v = buf[j];
    const bool ok = v != KEY_NONE;
    out_dist[(size_t)q * k + j] = ok ? key_score(v) : __int_as_float(0x7f800000);
    out_row[(size_t)q * k + j] = ok ? (long long)key_row(v) : -1ll;
    cnt += ok;
  }
  // count = number of valid keys among the first k
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) out_count[q] = s_cnt;
}

int launch_merge_shards(const uint64_t* keys, int world, int nq, int k, float* out_dist, long long* out_row,
                        int* out_count, cudaStream_t st) {
  if (nq <= 0) return 0;
  int total = world * k, n2 = 32;
  while (n2 < total) n2 <<= 1;
  const size_t smem = (size_t)n2 * 8;
  if (smem > 200 * 1024) return fail(6, "merge_shards: world*k too large for one CTA");
  merge_shards_kernel<<<nq, 256, smem, st>>>(keys, world, nq, k, out_dist, out_row, out_count);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- K7 / K5: exact distances of explicit (query, row) pairs -------------------------------------
// One warp per pair. queries [b x dp], rows [b x m] (0xFFFFFFFF = skip), out [b x m].
__global__ void __launch_bounds__(256) batch_distance_kernel(const float* __restrict__ vec, int dp, int d,
                                                             long long n_rows, int metric, int arith,
                                                             const float* __restrict__ queries, int qstride,
                                                             int b, const uint32_t* __restrict__ rows, int m,
                                                             float* __restrict__ out) {
  __shared__ __align__(16) double scratch_all[8 * (EXACT_SCRATCH_BYTES / 8)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* scratch = scratch_all + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  const long long total = (long long)b * m;
  for (long long pair = (long long)blockIdx.x * 8 + warp; pair < total; pair += (long long)gridDim.x * 8) {
    const int qi = (int)(pair / m);
    const uint32_t row = rows[pair];
    float dist = __int_as_float(0x7f800000);
    if (row != 0xFFFFFFFFu && (long long)row < n_rows) {
      dist = exact_distance_warp(metric, arith, queries + (size_t)qi * qstride, vec + (size_t)row * dp, d, scratch);
    }
    if (lane == 0) out[pair] = dist;
  }
}

int launch_batch_distance(const float* vec, int dp, int d, long long n_rows, int metric, int arith,
                          const float* queries, int qstride, int b, const uint32_t* rows, int m, float* out,
                          cudaStream_t st) {
  const long long total = (long long)b * m;
  if (total <= 0) return 0;
  long long blocks = (total + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  batch_distance_kernel<<<(int)blocks, 256, 0, st>>>(vec, dp, d, n_rows, metric, arith, queries, qstride, b, rows, m,
                                                      out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qg
