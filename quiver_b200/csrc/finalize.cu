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
#include "exact.cuh"
#include "finalize.cuh"
#include "scan.cuh"
#include "select.cuh"

namespace qg {

constexpr int FIN_THREADS = 256;
constexpr int FIN_WARPS = FIN_THREADS / 32;

__host__ __device__ inline size_t finalize_smem(int kp) {
  const int slots = pool_slots(kp);
  return (size_t)slots * 8 * 2 + (size_t)FIN_WARPS * EXACT_SCRATCH_BYTES + 64;
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

__global__ void __launch_bounds__(FIN_THREADS, 1) finalize_kernel(const FinalizeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int q = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kp = p.kp, nb = p.nb, k = p.k;
  const int slots = pool_slots(kp);
  const int highwater = slots - FIN_THREADS;

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

  if (tid == 0) {
    s_cnt = 0;
    s_tau = __int_as_float(0x7f800000);
    s_extra = 0;
    s_bad = 0;
  }
  if (warp == 0) {  // |q|^2 for the dot-product bound
    double s = 0.0;
    for (int i = lane; i < p.d; i += 32) s += (double)qv[i] * (double)qv[i];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_qn2 = s;
  }
  __syncthreads();

  PoolRef pr{pool, &s_cnt, &s_tau};

  // ---- 1. best kp scan keys among nb*kp, rank-major so the threshold tightens early -------
  const int total = nb * kp;
  for (int base = 0; base < total; base += FIN_THREADS) {
    const int idx = base + tid;
    uint64_t key = KEY_NONE;
    if (idx < total) key = part[(size_t)(idx % nb) * kp + (idx / nb)];
    const bool pass = (key != KEY_NONE) && (key_score(key) <= s_tau);
    warp_append(pr, pass, key);
    __syncthreads();
    if (s_cnt >= highwater) block_prune(pr, kp);
  }
  block_prune(pr, kp);
  const int ncand = s_cnt;
  const uint64_t last_key = ncand > 0 ? pool[ncand - 1] : 0ull;

  // ---- 2. exact re-rank --------------------------------------------------------------------
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

  // ---- 3. certificate and widening ------------------------------------------------------------
  if (ncand >= k && ncand > 0) {
    const float E = key_score(ex[k - 1]);
    const float T = score_upper_bound(p.metric, p.mode, p.cosine, E, (double)p.gamma, s_qn2,
                                      (double)(p.max_norm2 ? *p.max_norm2 : 0.f));
    for (int base = 0; base < total; base += FIN_THREADS) {
      const int idx = base + tid;
      if (idx < total) {
        const int b = idx % nb, j = idx / nb;
        const uint64_t key = part[(size_t)b * kp + j];
        if (key != KEY_NONE) {
          const bool within = key_score(key) <= T;
          if (within && key > last_key) {
            const int pos = atomicAdd(&s_extra, 1);
            if (ncand + pos < slots) pool[ncand + pos] = key;
          }
          if (within && j == kp - 1) s_bad = 1;  // a full list may have dropped rows <= T
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
      if (part[(size_t)b * kp + kp - 1] != KEY_NONE) s_bad = 1;
  }
  __syncthreads();

  // ---- 4. results ------------------------------------------------------------------------------
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

int finalize_set_attributes() {
  QG_CUDA_OK(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)finalize_smem(1024)));
  return 0;
}

int launch_finalize(const FinalizeParams& p, int nq, cudaStream_t st) {
  if (nq <= 0) return 0;
  finalize_kernel<<<nq, FIN_THREADS, finalize_smem(p.kp), st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
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
  static bool attr_done = false;
  if (!attr_done) {
    QG_CUDA_OK(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
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
