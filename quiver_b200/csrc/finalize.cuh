// finalize.cuh — parameters of the merge + exact re-rank + certificate kernel (finalize.cu),
// of the shard merge kernel, and of the exhaustive fallback (exhaustive.cu).
#pragma once
#include "common.cuh"

namespace qg {

struct FinalizeParams {
  const uint64_t* partial;  // [nq][nb][kp] sorted scan keys per scan CTA
  int nb, kp;
  const float* vec;         // [rows x dp]
  int dp, d;
  const float* queries;     // [nq x dp]
  const float* negatives;   // [nq x dp] or nullptr
  int metric, arith, mode, cosine;
  int k;
  float gamma;              // relative error bound of the fp32 scan sum
  const float* max_norm2;   // device scalar: max |x|^2 over the corpus (dot-product bound)
  float* out_dist;          // [nq x k]
  float* out_negdist;       // [nq x k] or nullptr
  long long* out_row;       // [nq x k]
  int* out_count;           // [nq]; -1 = not certified, caller must use the exhaustive path
  uint64_t* out_keys;       // [nq x k] shard mode (then out_dist/out_row/out_count unused)
  long long row_base;
};

int launch_finalize(const FinalizeParams& p, int nq, cudaStream_t st);

// Tensor-core regime (tc_scan.cu): one unsorted candidate list per query (keys admitted by the
// threshold tau), instead of per-CTA sorted lists.
struct FinalizeCandParams {
  const uint64_t* cand;     // [nq][cap] scan keys (score image << 32 | row), unordered
  const int* cand_cnt;      // [nq] appended keys (may exceed cap = overflow)
  int cap;
  const float* tau;         // [nq] admission threshold the scan used (+inf = every row admitted)
  int kp;                   // candidates re-ranked before the certificate (pow2 >= k + margin)
  int* reset_cnt;           // optional: [nq] candidate counters zeroed for the next pass (hoisted sample stage) ...
  int* reset_work;          // ... and the main scans' tile-claim counters
  int n_reset_work;         // how many of them (one per pass of the group this launch finalizes)
  double tc_gamma;          // bound on |tensor-core dot - exact dot| / (|q| |x|)
  double tc_norm_gamma;     // raw L2 scan: bound on the extra accumulation error / |x|^2 (the norm rides in the MMA), else 0
  unsigned long long* dbg;  // optional [16] phase time stamps of CTA 0 (development aid), else nullptr
  int lane_rerank;          // exact re-rank with one lane per candidate (needs dp % 4 == 0) instead of one warp per candidate
  FinalizeParams base;      // vec, dp, d, queries, negatives, metric, arith, mode, cosine, k, gamma,
                            // max_norm2, outputs, row_base (partial / nb unused)
};
int launch_finalize_cand(const FinalizeCandParams& p, int nq, cudaStream_t st);
int finalize_set_attributes();

int launch_merge_shards(const uint64_t* keys, int world, int nq, int k, float* out_dist, long long* out_row,
                        int* out_count, cudaStream_t st);

// Exhaustive exact path: exact distance of every passing row, full sort, first k.
struct ExhaustiveWork {
  uint64_t* keys_a = nullptr;
  uint64_t* keys_b = nullptr;
  void* temp = nullptr;
  size_t temp_bytes = 0;
  long long cap_rows = 0;
};
int exhaustive_search(ExhaustiveWork& w, const float* vec, long long rows, int dp, int d, const uint32_t* mask,
                      const float* query, const float* negative, int metric, int arith, long long k,
                      float* out_dist, float* out_negdist, long long* out_row, int* out_count, uint64_t* out_keys,
                      long long row_base, cudaStream_t st);
void exhaustive_free(ExhaustiveWork& w);

// Exact distances of explicit (query, row) pairs — HNSW neighbour batches, negatives.
int launch_batch_distance(const float* vec, int dp, int d, long long n_rows, int metric, int arith,
                          const float* queries, int qstride, int b, const uint32_t* rows, int m, float* out,
                          cudaStream_t st);

}  // namespace qg
