// hnsw.cuh — device-resident HNSW graph and the persistent search kernel (hnsw.cu).
#pragma once
#include "../../include/quiver_gpu.h"
#include "common.cuh"

namespace qg {

constexpr int HNSW_TOUCH_CAP = 8192;  // visited-bitset words a walk may touch before the clear falls back to a full wipe

// The reference's graph (pkg/hnsw/hnsw.go:44-82) as flat device arrays; node id = row of the index.
struct HnswDevGraph {
  long long n_nodes;
  int m, max_m0, entry_point, current_level;
  const int32_t* level;         // [n] node level, -1 = deleted
  const uint32_t* adj0;         // [n x max_m0], 0xFFFFFFFF padded
  const long long* upper_off;   // [n + 1]
  const uint32_t* upper_adj;    // per node: level[i] blocks of m entries
};

size_t hnsw_workspace_bytes(long long n_nodes, int sm_count);
// One warp per query; results of the graph walk only (the caller supplements under-filled queries with the
// exact pass, hnsw.go:676-710). out_count[i] = -1: the query's candidate heap outgrew shared memory.
int launch_hnsw_search(const HnswDevGraph& g, const float* vec, int dp, int d, int metric, int arith,
                       const float* d_queries, int nq, int kk, int ef0, void* workspace, int sm_count, uint32_t* out_idx,
                       float* out_dist, int* out_count, long long* out_evals, cudaStream_t st);

}  // namespace qg
