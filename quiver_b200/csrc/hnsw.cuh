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

// Graph construction (batched inserts, see hnsw.cu): one launch searches the committed graph for every node of
// the batch and writes the node's own lists + one reverse-link record per chosen neighbour; after the records
// have been sorted by key, the link launch merges them into the neighbours' lists.
int launch_hnsw_insert_batch(const HnswDevGraph& g, uint32_t* adj0_w, uint32_t* upper_w, const float* vec, int dp, int d,
                             int metric, int arith, int ef_c, long long first, int count, void* workspace, int sm_count,
                             unsigned long long* rev_key, float* rev_dist, unsigned int* rev_count, unsigned int rev_cap,
                             cudaStream_t st);
int launch_hnsw_link_batch(const HnswDevGraph& g, uint32_t* adj0_w, uint32_t* upper_w, const float* vec, int dp, int d,
                           int metric, int arith, const unsigned long long* sorted_keys, const unsigned int* perm,
                           const float* rev_dist, unsigned int n_rec, cudaStream_t st);
// sort of the reverse-link records by key (cub::DeviceRadixSort::SortPairs, perm = record index)
int hnsw_sort_records(const unsigned long long* keys, unsigned long long* keys_out, unsigned int* perm_in,
                      unsigned int* perm_out, unsigned int n, void** temp, size_t* temp_bytes, cudaStream_t st);

}  // namespace qg
