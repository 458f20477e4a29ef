// exhaustive.cu — the reference algorithm verbatim on the device: the exact distance of EVERY
// passing row in the reference's arithmetic, a full sort by (distance, row), the first k
// (pkg/hybrid/exact.go:114-129). Used when k is too large for the candidate pools
// (Collection.Search asks for k = Index.Size() when filters are present,
// pkg/core/collection.go:679-682) and when the flat scan could not certify a result
// (finalize.cu). The sort is cub::DeviceRadixSort — a library call on a cold path.
#include <cub/device/device_radix_sort.cuh>

#include "exact.cuh"
#include "finalize.cuh"

namespace qg {

__global__ void __launch_bounds__(256) exact_all_kernel(const float* __restrict__ vec, long long rows, int dp, int d,
                                                        const uint32_t* __restrict__ mask,
                                                        const float* __restrict__ query, int metric, int arith,
                                                        uint64_t* __restrict__ keys) {
  __shared__ __align__(16) double scratch_all[8 * (EXACT_SCRATCH_BYTES / 8)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* scratch = scratch_all + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
    uint64_t key = KEY_NONE;
    const bool pass = mask == nullptr || ((mask[r >> 5] >> (r & 31)) & 1u);
    if (pass) {
      const float dist = exact_distance_warp(metric, arith, query, vec + (size_t)r * dp, d, scratch);
      key = make_key(dist, (uint32_t)r);
    }
    if (lane == 0) keys[r] = key;
  }
}

__global__ void __launch_bounds__(256) emit_sorted_kernel(const uint64_t* __restrict__ sorted, long long rows,
                                                          long long k, const float* __restrict__ vec, int dp, int d,
                                                          const float* __restrict__ negative, int metric, int arith,
                                                          float* out_dist, float* out_negdist, long long* out_row,
                                                          int* out_count, uint64_t* out_keys, long long row_base) {
  __shared__ __align__(16) double scratch_all[8 * (EXACT_SCRATCH_BYTES / 8)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* scratch = scratch_all + (size_t)warp * (EXACT_SCRATCH_BYTES / 8);
  for (long long j = (long long)blockIdx.x * 8 + warp; j < k; j += (long long)gridDim.x * 8) {
    const uint64_t key = j < rows ? sorted[j] : KEY_NONE;
    const bool ok = key != KEY_NONE;
    if (out_keys != nullptr) {
      if (lane == 0)
        out_keys[j] = ok ? ((key & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)(row_base + key_row(key))) : KEY_NONE;
      continue;
    }
    float nd = __int_as_float(0x7f800000);
    if (ok && out_negdist != nullptr)
      nd = exact_distance_warp(metric, arith, vec + (size_t)key_row(key) * dp, negative, d, scratch);
    if (lane == 0) {
      out_dist[j] = ok ? key_score(key) : __int_as_float(0x7f800000);
      out_row[j] = ok ? (long long)key_row(key) + row_base : -1ll;
      if (out_negdist != nullptr) out_negdist[j] = nd;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_count != nullptr) {
    // number of valid keys among the first k: binary search for the first KEY_NONE
    long long lo = 0, hi = k < rows ? k : rows;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (sorted[mid] == KEY_NONE) hi = mid;
      else lo = mid + 1;
    }
    *out_count = (int)lo;
  }
}

void exhaustive_free(ExhaustiveWork& w) {
  if (w.keys_a) cudaFree(w.keys_a);
  if (w.keys_b) cudaFree(w.keys_b);
  if (w.temp) cudaFree(w.temp);
  w = ExhaustiveWork();
}

int exhaustive_search(ExhaustiveWork& w, const float* vec, long long rows, int dp, int d, const uint32_t* mask,
                      const float* query, const float* negative, int metric, int arith, long long k,
                      float* out_dist, float* out_negdist, long long* out_row, int* out_count, uint64_t* out_keys,
                      long long row_base, cudaStream_t st) {
  if (rows <= 0 || k <= 0) return 0;
  if (rows > 0x7fffffffll) return fail(6, "exhaustive path supports at most 2^31-1 rows per index");
  size_t need = 0;
  QG_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, need, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)rows, 0,
                                            64, st));
  if (w.cap_rows < rows || w.temp_bytes < need) {
    QG_CUDA_OK(cudaStreamSynchronize(st));
    exhaustive_free(w);
    QG_CUDA_OK(cudaMalloc(&w.keys_a, (size_t)rows * 8));
    QG_CUDA_OK(cudaMalloc(&w.keys_b, (size_t)rows * 8));
    QG_CUDA_OK(cudaMalloc(&w.temp, need));
    w.cap_rows = rows;
    w.temp_bytes = need;
  }
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  exact_all_kernel<<<(int)blocks, 256, 0, st>>>(vec, rows, dp, d, mask, query, metric, arith, w.keys_a);
  QG_CUDA_OK(cudaGetLastError());
  size_t tb = w.temp_bytes;
  QG_CUDA_OK(cub::DeviceRadixSort::SortKeys(w.temp, tb, w.keys_a, w.keys_b, (int)rows, 0, 64, st));
  long long eb = (k + 7) / 8;
  if (eb > 148 * 8) eb = 148 * 8;
  emit_sorted_kernel<<<(int)eb, 256, 0, st>>>(w.keys_b, rows, k, vec, dp, d, negative, metric, arith, out_dist,
                                              out_negdist, out_row, out_count, out_keys, row_base);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qg
