// compact.cu — see compact.cuh. HBM-bound byte moving: every array is read once and its live part written once.
#include "compact.cuh"

namespace qg {

static constexpr int SCAN_THREADS = 1024;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // counts per block

// Block-wide exclusive scan of one value per thread (1024 threads); returns the block total through `total`.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = warp_sums[lane];
    uint32_t si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    warp_sums[lane] = si - s;  // exclusive warp offsets
    if (lane == 31) warp_sums[32] = si;
  }
  __syncthreads();
  const uint32_t ex = warp_sums[warp] + inc - v;
  *total = warp_sums[32];
  __syncthreads();  // warp_sums is reused by the caller's next round
  return ex;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint32_t* __restrict__ in, long long n,
                                                                       uint32_t* __restrict__ bsum) {
  __shared__ uint32_t warp_sums[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  uint32_t total;
  block_exclusive_scan(s, warp_sums, &total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// One block walks the block sums in rounds of 1024 with a running carry; bsum becomes exclusive, the grand
// total lands in *total_out.
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(uint32_t* bsum, long long n_blocks,
                                                                  uint32_t* total_out) {
  __shared__ uint32_t warp_sums[33];
  uint32_t carry = 0;
  for (long long b0 = 0; b0 < n_blocks; b0 += SCAN_THREADS) {
    const long long i = b0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? bsum[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, warp_sums, &total);
    if (i < n_blocks) bsum[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, long long n,
                                                                  const uint32_t* __restrict__ bsum,
                                                                  uint32_t* __restrict__ out) {
  __shared__ uint32_t warp_sums[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, warp_sums, &total) + bsum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

int launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, long long n, uint32_t* block_tmp, cudaStream_t st) {
  if (n <= 0) {
    QG_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(uint32_t), st));
    return 0;
  }
  const long long n_blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n_blocks > 0x7FFFFFFFll) return fail(1, "scan: too many elements");
  scan_block_sums_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, st>>>(in, n, block_tmp);
  QG_CUDA_OK(cudaGetLastError());
  scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(block_tmp, n_blocks, out + n);
  QG_CUDA_OK(cudaGetLastError());
  scan_apply_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, st>>>(in, n, block_tmp, out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void word_popcount_kernel(const uint32_t* __restrict__ live, long long n_words,
                                     uint32_t* __restrict__ cnt) {
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
       w += (long long)gridDim.x * blockDim.x)
    cnt[w] = (uint32_t)__popc(live[w]);
}

int launch_word_popcount(const uint32_t* live, long long n_words, uint32_t* cnt, cudaStream_t st) {
  if (n_words <= 0) return 0;
  long long blocks = (n_words + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  word_popcount_kernel<<<(int)blocks, 256, 0, st>>>(live, n_words, cnt);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void compact_map_kernel(const uint32_t* __restrict__ live, const uint32_t* __restrict__ word_off,
                                   long long n_rows, uint32_t* __restrict__ map) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (long long)gridDim.x * blockDim.x) {
    const uint32_t w = live[r >> 5];
    const uint32_t bit = 1u << (r & 31);
    map[r] = (w & bit) ? word_off[r >> 5] + (uint32_t)__popc(w & (bit - 1u)) : 0xFFFFFFFFu;
  }
}

int launch_compact_map(const uint32_t* live, const uint32_t* word_off, long long n_rows, uint32_t* map,
                       cudaStream_t st) {
  if (n_rows <= 0) return 0;
  long long blocks = (n_rows + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  compact_map_kernel<<<(int)blocks, 256, 0, st>>>(live, word_off, n_rows, map);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// The map as the caller receives it: int64, -1 for a deleted row.
__global__ void widen_map_kernel(const uint32_t* __restrict__ map, long long n, long long* __restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const uint32_t j = map[r];
    out[r] = j == 0xFFFFFFFFu ? -1ll : (long long)j;
  }
}

int launch_widen_map(const uint32_t* map, long long n, long long* out, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  widen_map_kernel<<<(int)blocks, 256, 0, st>>>(map, n, out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// One warp per group of 32 consecutive old rows. The lanes read the group's 32 map entries in one coalesced
// load and move the per-row scalars; the live rows of a group land on consecutive new rows, and their
// vectors are moved four rows at a time (four independent 128-bit loads per lane in flight before the first
// store) — a dead row costs nothing, and the latency of the map load is paid once per 32 rows.
static constexpr int COMPACT_UNROLL = 4;

// Moves up to COMPACT_UNROLL rows of the fp32 array and of the bf16 copy: all loads of a 32-lane slice (both
// arrays, every row) are issued before the first store.
__device__ __forceinline__ void move_rows(const float4* __restrict__ src, float4* __restrict__ dst, int v4,
                                          const uint4* __restrict__ src16, uint4* __restrict__ dst16, int h8,
                                          const long long (&r)[COMPACT_UNROLL], const uint32_t (&j)[COMPACT_UNROLL],
                                          int cnt, int lane) {
  const int lim = v4 > h8 ? v4 : h8;
  for (int i0 = 0; i0 < lim; i0 += 32) {
    const int i = i0 + lane;
    const bool in32 = i < v4, in16 = i < h8;
    float4 t[COMPACT_UNROLL];
    uint4 h[COMPACT_UNROLL];
#pragma unroll
    for (int u = 0; u < COMPACT_UNROLL; ++u) {
      if (u < cnt && in32) t[u] = src[(size_t)r[u] * v4 + i];
      if (u < cnt && in16) h[u] = src16[(size_t)r[u] * h8 + i];
    }
#pragma unroll
    for (int u = 0; u < COMPACT_UNROLL; ++u) {
      if (u < cnt && in32) dst[(size_t)j[u] * v4 + i] = t[u];
      if (u < cnt && in16) dst16[(size_t)j[u] * h8 + i] = h[u];
    }
  }
}

// 80 registers = 3 resident CTAs per SM; capping at 64 for a fourth CTA spills and measured slower (1.35 vs 1.28 ms).
__global__ void __launch_bounds__(256) compact_rows_kernel(CompactRowsArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int v4 = a.dp >> 2, h8 = a.vec16 ? a.dp16 >> 3 : 0;  // the bf16 row is never longer than the fp32 row in 16-byte units
  const float4* src = reinterpret_cast<const float4*>(a.vec);
  float4* dst = reinterpret_cast<float4*>(a.vec_out);
  const uint4* src16 = reinterpret_cast<const uint4*>(a.vec16);
  uint4* dst16 = reinterpret_cast<uint4*>(a.vec16_out);
  for (long long base = warp0 * 32; base < a.n_rows; base += (long long)gridDim.x * 8 * 32) {
    const long long mine = base + lane;
    const uint32_t jm = mine < a.n_rows ? a.map[mine] : 0xFFFFFFFFu;
    const bool alive = jm != 0xFFFFFFFFu;
    if (alive) {
      a.inv_norm_out[jm] = a.inv_norm[mine];
      a.norm2_out[jm] = a.norm2[mine];
      a.unit_bias_out[jm] = a.unit_bias[mine];
    }
    unsigned m = __ballot_sync(0xffffffffu, alive);
    while (m) {
      long long r[COMPACT_UNROLL];
      uint32_t j[COMPACT_UNROLL];
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < COMPACT_UNROLL; ++u) {
        r[u] = 0;
        j[u] = 0;
        if (m) {
          const int s = __ffs(m) - 1;
          m &= m - 1;
          r[u] = base + s;
          j[u] = __shfl_sync(0xffffffffu, jm, s);
          cnt = u + 1;
        }
      }
      move_rows(src, dst, v4, src16, dst16, h8, r, j, cnt, lane);
    }
  }
}

int launch_compact_rows(const CompactRowsArgs& a, int sm_count, cudaStream_t st) {
  if (a.n_rows <= 0) return 0;
  if ((a.dp & 3) || (a.vec16 && (a.dp16 & 7))) return fail(1, "compact: row pitch is not 16-byte aligned");
  long long blocks = (a.n_rows + 255) / 256;  // 8 warps x 32 rows per CTA pass
  const long long cap = (long long)sm_count * 8;  // a multiple of the SM count; the warps stride over the row groups
  if (blocks > cap) blocks = cap;
  compact_rows_kernel<<<(int)blocks, 256, 0, st>>>(a);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void compact_column_kernel(CompactColArgs a) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.n;
       r += (long long)gridDim.x * blockDim.x) {
    const uint32_t j = a.map[r];
    if (j == 0xFFFFFFFFu) continue;
    a.kind_out[j] = a.kind[r];
    a.num_out[j] = a.num[r];
    a.scode_out[j] = a.scode[r];
    a.fcode_out[j] = a.fcode[r];
    if (a.arr_off) a.arr_cnt_out[j] = (uint32_t)(a.arr_off[r + 1] - a.arr_off[r]);
  }
}

int launch_compact_column(const CompactColArgs& a, cudaStream_t st) {
  if (a.n <= 0) return 0;
  long long blocks = (a.n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  compact_column_kernel<<<(int)blocks, 256, 0, st>>>(a);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) compact_elems_kernel(const uint32_t* __restrict__ map, long long n,
                                                            const int32_t* __restrict__ arr_off,
                                                            const int32_t* __restrict__ arr_code,
                                                            const uint32_t* __restrict__ new_off,
                                                            int32_t* __restrict__ code_out) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (long long)gridDim.x * 8) {
    const uint32_t j = map[r];
    if (j == 0xFFFFFFFFu) continue;
    const int b = arr_off[r], e = arr_off[r + 1];
    const uint32_t o = new_off[j];
    for (int i = b + lane; i < e; i += 32) code_out[o + (uint32_t)(i - b)] = arr_code[i];
  }
}

int launch_compact_elems(const uint32_t* map, long long n, const int32_t* arr_off, const int32_t* arr_code,
                         const uint32_t* new_off, int32_t* code_out, int sm_count, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 7) / 8;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  compact_elems_kernel<<<(int)blocks, 256, 0, st>>>(map, n, arr_off, arr_code, new_off, code_out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qg
