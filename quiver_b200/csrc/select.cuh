// select.cuh — CTA-level top-k' selection shared by the scan kernels and the merge kernel.
//
// Replaces the reference's "collect all N results, sort.Sort, slice to k"
// (pkg/hybrid/exact.go:114-129) with a threshold filter: a CTA keeps, per query, an
// unsorted candidate pool in shared memory plus a running threshold tau = score of the
// kp-th best candidate seen so far. A row is appended only when score <= tau, so after a
// short warm-up appends are rare. When a pool reaches its high-water mark the whole CTA
// sorts it (bitonic, shared memory), keeps the best kp entries and tightens tau.
//
// Keys are 64-bit (ordered-float score << 32 | row), so the order is the deterministic
// (score, row) order used everywhere in this library.
#pragma once
#include "common.cuh"

namespace qg {

// Pool geometry for a candidate count kp (power of two, 32..1024) and `nw` appending warps,
// each of which may append up to 32 keys per query between two looks at the prune flag.
__host__ __device__ constexpr int pool_slots(int kp) { return kp <= 128 ? 512 : (kp <= 512 ? 1024 : 2048); }
__host__ __device__ constexpr int pool_highwater(int kp, int nw) { return pool_slots(kp) - nw * 32; }

struct PoolRef {
  uint64_t* keys;  // [slots]
  int* cnt;        // number of valid keys (may transiently exceed the high-water mark)
  float* tau;      // accept score <= tau
};

// Bitonic sort (ascending) of keys[0..n2), n2 a power of two, by all threads of the CTA.
// Must be called by every thread of the block.
__device__ __forceinline__ void block_bitonic_sort(uint64_t* keys, int n2) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n2 >> 1); t += nt) {
        int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int hi = lo | stride;
        bool up = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Sort the pool, keep the best kp keys, update cnt and tau. All threads of the block call.
__device__ __forceinline__ void block_prune(PoolRef p, int kp) {
  __syncthreads();
  int n = *p.cnt;
  int n2 = next_pow2(n);
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) p.keys[i] = KEY_NONE;
  block_bitonic_sort(p.keys, n2);
  if (threadIdx.x == 0) {
    int keep = n < kp ? n : kp;
    *p.cnt = keep;
    *p.tau = (keep == kp) ? key_score(p.keys[kp - 1]) : __int_as_float(0x7f800000);
  }
  __syncthreads();
}

// Warp-aggregated append of the lanes whose `pass` is set. Returns the pool count after the
// append (warp-uniform) so the caller can raise the prune flag. All 32 lanes must call.
__device__ __forceinline__ int warp_append(PoolRef p, bool pass, uint64_t key) {
  const unsigned m = __ballot_sync(0xffffffffu, pass);
  if (m == 0) return 0;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(p.cnt, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pass) p.keys[base + __popc(m & ((1u << lane) - 1u))] = key;
  return base + __popc(m);
}

}  // namespace qg
