// Micro-benchmark: streaming a row-major fp32 matrix (1M x 128 = 512 MB) into shared memory with 1-D bulk
// copies, organised the way the flat scan does it (every warp owns a private ring and refills it itself)
// against a CTA-wide ring filled by one producer thread. No arithmetic: the consumers read one word per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_stream ring_stream.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <ctime>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: private ring per warp (lane 0 refills). wait_all = 1: all 32 lanes wait on the barrier, 0: lane 0 only
__global__ void __launch_bounds__(1024, 1) warp_rings(const float* base, long long n_tiles, int tile_bytes, int S, int wait_all, int order, unsigned long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* ring = smem + (size_t)warp * S * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * S * tile_bytes) + warp * S;
  if (lane == 0) for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long gws = (long long)gridDim.x * nw;
  // order 0: tile = it * gws + cta * nw + warp (neighbouring warps read neighbouring tiles)
  // order 1: every warp owns one contiguous slab of the matrix
  const long long gw = (long long)blockIdx.x * nw + warp;
  const long long per = (n_tiles + gws - 1) / gws;
  auto tile_of = [&](long long it) -> long long { return order == 0 ? it * gws + gw : (it < per ? gw * per + it : n_tiles); };
  auto issue = [&](long long it) {
    const long long t = tile_of(it);
    if (t >= n_tiles || lane != 0) return;
    const int s = (int)(it % S);
    mbar_expect(&bars[s], tile_bytes);
    bulk1d(ring + (size_t)s * tile_bytes, base + (size_t)t * (tile_bytes / 4), tile_bytes, &bars[s]);
  };
  for (int s = 0; s < S - 1; ++s) issue(s);
  unsigned long long acc = 0;
  for (long long it = 0;; ++it) {
    const long long t = tile_of(it);
    if (t >= n_tiles) break;
    issue(it + S - 1);
    const int s = (int)(it % S);
    if (wait_all || lane == 0) mbar_wait(&bars[s], (uint32_t)((it / S) & 1));
    __syncwarp();
    acc += *reinterpret_cast<volatile unsigned int*>(ring + (size_t)s * tile_bytes + lane * 16);
    __syncwarp();
  }
  if (acc == 0x1234567) *sink = acc;
}

// mode 1: one ring for the CTA, thread 0 of warp 0 produces, the other warps consume stages round robin
__global__ void __launch_bounds__(1024, 1) cta_ring(const float* base, long long n_tiles, int tile_bytes, int S, unsigned long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nc = (blockDim.x >> 5) - 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * tile_bytes);
  uint64_t* empty = full + S;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;  // tiles of this CTA: t = blockIdx + i * grid
  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (long long i = 0; i < per; ++i) {
        const long long t = i * gridDim.x + blockIdx.x;
        if (t >= n_tiles) break;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect(&full[s], tile_bytes);
        bulk1d(smem + (size_t)s * tile_bytes, base + (size_t)t * (tile_bytes / 4), tile_bytes, &full[s]);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // consumer c takes local tiles c, c + nc, ... ; stage of local tile i is i % S, phase (i / S) & 1
    unsigned long long acc = 0;
    for (long long i = warp - 1; i < per; i += nc) {
      const long long t = i * gridDim.x + blockIdx.x;
      if (t >= n_tiles) break;
      const int s = (int)(i % S);
      mbar_wait(&full[s], (uint32_t)((i / S) & 1));
      acc += *reinterpret_cast<volatile unsigned int*>(smem + (size_t)s * tile_bytes + lane * 16);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

__global__ void fill_kernel(float* p, size_t n, int mode) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = mode == 1 ? (float)(h % 218u) : (h >> 8) * (1.0f / 16777216.0f);  // small integers / U[0,1)
  }
}

int main(int argc, char** argv) {
  const long long rows = 1000000; const int dp = 128;
  const size_t bytes = (size_t)rows * dp * 4;
  float* d; cudaMalloc(&d, bytes + (1 << 20)); cudaMemset(d, 0, bytes);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(warp_rings, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(cta_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* what, float best) {
    cudaError_t err = cudaGetLastError();
    printf("%s  %.1f us  %.0f GB/s %s\n", what, best * 1e3, bytes / (best * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  char what[256];
  if (argc >= 8 && atoi(argv[7]) > 0) {
    fill_kernel<<<1184, 256>>>(d, bytes / 4, atoi(argv[7]));
    cudaDeviceSynchronize();
    printf("buffer filled with %s\n", atoi(argv[7]) == 1 ? "integers 0..217" : "U[0,1)");
  }
  if (argc >= 7) {  // one configuration: grid nw tile S wait_all order [fill]
    const int grid = atoi(argv[1]), nw = atoi(argv[2]), tile = atoi(argv[3]), S = atoi(argv[4]), wait_all = atoi(argv[5]), order = atoi(argv[6]);
    const size_t smem = (size_t)nw * S * tile + (size_t)nw * S * 8 + 64;
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      warp_rings<<<grid, nw * 32, smem>>>(d, (long long)(bytes / tile), tile, S, wait_all, order, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    snprintf(what, sizeof what, "warp_rings grid=%3d nw=%2d tile=%5d S=%d wait_all=%d order=%d", grid, nw, tile, S, wait_all, order);
    report(what, best);
    // what does a launch cost when it is not back to back with itself over the same buffer?
    float* d2; cudaMalloc(&d2, bytes + (1 << 20)); cudaMemset(d2, 0, bytes);
    float* small; cudaMalloc(&small, 64 << 20); cudaMemset(small, 0, 64 << 20);
    for (int scen = 0; scen < 5; ++scen) {
      printf("scenario %d (%s):", scen, scen == 0 ? "same buffer back to back" : scen == 1 ? "two buffers alternating" : scen == 2 ? "same buffer, 64 MB memset between" : scen == 3 ? "same buffer, 300 us host gap" : "same buffer, 2 ms host gap");
      for (int rep = 0; rep < 8; ++rep) {
        if (scen == 2) cudaMemsetAsync(small, rep, 64 << 20);
        if (scen == 3 || scen == 4) { cudaDeviceSynchronize(); timespec ts{0, scen == 3 ? 300000 : 2000000}; nanosleep(&ts, nullptr); }
        cudaEventRecord(e0);
        warp_rings<<<grid, nw * 32, smem>>>((scen == 1 && (rep & 1)) ? d2 : d, (long long)(bytes / tile), tile, S, wait_all, order, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(" %.1f", ms * 1e3);
      }
      printf("\n");
    }
    return 0;
  }
  for (int grid : {148, 296}) for (int nw : {4, 8, 16, 32}) for (int tile : {2048, 4096, 8192, 16384}) for (int S : {2, 3, 4, 6, 8}) for (int wait_all : {1, 0}) for (int order : {0, 1}) {
    const size_t smem = (size_t)nw * S * tile + (size_t)nw * S * 8 + 64;
    const size_t cap = grid == 148 ? 226 * 1024 : 112 * 1024;
    if (smem > cap) continue;
    if ((size_t)nw * (S - 1) * tile * (grid / 148) < 48 * 1024) continue;  // too little in flight to matter
    if (wait_all == 0 && !(nw == 16 && tile == 4096)) continue;
    if (order == 1 && !((nw == 16 && tile == 4096) || (nw == 8 && tile == 8192))) continue;
    const long long n_tiles = bytes / tile;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      warp_rings<<<grid, nw * 32, smem>>>(d, n_tiles, tile, S, wait_all, order, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    snprintf(what, sizeof what, "warp_rings grid=%3d nw=%2d tile=%5d S=%d in_flight/SM=%3zu KB wait_all=%d order=%d", grid, nw, tile, S, (size_t)nw * (S - 1) * tile * (grid / 148) / 1024, wait_all, order);
    report(what, best);
  }
  for (int nc : {4, 8, 16}) for (int tile : {4096, 8192, 16384}) for (int S : {8, 12, 16, 24, 32}) {
    const size_t smem = (size_t)S * tile + (size_t)S * 16 + 64;
    if (smem > 226 * 1024 || (size_t)S * tile < 64 * 1024) continue;
    const long long n_tiles = bytes / tile;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      cta_ring<<<148, (nc + 1) * 32, smem>>>(d, n_tiles, tile, S, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    snprintf(what, sizeof what, "cta_ring   consumers=%2d tile=%5d S=%2d in_flight/SM=%3zu KB", nc, tile, S, (size_t)S * tile / 1024);
    report(what, best);
  }
  return 0;
}
