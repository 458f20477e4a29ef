// Micro-benchmark: tcgen05.ld throughput per SM for several shapes and warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld tools/micro/tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(slot);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}

template <int SHAPE>
__device__ __forceinline__ uint32_t do_ld(uint32_t addr) {
  uint32_t acc = 0;
  if (SHAPE == 0) {  // 32x32b.x16
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (SHAPE == 1) {  // 32x32b.x32
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  } else if (SHAPE == 2) {  // 16x256b.x4 : 16 regs
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (SHAPE == 3) {  // 16x128b.x8 : 16 regs
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else if (SHAPE == 4) {  // 4 x (32x32b.x16) issued back to back, one wait
    uint32_t r[4][16];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]), "=r"(r[k][6]),
                     "=r"(r[k][7]), "=r"(r[k][8]), "=r"(r[k][9]), "=r"(r[k][10]), "=r"(r[k][11]), "=r"(r[k][12]),
                     "=r"(r[k][13]), "=r"(r[k][14]), "=r"(r[k][15])
                   : "r"(addr + k * 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= r[k][i];
  }
  return acc;
}

template <int SHAPE>
__global__ void bench(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  const uint32_t addr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) acc ^= do_ld<SHAPE>(addr + (uint32_t)((i & 1) * 0));
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

template <int SHAPE>
void run(const char* name, int bytes_per_warp_ld) {
  unsigned long long* d_cyc;
  uint32_t* d_sink;
  cudaMalloc(&d_cyc, 8 * 148);
  cudaMalloc(&d_sink, 4 * 148 * 1024);
  for (int warps : {1, 4, 8, 16}) {
    if (warps * 64 > 512 * 4 / 4 && false) continue;
    const int iters = 2000;
    bench<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
    bench<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double bpc = (double)bytes_per_warp_ld * warps * iters / (double)h[0];
    printf("%-28s warps %2d: %8llu cycles for %d loads/warp -> %.1f B/cycle/SM (%s)\n", name, warps, h[0], iters, bpc,
           cudaGetErrorString(e));
  }
  cudaFree(d_cyc);
  cudaFree(d_sink);
}

int main() {
  run<0>("32x32b.x16 (2 KB)", 2048);
  run<1>("32x32b.x32 (4 KB)", 4096);
  run<2>("16x256b.x4 (2 KB)", 2048);
  run<3>("16x128b.x8 (2 KB)", 2048);
  run<4>("4 x 32x32b.x16, one wait (8 KB)", 8192);
  return 0;
}
