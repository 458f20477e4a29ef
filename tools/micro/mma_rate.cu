// Micro-benchmark: tcgen05.mma issue/execute rate, kind::f16 (bf16), M = 128, A in tensor memory (TS) or in
// shared memory (SS), for several N and numbers of independent accumulators (dependency chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_rate.bin tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc),
               "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc),
               "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// CHAINS independent accumulators used round-robin; every MMA accumulates into its chain's D. The issue
// loop is unrolled 16x inside one elected region so that descriptor arithmetic stays off the critical path.
template <int N, int CHAINS, bool SS>
__global__ void bench(int iters, unsigned long long* out, int delay = 0) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (warp == 0) {
    asm volatile("fence.proxy.async.shared::cta;");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = make_sdesc(smem_u32(smem));
    const uint64_t adesc = make_sdesc(smem_u32(smem + 32768));
    const uint32_t d_base = tmem + 128;  // A occupies columns 0..127
    const long long t0 = clock64();
    for (int it = 0; it < iters / 16; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = d_base + (uint32_t)((i % CHAINS) * N);
          const uint32_t koff = (uint32_t)((i & 3) * 2);
          if (SS) umma_ss(d, adesc + koff, bdesc + koff, idesc, 1u);
          else umma_ts(d, tmem + (uint32_t)((i & 7) * 8), bdesc + koff, idesc, 1u);
        }
      }
      __syncwarp();
      if (delay > 0) {  // the issuer is busy elsewhere (barrier polls) between groups of 16 MMAs
        const long long until = clock64() + delay;
        while (clock64() < until) {}
      }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
      out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int CHAINS, bool SS>
void run(unsigned long long* d_out) {
  if (128 + CHAINS * N > 512) return;
  const int iters = 4096;
  cudaFuncSetAttribute(bench<N, CHAINS, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  bench<N, CHAINS, SS><<<148, 128, 70 * 1024>>>(iters, d_out);
  bench<N, CHAINS, SS><<<148, 128, 70 * 1024>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s M=128 N=%3d K=16 chains %d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", SS ? "SS" : "TS", N,
         CHAINS, (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <bool SS>
void run_all(unsigned long long* d_out) {
  run<32, 1, SS>(d_out); run<32, 2, SS>(d_out); run<32, 4, SS>(d_out);
  run<64, 1, SS>(d_out); run<64, 2, SS>(d_out); run<64, 4, SS>(d_out);
  run<96, 1, SS>(d_out); run<96, 2, SS>(d_out); run<96, 4, SS>(d_out);
  run<128, 1, SS>(d_out); run<128, 2, SS>(d_out);
  run<192, 1, SS>(d_out); run<192, 2, SS>(d_out);
  run<256, 1, SS>(d_out);
}

// How deep is the MMA queue? 16 MMAs (N = 64: 512 cycles of tensor work) then `delay` cycles in which the
// issuer does something else: if the total per group stays 512 the queue absorbs the gap.
void run_delay(unsigned long long* d_out) {
  const int iters = 4096;
  cudaFuncSetAttribute(bench<64, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int delay : {0, 100, 200, 300, 400, 600}) {
    bench<64, 2, false><<<148, 128, 70 * 1024>>>(iters, d_out, delay);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("TS N=64 groups of 16 MMAs + %3d idle issuer cycles: %.1f cycles per group (tensor work 512)\n", delay,
           (double)h[1] / (iters / 16));
  }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 16 * 148);
  run_delay(d_out);
  run_all<false>(d_out);
  run_all<true>(d_out);
  cudaFree(d_out);
  return 0;
}
