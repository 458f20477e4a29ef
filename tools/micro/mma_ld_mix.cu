// Micro-benchmark: does tcgen05.ld traffic from the epilogue warps slow the tensor pipe down (and vice
// versa)? Warp 0 issues TS-form kind::f16 MMAs (M = 128, N = 128, K = 16) back to back into three
// accumulator units while LDW warps read accumulator columns with 4 x tcgen05.ld.32x32b.x16 + one wait in
// a loop. Reported: cycles per MMA and tcgen05.ld bytes per cycle per SM, alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mma_ld_mix.bin tools/micro/mma_ld_mix.cu
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

// MMAS: number of MMAs the issuer runs (0 = none); ld warps loop `ld_iters` times (0 = none), or — when
// ld_iters < 0 — until the issuer is done.
__global__ void bench(int mmas, int ld_iters, int ld_mode, unsigned long long* out, uint32_t* sink) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int done_flag;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    done_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (warp == 0) {
    if (mmas > 0) {
      asm volatile("fence.proxy.async.shared::cta;");
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t bdesc = make_sdesc(smem_u32(smem));
      const uint32_t d_base = tmem + 128;
      const long long t0 = clock64();
      for (int it = 0; it < mmas / 8; ++it) {
        if (elect_one()) {
          const uint32_t d = d_base + (uint32_t)((it % 3) * 128);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            umma_ts(d, tmem + (uint32_t)(i * 8), bdesc + (uint64_t)((i & 3) * 2), idesc, i != 0 ? 1u : 0u);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (threadIdx.x == 0) out[blockIdx.x * 4 + 0] = (unsigned long long)(t2 - t0);
    }
    if (threadIdx.x == 0) done_flag = 1;
  } else if (warp >= 4) {
    const int e = warp - 4;
    // a warp may only touch TMEM lanes 32 * (warp % 4) ..; the warps of a quarter take different column ranges
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (uint32_t)((e >> 2) * 64 % 384);
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    for (int i = 0; ld_iters < 0 ? !done_flag : i < ld_iters; ++i) {
      uint32_t r[4][16];
      if (ld_mode == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]), "=r"(r[k][6]),
                         "=r"(r[k][7]), "=r"(r[k][8]), "=r"(r[k][9]), "=r"(r[k][10]), "=r"(r[k][11]), "=r"(r[k][12]),
                         "=r"(r[k][13]), "=r"(r[k][14]), "=r"(r[k][15])
                       : "r"(addr + k * 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(r[k][0]), "=r"(r[k][1]), "=r"(r[k][2]), "=r"(r[k][3]), "=r"(r[k][4]), "=r"(r[k][5]), "=r"(r[k][6]),
                         "=r"(r[k][7]), "=r"(r[k][8]), "=r"(r[k][9]), "=r"(r[k][10]), "=r"(r[k][11]), "=r"(r[k][12]),
                         "=r"(r[k][13]), "=r"(r[k][14]), "=r"(r[k][15])
                       : "r"(addr + k * 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[k][j];
      ++n;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      out[blockIdx.x * 4 + 1] = (unsigned long long)(t1 - t0);
      if (e == 0) out[blockIdx.x * 4 + 2] = (unsigned long long)n;
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 32 * 148);
  cudaMalloc(&d_sink, 4 * 148 * 1024);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int mmas = 8192;
  for (int ld_mode = 0; ld_mode < 2; ++ld_mode) {
    for (int ldw : {0, 4, 8, 16}) {
      for (int with_mma = 0; with_mma < 2; ++with_mma) {
        if (!with_mma && ldw == 0) continue;
        if (ld_mode == 1 && ldw == 0) continue;
        cudaMemset(d_out, 0, 32 * 148);
        for (int rep = 0; rep < 2; ++rep)
          bench<<<148, 128 + ldw * 32, 70 * 1024>>>(with_mma ? mmas : 0, ldw == 0 ? 0 : (with_mma ? -1 : 4000), ld_mode, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[4];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("ld warps %2d (%s)%s: ", ldw, ld_mode ? "wait after every x16" : "4 x x16 then one wait", with_mma ? " + MMA stream" : "             ");
        if (with_mma) printf("%.1f cycles per MMA (ideal 64); ", (double)h[0] / mmas);
        if (ldw) printf("tcgen05.ld %.1f B/cycle/SM (%llu loops of 8 KB per warp in %llu cycles)", 8192.0 * ldw * (double)h[2] / (double)h[1], h[2], h[1]);
        printf(" %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
