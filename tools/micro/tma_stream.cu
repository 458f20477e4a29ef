// Micro-benchmark: how fast can one CTA per SM stream a row-major [rows x 128] fp32 matrix into
// shared memory (a) with 2-D tiled TMA boxes of 32 floats x R rows (the UMMA K-major SW128 operand
// shape) and (b) with 1-D bulk copies of the same number of contiguous bytes, as a function of the
// number of stages in flight.  nvcc -arch=sm_100a -o tma_stream tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: 2-D boxes (32 floats x R rows), 4 boxes per tile; mode 1: 1-D bulk of the same bytes (contiguous)
template <int MODE>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const float* base, long long n_tiles, int R, int stages, unsigned long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int stage_bytes = R * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // producer
    int s = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect(&full[s], stage_bytes);
        if (MODE == 0) tma2d(smem + (size_t)s * stage_bytes, &tm, kb * 32, (int)(t * R), &full[s]);
        else bulk1d(smem + (size_t)s * stage_bytes, base + ((size_t)t * 4 + kb) * (stage_bytes / 4), stage_bytes, &full[s]);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (threadIdx.x == 32) {  // consumer: frees the stage as soon as it has landed
    int s = 0; uint32_t ph = 0; unsigned long long acc = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&full[s], ph);
        acc += *reinterpret_cast<volatile unsigned int*>(smem + (size_t)s * stage_bytes);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

int main() {
  const long long rows = 1000000; const int dp = 128;
  float* d; cudaMalloc(&d, (size_t)rows * dp * 4 + (1 << 20)); cudaMemset(d, 0, (size_t)rows * dp * 4);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Enc enc = (Enc)fn;
  cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int R : {64, 128}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)dp * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    for (int l2 = 0; l2 < 2; ++l2) {
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const long long n_tiles = rows / R;
      for (int mode = 0; mode < 2; ++mode) {
        if (mode == 1 && l2 == 1) continue;
        for (int stages : {4, 8, 12, 16, 24}) {
          const int stage_bytes = R * 128;
          const size_t smem = (size_t)stages * stage_bytes + stages * 16 + 64;
          if (smem > 226 * 1024) continue;
          float best = 1e9;
          for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) stream_kernel<0><<<148, 64, smem>>>(tm, d, n_tiles, R, stages, sink);
            else stream_kernel<1><<<148, 64, smem>>>(tm, d, n_tiles, R, stages, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
          }
          cudaError_t err = cudaGetLastError();
          printf("R=%3d %s l2promo=%d stages=%2d in_flight=%4d KB  %.1f us  %.0f GB/s %s\n", R, mode ? "bulk1d" : "tma2d ", l2, stages, stages * stage_bytes / 1024, best * 1e3, 512e6 / (best * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
      }
    }
  }
  return 0;
}
