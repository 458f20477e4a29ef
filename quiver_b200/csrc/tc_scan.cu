// tc_scan.cu — see tc_scan.cuh. sm_100a only: TMA (cp.async.bulk.tensor), tcgen05.mma kind::tf32,
// TMEM accumulators, tcgen05.ld epilogue.
//
// One pass serves up to 256 queries (the MMA N dimension). Three kernels per pass:
//   1. tc_scan_kernel<MODE, SAMPLE=true>   scores of a strided sample of corpus tiles; per tile and
//      query the two smallest scores are kept;
//   2. tc_tau_kernel                       per query, the TC_SAMPLE_RANK-th smallest sampled score
//      becomes the admission threshold tau (expected ~256 corpus rows pass per query);
//   3. tc_scan_kernel<MODE, SAMPLE=false>  persistent scan of every tile: 128 x N accumulator tile
//      in TMEM -> score -> `score <= tau` -> (rare) append of the (score,row) key to the query's
//      candidate list in global memory.
// finalize_cand_kernel (finalize.cu) then re-ranks the candidates exactly and certifies the result:
// tau is only a performance heuristic, never a correctness assumption.
//
// Pipelines (mbarriers): full/empty ring of A stages between the TMA warp and the MMA thread;
// tmem_full/tmem_empty over two accumulator buffers between the MMA thread and the epilogue warps.
#include <cuda.h>

#include "scan.cuh"
#include "tc_scan.cuh"

namespace qg {

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

struct TcKParams {
  long long n_rows;
  long long n_tiles;
  int n_cols, kb, stages, nq;
  uint32_t idesc;
  int cosine;
  const float* row_norm2;
  const float* inv_norm;
  const uint32_t* mask;
  const float* queries;  // [nq x dp] (query norms for cosine)
  int dp;
  uint32_t* sample;
  int n_sample;
  const float* tau;
  uint64_t* cand;
  int* cand_cnt;
};

// shared memory carve-up (offsets from the 1024-aligned base)
struct TcSmem {
  int off_b, off_a, off_bars, off_tmem, off_tau, off_rnq, off_scratch, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int n_cols, int kb, int stages) {
  TcSmem s;
  s.off_b = 0;
  s.off_a = s.off_b + kb * n_cols * 128;
  s.off_bars = s.off_a + stages * TC_STAGE_BYTES;
  s.off_tmem = s.off_bars + (2 * stages + 5) * 8;
  s.off_tau = (s.off_tmem + 4 + 15) & ~15;
  s.off_rnq = s.off_tau + TC_MAX_COLS * 4;
  s.off_scratch = s.off_rnq + TC_MAX_COLS * 4;
  s.total = s.off_scratch + 4 * TC_MAX_COLS * 2 * 4;
  return s;
}

template <int MODE, bool SAMPLE>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_scan_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const TcKParams p) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                         ~(uintptr_t)1023);
  const TcSmem L = tc_smem_layout(p.n_cols, p.kb, p.stages);
  unsigned char* sB = smem + L.off_b;
  unsigned char* sA = smem + L.off_a;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bars);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.off_tmem);
  float* s_tau = reinterpret_cast<float*>(smem + L.off_tau);
  float* s_rnq = reinterpret_cast<float*>(smem + L.off_rnq);
  uint32_t* s_scratch = reinterpret_cast<uint32_t*>(smem + L.off_scratch);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = p.n_cols, kb = p.kb, S = p.stages;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * n_cols) tmem_cols <<= 1;

  // work list of this CTA: MAIN = tiles b, b+G, ...; SAMPLE = sampled tiles s = b, b+G, ... (< n_sample)
  const long long n_work = SAMPLE ? (long long)p.n_sample : p.n_tiles;
  auto tile_of = [&](long long w) -> long long {
    return SAMPLE ? (w * p.n_tiles) / p.n_sample : w;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8);
    mbar_init(&tmem_empty[1], 8);
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  // thresholds and query norms
  for (int i = threadIdx.x; i < TC_MAX_COLS; i += blockDim.x) {
    float t = __int_as_float(0x7f800000);
    if (!SAMPLE && i < p.nq) t = p.tau[i];
    if (!SAMPLE && i >= p.nq) t = -__int_as_float(0x7f800000);  // padded columns never pass
    s_tau[i] = t;
  }
  if (MODE == MODE_DOT) {
    for (int qi = warp; qi < TC_MAX_COLS; qi += TC_THREADS / 32) {
      float s = 0.f;
      if (qi < p.nq && p.cosine) {
        const float* qv = p.queries + (size_t)qi * p.dp;
        for (int i = lane; i < p.dp; i += 32) s = fmaf(qv[i], qv[i], s);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      if (lane == 0) s_rnq[qi] = p.cosine ? (s > 0.f ? rsqrtf(s) : 0.f) : 1.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, (uint32_t)(kb * n_cols * 128));
      for (int kbi = 0; kbi < kb; ++kbi) tma_load_2d(sB + (size_t)kbi * n_cols * 128, &tm_b, kbi * TC_KBLOCK, 0, b_full);
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
        const long long tile = tile_of(w);
        for (int kbi = 0; kbi < kb; ++kbi) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
          tma_load_2d(sA + (size_t)stage * TC_STAGE_BYTES, &tm_a, kbi * TC_KBLOCK, (int)(tile * TC_TILE_ROWS),
                      &full[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(b_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * n_cols);
        for (int kbi = 0; kbi < kb; ++kbi) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * TC_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)kbi * n_cols * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(d_tmem, make_sdesc(a_addr + k * 32), make_sdesc(b_addr + k * 32), p.idesc,
                      (uint32_t)((kbi | k) != 0));
          }
          umma_commit(&empty[stage]);  // frees the A stage once these MMAs have read it
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps =====
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = (warp - 2) >> 2;      // interleaved 16-column chunks
    const int n_chunks = n_cols >> 4;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    long long it = 0;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const long long tile = tile_of(w);
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const long long row = tile * TC_TILE_ROWS + quarter * 32 + lane;
      bool valid = row < p.n_rows;
      if (valid && p.mask != nullptr) valid = (__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u;
      float xn = 0.f;
      if (valid) xn = (MODE == MODE_L2) ? __ldg(p.row_norm2 + row) : (p.inv_norm ? __ldg(p.inv_norm + row) : 1.f);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();

      for (int c = half; c < n_chunks; c += 2) {
        float a[16];
        tmem_ld16(tmem_base + lane_base + (uint32_t)(acc * n_cols + c * 16), a);
        float v[16];
        if (MODE == MODE_L2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(-2.f, a[j], xn);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(-a[j], xn * s_rnq[c * 16 + j], 1.0f);
        }
        if (SAMPLE) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t key = valid ? f32_to_ordered(v[j]) : 0xFFFFFFFFu;
            const uint32_t m1 = __reduce_min_sync(0xffffffffu, key);
            const unsigned who = __ballot_sync(0xffffffffu, key == m1);
            const uint32_t key2 = (lane == (__ffs(who) - 1)) ? 0xFFFFFFFFu : key;
            const uint32_t m2 = __reduce_min_sync(0xffffffffu, key2);
            if (lane == 0) {
              s_scratch[(quarter * TC_MAX_COLS + c * 16 + j) * 2 + 0] = m1;
              s_scratch[(quarter * TC_MAX_COLS + c * 16 + j) * 2 + 1] = m2;
            }
          }
        } else {
          const float4* t4 = reinterpret_cast<const float4*>(s_tau + c * 16);
          float tq[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 t = t4[j4];
            tq[j4 * 4 + 0] = t.x;
            tq[j4 * 4 + 1] = t.y;
            tq[j4 * 4 + 2] = t.z;
            tq[j4 * 4 + 3] = t.w;
          }
          float m = __int_as_float(0x7f800000);
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fminf(m, v[j] - tq[j]);
          const bool hit = valid && (m <= 0.f);
          if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int qi = c * 16 + j;
              const bool pass = valid && (v[j] <= tq[j]);
              const unsigned mk = __ballot_sync(0xffffffffu, pass);
              if (mk != 0u) {
                const int leader = __ffs(mk) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(p.cand_cnt + qi, __popc(mk));
                base = __shfl_sync(0xffffffffu, base, leader);
                const int pos = base + __popc(mk & ((1u << lane) - 1u));
                if (pass && pos < TC_CAND_CAP) p.cand[(size_t)qi * TC_CAND_CAP + pos] = make_key(v[j], (uint32_t)row);
              }
            }
          }
        }
      }
      // this warp is done with the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);

      if (SAMPLE) {
        named_bar_sync(1, 256);
        const int e = threadIdx.x - 64;  // 0..255 over the epilogue warps
        if (e < n_cols) {
          uint32_t b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu;
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const uint32_t x = s_scratch[(qd * TC_MAX_COLS + e) * 2 + r];
              if (x < b0) {
                b1 = b0;
                b0 = x;
              } else if (x < b1) {
                b1 = x;
              }
            }
          }
          uint32_t* out = p.sample + ((size_t)e * p.n_sample + (size_t)w) * 2;
          out[0] = b0;
          out[1] = b1;
        }
        named_bar_sync(1, 256);
      }
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Threshold per query: the TC_SAMPLE_RANK-th smallest sampled score (+inf when the sample holds
// fewer valid scores). One warp per query; bitwise binary search over the ordered-float images.
// Also clears the candidate counters of the pass.
__global__ void __launch_bounds__(32) tc_tau_kernel(const uint32_t* __restrict__ sample, int n_sample, int rank,
                                                    float* __restrict__ tau, int* __restrict__ cand_cnt) {
  const int q = blockIdx.x, lane = threadIdx.x;
  const uint32_t* vals = sample + (size_t)q * n_sample * 2;
  const int n = n_sample * 2;
  uint32_t prefix = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t probe = prefix | ((1u << bit) - 1u);  // largest value with this prefix and the bit clear
    int c = 0;
    for (int i = lane; i < n; i += 32) c += (vals[i] <= probe) && (vals[i] != 0xFFFFFFFFu);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < rank) prefix |= (1u << bit);
  }
  if (lane == 0) {
    // prefix == 0xFFFFFFFF: fewer than `rank` valid samples -> admit everything
    tau[q] = (prefix == 0xFFFFFFFFu) ? __int_as_float(0x7f800000) : ordered_to_f32(prefix);
    cand_cnt[q] = 0;
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static bool g_encode_tried = false;

int tc_available() {
  if (!g_encode_tried) {
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return g_encode ? 0 : -1;
}

static int encode_map(CUtensorMap* map, const float* base, long long rows, int dp, int box_rows) {
  if (tc_available() != 0) return fail(4, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)dp * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_KBLOCK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(4, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

int tc_plan(int dp, int nq, TcPlan* out) {
  if (dp < 4 || dp % 4 != 0) return -1;
  const int kb = (dp + TC_KBLOCK - 1) / TC_KBLOCK;
  const int budget = 227 * 1024 - 1024;  // alignment slack
  int n_cols = std::min(TC_MAX_COLS, ((nq + 15) / 16) * 16);
  for (;;) {
    if (n_cols < 16) return -1;
    const TcSmem fixed = tc_smem_layout(n_cols, kb, 0);
    const int left = budget - fixed.total;
    int stages = left / (TC_STAGE_BYTES + 16);
    // at least 64 KB of corpus loads in flight per SM (Little's law at ~6.5 TB/s over 148 SMs)
    if (stages >= 4) {
      stages = std::min(stages, 12);
      out->n_cols = n_cols;
      out->kb = kb;
      out->stages = stages;
      out->smem = tc_smem_layout(n_cols, kb, stages).total + 1024;
      return 0;
    }
    n_cols -= 16;
  }
}

template <int MODE, bool SAMPLE>
static int set_attr_one() {
  QG_CUDA_OK(cudaFuncSetAttribute(tc_scan_kernel<MODE, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  return 0;
}

int tc_set_attributes() {
  if (int e = set_attr_one<MODE_L2, false>()) return e;
  if (int e = set_attr_one<MODE_L2, true>()) return e;
  if (int e = set_attr_one<MODE_DOT, false>()) return e;
  if (int e = set_attr_one<MODE_DOT, true>()) return e;
  return 0;
}

int launch_tc_pass(const TcPlan& plan, const TcArgs& a, int sm_count, cudaStream_t st, int* launches) {
  CUtensorMap tm_a, tm_b;
  if (int rc = encode_map(&tm_a, a.vec, a.n_rows, a.dp, TC_TILE_ROWS)) return rc;
  if (int rc = encode_map(&tm_b, a.queries, a.nq, a.dp, plan.n_cols)) return rc;
  TcKParams p{};
  p.n_rows = a.n_rows;
  p.n_tiles = (a.n_rows + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
  p.n_cols = plan.n_cols;
  p.kb = plan.kb;
  p.stages = plan.stages;
  p.nq = a.nq;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(plan.n_cols >> 3) << 17) |
            ((uint32_t)(TC_TILE_ROWS >> 4) << 24);
  p.cosine = a.cosine;
  p.row_norm2 = a.row_norm2;
  p.inv_norm = a.inv_norm;
  p.mask = a.mask;
  p.queries = a.queries;
  p.dp = a.dp;
  p.sample = a.sample;
  p.n_sample = a.n_sample;
  p.tau = a.tau;
  p.cand = a.cand;
  p.cand_cnt = a.cand_cnt;
  const int grid_s = (int)std::min<long long>(sm_count, a.n_sample);
  const int grid_m = (int)std::min<long long>(sm_count, p.n_tiles);
  if (a.mode == MODE_L2) {
    tc_scan_kernel<MODE_L2, true><<<grid_s, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  } else {
    tc_scan_kernel<MODE_DOT, true><<<grid_s, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  }
  QG_CUDA_OK(cudaGetLastError());
  tc_tau_kernel<<<a.nq, 32, 0, st>>>(a.sample, a.n_sample, TC_SAMPLE_RANK, a.tau, a.cand_cnt);
  QG_CUDA_OK(cudaGetLastError());
  if (a.mode == MODE_L2) {
    tc_scan_kernel<MODE_L2, false><<<grid_m, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  } else {
    tc_scan_kernel<MODE_DOT, false><<<grid_m, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  }
  QG_CUDA_OK(cudaGetLastError());
  if (launches) *launches += 3;
  return 0;
}

}  // namespace qg
