// tc_scan.cu — see tc_scan.cuh. sm_100a only: TMA (cp.async.bulk.tensor), tcgen05.mma kind::f16 (bf16
// stream) / kind::tf32, TMEM accumulators, tcgen05.ld epilogue.
//
// One pass serves up to 256 queries. Kernels:
//   1. tc_ts_kernel<.., SAMPLE=true>   scores of a strided sample of corpus tiles; per tile and query
//      the minimum of each half of the tile's chunks is kept. For a multi-pass search it runs ONCE
//      for all passes (grid.y = pass);
//   2. tc_tau_kernel                   per query, the sample_rank-th smallest sampled score becomes the
//      admission threshold tau (expected ~256 corpus rows pass per query); also once per search;
//   3. tc_ts_kernel<.., SAMPLE=false>  persistent scan of every tile of one pass: queries resident in
//      TMEM (A operand), corpus tiles through a TMA ring (B operand), one 128-column accumulator unit
//      per (tile, query block) -> score -> `score <= tau` -> (rare) append of the (score,row) key to
//      the query's candidate list in global memory.
//   (tc_scan_kernel is the SS form for dims too large to keep the queries in TMEM.)
// finalize_cand_kernel (finalize.cu) then re-ranks the candidates exactly and certifies the result
// (once per group of TC_PASS_GROUP passes): tau is only a performance heuristic, never a correctness
// assumption. All launches of a search are chained with programmatic dependent launch.
//
// Pipelines (mbarriers): full/empty ring of corpus stages between the TMA warp and the MMA issuer;
// tmem_full/tmem_empty over 2-3 accumulator units between the MMA issuer and the two epilogue groups;
// xs_full/xs_empty ring of per-tile row terms (masked / non-raw scans only).
#include <cuda.h>
#include <cuda_bf16.h>

#include "scan.cuh"
#include "tc_scan.cuh"
#include "misc.cuh"
#include <cstdlib>

namespace qg {

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One lane of a CONVERGED warp. The single-thread instructions (TMA, tcgen05.mma, tcgen05.commit) are
// issued under this predicate from warp-uniform loops; wrapping a whole loop in `if (lane == 0)`
// instead makes the warp divergent and ptxas then serialises every uniform-datapath instruction
// behind an ELECT / BRA.U.ANY loop (measured: ~90 cycles per MMA issue).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
// A operand from tensor memory (the resident query block), B from shared memory.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with bf16 operands (kind::f16): K = 16 elements per instruction, twice the tf32 rate.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
      "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])),
      "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])),
      "r"(__float_as_uint(r[11])), "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])),
      "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
    // ===== TMA producer (converged warp, one elected lane issues) =====
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)(kb * n_cols * 128));
      for (int kbi = 0; kbi < kb; ++kbi) tma_load_2d(sB + (size_t)kbi * n_cols * 128, &tm_b, kbi * TC_KBLOCK, 0, b_full);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (long long w = blockIdx.x; w < n_work; w += gridDim.x) {
      const long long tile = tile_of(w);
      for (int kbi = 0; kbi < kb; ++kbi) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
          tma_load_2d(sA + (size_t)stage * TC_STAGE_BYTES, &tm_a, kbi * TC_KBLOCK, (int)(tile * TC_TILE_ROWS),
                      &full[stage]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, one elected lane issues) =====
    mbar_wait(b_full, 0);
    tc_fence_after();
    const uint64_t adesc0 = make_sdesc(smem_u32(sA));
    const uint64_t bdesc0 = make_sdesc(smem_u32(sB));
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
        if (elect_one()) {
          const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)stage * (TC_STAGE_BYTES >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)kbi * (uint32_t)((n_cols * 128) >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (uint32_t)((kbi | k) != 0));
          umma_commit(&empty[stage]);  // frees the A stage once these MMAs have read it
          if (kbi == kb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
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


// ================================================================================================
// TS variant (dp <= 256): the QUERIES are the A operand, resident in tensor memory for the whole
// pass (tcgen05.st once per CTA), and the corpus tile is the B operand streamed through shared
// memory. D[query lane][corpus row column] — so every epilogue thread owns one query: its threshold
// lives in a register, a tile's per-row terms are broadcast from shared memory, and all of shared
// memory is left to the corpus ring (~190 KB in flight per SM, which the HBM pipe needs).
//   TMEM columns: [0, nblk*kb*32) query blocks | then (blk*2 + acc) * 64 accumulators
//   warps: 0 TMA (corpus tiles + the tile's per-row terms), 1 MMA + TMEM alloc, 2..9 epilogue
// ================================================================================================
// How the roles of the TS kernel wait on their barriers. mbarrier.try_wait may park a thread for a while before it
// looks again; test_wait in a loop notices the phase flip at once but spends issue slots. QG_TC_POLL: bit 0 = the
// control warps (TMA producer, MMA issuer) poll, bit 1 = the epilogue warps poll.
#ifndef QG_TC_POLL
#define QG_TC_POLL 0
#endif
__device__ __forceinline__ void ts_wait_ctrl(uint64_t* bar, uint32_t parity) {
  if (QG_TC_POLL & 1) mbar_poll(bar, parity);
  else mbar_wait(bar, parity);
}
__device__ __forceinline__ void ts_wait_epi(uint64_t* bar, uint32_t parity) {
  if (QG_TC_POLL & 2) mbar_poll(bar, parity);
  else mbar_wait(bar, parity);
}
constexpr int TS_KSTEP_BYTES = 128;               // one swizzle row: 32 floats
constexpr int TS_EPI_WARPS = 16;                 // epilogue warps: four per TMEM lane quarter
constexpr int TS_THREADS = 128 + TS_EPI_WARPS * 32;  // warp group 0: TMA producer, MMA issuer, two idle warps
constexpr int TS_REGS_CTRL = 56;                 // setmaxnreg: control warp group gives registers away ...
constexpr int TS_REGS_EPI = 104;                 // ... to the epilogue warp groups (64 accumulator registers each)
constexpr int TS_MAX_ACC = 4;        // accumulator buffers in tensor memory (as many as fit beside the queries)
constexpr int TS_CLAIM = 2;          // corpus tiles per work claim of the main scan
constexpr int TS_XS = 8;                          // ring of per-tile row-term buffers
// corpus rows per tile = MMA N = 128. The accumulator of one (tile, query block) pair is a *unit* of
// 128 TMEM columns; as many unit buffers as fit beside the resident queries rotate (2 or 3).
__host__ __device__ constexpr int ts_rows(int /*nblk*/) { return 128; }

struct TsKParams {
  long long n_rows;
  long long n_tiles;
  long long n_sample_from;  // tiles the sample is drawn from (raw scan: full tiles only, the out-of-bounds rows of a partial tile read as zeros)
  int kb, ksteps, a_cols, stages, nq, nblk;
  uint32_t idesc;
  int cosine;
  const float* bias;   // [rows padded to 128] additive row term: |x|^2 (L2) or 1 (dot); +inf = row excluded
  const float* sc;     // [rows padded to 128] 1/|x| (cosine) or nullptr
  const uint32_t* apack;  // this pass's queries in tensor-memory order (tc_pack_kernel)
  long long pass_stride;  // sample stage over a whole search: 32-bit words between two passes' query blocks
                          // (blockIdx.y = pass); nq then counts all queries of the search
  int* work_counter;      // main scan: tiles beyond the first of each CTA are claimed here (zeroed by tc_tau_kernel)
  int dp;
  uint32_t* sample;  // [n_cols][n_sample][2]: one minimum per tile and chunk parity
  int n_sample;
  const float* tau;
  uint64_t* cand;
  int* cand_cnt;
  unsigned long long* dbg;  // optional per-role cycle counters of CTA 0 (development aid), else nullptr
};

struct TsSmem {
  int off_ring, off_bias, off_sc, off_bars, off_tmem, off_tags, total;
};
__host__ __device__ inline TsSmem ts_smem_layout(int stages, int kb, int rows) {
  TsSmem s;
  s.off_ring = 0;
  s.off_bias = stages * kb * rows * TS_KSTEP_BYTES;  // one stage = one whole tile (kb k-blocks)
  s.off_sc = s.off_bias + TS_XS * rows * 4;
  s.off_bars = s.off_sc + TS_XS * rows * 4;
  s.off_tmem = s.off_bars + (2 * stages + 2 * TS_MAX_ACC + 2 * TS_XS + 1) * 8;
  // int ring_tag[stages], xs_work[TS_XS], acc_work[TS_MAX_ACC]: the work index travelling with a ring
  // stage / row-term slot / accumulator buffer (-1 = end of work)
  s.off_tags = s.off_tmem + 16;
  s.total = s.off_tags + (stages + TS_XS + TS_MAX_ACC) * 4;
  return s;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// The per-role cycle counters and wall-clock stamps below cost ~35 predicated instructions per tile and
// epilogue warp even when switched off at run time, so they are compiled in only with
// -DQG_TC_INSTRUMENT (make TC_INSTRUMENT=1; tools/tc_timing.py needs that build).
#ifdef QG_TC_INSTRUMENT
constexpr bool TS_INSTRUMENT = true;
#else
constexpr bool TS_INSTRUMENT = false;
#endif

// wall-clock stamps of the first and the last CTA (debug counters 40..47 and 48..55)
__device__ __forceinline__ void dbg_stamp(unsigned long long* dbg, int slot) {
  if (!TS_INSTRUMENT || dbg == nullptr) return;
  if (blockIdx.x == 0) dbg[40 + slot] = global_ns();
  else if (blockIdx.x == gridDim.x - 1) dbg[48 + slot] = global_ns();
}

// Cycle counters of CTA 0 (development aid): laps are added straight to global memory by one lane, so
// a production launch (dbg == nullptr) carries two dead registers instead of a dozen live counters.
struct DbgClock {
  unsigned long long* d;
  long long t0;
  __device__ __forceinline__ void start() {
    if (TS_INSTRUMENT && d != nullptr) t0 = clock64();
  }
  __device__ __forceinline__ void lap(int slot) {
    if (TS_INSTRUMENT && d != nullptr) {
      const long long t1 = clock64();
      d[slot] += (unsigned long long)(t1 - t0);
      t0 = t1;
    }
  }
};

__device__ __forceinline__ int lds_s32(const volatile int* p) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const int*>(p))) : "memory");
  return v;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// KB > 0: compile-time number of 128-byte k-blocks per row (MMA issue loop fully unrolled); KB == 0: p.kb.
// BF16: the corpus stream is the bf16 copy of the index (64 elements per k-block, kind::f16 MMA) and the
// resident queries are rounded to bf16 — the scores only select candidates, the re-rank stays exact.
// RAW (BF16 only, no mask): the bf16 rows carry the pieces of -|x|^2 / 2 behind the vector and the
// queries carry ones there, so the accumulator already holds q.x - |x|^2 / 2 (L2) or q.x (dot, cosine on
// normalised rows): no row-term ring, the epilogue is a max tree over the raw accumulator.
template <int MODE, bool SAMPLE, int NBLK, int KB, bool BF16, bool RAW>
__global__ void __launch_bounds__(TS_THREADS, 1)
    tc_ts_kernel(const __grid_constant__ CUtensorMap tm_x, const TsKParams p) {
  constexpr int ROWS = ts_rows(NBLK);
  constexpr int KELEMS = BF16 ? 64 : 32;               // elements per 128-byte k-block
  constexpr int KBLOCK_BYTES = ROWS * TS_KSTEP_BYTES;  // one 32-float block of a tile
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) &
                                                         ~(uintptr_t)1023);
  const int kb = KB > 0 ? KB : p.kb;
  const int S = p.stages;
  const TsSmem L = ts_smem_layout(S, kb, ROWS);
  unsigned char* ring = smem + L.off_ring;
  float* xs_bias = reinterpret_cast<float*>(smem + L.off_bias);
  float* xs_sc = reinterpret_cast<float*>(smem + L.off_sc);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bars);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = empty + S;
  uint64_t* tmem_empty = tmem_full + TS_MAX_ACC;
  uint64_t* xs_full = tmem_empty + TS_MAX_ACC;
  uint64_t* xs_empty = xs_full + TS_XS;
  uint64_t* a_ready = xs_empty + TS_XS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.off_tmem);
  volatile int* ring_tag = reinterpret_cast<volatile int*>(smem + L.off_tags);
  volatile int* xs_work = ring_tag + S;
  volatile int* acc_work = xs_work + TS_XS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_cols = p.a_cols;                   // TMEM columns of one query block
  const int ksteps = p.ksteps;                   // MMA k-steps per row (the last k-block may be partial)
  const int tile_bytes = kb * KBLOCK_BYTES;      // one ring stage = one whole tile
  const int d_off = NBLK * a_cols;               // first accumulator column
  // Accumulator units: one (tile, query block) pair = ROWS columns. Unit u of iteration `it` is
  // u = it (one block) or 2 * it + blk (two blocks); it lives in buffer u % n_acc and is drained by
  // epilogue group u & 1 — the tile-parity ping-pong half (one block) or the block's warps (two blocks).
  const int n_acc = min(TS_MAX_ACC, (512 - d_off) / ROWS);
  const long long n_work = SAMPLE ? (long long)p.n_sample : p.n_tiles;
  const int py = SAMPLE ? (int)blockIdx.y : 0;  // pass of the search this CTA samples for
  auto tile_of = [&](long long w) -> long long { return SAMPLE ? (w * p.n_sample_from) / p.n_sample : w; };

  // Chained launch: the successor may be launched right away; this kernel's own setup (barriers, TMEM,
  // resident queries) does not depend on the predecessor, everything after pdl_wait() may.
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    dbg_stamp(p.dbg, 0);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < TS_MAX_ACC; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], TS_EPI_WARPS / 2);  // a unit is drained by one group of 8 warps
    }
    for (int i = 0; i < TS_XS; ++i) {
      mbar_init(&xs_full[i], 1);
      mbar_init(&xs_empty[i], NBLK == 2 ? TS_EPI_WARPS : TS_EPI_WARPS / 2);  // both blocks read a tile's row terms
    }
    mbar_init(a_ready, TS_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 1);

  // register budgets are set inside the role branches (after a merge point ptxas assumes the lower one)
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TS_REGS_CTRL));
  if (warp == 0) {
    // ===== TMA producer (converged warp, one elected lane issues): whole corpus tiles and the
    //       tile's per-row terms =====
    int stage = 0;
    uint32_t phase = 0;
    long long it = 0;
    const bool with_sc = MODE != MODE_L2 && p.sc != nullptr;
    const uint32_t xs_bytes = (uint32_t)(ROWS * 4 * (with_sc ? 2 : 1));
    DbgClock clk{(p.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr, 0};
    const long long t_prod_begin = TS_INSTRUMENT ? clock64() : 0;
    // the threshold kernel zeroes the work counter (main scan); the previous pass's finalize still reads
    // nothing this kernel writes before this point
    pdl_wait();
    int claim_a = 0, claim_b = 0;
    if (!SAMPLE && lane == 0) {
      claim_a = atomicAdd(p.work_counter, TS_CLAIM);
      claim_b = atomicAdd(p.work_counter, TS_CLAIM);
    }
    // Work distribution of the main scan: tiles are claimed from a global counter, TS_CLAIM at a time
    // (SMs stream at visibly different rates; a static split leaves the fast ones idle for the last
    // fifth of the kernel). A claim's round trip is longer than a tile, so two claims are kept in
    // flight in two different registers (the scoreboard is per register, not per lane). The sample
    // kernel keeps the static split (its slots are indexed by w). The end of work travels through the
    // rings as a tag: once through the corpus ring (MMA issuer) and through two consecutive slots of
    // the row-term ring (one per epilogue half).
    auto push = [&](long long w, bool live, bool to_ring) {
      const long long tile = live ? tile_of(w) : 0;
      if (!RAW) {
        const int xb = (int)(it % TS_XS);
        const uint32_t xphase = (uint32_t)((it / TS_XS) & 1);
        clk.start();
        ts_wait_ctrl(&xs_empty[xb], xphase ^ 1u);
        clk.lap(1);
        if (elect_one()) {
          xs_work[xb] = live ? (int)w : -1;
          if (live) {
            mbar_arrive_expect_tx(&xs_full[xb], xs_bytes);
            bulk_g2s(xs_bias + xb * ROWS, p.bias + tile * ROWS, ROWS * 4, &xs_full[xb]);
            if (with_sc) bulk_g2s(xs_sc + xb * ROWS, p.sc + tile * ROWS, ROWS * 4, &xs_full[xb]);
          } else {
            mbar_arrive(&xs_full[xb]);
          }
        }
        __syncwarp();
      }
      if (to_ring) {
        clk.start();
        ts_wait_ctrl(&empty[stage], phase ^ 1u);
        clk.lap(2);
        if (elect_one()) {
          ring_tag[stage] = live ? (int)w : -1;
          if (live) {
            unsigned char* dst = ring + (size_t)stage * tile_bytes;
            mbar_arrive_expect_tx(&full[stage], (uint32_t)tile_bytes);
            for (int kbi = 0; kbi < kb; ++kbi)
              tma_load_2d(dst + (size_t)kbi * KBLOCK_BYTES, &tm_x, kbi * KELEMS, (int)(tile * ROWS), &full[stage]);
          } else {
            mbar_arrive(&full[stage]);
          }
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ++it;
    };
    if (SAMPLE) {
      for (long long w = blockIdx.x; w < n_work; w += gridDim.x) push(w, true, true);
    } else {
      bool more = true;
      auto run_claim = [&](int& claim) {
        const int base = __shfl_sync(0xffffffffu, claim, 0);  // waits for this claim's register only
        if (base >= n_work) {
          more = false;
          return;
        }
        if (lane == 0) claim = atomicAdd(p.work_counter, TS_CLAIM);  // re-armed two claims ahead
        for (int j = 0; j < TS_CLAIM && base + j < n_work; ++j) push(base + j, true, true);
      };
      while (more) {
        run_claim(claim_a);
        if (more) run_claim(claim_b);
      }
    }
    push(0, false, true);
    if (!RAW && NBLK == 1) push(0, false, false);  // one end tag per ping-pong half in the row-term ring
    if (TS_INSTRUMENT && p.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
      p.dbg[0] = (unsigned long long)(clock64() - t_prod_begin);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, one elected lane issues) =====
    DbgClock clk{(p.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr, 0};
    const long long t_mma_begin = TS_INSTRUMENT ? clock64() : 0;
    ts_wait_ctrl(a_ready, 0);
    tc_fence_after();
    if (lane == 0) dbg_stamp(p.dbg, 3);
    const uint64_t desc0 = make_sdesc(smem_u32(ring));
    const uint32_t idesc = p.idesc;
    int stage = 0;
    uint32_t phase = 0;
    long long it = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (;; ++it) {
      clk.start();
      ts_wait_ctrl(&full[stage], phase);
      clk.lap(5);
      const int wtag = lds_s32(&ring_tag[stage]);
      if (wtag < 0) break;  // end of work
      const uint64_t bdesc = desc0 + (uint64_t)((uint32_t)stage * (uint32_t)(tile_bytes >> 4));
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {  // one accumulator unit per resident query block
        ts_wait_ctrl(&tmem_empty[acc], acc_phase ^ 1u);
        clk.lap(4);
        tc_fence_after();
        if (elect_one()) {
          if (RAW) {  // the work index travels with the accumulator buffer
            acc_work[acc] = wtag;
            __threadfence_block();
          }
          const uint32_t du = tmem_base + (uint32_t)(d_off + acc * ROWS);
          const uint32_t au = tmem_base + (uint32_t)(blk * a_cols);
          if (KB > 0) {
#pragma unroll
            for (int kbi = 0; kbi < (KB > 0 ? KB : 1); ++kbi) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kbi * 4 + k >= ksteps) continue;
                const uint64_t bd = bdesc + (uint64_t)(kbi * (KBLOCK_BYTES >> 4) + k * 2);
                const uint32_t aa = au + (uint32_t)(kbi * TC_KBLOCK + k * 8);
                if (BF16) umma_bf16_ts(du, aa, bd, idesc, (kbi | k) != 0 ? 1u : 0u);
                else umma_tf32_ts(du, aa, bd, idesc, (kbi | k) != 0 ? 1u : 0u);
              }
            }
          } else {
            uint64_t bd0 = bdesc;
            uint32_t a0 = au;
            for (int kbi = 0; kbi < kb; ++kbi) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kbi * 4 + k >= ksteps) continue;
                if (BF16) umma_bf16_ts(du, a0 + k * 8, bd0 + (uint64_t)(k * 2), idesc, (kbi | k) != 0 ? 1u : 0u);
                else umma_tf32_ts(du, a0 + k * 8, bd0 + (uint64_t)(k * 2), idesc, (kbi | k) != 0 ? 1u : 0u);
              }
              bd0 += (uint64_t)(KBLOCK_BYTES >> 4);
              a0 += TC_KBLOCK;
            }
          }
          if (blk == NBLK - 1) umma_commit(&empty[stage]);  // the stage is free once the last block has read it
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == n_acc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if (++stage == S) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (RAW) {
      // end of work for the two epilogue halves: the next two accumulator buffers carry the tag
      for (int e = 0; e < 2; ++e) {
        ts_wait_ctrl(&tmem_empty[acc], acc_phase ^ 1u);
        if (elect_one()) {
          acc_work[acc] = -1;
          __threadfence_block();
          mbar_arrive(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == n_acc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    if (lane == 0) dbg_stamp(p.dbg, 5);
    if (TS_INSTRUMENT && p.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
      p.dbg[3] = (unsigned long long)(clock64() - t_mma_begin);
    }
  } else if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TS_REGS_EPI));
    // ===== epilogue warps: one query per thread; two ping-pong halves of 8 warps alternate tiles =====
    // A tile's epilogue is a latency chain (barrier waits, tcgen05.ld, the dependent min tree), so
    // even tiles go to half 0 (accumulator buffer 0) and odd tiles to half 1 (buffer 1): the two
    // chains overlap. 16 warps = 4 per TMEM lane quarter (a warp may only touch lanes
    // 32*(warp%4)..+31); within a half:
    //   NBLK == 2: the two warps of a quarter take one query block each, all 4 chunks of the 64-row tile
    //   NBLK == 1: they split the 8 chunks of the 128-row tile (even / odd chunks)
    const int quarter = warp & 3;
    const int grp = (warp - 4) >> 2;                     // 0..3
    const int pp = grp & 1;                              // epilogue group: drains the units with u & 1 == pp
    const int sub = grp >> 1;                            // which half of the unit's chunks (even / odd)
    const int blk = NBLK == 2 ? pp : 0;                  // query block of this warp
    const int q = py * (NBLK * 128) + blk * 128 + quarter * 32 + lane;  // query index (within p.nq)
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    constexpr int NCH = 4;                               // chunks per warp and unit (8 chunks of 16 rows, two warps)
    auto chunk_of = [&](int ci) -> int { return sub + 2 * ci; };

    // the sample stage may directly follow the kernel that packs the queries; the main scan's
    // predecessor (threshold kernel) never touches them, so its query load overlaps that kernel
    if (SAMPLE) pdl_wait();
    // ---- resident query block -> tensor memory (cosine: pre-scaled by 1/|q|) ----
    // (tc_pack_kernel laid them out so that a warp reads 2 KB contiguous per 16 columns; the 16 warps
    //  split the 16-column groups of their query block between them)
    {
      const int nch = a_cols / 16;
      const int cpart = NBLK == 2 ? sub : grp, nparts = NBLK == 2 ? 2 : 4;
      const uint4* ap = reinterpret_cast<const uint4*>(p.apack + (size_t)py * (size_t)p.pass_stride) +
                        ((size_t)blk * nch * 128 + (size_t)(quarter * 32 + lane)) * 4;
      for (int c = cpart; c < nch; c += nparts) {
        float v[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const uint4 t = __ldg(ap + (size_t)c * 128 * 4 + j4);
          v[j4 * 4 + 0] = __uint_as_float(t.x);
          v[j4 * 4 + 1] = __uint_as_float(t.y);
          v[j4 * 4 + 2] = __uint_as_float(t.z);
          v[j4 * 4 + 3] = __uint_as_float(t.w);
        }
        tmem_st16(tmem_base + lane_base + (uint32_t)(blk * a_cols + c * 16), v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
    if (warp == 4 && lane == 0) dbg_stamp(p.dbg, 2);

    pdl_wait();  // thresholds (main scan) / the buffers the previous pass still reads (sample stage)
    float tau_me = -__int_as_float(0x7f800000);
    if (!SAMPLE && q < p.nq) tau_me = __ldcg(p.tau + q);
    // raw scan: a row is admitted when its accumulator value a' >= theta. L2: score = -2 a' exactly, so
    // theta = -tau / 2; dot / cosine: score = 1 - a' is rounded, so theta is lowered by a few ulps (a
    // few more rows are admitted; every admitted row carries its own score).
    float theta_me = __int_as_float(0x7f800000);
    if (RAW && !SAMPLE && q < p.nq)
      theta_me = MODE == MODE_L2 ? -0.5f * tau_me : (1.f - tau_me) - 4e-7f * (1.f + fabsf(tau_me));
    auto raw_score = [](float a) -> float { return MODE == MODE_L2 ? -2.f * a : 1.f - a; };
    const bool has_sc = MODE != MODE_L2 && p.sc != nullptr;
    const uint32_t bias_s = smem_u32(xs_bias), sc_s = smem_u32(xs_sc);
    // a lane's hit is parked in registers and published one tile later, so the round trip of the
    // global atomic that claims its slot never sits on the tile's critical path
    uint64_t pend_key = 0, prev_key = 0;
    int prev_pos = 0;
    bool has_pend = false, prev_has = false;
    DbgClock clk{(p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && warp < 8) ? p.dbg + 8 + (warp - 4) * 8 : nullptr, 0};
    const long long t_epi_begin = TS_INSTRUMENT ? clock64() : 0;

    int acc = pp;  // unit u uses buffer u % n_acc; this group drains every second unit
    uint32_t acc_phase = 0;
    const bool warp_idle = (q - lane) >= p.nq;  // warp-uniform: the warp's 32 queries are consecutive
    // one block: units are tiles and the group sees every second tile; two blocks: every tile, its block
    for (long long it = (NBLK == 2 ? 0 : pp);; it += (NBLK == 2 ? 1 : 2)) {
      const int xb = (int)(it % TS_XS);
      const uint32_t xphase = (uint32_t)((it / TS_XS) & 1);
      clk.start();
      long long w;
      if (RAW) {
        ts_wait_epi(&tmem_full[acc], acc_phase);
        w = lds_s32(&acc_work[acc]);
        if (w < 0) break;  // end of work
        clk.lap(2);
      } else {
        ts_wait_epi(&xs_full[xb], xphase);
        w = lds_s32(&xs_work[xb]);
        if (w < 0) break;  // end of work
        clk.lap(1);
        ts_wait_epi(&tmem_full[acc], acc_phase);
        clk.lap(2);
      }
      const long long tile = tile_of(w);
      if (it == 0 && warp == 4 && lane == 0) dbg_stamp(p.dbg, 4);
      tc_fence_after();
      if (NBLK == 1 && warp_idle) {  // (one-block kernels only: the branch costs the two-block kernel 3 %)
        // none of this warp's 32 queries exists (a batch smaller than the query block: a single query leaves
        // 15 of the 16 epilogue warps without one): hand the buffers straight back, read and score nothing
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tmem_empty[acc]);
          if (!RAW) mbar_arrive(&xs_empty[xb]);
        }
        acc += 2;
        if (acc >= n_acc) {
          acc -= n_acc;
          acc_phase ^= 1u;
        }
        continue;
      }
      const uint32_t d_addr = tmem_base + lane_base + (uint32_t)(d_off + acc * ROWS);
      // all of this warp's accumulator chunks are requested before the first one is consumed
      uint32_t araw[NCH][16];
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) tmem_ld16_issue(d_addr + (uint32_t)(chunk_of(ci) * 16), araw[ci]);
      tmem_ld_wait();
      clk.lap(3);
      // the accumulator now lives in registers: hand the buffer back before the min tree runs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc += 2;
      if (acc >= n_acc) {
        acc -= n_acc;
        acc_phase ^= 1u;
      }
      float tile_min;  // smallest score of the warp's rows of this tile (sample stage)
      if (RAW) {
        // the accumulator is the (negated, scaled) score: a max tree per chunk, one vote per tile
        float cmax[NCH];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const uint32_t(&v)[16] = araw[ci];
          float m = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
          for (int j = 2; j < 16; ++j) m = fmaxf(m, __uint_as_float(v[j]));
          cmax[ci] = m;
        }
        float tile_max = cmax[0];
#pragma unroll
        for (int ci = 1; ci < NCH; ++ci) tile_max = fmaxf(tile_max, cmax[ci]);
        tile_min = raw_score(tile_max);
        if (!SAMPLE) {
          const bool lane_hit = tile_max >= theta_me;
          if (__any_sync(0xffffffffu, lane_hit)) {
            if (lane_hit) {  // usually a single lane with a single admitted row
#pragma unroll
              for (int ci = 0; ci < NCH; ++ci) {
                if (cmax[ci] >= theta_me) {
                  int n_hit = 0, j_hit = 0;
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const bool h = __uint_as_float(araw[ci][j]) >= theta_me;
                    n_hit += h ? 1 : 0;
                    j_hit = h ? j : j_hit;
                  }
                  const long long row0 = tile * ROWS + chunk_of(ci) * 16;
                  if (n_hit == 1 && !has_pend) {
                    if (row0 + j_hit < p.n_rows) {  // rows past the end of the corpus read as zeros
                      pend_key = make_key(raw_score(cmax[ci]), (uint32_t)(row0 + j_hit));
                      has_pend = true;
                    }
                  } else {  // several admitted rows in one tile for this query (rare): publish directly
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float aj = __uint_as_float(araw[ci][j]);
                      if (aj >= theta_me && row0 + j < p.n_rows) {
                        const int pos = atomicAdd(p.cand_cnt + q, 1);
                        if (pos < TC_CAND_CAP)
                          p.cand[(size_t)q * TC_CAND_CAP + pos] = make_key(raw_score(aj), (uint32_t)(row0 + j));
                      }
                    }
                  }
                }
              }
            }
          }
        }
      } else {
        // scores of the warp's NCH chunks (in place) and their minima: independent chains, one vote per tile
        float cmin[NCH];
  #pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const uint32_t boff = (uint32_t)((xb * ROWS + chunk_of(ci) * 16) * 4);
          uint32_t(&v)[16] = araw[ci];
  #pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = lds128(bias_s + boff + j4 * 16);
            const float a0 = __uint_as_float(v[j4 * 4 + 0]), a1 = __uint_as_float(v[j4 * 4 + 1]);
            const float a2 = __uint_as_float(v[j4 * 4 + 2]), a3 = __uint_as_float(v[j4 * 4 + 3]);
            float r0, r1, r2, r3;
            if (MODE == MODE_L2) {
              r0 = fmaf(-2.f, a0, bb.x);
              r1 = fmaf(-2.f, a1, bb.y);
              r2 = fmaf(-2.f, a2, bb.z);
              r3 = fmaf(-2.f, a3, bb.w);
            } else {
              float4 ss = make_float4(1.f, 1.f, 1.f, 1.f);
              if (has_sc) ss = lds128(sc_s + boff + j4 * 16);
              r0 = fmaf(-a0, ss.x, bb.x);
              r1 = fmaf(-a1, ss.y, bb.y);
              r2 = fmaf(-a2, ss.z, bb.z);
              r3 = fmaf(-a3, ss.w, bb.w);
            }
            v[j4 * 4 + 0] = __float_as_uint(r0);
            v[j4 * 4 + 1] = __float_as_uint(r1);
            v[j4 * 4 + 2] = __float_as_uint(r2);
            v[j4 * 4 + 3] = __float_as_uint(r3);
            // three-input minima (FMNMX3): two values per instruction on the half-rate ALU pipe
            if (j4 == 0) cmin[ci] = fminf(fminf(r0, r1), r2);
            else cmin[ci] = fminf(fminf(cmin[ci], r0), fminf(r1, r2));
            cmin[ci] = fminf(cmin[ci], r3);
          }
        }
        tile_min = cmin[0];
  #pragma unroll
        for (int ci = 1; ci < NCH; ++ci) tile_min = fminf(tile_min, cmin[ci]);
        if (!SAMPLE) {
          const bool lane_hit = tile_min <= tau_me;
          if (__any_sync(0xffffffffu, lane_hit)) {
            if (lane_hit) {  // usually a single lane with a single admitted row
  #pragma unroll
              for (int ci = 0; ci < NCH; ++ci) {
                if (cmin[ci] <= tau_me) {
                  // branch-free count of admitted rows of the chunk and the index of the last one; a
                  // single admitted row is the chunk minimum itself
                  int n_hit = 0, j_hit = 0;
  #pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const bool h = __uint_as_float(araw[ci][j]) <= tau_me;
                    n_hit += h ? 1 : 0;
                    j_hit = h ? j : j_hit;
                  }
                  const uint32_t row0 = (uint32_t)(tile * ROWS + chunk_of(ci) * 16);
                  if (n_hit == 1 && !has_pend) {
                    pend_key = make_key(cmin[ci], row0 + (uint32_t)j_hit);
                    has_pend = true;
                  } else {  // several admitted rows in one tile for this query (rare): publish directly
  #pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float vj = __uint_as_float(araw[ci][j]);
                      if (vj <= tau_me) {
                        const int pos = atomicAdd(p.cand_cnt + q, 1);
                        if (pos < TC_CAND_CAP) p.cand[(size_t)q * TC_CAND_CAP + pos] = make_key(vj, row0 + (uint32_t)j);
                      }
                    }
                  }
                }
              }
            }
          }
        }
      }
      clk.lap(4);
      if (!SAMPLE) {
        if (prev_has && prev_pos < TC_CAND_CAP) p.cand[(size_t)q * TC_CAND_CAP + prev_pos] = prev_key;
        prev_has = has_pend;
        if (has_pend) {
          prev_pos = atomicAdd(p.cand_cnt + q, 1);
          prev_key = pend_key;
          has_pend = false;
        }
      }
      if (!RAW) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&xs_empty[xb]);
      }
      clk.lap(5);
      if (SAMPLE && q < p.nq) {
        // minimum score of this tile for this query: sample[q][w] (NBLK == 2) or sample[q][w][group]
        const uint32_t mn = tile_min == __int_as_float(0x7f800000) ? 0xFFFFFFFFu : f32_to_ordered(tile_min);
        p.sample[((size_t)q * p.n_sample + (size_t)w) * 2 + sub] = mn;
      }
    }
    if (!SAMPLE && prev_has && prev_pos < TC_CAND_CAP) p.cand[(size_t)q * TC_CAND_CAP + prev_pos] = prev_key;
    if (TS_INSTRUMENT && p.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
      if (warp < 8) p.dbg[8 + (warp - 4) * 8] = (unsigned long long)(clock64() - t_epi_begin);  // one warp per lane quarter reports
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 6);
}

// Threshold per query: the rank-th smallest sampled score (+inf when the sample holds
// fewer valid scores). One warp per query; bitwise binary search over the ordered-float images.
// Also clears the candidate counters of the pass.
// Pop the warp-wide minimum `rank` times from per-lane ascending lists best[0..R); returns the last
// popped value (0xFFFFFFFF when the lists run dry).
template <int R>
__device__ __forceinline__ uint32_t warp_pop_rank(uint32_t (&best)[R], int rank, uint32_t* popped /*[rank], lane 0*/) {
  const int lane = threadIdx.x & 31;
  uint32_t m = 0xFFFFFFFFu;
  for (int t = 0; t < rank; ++t) {
    m = __reduce_min_sync(0xffffffffu, best[0]);
    const unsigned who = __ballot_sync(0xffffffffu, best[0] == m);
    if (lane == __ffs(who) - 1) {
#pragma unroll
      for (int r = 0; r + 1 < R; ++r) best[r] = best[r + 1];
      best[R - 1] = 0xFFFFFFFFu;
    }
    if (popped != nullptr && lane == 0) popped[t] = m;
  }
  return m;
}

constexpr int TAU_THREADS = 128;

__global__ void __launch_bounds__(TAU_THREADS) tc_tau_kernel(const uint32_t* __restrict__ sample, int n_vals, int rank,
                                                             float* __restrict__ tau, int* __restrict__ cand_cnt,
                                                             int n_cnt, int* __restrict__ work_counter,
                                                             int n_work_counters) {
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  pdl_wait();  // the sample comes from the kernel just before; tau / cand_cnt are read by the one before that
  const uint32_t* vals = sample + (size_t)q * n_vals;
  constexpr int R = TC_SAMPLE_RANK_MAX;
  __shared__ uint32_t s_part[TAU_THREADS / 32][R];
  uint32_t best[R];  // this lane's R smallest values, ascending
#pragma unroll
  for (int r = 0; r < R; ++r) best[r] = 0xFFFFFFFFu;
  for (int i = tid; i < n_vals; i += TAU_THREADS) {
    uint32_t x = __ldcg(vals + i);
    if (x < best[R - 1]) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t lo = min(best[r], x);
        x = max(best[r], x);
        best[r] = lo;
      }
    }
  }
  if (lane < R) s_part[warp][lane] = 0xFFFFFFFFu;
  __syncwarp();
  warp_pop_rank<R>(best, rank, s_part[warp]);  // the `rank` smallest of this warp's share, ascending
  __syncthreads();
  if (warp == 0) {
    // 4 warps x 16 ascending values: every lane takes two consecutive ones
    static_assert((TAU_THREADS / 32) * R == 64, "final merge assumes 64 partial values");
    uint32_t two[2];
    two[0] = s_part[lane / 8][(lane % 8) * 2];
    two[1] = s_part[lane / 8][(lane % 8) * 2 + 1];
    const uint32_t m = warp_pop_rank<2>(two, rank, nullptr);
    if (lane == 0) {
      // 0xFFFFFFFF: fewer than `rank` valid samples -> admit everything
      tau[q] = (m == 0xFFFFFFFFu) ? __int_as_float(0x7f800000) : ordered_to_f32(m);
      if (q < n_cnt) cand_cnt[q] = 0;  // the first pass's counters (finalize_cand re-zeroes them for the next)
      if (q < n_work_counters && work_counter != nullptr) work_counter[q] = 0;
    }
  }
}

// Per-search row-term column for the TS variant when a mask (tombstones / filter) is active.
__global__ void __launch_bounds__(256) tc_bias_kernel(const uint32_t* __restrict__ mask, const float* __restrict__ norm2,
                                                      long long n_rows, long long n_pad, float* __restrict__ bias) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad;
       i += (long long)gridDim.x * blockDim.x) {
    bool ok = i < n_rows;
    if (ok && mask != nullptr) ok = (__ldg(mask + (i >> 5)) >> (i & 31)) & 1u;
    bias[i] = ok ? (norm2 ? __ldg(norm2 + i) : 1.0f) : __int_as_float(0x7f800000);
  }
}

int launch_tc_bias(const uint32_t* mask, const float* norm2, long long n_rows, long long n_pad, float* bias,
                   cudaStream_t st) {
  if (n_pad <= 0) return 0;
  long long blocks = (n_pad + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  tc_bias_kernel<<<(int)blocks, 256, 0, st>>>(mask, norm2, n_rows, n_pad, bias);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// Queries of a search in tensor-memory order, one block of n_cols queries per pass:
//   apack[pass][blk][c][r][j] = 32-bit column c*16+j of query pass*n_cols + blk*128 + r
// (fp32 value, or two consecutive bf16 values, low half first; cosine: pre-scaled by 1/|q|; zero
// beyond nq and beyond dp). One warp per query slot.
__global__ void __launch_bounds__(256) tc_pack_kernel(const float* __restrict__ queries, int nq, int dp, int d, int ones,
                                                      int n_cols, int a_cols, int bf16, int cosine,
                                                      uint32_t* __restrict__ apack, long long n_slots) {
  const int lane = threadIdx.x & 31;
  const long long slot = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= n_slots) return;
  const long long pass = slot / n_cols;
  const int qi = (int)(slot - pass * n_cols), blk = qi >> 7, r = qi & 127;
  const bool valid = slot < nq;
  const float* qv = queries + (size_t)slot * dp;
  float rnq = 1.f;
  if (cosine && valid) {
    float s = 0.f;
    for (int e = lane; e < dp; e += 32) s = fmaf(qv[e], qv[e], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    rnq = s > 0.f ? rsqrtf(s) : 0.f;
  }
  const int nch = a_cols >> 4;
  uint32_t* dst = apack + (size_t)pass * n_cols * a_cols;
  for (int col = lane; col < a_cols; col += 32) {
    uint32_t bits = 0;
    if (valid) {
      if (bf16) {
        // elements d .. d + ones - 1 face the norm columns of the bf16 rows (raw L2 scan)
        const int e = col * 2;
        const float a = e < d ? qv[e] * rnq : (e < d + ones ? 1.f : 0.f);
        const float b = e + 1 < d ? qv[e + 1] * rnq : (e + 1 < d + ones ? 1.f : 0.f);
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        bits = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        bits = col < dp ? __float_as_uint(qv[col] * rnq) : 0u;
      }
    }
    dst[((size_t)(blk * nch + (col >> 4)) * 128 + r) * 16 + (col & 15)] = bits;
  }
}

size_t tc_pack_bytes(const TcPlan& plan, int nq) {
  const long long passes = (nq + plan.n_cols - 1) / plan.n_cols;
  return (size_t)passes * plan.n_cols * plan.a_cols * 4;
}

int launch_tc_pack(const TcPlan& plan, const float* queries, int nq, int dp, int d, int ones, int cosine, void* apack,
                   cudaStream_t st) {
  if (plan.variant != 1 || nq <= 0) return 0;
  const long long passes = (nq + plan.n_cols - 1) / plan.n_cols;
  const long long n_slots = passes * plan.n_cols;
  tc_pack_kernel<<<(int)((n_slots + 7) / 8), 256, 0, st>>>(queries, nq, dp, d, plan.bf16 ? ones : 0, plan.n_cols,
                                                           plan.a_cols, plan.bf16, cosine, (uint32_t*)apack, n_slots);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// bf16 copy of the corpus for the tensor-core stream (round to nearest even): the d vector elements
// (cosine: scaled by 1/|x|), then for L2 the three bf16 pieces of -|x|^2 / 2 (their sum is exact),
// zero padded to dp16.
__global__ void __launch_bounds__(256) tc_to_bf16_kernel(const float* __restrict__ vec, long long row0, long long n,
                                                         int dp, int d, int dp16, const float* __restrict__ norm2,
                                                         const float* __restrict__ inv_norm,
                                                         __nv_bfloat16* __restrict__ out) {
  const long long total = n * (dp16 / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (dp16 / 2);
    const int c = (int)(i - r * (dp16 / 2)) * 2;
    const float* src = vec + (size_t)(row0 + r) * dp;
    const float sc = inv_norm ? __ldg(inv_norm + row0 + r) : 1.f;
    float v[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int e = c + t;
      float x = 0.f;
      if (e < d) {
        x = src[e] * sc;
      } else if (norm2 != nullptr && e < d + 3) {
        const float s0 = -0.5f * __ldg(norm2 + row0 + r);
        const float hi = __bfloat162float(__float2bfloat16_rn(s0));
        const float r1 = s0 - hi;
        const float mid = __bfloat162float(__float2bfloat16_rn(r1));
        x = e == d ? hi : (e == d + 1 ? mid : r1 - mid);
      }
      v[t] = x;
    }
    reinterpret_cast<__nv_bfloat162*>(out + (size_t)(row0 + r) * dp16)[c / 2] = __floats2bfloat162_rn(v[0], v[1]);
  }
}

int launch_tc_to_bf16(const float* vec, long long row0, long long n, int dp, int d, int dp16, const float* norm2,
                      const float* inv_norm, void* out, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n * (dp16 / 2) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tc_to_bf16_kernel<<<(int)blocks, 256, 0, st>>>(vec, row0, n, dp, d, dp16, norm2, inv_norm,
                                                 reinterpret_cast<__nv_bfloat16*>(out));
  QG_CUDA_OK(cudaGetLastError());
  return 0;
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

static int encode_map(CUtensorMap* map, const void* base, long long rows, int dp, int box_rows, bool bf16 = false) {
  if (tc_available() != 0) return fail(4, "cuTensorMapEncodeTiled is not available from this driver");
  const int esz = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)dp * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(4, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}


int tc_plan(int dp, int d16, int nq, bool bf16, TcPlan* out) {
  if (dp < 4 || dp % 4 != 0) return -1;
  int kb = (dp + TC_KBLOCK - 1) / TC_KBLOCK;
  const int budget = 227 * 1024 - 1024;  // alignment slack
  out->bf16 = 0;
  if (bf16) {
    // bf16 stream: 16 elements per MMA k-step, 64 per 128-byte k-block; d16 counts the vector and the
    // norm columns. 512 TMEM columns = nblk * a_cols (queries) + at least two accumulator buffers.
    const int ksteps = (d16 + 15) / 16;
    const int kb16 = (ksteps + 3) / 4;
    const int a_cols = (ksteps * 8 + 15) & ~15;
    if (a_cols <= 256) {
      const int nblk = (a_cols <= 128 && nq > 128) ? 2 : 1;
      const int rows = ts_rows(nblk);
      int stages = (budget - ts_smem_layout(0, kb16, rows).total) / (kb16 * rows * TS_KSTEP_BYTES + 16);
      stages = std::min(stages, 12);
      if (stages >= 2) {
        out->variant = 1;
        out->bf16 = 1;
        out->nblk = nblk;
        out->n_cols = 128 * nblk;
        out->kb = kb16;
        out->ksteps = ksteps;
        out->a_cols = a_cols;
        out->stages = stages;
        out->tile_rows = rows;
        out->sample_vals = 2;
        out->smem = ts_smem_layout(stages, kb16, rows).total + 1024;
        return 0;
      }
    }
  }
  // TS variant: queries resident in tensor memory. 512 columns = nblk * kb * 32 (queries) +
  // nblk * 2 * 64 (double-buffered accumulators).
  if (kb <= 8) {
    const int nblk = (kb <= 4 && nq > 128) ? 2 : 1;
    const int rows = ts_rows(nblk);
    int stages = (budget - ts_smem_layout(0, kb, rows).total) / (kb * rows * TS_KSTEP_BYTES + 16);
    stages = std::min(stages, 12);
    if (stages >= 2) {
      out->variant = 1;
      out->nblk = nblk;
      out->n_cols = 128 * nblk;
      out->kb = kb;
      out->ksteps = kb * 4;
      out->a_cols = kb * TC_KBLOCK;
      out->stages = stages;
      out->tile_rows = rows;
      out->sample_vals = 2;
      out->smem = ts_smem_layout(stages, kb, rows).total + 1024;
      return 0;
    }
  }
  // SS variant: queries in shared memory next to the corpus ring; fewer query columns leave more
  // stages (bytes in flight), so pick the column count that maximises columns * min(1, stages / 11).
  int best_cols = 0, best_stages = 0;
  double best = 0.0;
  for (int n_cols = 16; n_cols <= std::min(TC_MAX_COLS, ((nq + 15) / 16) * 16); n_cols += 16) {
    const int left = budget - tc_smem_layout(n_cols, kb, 0).total;
    int stages = left / (TC_STAGE_BYTES + 16);
    if (stages < 4) break;
    stages = std::min(stages, 12);
    const double score = n_cols * std::min(1.0, stages / 11.0);
    if (score > best) {
      best = score;
      best_cols = n_cols;
      best_stages = stages;
    }
  }
  if (best_cols == 0) return -1;
  out->variant = 0;
  out->nblk = 0;
  out->n_cols = best_cols;
  out->kb = kb;
  out->stages = best_stages;
  out->tile_rows = TC_TILE_ROWS;
  out->sample_vals = 2;
  out->smem = tc_smem_layout(best_cols, kb, best_stages).total + 1024;
  return 0;
}

template <int MODE, bool SAMPLE>
static int set_attr_one() {
  QG_CUDA_OK(cudaFuncSetAttribute(tc_scan_kernel<MODE, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  return 0;
}

typedef void (*TsKernelFn)(const CUtensorMap, const TsKParams);

template <int MODE, bool SAMPLE, int NBLK, bool BF16, bool RAW>
static TsKernelFn ts_kernel_kb(int kb) {
  switch (kb) {
    case 1: return tc_ts_kernel<MODE, SAMPLE, NBLK, 1, BF16, RAW>;
    case 2: return tc_ts_kernel<MODE, SAMPLE, NBLK, 2, BF16, RAW>;
    case 3: return tc_ts_kernel<MODE, SAMPLE, NBLK, 3, BF16, RAW>;
    case 4: return tc_ts_kernel<MODE, SAMPLE, NBLK, 4, BF16, RAW>;
    default: return tc_ts_kernel<MODE, SAMPLE, NBLK, 0, BF16, RAW>;
  }
}

template <int MODE, bool BF16, bool RAW>
static TsKernelFn ts_kernel_ms(bool sample, int nblk, int kb) {
  if (sample) return nblk == 2 ? ts_kernel_kb<MODE, true, 2, BF16, RAW>(kb) : ts_kernel_kb<MODE, true, 1, BF16, RAW>(kb);
  return nblk == 2 ? ts_kernel_kb<MODE, false, 2, BF16, RAW>(kb) : ts_kernel_kb<MODE, false, 1, BF16, RAW>(kb);
}

// raw kernels exist for the bf16 stream only
static TsKernelFn ts_kernel(int mode, bool sample, int nblk, int kb, bool bf16, bool raw) {
  if (mode == MODE_L2) {
    if (!bf16) return ts_kernel_ms<MODE_L2, false, false>(sample, nblk, kb);
    return raw ? ts_kernel_ms<MODE_L2, true, true>(sample, nblk, kb) : ts_kernel_ms<MODE_L2, true, false>(sample, nblk, kb);
  }
  if (!bf16) return ts_kernel_ms<MODE_DOT, false, false>(sample, nblk, kb);
  return raw ? ts_kernel_ms<MODE_DOT, true, true>(sample, nblk, kb) : ts_kernel_ms<MODE_DOT, true, false>(sample, nblk, kb);
}

static int set_attr_ts() {
  for (int mode : {MODE_L2, MODE_DOT})
    for (int sample = 0; sample < 2; ++sample)
      for (int nblk = 1; nblk <= 2; ++nblk)
        for (int kb = 0; kb <= 4; ++kb)
          for (int bf = 0; bf < 3; ++bf)  // tf32, bf16 with row terms, bf16 raw
            QG_CUDA_OK(cudaFuncSetAttribute(ts_kernel(mode, sample != 0, nblk, kb == 0 ? 9 : kb, bf != 0, bf == 2),
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return 0;
}

int tc_set_attributes() {
  QG_CUDA_OK(cudaFuncSetAttribute(tc_tau_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
  QG_CUDA_OK(cudaFuncSetAttribute(tc_bias_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
  if (int e = set_attr_ts()) return e;
  if (int e = set_attr_one<MODE_L2, false>()) return e;
  if (int e = set_attr_one<MODE_L2, true>()) return e;
  if (int e = set_attr_one<MODE_DOT, false>()) return e;
  if (int e = set_attr_one<MODE_DOT, true>()) return e;
  return 0;
}

static int launch_ts_pass(const TcPlan& plan, const TcArgs& a, int sm_count, cudaStream_t st, int* launches,
                          const TcStageHook* hook) {
  const int rows = ts_rows(plan.nblk);
  const bool bf16 = plan.bf16 != 0;
  const bool raw = bf16 && a.raw != 0;
  const int sample_rank = std::min(TC_SAMPLE_RANK_MAX, std::max(1, a.sample_rank));
  CUtensorMap tm_x;
  if (bf16) {
    if (a.vec16 == nullptr) return fail(1, "tensor-core bf16 pass needs the bf16 copy of the corpus");
    if (int rc = encode_map(&tm_x, a.vec16, a.n_rows, a.dp16, rows, true)) return rc;
  } else {
    if (int rc = encode_map(&tm_x, a.vec, a.n_rows, a.dp, rows)) return rc;
  }
  TsKParams p{};
  p.n_rows = a.n_rows;
  p.n_tiles = (a.n_rows + rows - 1) / rows;
  p.n_sample_from = raw ? a.n_rows / rows : p.n_tiles;
  p.kb = plan.kb;
  p.ksteps = plan.ksteps;
  p.a_cols = plan.a_cols;
  p.stages = plan.stages;
  p.nq = a.nq;
  p.nblk = plan.nblk;
  if (raw && p.n_sample_from <= 0) return fail(1, "raw tensor-core scan needs at least one full tile");
  const uint32_t fmt = bf16 ? 1u : 2u;  // F16F32Format: BF16 = 1, TF32 = 2
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(rows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.cosine = a.cosine;
  p.bias = a.bias;
  p.sc = (a.cosine && !bf16) ? a.inv_norm : nullptr;  // the bf16 rows of a cosine index are stored normalised
  p.apack = static_cast<const uint32_t*>(a.apack);
  p.work_counter = a.work_counter;
  p.dp = a.dp;
  p.sample = a.sample;
  p.n_sample = a.n_sample;
  p.tau = a.tau;
  p.cand = a.cand;
  p.cand_cnt = a.cand_cnt;
  p.dbg = nullptr;
  if (a.apack == nullptr || a.work_counter == nullptr) return fail(1, "tensor-core TS pass needs packed queries and a work counter");
  if (!raw && a.bias == nullptr) return fail(1, "tensor-core TS pass needs the per-row bias column");
  const int grid_m = (int)std::min<long long>(sm_count, p.n_tiles);
  if (a.sample_only || !a.presampled) {
    // sample + threshold: for this pass, or (sample_only) for all passes of a search in one launch each
    // — the sample depends on nothing but the queries, and hoisting it out of the pass loop pays its
    // launch, TMEM allocation and pipeline ramp once per search instead of once per 256 queries
    const int n_pass = a.sample_only ? (a.nq + plan.n_cols - 1) / plan.n_cols : 1;
    int gx = (int)std::min<long long>(sm_count, a.n_sample);
    if (n_pass > 1) gx = (int)std::max<long long>(1, std::min<long long>(a.n_sample, sm_count / n_pass));
    p.pass_stride = (long long)plan.n_cols * plan.a_cols;
    if (hook) hook->fn(hook->ctx, 0, 1, st);
    QG_CUDA_OK(launch_chained(ts_kernel(a.mode, true, plan.nblk, plan.kb, bf16, raw), dim3(gx, n_pass), dim3(TS_THREADS),
                              (size_t)plan.smem, st, tm_x, p));
    QG_CUDA_OK(launch_chained(tc_tau_kernel, dim3(a.nq), dim3(TAU_THREADS), (size_t)0, st, (const uint32_t*)a.sample,
                              a.n_sample * plan.sample_vals, sample_rank, a.tau, a.cand_cnt,
                              a.sample_only ? a.n_cnt : plan.n_cols, a.work_counter,
                              a.sample_only ? std::max(1, a.n_cnt / plan.n_cols) : 1));
    if (std::getenv("QG_TC_NOHIT"))  // development aid: a scan that admits nothing
      launch_fill_f32(a.tau, a.nq, -__builtin_huge_valf(), st);
    if (hook) hook->fn(hook->ctx, 0, 0, st);
    if (launches) *launches += 2;
    if (a.sample_only) return 0;
  }
  p.dbg = a.dbg;
  if (hook) hook->fn(hook->ctx, 1, 1, st);
  QG_CUDA_OK(launch_chained(ts_kernel(a.mode, false, plan.nblk, plan.kb, bf16, raw), dim3(grid_m), dim3(TS_THREADS),
                            (size_t)plan.smem, st, tm_x, p));
  if (hook) hook->fn(hook->ctx, 1, 0, st);
  if (launches) *launches += 1;
  return 0;
}

int launch_tc_pass(const TcPlan& plan, const TcArgs& a, int sm_count, cudaStream_t st, int* launches,
                   const TcStageHook* hook) {
  if (plan.variant == 1) return launch_ts_pass(plan, a, sm_count, st, launches, hook);
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
  if (hook) hook->fn(hook->ctx, 0, 1, st);
  if (a.mode == MODE_L2) {
    tc_scan_kernel<MODE_L2, true><<<grid_s, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  } else {
    tc_scan_kernel<MODE_DOT, true><<<grid_s, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  }
  QG_CUDA_OK(cudaGetLastError());
  tc_tau_kernel<<<a.nq, TAU_THREADS, 0, st>>>(a.sample, a.n_sample * 2,
                                              std::min(TC_SAMPLE_RANK_MAX, std::max(1, a.sample_rank)), a.tau, a.cand_cnt,
                                              a.nq, nullptr, 0);
  QG_CUDA_OK(cudaGetLastError());
  if (hook) hook->fn(hook->ctx, 0, 0, st);
  if (hook) hook->fn(hook->ctx, 1, 1, st);
  if (a.mode == MODE_L2) {
    tc_scan_kernel<MODE_L2, false><<<grid_m, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  } else {
    tc_scan_kernel<MODE_DOT, false><<<grid_m, TC_THREADS, plan.smem, st>>>(tm_a, tm_b, p);
  }
  QG_CUDA_OK(cudaGetLastError());
  if (hook) hook->fn(hook->ctx, 1, 0, st);
  if (launches) *launches += 3;
  return 0;
}

}  // namespace qg
