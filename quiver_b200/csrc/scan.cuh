// scan.cuh — flat-scan kernels: distance scan fused with CTA-level top-k' selection.
//
// Replaces the hot loop of hybrid.ExactIndex.Search (reference pkg/hybrid/exact.go:114-129:
// one distFunc call per stored vector, then a full sort) for small query batches.
//
// Data movement: the corpus is row-major fp32 [rows x D] in HBM. Every warp owns a private
// ring of STAGES shared-memory tiles; lane 0 fills them with 1-D bulk async copies
// (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), so a tile of RT consecutive
// rows is one contiguous RT*D*4-byte transfer and many KB per SM stay in flight without
// holding registers. In gather mode (prefiltered scans, HNSW-style row lists) the tile is
// assembled from one bulk copy per listed row instead.
//
// Arithmetic: each row is split over G lanes (G*16 contiguous bytes per quarter-warp =>
// conflict-free LDS.128); queries live in registers; a transposing butterfly reduces RB
// rows at once (about one SHFL per row per query). The fp32 score is only used to SELECT
// candidates; the distances returned to the caller are recomputed in the reference's own
// arithmetic by the finalize kernel (finalize.cu).
#pragma once
#include "common.cuh"
#include "select.cuh"

namespace qg {

enum { MODE_L2 = 0, MODE_DOT = 1, MODE_L1 = 2 };

struct ScanParams {
  const float* vec;        // [rows x dp]
  const float* inv_norm;   // [rows] 1/|x| (0 for zero rows) when cosine, else nullptr
  const uint32_t* mask;    // row-pass bitmask (live & filter), nullptr = every row passes
  const uint32_t* gather;  // row list in gather mode, else nullptr
  long long n_items;       // rows (dense) or list length (gather)
  const float* queries;    // [nq x dp], device
  int nq;                  // queries served by this pass (<= QB)
  int kp;                  // candidates kept per query per CTA (power of two, 32..1024)
  int cosine;              // MODE_DOT only: 1 = cosine score, 0 = 1 - dot
  uint64_t* partial;       // [nq][gridDim.x][kp] sorted keys (KEY_NONE padded)
  // generic kernel only
  int dp;                  // padded dimension (multiple of 4)
  int tile_rows;           // rows per tile (<= 32)
  int stages;              // ring depth
};

constexpr int SCAN_NW = 8;  // warps per CTA of the fast kernels
#ifndef QG_SCAN_TILE_TARGET
#define QG_SCAN_TILE_TARGET 8192  // bytes per bulk copy (tile) the geometry aims for
#endif

template <int D>
struct ScanGeom {
  static_assert(D % 32 == 0, "fast kernels need D % 32 == 0");
  static constexpr int C = D / 4;  // float4 chunks per row
  static constexpr int G = (C % 32 == 0) ? 32 : ((C % 16 == 0) ? 16 : 8);  // lanes per row
  static constexpr int CPL = C / G;                                         // chunks per lane
  static constexpr int RPP = 32 / G;                                        // rows per pass
  static constexpr int ROW_BYTES = D * 4;
  // passes (= accumulators per lane per query) chosen so that a tile is about 4 KB
  static constexpr int RB_RAW = 4096 / (ROW_BYTES * RPP);
  static constexpr int RB = RB_RAW >= 8 ? 8 : (RB_RAW >= 4 ? 4 : (RB_RAW >= 2 ? 2 : 1));
  static constexpr int RS = RPP * RB;  // rows per sub-batch (one butterfly + one filter step)
  // sub-batches per tile; a tile never exceeds 32 rows (one gather copy per lane, and the pool
  // head-room of select.cuh assumes at most 32 appends per warp per query between flag checks)
  static constexpr int NSUB_CAP = 32 / RS;
  static constexpr int NSUB_RAW0 = QG_SCAN_TILE_TARGET / (RS * ROW_BYTES);
  static constexpr int NSUB_RAW = NSUB_RAW0 < NSUB_CAP ? NSUB_RAW0 : NSUB_CAP;
  static constexpr int NSUB = NSUB_RAW >= 4 ? 4 : (NSUB_RAW >= 2 ? 2 : 1);
  static constexpr int RT = RS * NSUB;  // rows per tile = one bulk copy
  static constexpr int TILE_BYTES = RT * ROW_BYTES;
  static constexpr int STAGES = TILE_BYTES <= 4096 ? 4 : (TILE_BYTES <= 8192 ? 3 : 2);
  static constexpr int REP = G / RB;  // lanes holding the same reduced row
  static_assert(RB <= G, "butterfly needs RB <= G");
  static_assert(RT <= 32, "a tile holds at most 32 rows");
};

// Reduce RB per-row accumulators over the G lanes of a row group. On return a[0] holds the
// total of pass p = (lane % G) / (G / RB); that value is replicated on G/RB lanes.
template <int RB, int G>
__device__ __forceinline__ float butterfly_reduce(float (&a)[RB], int lane) {
  int nv = RB;
#pragma unroll
  for (int o = G / 2; o >= 1; o >>= 1) {
    if (nv > 1) {
      const int half = nv / 2;
      const bool hi = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < RB / 2; ++i) {
        if (i < half) {
          float send = hi ? a[i] : a[i + half];
          float keep = hi ? a[i + half] : a[i];
          a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      nv = half;
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
    }
  }
  return a[0];
}

template <int MODE>
__device__ __forceinline__ float accum4(float acc, const float4& x, const float4& q) {
  if (MODE == MODE_L2) {
    float d0 = x.x - q.x, d1 = x.y - q.y, d2 = x.z - q.z, d3 = x.w - q.w;
    acc = fmaf(d0, d0, acc);
    acc = fmaf(d1, d1, acc);
    acc = fmaf(d2, d2, acc);
    acc = fmaf(d3, d3, acc);
  } else if (MODE == MODE_DOT) {
    acc = fmaf(x.x, q.x, acc);
    acc = fmaf(x.y, q.y, acc);
    acc = fmaf(x.z, q.z, acc);
    acc = fmaf(x.w, q.w, acc);
  } else {
    acc += fabsf(x.x - q.x);
    acc += fabsf(x.y - q.y);
    acc += fabsf(x.z - q.z);
    acc += fabsf(x.w - q.w);
  }
  return acc;
}

// Shared-memory control block common to both scan kernels.
struct ScanCtl {
  int prune_flag;
  int done_warps;
};

// Everything after the distance arithmetic: threshold filter, pool append, cooperative
// prune protocol. `pools` are the per-query pools in shared memory.
template <int QB>
struct PoolSet {
  uint64_t* keys;  // [QB][slots]
  int* cnt;        // [QB]
  float* tau;      // [QB]
  int slots;
  __device__ __forceinline__ PoolRef ref(int qi) const {
    return PoolRef{keys + (size_t)qi * slots, cnt + qi, tau + qi};
  }
};

// Called by every thread of the CTA once all of them have seen the prune flag.
template <int QB>
__device__ __forceinline__ void cta_prune_all(const PoolSet<QB>& ps, ScanCtl* ctl, int nq, int kp) {
  __syncthreads();
  for (int qi = 0; qi < nq; ++qi) {
    if (ps.cnt[qi] > kp) block_prune(ps.ref(qi), kp);
  }
  if (threadIdx.x == 0) st_volatile_s32(&ctl->prune_flag, 0);
  __syncthreads();
}

// End-of-scan protocol: wait until every warp has finished appending (taking part in any
// prune still requested), then sort every pool and write the CTA's partial lists.
template <int QB>
__device__ __forceinline__ void cta_finish(const PoolSet<QB>& ps, ScanCtl* ctl, int nq, int kp, int nwarps,
                                           uint64_t* partial) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(&ctl->done_warps, 1);
  }
  for (;;) {
    int d = ld_volatile_s32(&ctl->done_warps);  // read BEFORE the flag
    int f = ld_volatile_s32(&ctl->prune_flag);
    d = __shfl_sync(0xffffffffu, d, 0);
    f = __shfl_sync(0xffffffffu, f, 0);
    if (f) {
      cta_prune_all<QB>(ps, ctl, nq, kp);
    } else if (d == nwarps) {
      break;
    } else {
      __nanosleep(64);
    }
  }
  __syncthreads();
  for (int qi = 0; qi < nq; ++qi) {
    block_prune(ps.ref(qi), kp);
    const int n = ps.cnt[qi];
    uint64_t* out = partial + ((size_t)qi * gridDim.x + blockIdx.x) * kp;
    const uint64_t* src = ps.keys + (size_t)qi * ps.slots;
    for (int i = threadIdx.x; i < kp; i += blockDim.x) out[i] = i < n ? src[i] : KEY_NONE;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Fast kernel: compile-time D (multiple of 32), queries in registers.
// Shared memory: ring [NW][STAGES][TILE_BYTES] | pool keys [QB][slots] | mbar [NW][STAGES]
//                | cnt[QB] | tau[QB] | ScanCtl
// ------------------------------------------------------------------------------------------
template <int D, int QB>
__host__ __device__ constexpr size_t scan_fast_smem(int kp) {
  using Gm = ScanGeom<D>;
  return (size_t)SCAN_NW * Gm::STAGES * Gm::TILE_BYTES + (size_t)QB * pool_slots(kp) * 8 +
         (size_t)SCAN_NW * Gm::STAGES * 8 + QB * 8 + sizeof(ScanCtl) + 64;
}

template <int D, int QB, int MODE>
__global__ void __launch_bounds__(SCAN_NW * 32, 1) scan_fast_kernel(const ScanParams p) {
  using Gm = ScanGeom<D>;
  constexpr int C = Gm::C, G = Gm::G, CPL = Gm::CPL, RPP = Gm::RPP, RB = Gm::RB, RT = Gm::RT;
  constexpr int S = Gm::STAGES, TILE_BYTES = Gm::TILE_BYTES, RS = Gm::RS, NSUB = Gm::NSUB;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slots = pool_slots(p.kp);
  const int highwater = pool_highwater(p.kp, SCAN_NW);

  unsigned char* ring = smem_raw + (size_t)warp * S * TILE_BYTES;
  uint64_t* pool_keys = reinterpret_cast<uint64_t*>(smem_raw + (size_t)SCAN_NW * S * TILE_BYTES);
  uint64_t* bars = pool_keys + (size_t)QB * slots;
  int* cnt = reinterpret_cast<int*>(bars + SCAN_NW * S);
  float* tau = reinterpret_cast<float*>(cnt + QB);
  ScanCtl* ctl = reinterpret_cast<ScanCtl*>(tau + QB);
  uint64_t* mybar = bars + warp * S;
  PoolSet<QB> ps{pool_keys, cnt, tau, slots};

  if (threadIdx.x == 0) {
    for (int i = 0; i < SCAN_NW * S; ++i) mbar_init(&bars[i], 1);
    for (int qi = 0; qi < QB; ++qi) {
      cnt[qi] = 0;
      tau[qi] = __int_as_float(0x7f800000);
    }
    ctl->prune_flag = 0;
    ctl->done_warps = 0;
    fence_mbar_init();
  }
  __syncthreads();

  const int grp = lane / G, gl = lane % G;

  // ---- queries into registers ------------------------------------------------------------
  float4 qreg[QB][CPL];
  float rnq[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qi < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)qi * D) + c * G + gl);
      qreg[qi][c] = v;
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s);
      s = fmaf(v.w, v.w, s);
    }
    rnq[qi] = 1.f;
    if (MODE == MODE_DOT) {
#pragma unroll
      for (int o = G / 2; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      rnq[qi] = p.cosine ? (s > 0.f ? rsqrtf(s) : 0.f) : 1.f;
    }
  }

  // ---- tile schedule ---------------------------------------------------------------------
  const long long n_items = p.n_items;
  const long long n_tiles = (n_items + RT - 1) / RT;
  const long long gws = (long long)gridDim.x * SCAN_NW;
  const long long gw = (long long)blockIdx.x * SCAN_NW + warp;
  const bool gather = p.gather != nullptr;

  auto issue = [&](long long it) {
    const long long t = it * gws + gw;
    if (t >= n_tiles) return;
    const int stage = (int)(it % S);
    const long long item0 = t * RT;
    const int rows = (int)((n_items - item0) < RT ? (n_items - item0) : RT);
    unsigned char* dst = ring + (size_t)stage * TILE_BYTES;
    if (!gather) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&mybar[stage], (uint32_t)rows * Gm::ROW_BYTES);
        bulk_g2s(dst, p.vec + (size_t)item0 * D, (uint32_t)rows * Gm::ROW_BYTES, &mybar[stage]);
      }
    } else {
      if (lane == 0) mbar_arrive_expect_tx(&mybar[stage], (uint32_t)rows * Gm::ROW_BYTES);
      __syncwarp();
      if (lane < rows) {
        const uint32_t r = __ldg(p.gather + item0 + lane);
        bulk_g2s(dst + (size_t)lane * Gm::ROW_BYTES, p.vec + (size_t)r * D, Gm::ROW_BYTES, &mybar[stage]);
      }
    }
  };

#pragma unroll 1
  for (int s = 0; s < S - 1; ++s) issue(s);

  // row this lane owns after the butterfly
  const int my_pass = gl / Gm::REP;
  const int my_row_in_tile = my_pass * RPP + grp;
  const bool my_unique = (gl % Gm::REP) == 0;

#pragma unroll 1
  for (long long it = 0;; ++it) {
    const long long t = it * gws + gw;
    if (t >= n_tiles) break;
    // checkpoint: take part in a prune if one was requested
    if (__shfl_sync(0xffffffffu, ld_volatile_s32(&ctl->prune_flag), 0)) cta_prune_all<QB>(ps, ctl, p.nq, p.kp);

    issue(it + S - 1);

    const int stage = (int)(it % S);
    const uint32_t parity = (uint32_t)((it / S) & 1);
    const long long item0 = t * RT;
    bool valid[NSUB];
    uint32_t my_row[NSUB];
    float rinv[NSUB];
#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb) {
      const long long my_item = item0 + sb * RS + my_row_in_tile;
      valid[sb] = my_unique && (my_item < n_items);
      my_row[sb] = 0;
      rinv[sb] = 1.f;
      if (valid[sb]) {
        my_row[sb] = gather ? __ldg(p.gather + my_item) : (uint32_t)my_item;
        if (p.mask != nullptr && !gather) valid[sb] = (__ldg(p.mask + (my_row[sb] >> 5)) >> (my_row[sb] & 31)) & 1u;
        if (MODE == MODE_DOT && p.inv_norm != nullptr) rinv[sb] = __ldg(p.inv_norm + my_row[sb]);
      }
    }

    float tau_r[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) tau_r[qi] = *reinterpret_cast<volatile float*>(&tau[qi]);

    mbar_wait(&mybar[stage], parity);
    const float4* tile = reinterpret_cast<const float4*>(ring + (size_t)stage * TILE_BYTES);

#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb) {
      const float4* tb = tile + sb * RS * C;
      float acc[QB][RB];
#pragma unroll
      for (int qi = 0; qi < QB; ++qi)
#pragma unroll
        for (int pr = 0; pr < RB; ++pr) acc[qi][pr] = 0.f;

#pragma unroll
      for (int pr = 0; pr < RB; ++pr) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const float4 x = tb[(pr * RPP + grp) * C + c * G + gl];
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) acc[qi][pr] = accum4<MODE>(acc[qi][pr], x, qreg[qi][c]);
        }
      }

#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        float v = butterfly_reduce<RB, G>(acc[qi], lane);
        if (qi < p.nq) {  // warp-uniform
          float score = v;
          if (MODE == MODE_DOT) score = 1.0f - v * (rinv[sb] * rnq[qi]);
          const bool pass = valid[sb] && (score <= tau_r[qi]);
          const int after = warp_append(ps.ref(qi), pass, make_key(score, my_row[sb]));
          if (after >= highwater && lane == 0) st_volatile_s32(&ctl->prune_flag, 1);
        }
      }
    }
    __syncwarp();
    fence_proxy_async();
  }

  cta_finish<QB>(ps, ctl, p.nq, p.kp, SCAN_NW, p.partial);
}

// ------------------------------------------------------------------------------------------
// Generic kernel: any padded dimension dp (multiple of 4), queries in shared memory, one
// row per warp pass (lanes stride over the row's float4 chunks). Handles MODE_L1 too.
// Shared memory: ring [NW][stages][tile_rows*dp*4] | queries [QB][dp] | pool keys | mbar
//                | cnt | tau | ScanCtl
// ------------------------------------------------------------------------------------------
template <int QB>
__host__ __device__ inline size_t scan_generic_smem(int nw, int stages, int tile_rows, int dp, int kp) {
  return (size_t)nw * stages * tile_rows * dp * 4 + (size_t)QB * dp * 4 + (size_t)QB * pool_slots(kp) * 8 +
         (size_t)nw * stages * 8 + QB * 8 + sizeof(ScanCtl) + 64;
}

template <int QB, int MODE>
__global__ void __launch_bounds__(SCAN_NW * 32, 1) scan_generic_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dp = p.dp, C = dp >> 2, RT = p.tile_rows, S = p.stages;
  const int row_bytes = dp * 4, tile_bytes = RT * row_bytes;
  const int slots = pool_slots(p.kp);
  const int highwater = pool_highwater(p.kp, nw);

  unsigned char* ring = smem_raw + (size_t)warp * S * tile_bytes;
  float* qs = reinterpret_cast<float*>(smem_raw + (size_t)nw * S * tile_bytes);
  uint64_t* pool_keys = reinterpret_cast<uint64_t*>(qs + (size_t)QB * dp);
  uint64_t* bars = pool_keys + (size_t)QB * slots;
  int* cnt = reinterpret_cast<int*>(bars + nw * S);
  float* tau = reinterpret_cast<float*>(cnt + QB);
  ScanCtl* ctl = reinterpret_cast<ScanCtl*>(tau + QB);
  uint64_t* mybar = bars + warp * S;
  PoolSet<QB> ps{pool_keys, cnt, tau, slots};

  if (threadIdx.x == 0) {
    for (int i = 0; i < nw * S; ++i) mbar_init(&bars[i], 1);
    for (int qi = 0; qi < QB; ++qi) {
      cnt[qi] = 0;
      tau[qi] = __int_as_float(0x7f800000);
    }
    ctl->prune_flag = 0;
    ctl->done_warps = 0;
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < QB * dp; i += blockDim.x) {
    const int qi = i / dp;
    qs[i] = qi < p.nq ? __ldg(p.queries + i) : 0.f;
  }
  __syncthreads();

  float rnq[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    rnq[qi] = 1.f;
    if (MODE == MODE_DOT && p.cosine) {
      float s = 0.f;
      for (int i = lane; i < dp; i += 32) s = fmaf(qs[qi * dp + i], qs[qi * dp + i], s);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      rnq[qi] = s > 0.f ? rsqrtf(s) : 0.f;
    }
  }

  const long long n_items = p.n_items;
  const long long n_tiles = (n_items + RT - 1) / RT;
  const long long gws = (long long)gridDim.x * nw;
  const long long gw = (long long)blockIdx.x * nw + warp;
  const bool gather = p.gather != nullptr;

  auto issue = [&](long long it) {
    const long long t = it * gws + gw;
    if (t >= n_tiles) return;
    const int stage = (int)(it % S);
    const long long item0 = t * RT;
    const int rows = (int)((n_items - item0) < RT ? (n_items - item0) : RT);
    unsigned char* dst = ring + (size_t)stage * tile_bytes;
    if (!gather) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&mybar[stage], (uint32_t)rows * row_bytes);
        bulk_g2s(dst, p.vec + (size_t)item0 * dp, (uint32_t)rows * row_bytes, &mybar[stage]);
      }
    } else {
      if (lane == 0) mbar_arrive_expect_tx(&mybar[stage], (uint32_t)rows * row_bytes);
      __syncwarp();
      if (lane < rows) {
        const uint32_t r = __ldg(p.gather + item0 + lane);
        bulk_g2s(dst + (size_t)lane * row_bytes, p.vec + (size_t)r * dp, row_bytes, &mybar[stage]);
      }
    }
  };

  for (int s = 0; s < S - 1; ++s) issue(s);

#pragma unroll 1
  for (long long it = 0;; ++it) {
    const long long t = it * gws + gw;
    if (t >= n_tiles) break;
    if (__shfl_sync(0xffffffffu, ld_volatile_s32(&ctl->prune_flag), 0)) cta_prune_all<QB>(ps, ctl, p.nq, p.kp);
    issue(it + S - 1);

    const int stage = (int)(it % S);
    const uint32_t parity = (uint32_t)((it / S) & 1);
    const long long item0 = t * RT;
    const long long my_item = item0 + lane;  // lane j keeps the result of tile row j
    bool valid = (lane < RT) && (my_item < n_items);
    uint32_t my_row = 0;
    float rinv = 1.f;
    if (valid) {
      my_row = gather ? __ldg(p.gather + my_item) : (uint32_t)my_item;
      if (p.mask != nullptr && !gather) valid = (__ldg(p.mask + (my_row >> 5)) >> (my_row & 31)) & 1u;
      if (MODE == MODE_DOT && p.inv_norm != nullptr) rinv = __ldg(p.inv_norm + my_row);
    }
    float tau_r[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) tau_r[qi] = *reinterpret_cast<volatile float*>(&tau[qi]);

    mbar_wait(&mybar[stage], parity);
    const float4* tb = reinterpret_cast<const float4*>(ring + (size_t)stage * tile_bytes);
    const float4* q4 = reinterpret_cast<const float4*>(qs);

    float mine[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) mine[qi] = 0.f;

    const int rows = (int)((n_items - item0) < RT ? (n_items - item0) : RT);
    for (int r = 0; r < rows; ++r) {
      float acc[QB];
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) acc[qi] = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float4 x = tb[r * C + c];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) acc[qi] = accum4<MODE>(acc[qi], x, q4[qi * C + c]);
      }
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        float v = acc[qi];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == r) mine[qi] = v;
      }
    }

#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
      if (qi < p.nq) {
        float score = mine[qi];
        if (MODE == MODE_DOT) score = 1.0f - score * (rinv * rnq[qi]);
        const bool pass = valid && (score <= tau_r[qi]);
        const int after = warp_append(ps.ref(qi), pass, make_key(score, my_row));
        if (after >= highwater && lane == 0) st_volatile_s32(&ctl->prune_flag, 1);
      }
    }
    __syncwarp();
    fence_proxy_async();
  }

  cta_finish<QB>(ps, ctl, p.nq, p.kp, nw, p.partial);
}

// Host-side launchers (scan_launch.cu / scan_fast_*.cu).
// Returns 0 when a fast kernel exists for this padded dimension, else -1.
int launch_scan_fast(int dp, int qb, int mode, const ScanParams& p, int grid, cudaStream_t st);
int launch_scan_generic(int qb, int mode, const ScanParams& p, int grid, int nw, cudaStream_t st);
int scan_fast_supported(int dp);
int scan_fast_tile_rows(int dp);
int scan_fast_max_qb(int dp);
int scan_fast_ring_bytes(int dp);
int scan_set_attributes();  // opt in to large dynamic shared memory for every instantiation

}  // namespace qg
