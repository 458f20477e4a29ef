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
  int dbg;                 // diagnostics (QG_SCAN_DBG): 1 = never admit a row, 2 = do not score at all
  unsigned long long* trace;  // diagnostics (QG_SCAN_TRACE): [cta][16] time stamps and counters, else nullptr
  // dense kernel, kp <= 128: threshold exchange between the CTAs of one launch (see scan_dense_kernel),
  // [QB][grid * warps] words of (epoch << 32 | score bits); nullptr = off
  unsigned long long* xchg;
  uint32_t epoch;             // distinguishes this launch's words from older ones (never 0)
  const float* pace;  // dense kernel: per-row array of the index read for pacing (see the kernel), or nullptr
};

// A partial list cut by a threshold that did not come from the CTA's own rows ends in this pseudo key:
// score = the threshold, row = XCHG_ROW. finalize.cu treats it like a full list's last key (rows above it
// may have been dropped) and never as a candidate.
constexpr uint32_t XCHG_ROW = 0xFFFFFFFFu;
constexpr int XCHG_MAX_CTAS = 148;  // launches with more CTAs, or query blocks above XCHG_MAX_QB, run without it
constexpr int XCHG_MAX_QB = 2;
constexpr uint32_t XCHG_FIRST_TILE = 2;  // a CTA's first fetch: when its warps reach this tile
constexpr int XCHG_FETCHES = 6;
constexpr int DEFER_ROWS = 64;           // single-query blocks: rows a warp scores before it must admit rows blindly          // at most this many fetches per query; fewer once one saw every publisher
constexpr int XCHG_MAX_KP = 128;     // r = kp / 32 <= 4 published scores per lane decide the bound

__device__ __forceinline__ unsigned long long scan_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

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
  int kept;    // cta_finish: keys that survive an exchanged threshold
  int prunes;  // mid-scan prunes (diagnostics)
  int x_busy, x_count;     // threshold exchange: a fetch is in flight (owned by one warp) / fetches completed
  int x_full;              // bit qi: a fetch for query qi saw the score of every publishing warp
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
  if (threadIdx.x == 0) {
    ctl->prunes++;
    st_volatile_s32(&ctl->prune_flag, 0);
  }
  __syncthreads();
}

// End-of-scan protocol: wait until every warp has finished appending (taking part in any
// prune still requested), then sort every pool and write the CTA's partial lists.
template <int QB>
__device__ __forceinline__ void cta_finish(const PoolSet<QB>& ps, ScanCtl* ctl, int nq, int kp, int nwarps,
                                           uint64_t* partial, const float* tau_ext = nullptr,
                                           uint64_t* scratch = nullptr) {
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
    if (tau_ext != nullptr && ps.cnt[qi] <= 64) {
      // The usual end of a scan that had an exchanged threshold: a few dozen keys. One warp ranks them (two keys
      // per lane, every key compared with every other through shuffles) and writes the list; no block barrier.
      if (threadIdx.x < 32) {
        const int n0 = ps.cnt[qi];
        const float ext = *reinterpret_cast<const volatile float*>(&tau_ext[qi]);
        const uint64_t* src0 = ps.keys + (size_t)qi * ps.slots;
        uint64_t k0 = lane < n0 ? src0[lane] : KEY_NONE, k1 = lane + 32 < n0 ? src0[lane + 32] : KEY_NONE;
        if (k0 != KEY_NONE && !(key_score(k0) <= ext)) k0 = KEY_NONE;  // above the threshold: cannot matter
        if (k1 != KEY_NONE && !(key_score(k1) <= ext)) k1 = KEY_NONE;
        int r0 = 0, r1 = 0;
        for (int j = 0; j < 32; ++j) {
          const uint64_t a = __shfl_sync(0xffffffffu, k0, j), b = __shfl_sync(0xffffffffu, k1, j);
          r0 += (a < k0) + (b < k0);
          r1 += (a < k1) + (b < k1);
        }
        const int n = __popc(__ballot_sync(0xffffffffu, k0 != KEY_NONE)) + __popc(__ballot_sync(0xffffffffu, k1 != KEY_NONE));
        uint64_t* out = partial + ((size_t)qi * gridDim.x + blockIdx.x) * kp;
        // keys are distinct (the row is part of the key), so the ranks of the kept ones are 0 .. n-1
        if (k0 != KEY_NONE && r0 < kp) out[r0] = k0;
        if (k1 != KEY_NONE && r1 < kp) out[r1] = k1;
        const uint64_t tail = (n < kp && ext < __int_as_float(0x7f800000)) ? make_key(ext, XCHG_ROW) : KEY_NONE;
        for (int i = n + lane; i < kp; i += 32) out[i] = i == kp - 1 ? tail : KEY_NONE;
      }
      continue;
    }
    if (tau_ext != nullptr && scratch != nullptr) {
      // keys admitted before the exchanged threshold arrived: only those at or below it can matter, and
      // sorting the few that are left is much cheaper than sorting the pool
      const float ext = *reinterpret_cast<const volatile float*>(&tau_ext[qi]);
      const int n0 = ps.cnt[qi];
      if (ext < __int_as_float(0x7f800000) && n0 > kp) {
        if (threadIdx.x == 0) ctl->kept = 0;
        __syncthreads();
        const uint64_t* src0 = ps.keys + (size_t)qi * ps.slots;
        for (int base = 0; base < n0; base += blockDim.x) {
          const int i = base + threadIdx.x;
          const uint64_t key = i < n0 ? src0[i] : KEY_NONE;
          const bool keep = i < n0 && key_score(key) <= ext;
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          if (m != 0) {
            const int ln = threadIdx.x & 31, leader = __ffs(m) - 1;
            int at = 0;
            if (ln == leader) at = atomicAdd(&ctl->kept, __popc(m));
            at = __shfl_sync(0xffffffffu, at, leader);
            if (keep) scratch[at + __popc(m & ((1u << ln) - 1u))] = key;
          }
        }
        __syncthreads();
        const int n1 = ctl->kept;
        uint64_t* dst0 = ps.keys + (size_t)qi * ps.slots;
        for (int i = threadIdx.x; i < n1; i += blockDim.x) dst0[i] = scratch[i];
        __syncthreads();
        if (threadIdx.x == 0) ps.cnt[qi] = n1;
        __syncthreads();
      }
    }
    block_prune(ps.ref(qi), kp);
    const int n = ps.cnt[qi];
    uint64_t* out = partial + ((size_t)qi * gridDim.x + blockIdx.x) * kp;
    const uint64_t* src = ps.keys + (size_t)qi * ps.slots;
    // fewer than kp keys although rows above an exchanged threshold were skipped: say so in the last slot
    uint64_t tail = KEY_NONE;
    if (tau_ext != nullptr && n < kp) {
      const float ext = *reinterpret_cast<const volatile float*>(&tau_ext[qi]);
      if (ext < __int_as_float(0x7f800000)) tail = make_key(ext, XCHG_ROW);
    }
    for (int i = threadIdx.x; i < kp; i += blockDim.x) out[i] = i < n ? src[i] : (i == kp - 1 ? tail : KEY_NONE);
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
    ctl->prunes = 0;
    ctl->x_busy = 0;
    ctl->x_count = 0;
    ctl->x_full = 0;
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

    mbar_poll(&mybar[stage], parity);  // polling: a thread parked by try_wait comes back late (see the dense kernel)
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
// Dense kernel: the fast kernel specialised for scans over consecutive rows (no row list) — the
// regime judged against the HBM roofline (one query, 1M x 128: the round-1 fast kernel spent ~580
// instructions per 16-row tile, two thirds of them loop overhead, at two warps per scheduler).
//  * NW = 16 warps per CTA for query blocks of one or two (4 KB tiles, three or four stages: the
//    same bytes in flight per SM, twice the warps to hide LDS / SHFL / FFMA latency), 8 otherwise;
//  * tile bookkeeping in 32-bit counters and one running pointer; the tile's mask bits come from one
//    warp-uniform word; every sub-batch of the tile is scored before a single vote decides whether
//    any lane has a candidate (the append path is rare once the thresholds have tightened);
//  * packed fp32 pairs (add.f32x2 / fma.f32x2: the same IEEE result per element, half the
//    instructions) while the accumulators fit in registers.
// Shared memory: ring [NW][STAGES][TILE_BYTES] | pool keys [QB][slots] | mbar [NW][STAGES]
//                | cnt[QB] | tau[QB] | ScanCtl
// ------------------------------------------------------------------------------------------
template <int D, int NW>
struct DenseGeom {
  using B = ScanGeom<D>;
  static constexpr int C = B::C, G = B::G, CPL = B::CPL, RPP = B::RPP, RB = B::RB, RS = B::RS, REP = B::REP;
  static constexpr int ROW_BYTES = B::ROW_BYTES;
  static constexpr int TILE_TARGET = NW > 8 ? 4096 : 8192;
  static constexpr int NSUB_RAW0 = TILE_TARGET / (RS * ROW_BYTES);
  static constexpr int NSUB_RAW = NSUB_RAW0 < B::NSUB_CAP ? NSUB_RAW0 : B::NSUB_CAP;
  static constexpr int NSUB = NSUB_RAW >= 4 ? 4 : (NSUB_RAW >= 2 ? 2 : 1);
  static constexpr int RT = RS * NSUB;
  static constexpr int TILE_BYTES = RT * ROW_BYTES;
  static constexpr int RING_BUDGET = 192 * 1024;
  static constexpr int S_RAW = RING_BUDGET / (NW * TILE_BYTES);
  static constexpr int STAGES = S_RAW >= 4 ? 4 : S_RAW;
  static_assert(RT <= 32 && (32 % RT) == 0, "a tile's mask bits must sit in one 32-bit word");
};

// 16 warps where three stages of 4 KB-target tiles fit the ring and the queries leave registers for it
// (8 warps with 6-8 KB tiles for the dimensions whose 16-warp tile is only 3 KB — 96, 192, 384, 768 — was measured
// no better: 12.5M x 96 754 vs 746 us, 1M x 96 88 vs 72 us, 2M x 768 899 vs 930 us.)
template <int D, int QB>
constexpr int dense_nw() {
  return (QB <= 2 && DenseGeom<D, 16>::S_RAW >= 3) ? 16 : 8;
}

// Pool size for `headroom` = keys the CTA's warps may still append after the prune flag went up
// (one tile per warp): a power of two, at least 512, with 1.5 * kp below the high-water mark.
__host__ __device__ constexpr int dense_pool_slots(int kp, int headroom, int qb) {
  (void)qb;
  int s = 512;
  while (s - headroom < kp + kp / 2) s <<= 1;
  return s;
}

template <int D, int QB>
__host__ __device__ constexpr size_t scan_dense_smem(int kp) {
  constexpr int NW = dense_nw<D, QB>();
  using Gm = DenseGeom<D, NW>;
  return (size_t)NW * Gm::STAGES * Gm::TILE_BYTES + (size_t)QB * dense_pool_slots(kp, NW * Gm::RT, QB) * 8 +
         (size_t)NW * Gm::STAGES * 8 + QB * 12 + sizeof(ScanCtl) + 64 +
         ((QB <= XCHG_MAX_QB && kp <= XCHG_MAX_KP) ? (size_t)XCHG_MAX_CTAS * NW * 8 + 32 : 0) +
         ((QB == 1 && Gm::RT <= 16 && kp <= XCHG_MAX_KP) ? (size_t)NW * DEFER_ROWS * 8 : 0);
}

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float f2_sum(uint64_t a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return lo + hi;
}

template <int D, int QB, int MODE, bool GATHER>
__global__ void __launch_bounds__(dense_nw<D, QB>() * 32, 1) scan_dense_kernel(const ScanParams p) {
  constexpr int NW = dense_nw<D, QB>();
  using Gm = DenseGeom<D, NW>;
  constexpr int C = Gm::C, G = Gm::G, CPL = Gm::CPL, RPP = Gm::RPP, RB = Gm::RB, RT = Gm::RT;
  constexpr int S = Gm::STAGES, TILE_BYTES = Gm::TILE_BYTES, RS = Gm::RS, NSUB = Gm::NSUB;
  constexpr bool PACK = (MODE != MODE_L1) && (QB * RB <= 16);
  static_assert(S >= 2, "the ring needs two stages");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* trace = p.trace ? p.trace + (size_t)blockIdx.x * 32 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = scan_global_ns();
  const int slots = dense_pool_slots(p.kp, NW * RT, QB);
  const int highwater = slots - NW * RT;

  unsigned char* ring = smem_raw + (size_t)warp * S * TILE_BYTES;
  uint64_t* pool_keys = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * S * TILE_BYTES);
  uint64_t* bars = pool_keys + (size_t)QB * slots;
  int* cnt = reinterpret_cast<int*>(bars + NW * S);
  float* tau = reinterpret_cast<float*>(cnt + QB);
  float* tau_ext = tau + QB;  // thresholds learnt from the other CTAs (only ever lowered, by warp 0)
  ScanCtl* ctl = reinterpret_cast<ScanCtl*>(tau_ext + QB);
  // exchange staging (query blocks of one or two): [grid * NW] words + its barrier, after the control block
  unsigned long long* xstage = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(ctl + 1) + 15) & ~(uintptr_t)15);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(xstage + XCHG_MAX_CTAS * NW);
  // single-query blocks: this warp's rows scored before any threshold was known (DEFER_ROWS keys)
  uint64_t* deferred = reinterpret_cast<uint64_t*>(xbar + 2) + (size_t)warp * DEFER_ROWS;
  uint64_t* mybar = bars + warp * S;
  PoolSet<QB> ps{pool_keys, cnt, tau, slots};

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * S; ++i) mbar_init(&bars[i], 1);
    if (QB <= XCHG_MAX_QB && p.kp <= XCHG_MAX_KP) mbar_init(xbar, 1);
    for (int qi = 0; qi < QB; ++qi) {
      cnt[qi] = 0;
      tau[qi] = __int_as_float(0x7f800000);
      tau_ext[qi] = __int_as_float(0x7f800000);
    }
    ctl->prune_flag = 0;
    ctl->done_warps = 0;
    ctl->prunes = 0;
    ctl->x_busy = 0;
    ctl->x_count = 0;
    ctl->x_full = 0;
    fence_mbar_init();
  }
  __syncthreads();

  const int grp = lane / G, gl = lane % G;

  // ---- queries into registers (L2 with packed pairs: negated, so that x + (-q) is the exact x - q) ----
  float4 qreg[QB][CPL];
  float rnq[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qi < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.queries + (size_t)qi * D) + c * G + gl);
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s);
      s = fmaf(v.w, v.w, s);
      if (PACK && MODE == MODE_L2) v = make_float4(-v.x, -v.y, -v.z, -v.w);
      qreg[qi][c] = v;
    }
    rnq[qi] = 1.f;
    if (MODE == MODE_DOT) {
#pragma unroll
      for (int o = G / 2; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      rnq[qi] = p.cosine ? (s > 0.f ? rsqrtf(s) : 0.f) : 1.f;
    }
  }

  // ---- tile schedule: warp gw owns the row blocks t = gw, gw + gws, ... (Handing the tiles out through a
  // per-CTA counter, or the last sixth of the corpus through a global one, evened out the finishing times of the
  // warps but was measured slower both times; what keeps the warps together is the pacing load below.) ----
  const long long n_items = p.n_items;
  const uint32_t n_tiles = (uint32_t)((n_items + RT - 1) / RT);
  const uint32_t gws = gridDim.x * NW;
  const uint32_t gw = blockIdx.x * NW + warp;
  uint32_t claimed = 0;
  auto claim = [&]() -> uint32_t { return (claimed++) * gws + gw; };
  // GATHER (scans over a row list, p.gather): a tile is RT listed rows, fetched with one bulk copy per row; lane l
  // holds the row number of the tile's row l. The numbers of a tile are loaded one iteration before the tile is
  // issued, so that the issue does not wait for them.
  auto load_ids = [&](uint32_t t) -> uint32_t {
    const long long item = (long long)t * RT + lane;
    return (GATHER && lane < RT && item < n_items) ? __ldg(p.gather + item) : 0u;
  };
  auto issue = [&](uint32_t t, int st, uint32_t ids) {
    if (t >= n_tiles) return;
    const long long left = n_items - (long long)t * RT;
    const uint32_t rows = (uint32_t)(left < RT ? left : RT);
    if (!GATHER) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&mybar[st], rows * Gm::ROW_BYTES);
        bulk_g2s(ring + (size_t)st * TILE_BYTES, p.vec + (size_t)t * RT * D, rows * Gm::ROW_BYTES, &mybar[st]);
      }
    } else {
      if (lane == 0) mbar_arrive_expect_tx(&mybar[st], rows * Gm::ROW_BYTES);
      __syncwarp();
      if (lane < rows)
        bulk_g2s(ring + (size_t)st * TILE_BYTES + (size_t)lane * Gm::ROW_BYTES, p.vec + (size_t)ids * D, Gm::ROW_BYTES,
                 &mybar[st]);
    }
  };
  if (trace && threadIdx.x == 0) trace[1] = scan_global_ns();
  uint32_t tq[S - 1];   // tiles claimed and in flight, oldest first
  uint32_t idq[S - 1];  // GATHER: their row numbers (lane l: row l of the tile)
#pragma unroll
  for (int s = 0; s < S - 1; ++s) {
    tq[s] = claim();
    idq[s] = load_ids(tq[s]);
    issue(tq[s], s, idq[s]);
  }
  uint32_t t_ahead = 0, id_ahead = 0;  // GATHER: the tile to be issued at the next iteration and its row numbers
  if (GATHER) {
    t_ahead = claim();
    id_ahead = load_ids(t_ahead);
  }

  // row this lane owns after the butterfly
  const int my_pass = gl / Gm::REP;
  const int my_row_in_tile = my_pass * RPP + grp;
  const bool my_unique = (gl % Gm::REP) == 0;

  // Threshold exchange. A CTA sees 1/148 of the rows, so its own kp-th best score is a loose threshold and the
  // first rows of every CTA are all admitted. Instead every warp publishes the best score of its FIRST tile
  // (one distinct row per warp, up to 2 368 per launch), and a little later one warp per CTA reads them all:
  // with r = kp / 32, the largest over the 32 lanes of each lane's r-th smallest published score is met by kp
  // distinct rows, hence a valid bound on the launch's kp-th best score — about the 0.5 % quantile after 8
  // rows per warp, where the CTA's own pool would still admit every row.
  const int xr = p.kp >> 5;  // 1, 2 or 4 (the host enables the exchange for kp <= 128 only)
  const int x_total = p.xchg != nullptr ? XCHG_FETCHES * p.nq : 0;  // fetches per launch
  bool x_pending = false;                                 // this warp issued a fetch and has to consume it
  constexpr bool XCH = QB <= XCHG_MAX_QB;                 // larger query blocks are compiled without the exchange
  bool x_open = XCH && p.xchg != nullptr;                 // this warp still looks for a fetch to do
  // Until a threshold exists every row would be admitted, and sixteen warps appending every row of their first
  // tiles to one pool is the most expensive part of a single-query scan (the pool overflows and is sorted while
  // all CTAs of the launch are at the same tile). A warp therefore parks the keys of its first tiles in its own
  // corner of shared memory and admits only the survivors once the exchanged threshold has arrived.
  constexpr bool DEFER = QB == 1 && RT <= 16;  // 32-row tiles (D = 32) need the shared memory for their pool
  constexpr int DEFER_TILES = DEFER_ROWS / RT;
  static_assert(DEFER_TILES >= 1, "a tile has at most 32 rows");
  bool deferring = DEFER && p.xchg != nullptr && p.kp <= XCHG_MAX_KP;
  int n_deferred = 0;  // tiles parked
  // admits the parked keys that pass the thresholds of the moment, one tile's worth between two looks at the
  // prune flag (the pool's head-room is one tile per warp)
  auto flush_deferred = [&]() {
    const int n = n_deferred * RT;
    for (int base = 0; base < n; base += RT) {
      const float thr = fminf(*reinterpret_cast<volatile float*>(&tau[0]), *reinterpret_cast<volatile float*>(&tau_ext[0]));
      const uint64_t key = lane < RT ? deferred[base + lane] : KEY_NONE;
      const bool pass = key != KEY_NONE && key_score(key) <= thr;
      const int after = warp_append(ps.ref(0), pass, key);
      if (after >= highwater && lane == 0) st_volatile_s32(&ctl->prune_flag, 1);
      // a look at the flag after EVERY chunk, the last one included: the caller appends the rows of its current
      // tile right after this returns, and a warp may add one tile's worth of keys between two looks
      __syncwarp();
      if (__shfl_sync(0xffffffffu, ld_volatile_s32(&ctl->prune_flag), 0)) cta_prune_all<QB>(ps, ctl, p.nq, p.kp);
    }
    n_deferred = 0;
  };
  uint32_t it = 0;
  int stage = 0;
  uint32_t parity = 0;
  bool first_tile = true;
#pragma unroll 1
  for (;; ++it) {
    const uint32_t t = tq[0];
    if (t >= n_tiles) break;  // a warp's tiles come in ascending order: nothing behind this one
    // checkpoint: take part in a prune if one was requested
    if (__shfl_sync(0xffffffffu, ld_volatile_s32(&ctl->prune_flag), 0)) cta_prune_all<QB>(ps, ctl, p.nq, p.kp);
    const uint32_t ids_cur = idq[0];
    {
#pragma unroll
      for (int s = 0; s + 1 < S - 1; ++s) {
        tq[s] = tq[s + 1];
        idq[s] = idq[s + 1];
      }
      const uint32_t tn = GATHER ? t_ahead : claim();
      const uint32_t idn = GATHER ? id_ahead : 0u;
      tq[S - 2] = tn;
      idq[S - 2] = idn;
      const int st_new = stage + (S - 1) >= S ? stage - 1 : stage + (S - 1);
      issue(tn, st_new, idn);
      if (GATHER) {
        t_ahead = claim();
        id_ahead = load_ids(t_ahead);
      }
    }

    // From its tile number XCHG_FIRST_TILE on, a warp that finds the staging area free (x_busy) fetches the
    // published scores of one query with one bulk copy and turns them into a threshold at the first later tile
    // at which the copy has landed — nothing waits for L2 or for another warp. Repeated until a fetch has seen
    // every publisher (at most XCHG_FETCHES times per query).
    if (XCH && x_pending && __all_sync(0xffffffffu, mbar_try_wait(xbar, (uint32_t)(ld_volatile_s32(&ctl->x_count) & 1)))) {
      const int xc = ld_volatile_s32(&ctl->x_count);
      const int qi = xc % p.nq;
      float best[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) best[jj] = __int_as_float(0x7f800000);
      int seen = 0;
      for (uint32_t i = lane; i < gws; i += 32) {
        const unsigned long long w = xstage[i];
        const bool fresh = (uint32_t)(w >> 32) == p.epoch;
        seen += fresh;
        float v = fresh ? __uint_as_float((uint32_t)w) : __int_as_float(0x7f800000);
        if (xr == 1) {
          best[0] = fminf(best[0], v);
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {  // keep the four smallest, ascending
            const float lo = fminf(best[jj], v);
            v = fmaxf(best[jj], v);
            best[jj] = lo;
          }
        }
      }
      float x = xr == 1 ? best[0] : (xr == 2 ? best[1] : best[3]);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
        seen += __shfl_xor_sync(0xffffffffu, seen, o);
      }
      __syncwarp();
      if (lane == 0) {
        volatile float* te = tau_ext + qi;
        if (trace && xc < 4) trace[8 + xc] = ((unsigned long long)it << 48) | ((unsigned long long)(uint32_t)cnt[0] << 24) | (uint32_t)seen;
        if (trace && !(*te < __int_as_float(0x7f800000)) && x < __int_as_float(0x7f800000)) trace[12] = scan_global_ns();
        if (x < *te) *te = x;
        // every warp that has a tile publishes once: with all of them seen, later fetches cannot improve on x
        int full = ld_volatile_s32(&ctl->x_full);
        if ((uint32_t)seen >= (n_tiles < gws ? n_tiles : gws)) full |= 1 << qi;
        st_volatile_s32(&ctl->x_full, full);
        st_volatile_s32(&ctl->x_count, xc + 1);  // also the phase of the staging barrier: one step per fetch
        __threadfence_block();
        st_volatile_s32(&ctl->x_busy, 0);
      }
      x_pending = false;
    } else if (XCH && !x_pending && x_open && it >= XCHG_FIRST_TILE) {
      int got = 0;  // 1 = this warp fetches now, -1 = the exchange is over: stop looking (a warp-local decision)
      if (lane == 0) {
        const int all_full = (1 << p.nq) - 1;
        if (ld_volatile_s32(&ctl->x_count) >= x_total || ld_volatile_s32(&ctl->x_full) == all_full) {
          got = -1;
        } else if (ld_volatile_s32(&ctl->x_busy) == 0 && atomicCAS(&ctl->x_busy, 0, 1) == 0) {
          // the owner before us may have finished the job between our look and our claim
          const int xc = ld_volatile_s32(&ctl->x_count);
          if (xc < x_total && ld_volatile_s32(&ctl->x_full) != all_full) {
            mbar_arrive_expect_tx(xbar, gws * 8u);
            bulk_g2s(xstage, p.xchg + (size_t)(xc % p.nq) * gws, gws * 8u, xbar);
            got = 1;
          } else {
            st_volatile_s32(&ctl->x_busy, 0);
          }
        }
      }
      got = __shfl_sync(0xffffffffu, got, 0);
      if (got > 0) x_pending = true;
      if (got < 0) x_open = false;
    }

    const long long item0 = (long long)t * RT;
    const long long left = n_items - item0;
    // bit r of `rowbits`: row r of the tile exists and passes the mask
    uint32_t rowbits = 0xffffffffu;
    if (!GATHER && p.mask != nullptr) rowbits = __ldg(p.mask + (item0 >> 5)) >> ((uint32_t)item0 & 31u);
    if (left < RT) rowbits &= (1u << (int)left) - 1u;
    uint32_t rowid[NSUB];  // index row of this lane's row in each sub-batch
    float rinv[NSUB];
#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb) {
      const int r = sb * RS + my_row_in_tile;
      rowid[sb] = GATHER ? __shfl_sync(0xffffffffu, ids_cur, r) : (uint32_t)item0 + (uint32_t)r;
      rinv[sb] = 1.f;
      if (MODE == MODE_DOT && p.inv_norm != nullptr) rinv[sb] = (r < left) ? __ldg(p.inv_norm + rowid[sb]) : 0.f;
    }
    // Pacing (L2 / plain dot product; the cosine scan gets the same effect from its 1/|x| load): one 4-byte
    // load per tile and lane from a per-row array of the index, which the warp must have received before it
    // votes on the tile. Without it a warp turns a landed tile around in ~0.2 us and the warps of the chip drift
    // up to 1.5x apart in row position; with a global-load round trip in every iteration they move through the
    // corpus as one front (first and last warp of a CTA 1 us apart instead of 8) and the same scan is 10 %
    // faster (1M x 128, one query: 98 -> 88.5 us). The value itself decides nothing (see its consumer below).
    float pace_v = 0.f;
    if (MODE != MODE_DOT || p.inv_norm == nullptr) {
      if (p.pace != nullptr && my_row_in_tile < left) pace_v = __ldg(p.pace + rowid[0]);
    }
    float tau_r[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi)
      tau_r[qi] = fminf(*reinterpret_cast<volatile float*>(&tau[qi]), *reinterpret_cast<volatile float*>(&tau_ext[qi]));

    // test_wait in a loop, not try_wait: a parked warp comes back late (1M x 128 cosine: 84 -> 77 us)
    mbar_poll(&mybar[stage], parity);
    const unsigned char* tile = ring + (size_t)stage * TILE_BYTES;
    if (trace && first_tile && threadIdx.x == 0) trace[2] = scan_global_ns();
    first_tile = false;
    if (p.dbg == 2) {
      __syncwarp();
      stage = (stage + 1 == S) ? 0 : stage + 1;
      parity ^= (stage == 0);
      continue;
    }

    float score[NSUB][QB];
#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb) {
      float acc[QB][RB];
      if constexpr (PACK) {
        const ulonglong2* tb = reinterpret_cast<const ulonglong2*>(tile) + sb * RS * C;
        uint64_t acc2[QB][RB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi)
#pragma unroll
          for (int pr = 0; pr < RB; ++pr) acc2[qi][pr] = 0ull;
#pragma unroll
        for (int pr = 0; pr < RB; ++pr) {
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const ulonglong2 x = tb[(pr * RPP + grp) * C + c * G + gl];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
              const uint64_t q01 = f2_pack(qreg[qi][c].x, qreg[qi][c].y), q23 = f2_pack(qreg[qi][c].z, qreg[qi][c].w);
              if (MODE == MODE_L2) {
                const uint64_t d01 = f2_add(x.x, q01), d23 = f2_add(x.y, q23);
                acc2[qi][pr] = f2_fma(d01, d01, acc2[qi][pr]);
                acc2[qi][pr] = f2_fma(d23, d23, acc2[qi][pr]);
              } else {
                acc2[qi][pr] = f2_fma(x.x, q01, acc2[qi][pr]);
                acc2[qi][pr] = f2_fma(x.y, q23, acc2[qi][pr]);
              }
            }
          }
        }
#pragma unroll
        for (int qi = 0; qi < QB; ++qi)
#pragma unroll
          for (int pr = 0; pr < RB; ++pr) acc[qi][pr] = f2_sum(acc2[qi][pr]);
      } else {
        const float4* tb = reinterpret_cast<const float4*>(tile) + sb * RS * C;
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
      }
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        const float v = butterfly_reduce<RB, G>(acc[qi], lane);
        score[sb][qi] = (MODE == MODE_DOT) ? 1.0f - v * (rinv[sb] * rnq[qi]) : v;
      }
    }

    // The consumer of the pacing load: the warp waits here until the value is back. (A consumer without an effect
    // is dropped by the assembler, hence a store behind a comparison with a NaN pattern that no sum of squares
    // produces; the word it would write is scratch that cta_finish initialises before use.)
    // Mixing in a score keeps the wait behind the arithmetic, where the load has had the whole tile to come back.
    if ((__float_as_uint(pace_v) ^ __float_as_uint(score[0][0])) == 0x7fc12345u && pace_v != pace_v)
      st_volatile_s32(&ctl->kept, 0);
    if (XCH && p.xchg != nullptr && it == 0) {  // publish the best score of this warp's first tile, per query
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) {
        if (qi < p.nq) {
          float m = __int_as_float(0x7f800000);
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb)
            if (my_unique && ((rowbits >> (sb * RS + my_row_in_tile)) & 1u)) m = fminf(m, score[sb][qi]);
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane == 0) __stcg(p.xchg + (size_t)qi * gws + gw, ((unsigned long long)p.epoch << 32) | __float_as_uint(m));
        }
      }
    }

    if (DEFER && deferring) {
      if (tau_r[0] < __int_as_float(0x7f800000)) {
        // a threshold has arrived: admit what survives of the parked tiles, then go on normally
        __syncwarp();
        flush_deferred();
        deferring = false;
      } else if (n_deferred < DEFER_TILES) {
        uint64_t* slot = deferred + n_deferred * RT;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {
          const int r = sb * RS + my_row_in_tile;
          if (my_unique)
            slot[r] = ((rowbits >> r) & 1u) ? make_key(score[sb][0], rowid[sb]) : KEY_NONE;
        }
        ++n_deferred;
        __syncwarp();
        stage = (stage + 1 == S) ? 0 : stage + 1;
        parity ^= (stage == 0);
        continue;
      } else {
        // no threshold after DEFER_TILES tiles (a small launch, or the exchange is off): admit everything
        __syncwarp();
        flush_deferred();
        deferring = false;
      }
    }

    // one vote per tile: does any lane hold a candidate for any query?
    bool hit = false;
#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb) {
      const bool live = my_unique && ((rowbits >> (sb * RS + my_row_in_tile)) & 1u);
#pragma unroll
      for (int qi = 0; qi < QB; ++qi) hit |= live && (qi < p.nq) && (score[sb][qi] <= tau_r[qi]);
    }
    if (__any_sync(0xffffffffu, hit) && p.dbg != 1) {
#pragma unroll
      for (int sb = 0; sb < NSUB; ++sb) {
        const bool live = my_unique && ((rowbits >> (sb * RS + my_row_in_tile)) & 1u);
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
          if (qi < p.nq) {  // warp-uniform
            const bool pass = live && (score[sb][qi] <= tau_r[qi]);
            const int after = warp_append(ps.ref(qi), pass, make_key(score[sb][qi], rowid[sb]));
            if (after >= highwater && lane == 0) st_volatile_s32(&ctl->prune_flag, 1);
          }
        }
      }
    }
    // every lane's reads of this stage precede lane 0's next bulk copy into it (issued one iteration later)
    __syncwarp();
    stage = (stage + 1 == S) ? 0 : stage + 1;
    parity ^= (stage == 0);
  }
  if (DEFER && n_deferred > 0) {
    __syncwarp();
    flush_deferred();
  }
  if (XCH && x_pending) {  // no copy may outlive the CTA
    mbar_wait(xbar, (uint32_t)(ld_volatile_s32(&ctl->x_count) & 1));
    x_pending = false;
  }
  if (trace && threadIdx.x == 0) {
    trace[6] = ((unsigned long long)(uint32_t)ctl->prunes << 32) | __float_as_uint(tau_ext[0]);
    trace[7] = ((unsigned long long)(uint32_t)cnt[0] << 32) | __float_as_uint(tau[0]);
  }
  if (trace && lane == 0) {
    const unsigned long long now = scan_global_ns();
    trace[16 + warp] = ((unsigned long long)it << 48) | (now & 0xffffffffffffull);  // tiles done, time
    atomicMax(&trace[3], now);  // last warp out of the tile loop
    atomicMin(&trace[5], now);  // first warp out
  }

  // the ring is idle by now (every copy issued was waited for): scratch for the final compaction
  cta_finish<QB>(ps, ctl, p.nq, p.kp, NW, p.partial, p.xchg != nullptr ? tau_ext : nullptr,
                 reinterpret_cast<uint64_t*>(smem_raw));
  if (trace && threadIdx.x == 0) trace[4] = scan_global_ns();
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
    ctl->prunes = 0;
    ctl->x_busy = 0;
    ctl->x_count = 0;
    ctl->x_full = 0;
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

    mbar_poll(&mybar[stage], parity);  // polling: a thread parked by try_wait comes back late (see the dense kernel)
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
// Dense kernel (no row list): geometry depends on the query block too (16 warps for blocks of one or two).
int launch_scan_dense(int dp, int qb, int mode, const ScanParams& p, int grid, cudaStream_t st);
int scan_dense_geometry(int dp, int qb, int kp, int* tile_rows, int* warps, size_t* smem);  // 0, or -1: no kernel
int launch_scan_generic(int qb, int mode, const ScanParams& p, int grid, int nw, cudaStream_t st);
int scan_fast_supported(int dp);
int scan_fast_tile_rows(int dp);
int scan_fast_max_qb(int dp);
int scan_fast_ring_bytes(int dp);
int scan_set_attributes();  // opt in to large dynamic shared memory for every instantiation

}  // namespace qg
