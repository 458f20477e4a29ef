// tc_scan.cuh — tensor-core regime of the exact scan (large query batches).
//
// For batches (Q >= 2 with a bf16 copy, else Q >= 8) the distance scan of ExactIndex.Search / HybridIndex.BatchSearch
// (reference pkg/hybrid/exact.go:114-129, hybrid_index.go:703-795: one goroutine per query, each
// a full pass over the corpus) is a dense [rows x d] x [d x Q] contraction. tc_scan.cu runs it on
// the 5th-generation tensor cores: TMA stages tiles of a bf16 copy of the corpus (128 rows x 64
// elements, 128-byte swizzle; the fp32 rows with kind::tf32 when there is no bf16 copy) in shared
// memory, the queries sit in tensor memory, one elected thread issues tcgen05.mma with the
// accumulators in TMEM, and the epilogue warps read them back with tcgen05.ld and keep only the rows
// whose score is below a per-query threshold. The low-precision scores only SELECT candidates;
// returned distances are recomputed in the reference's arithmetic and the selection is certified
// (finalize.cu).
#pragma once
#include "common.cuh"

namespace qg {

constexpr int TC_TILE_ROWS = 128;       // MMA M
constexpr int TC_KBLOCK = 32;           // floats per 128-byte swizzle row
constexpr int TC_STAGE_BYTES = TC_TILE_ROWS * TC_KBLOCK * 4;
constexpr int TC_MAX_COLS = 256;        // MMA N (queries per pass)
constexpr int TC_THREADS = 320;         // SS form: warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_SAMPLE_RANK_MAX = 16;  // largest order statistic of the sample the threshold kernel can take
// Order statistic used as threshold: the spread of the admitted count is Gamma(rank) / rank, so larger k
// (fewer admitted rows per wanted row) takes a higher rank from a proportionally larger sample.
// Rank 8 left a measurable lower tail at k = 10 (1 query in 10 000 admitted 37 rows instead of ~256 and
// could not be certified); rank 12 puts that below 1e-6 per query for a sample 1.5x as large.
inline int tc_sample_rank(int k) { return k > 32 ? 16 : 12; }
constexpr int TC_CAND_CAP = 2048;       // candidate capacity per query per pass
constexpr int TC_PASS_GROUP = 8;        // passes of a multi-pass search finalized by one finalize_cand launch

struct TcPlan {
  int variant;      // 1 = TS (queries resident in tensor memory, dp <= 256), 0 = SS (queries in shared memory)
  int bf16;         // TS only: stream the bf16 copy of the corpus (kind::f16) instead of the fp32 rows (kind::tf32)
  int nblk;         // TS: resident blocks of 128 queries (1 or 2)
  int n_cols;       // queries per pass
  int kb;           // 128-byte k-blocks per row (32 floats or 64 bf16 elements)
  int ksteps;       // TS: MMA k-steps per row (8 floats or 16 bf16 elements each)
  int a_cols;       // TS: tensor-memory columns of one resident query block
  int stages;       // corpus ring depth
  int tile_rows;    // corpus rows per tile (sample granularity)
  int sample_vals;  // sampled scores kept per tile and query
  int smem;         // dynamic shared memory bytes
};

// Columns the bf16 copy of a row carries after the d vector elements: for L2 the three bf16 pieces of
// -|x|^2 / 2 (hi + mid + lo is exact to 24 bits), so that the MMA itself produces q.x - |x|^2 / 2 and the
// epilogue of an unmasked scan is a bare max tree ("raw" scan). They are only added where they fit the
// padding of the last 128-byte k-block (d = 96, 100, 200 ...): a whole extra k-block for three columns
// (d = 128) costs more shared-memory traffic than the lighter epilogue returns (measured: 126 us against
// 120 us per 256-query pass over 1M x 128). Dot / cosine rows need none (cosine rows are stored
// normalised), so their unmasked scans are always raw.
inline int tc_extra_cols(int d, int mode_l2) {
  if (!mode_l2) return 0;
  return (d + 3 + 63) / 64 == (d + 63) / 64 ? 3 : 0;
}
inline bool tc_raw_supported(int d, int mode_l2) { return !mode_l2 || tc_extra_cols(d, mode_l2) == 3; }
inline int tc_dp16(int d, int mode_l2) { return (d + tc_extra_cols(d, mode_l2) + 7) & ~7; }

// Geometry for a padded dimension dp (fp32 rows), d16 used elements per bf16 row (vector + extra
// columns) and nq queries; returns 0 when the tensor path fits.
int tc_plan(int dp, int d16, int nq, bool bf16, TcPlan* out);

struct TcArgs {
  const float* vec;         // [n_rows x dp]
  const void* vec16;        // [n_rows x dp16] bf16 copy (plan.bf16), else nullptr
  int dp16;
  int raw;                  // plan.bf16 only, no mask: scores come straight out of the MMA (no row-term column)
  // TS variant, multi-pass searches: launch_tc_pass with sample_only = 1 runs the sample + threshold stage
  // for ALL nq queries of the search (apack / sample / tau sized for every pass); the per-pass calls then
  // set presampled = 1 and point apack / tau at their pass.
  int sample_only, presampled;
  int n_cnt;                // sample_only: candidate counters (and n_cnt / n_cols work counters) the threshold
                            // kernel zeroes for the first group of passes
  long long n_rows;
  int dp;
  const float* row_norm2;   // [n_rows] |x|^2 (L2 scores)
  const float* inv_norm;    // [n_rows] 1/|x| (cosine), else nullptr
  const uint32_t* mask;     // row-pass bits or nullptr
  const float* bias;        // TS variant: [rows padded to 64] |x|^2 (L2) or 1 (dot), +inf for rows that
                            // may not match (mask already applied) and for the padding
  const float* queries;     // [nq x dp] device
  const void* apack;        // TS variant: this pass's block of the packed queries (launch_tc_pack)
  int* work_counter;        // TS variant: one int of scratch (tile claims of the main scan)
  int nq;
  int mode;                 // MODE_L2 / MODE_DOT (scan.cuh)
  int cosine;
  // sample stage
  uint32_t* sample;         // [n_cols][n_sample][plan.sample_vals] ordered-float images
  int n_sample;             // sampled tiles
  int sample_rank;          // order statistic of the sample that becomes the threshold (<= TC_SAMPLE_RANK_MAX)
  float* tau;               // [n_cols] thresholds (written by the threshold kernel)
  // main stage
  uint64_t* cand;           // [n_cols][TC_CAND_CAP] scan keys
  int* cand_cnt;            // [n_cols]
  unsigned long long* dbg;  // optional [64] per-role cycle counters of CTA 0 of the main scan, else nullptr
};

// Optional hook called on the launching stream around the two stages of a pass (profiling):
// stage 0 = sample + threshold kernels, stage 1 = main scan kernel; begin = 1 / end = 0.
struct TcStageHook {
  void (*fn)(void* ctx, int stage, int begin, cudaStream_t st);
  void* ctx;
};
// Enqueue sample -> threshold -> main scan for one pass of args.nq <= plan.n_cols queries.
int launch_tc_pass(const TcPlan& plan, const TcArgs& args, int sm_count, cudaStream_t st, int* launches,
                   const TcStageHook* hook = nullptr);
int tc_set_attributes();
// TS variant: all nq queries of a search in tensor-memory order, one block of plan.n_cols queries per
// pass (pass i starts at byte i * tc_pack_bytes(plan, plan.n_cols)). Cosine queries are pre-scaled.
// `ones`: query elements d .. d + ones - 1 are set to 1 (raw L2 scan: they pick up the norm columns).
size_t tc_pack_bytes(const TcPlan& plan, int nq);
int launch_tc_pack(const TcPlan& plan, const float* queries, int nq, int dp, int d, int ones, int cosine, void* apack,
                   cudaStream_t st);
// out[row0 + r][0..d) = bf16(vec[row0 + r][i] * (inv_norm ? inv_norm[row0 + r] : 1)), then the
// tc_extra_cols() columns (norm2 != nullptr: the three pieces of -norm2 / 2), zero padded to dp16.
int launch_tc_to_bf16(const float* vec, long long row0, long long n, int dp, int d, int dp16, const float* norm2,
                      const float* inv_norm, void* out, cudaStream_t st);
// bias[i] = row i passes (mask nullptr = all rows < n_rows) ? (norm2 ? norm2[i] : 1) : +inf, for i < n_pad.
int launch_tc_bias(const uint32_t* mask, const float* norm2, long long n_rows, long long n_pad, float* bias,
                   cudaStream_t st);
// 0 when the driver entry point for tensor maps is available.
int tc_available();

}  // namespace qg
