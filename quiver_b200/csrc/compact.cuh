// compact.cuh — tombstone compaction: dead rows are squeezed out of every per-row array of an index
// (the reference simply deletes the map entry, exact.go:61-70 / hybrid_index.go:244-290,
// so a long-lived index never carries dead rows — this is what gives the device index the same property).
#pragma once
#include "common.cuh"

namespace qg {

// Exclusive prefix sum of n 32-bit counts; out has n + 1 entries, out[n] = total.
// block_tmp: at least ceil(n / 4096) + 1 words of scratch.
int launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, long long n, uint32_t* block_tmp, cudaStream_t st);
inline long long scan_tmp_words(long long n) { return (n + 4095) / 4096 + 1; }

// cnt[w] = popcount(live[w]) for the n_words words of the live mask.
int launch_word_popcount(const uint32_t* live, long long n_words, uint32_t* cnt, cudaStream_t st);

// map[r] = new row of old row r (word_off[r / 32] + rank of r inside its word), 0xFFFFFFFF for a dead row.
int launch_compact_map(const uint32_t* live, const uint32_t* word_off, long long n_rows, uint32_t* map,
                       cudaStream_t st);

// out[r] = map[r] widened to int64, -1 for a dead row (the form qg_index_compact hands to its caller).
int launch_widen_map(const uint32_t* map, long long n, long long* out, cudaStream_t st);

struct CompactRowsArgs {
  const uint32_t* map;  // old row -> new row
  long long n_rows;     // old rows
  const float* vec;     float* vec_out;      int dp;    // fp32 rows, dp % 4 == 0
  const void* vec16;    void* vec16_out;     int dp16;  // bf16 rows (dp16 % 8 == 0) or nullptr
  const float* inv_norm; float* inv_norm_out;
  const float* norm2;    float* norm2_out;
  const float* unit_bias; float* unit_bias_out;
};
// One warp per 32 old rows: live rows are copied to their new position in every array.
int launch_compact_rows(const CompactRowsArgs& a, int sm_count, cudaStream_t st);

struct CompactColArgs {
  const uint32_t* map;
  long long n;  // old rows the column covers
  const uint8_t* kind;   uint8_t* kind_out;
  const double* num;     double* num_out;
  const int32_t* scode;  int32_t* scode_out;
  const int32_t* fcode;  int32_t* fcode_out;
  const int32_t* arr_off;  // nullable: CSR offsets of the old rows
  uint32_t* arr_cnt_out;   // nullable: element count per NEW row (input of the offset scan)
};
int launch_compact_column(const CompactColArgs& a, cudaStream_t st);

// Elements of each live old row copied to new_off[new row] .. (one warp per old row).
int launch_compact_elems(const uint32_t* map, long long n, const int32_t* arr_off, const int32_t* arr_code,
                         const uint32_t* new_off, int32_t* code_out, int sm_count, cudaStream_t st);

}  // namespace qg
