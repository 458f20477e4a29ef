// misc.cuh — small kernels around the scan: row norms, synthetic fill, predicate masks,
// mask combination / compaction, tombstones.
#pragma once
#include "../../include/quiver_gpu.h"
#include "common.cuh"

namespace qg {

struct FacetColDev {
  const uint8_t* kind;
  const double* num;
  const int32_t* scode;
  const int32_t* fcode;
  const int32_t* arr_off;   // [rows + 1] CSR offsets of the array-valued rows' elements, or nullptr
  const int32_t* arr_code;  // element codes
};

struct FilterProgDev {
  const qg_pred* preds;
  int n_preds;
  const qg_clause* clauses;
  const int32_t* iset;
  const double* fset;
};

// 1/|x| per row (0 for zero rows), |x|^2 per row, 1.0 into unit_bias, and the running max |x|^2
// (atomic, device scalar).
int launch_row_norms(const float* vec, long long row0, long long n, int dp, int d, float* inv_norm, float* norm2,
                     float* unit_bias, float* max_norm2, cudaStream_t st);
int launch_fill_f32(float* p, long long n, float v, cudaStream_t st);
// rows [row0, row0+n) of the index filled from the generator (kind 0..3).
int launch_synth_fill(float* vec, long long row0, long long n, int dp, int d, int kind, uint64_t seed,
                      long long global_row0, cudaStream_t st);
// mask words [ceil(n/32)]: bit = row matches every predicate. matches: device counter (zeroed by the call).
int launch_filter_eval(const FacetColDev* cols, FilterProgDev prog, long long n, uint32_t* mask,
                       unsigned long long* matches, cudaStream_t st);
// out = a & b over n_words words (b may be nullptr => copy); count: device counter of set bits (zeroed).
int launch_mask_and(const uint32_t* a, const uint32_t* b, long long n_rows, uint32_t* out, unsigned long long* count,
                    cudaStream_t st);
// row ids of the set bits of mask -> list (unordered between warps); n_out: device counter (zeroed).
int launch_mask_compact(const uint32_t* mask, long long n_rows, uint32_t* list, unsigned long long* n_out,
                        cudaStream_t st);
// clear the live bit of each listed row; n_cleared counts the rows that were live.
int launch_tombstone(uint32_t* live, const long long* rows, long long n, long long n_rows,
                     unsigned long long* n_cleared, cudaStream_t st);
// set live bits for rows [row0, row0+n)
int launch_set_live(uint32_t* live, long long row0, long long n, cudaStream_t st);
// gather rows into a dense [n x d] buffer (fetch)
int launch_fetch_rows(const float* vec, int dp, int d, long long n_rows, const long long* rows, long long n,
                      float* out, cudaStream_t st);

}  // namespace qg
