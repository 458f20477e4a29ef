// misc.cu — see misc.cuh.
#include "misc.cuh"
#include "synth.cuh"

namespace qg {

// ---- K9: row norms -------------------------------------------------------------------------------
// The reference recomputes |b|^2 inside every CosineDistance call (distances.go:21); here it is
// computed once at upload. Only the fp32 scan score uses this column; returned distances are
// re-derived exactly by finalize.cu.
__global__ void __launch_bounds__(256) row_norms_kernel(const float* __restrict__ vec, long long row0, long long n,
                                                        int dp, int d, float* __restrict__ inv_norm,
                                                        float* __restrict__ norm2, float* __restrict__ unit_bias,
                                                        float* max_norm2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float local_max = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < n; r += (long long)gridDim.x * 8) {
    const float* x = vec + (size_t)(row0 + r) * dp;
    double s = 0.0;
    for (int i = lane; i < d; i += 32) s += (double)x[i] * (double)x[i];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float sf = (float)s;
    if (lane == 0) {
      inv_norm[row0 + r] = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.f;
      norm2[row0 + r] = sf;
      unit_bias[row0 + r] = 1.0f;
    }
    local_max = fmaxf(local_max, sf);
  }
  if (lane == 0 && local_max > 0.f) {
    // non-negative floats order like their bit patterns; round the bound up
    atomicMax(reinterpret_cast<unsigned int*>(max_norm2), __float_as_uint(local_max) + 1u);
  }
}

int launch_row_norms(const float* vec, long long row0, long long n, int dp, int d, float* inv_norm, float* norm2,
                     float* unit_bias, float* max_norm2, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  row_norms_kernel<<<(int)blocks, 256, 0, st>>>(vec, row0, n, dp, d, inv_norm, norm2, unit_bias, max_norm2);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

int launch_fill_f32(float* p, long long n, float v, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill_f32_kernel<<<(int)blocks, 256, 0, st>>>(p, n, v);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- synthetic fill -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) synth_fill_kernel(float* __restrict__ vec, long long row0, long long n, int dp,
                                                         int d, int kind, uint64_t seed, long long global_row0) {
  const long long total = n * dp;
  const int k = kind == 3 ? 2 : kind;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dp;
    const int c = (int)(i - r * dp);
    vec[(size_t)(row0 + r) * dp + c] = c < d ? synth_raw(k, seed, (uint64_t)(global_row0 + r), (uint32_t)c) : 0.f;
  }
}

// kind 3: divide every row by its L2 norm, norm accumulated sequentially in float64 (the
// products are exact, so this matches oracle/synth.h bit for bit).
__global__ void __launch_bounds__(128) synth_normalize_kernel(float* __restrict__ vec, long long row0, long long n,
                                                              int dp, int d) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (long long)gridDim.x * blockDim.x) {
    float* x = vec + (size_t)(row0 + r) * dp;
    double n2 = 0.0;
    for (int c = 0; c < d; ++c) n2 = __dadd_rn(n2, (double)x[c] * (double)x[c]);
    if (n2 > 0.0) {
      const double nrm = sqrt(n2);
      for (int c = 0; c < d; ++c) x[c] = (float)((double)x[c] / nrm);
    }
  }
}

int launch_synth_fill(float* vec, long long row0, long long n, int dp, int d, int kind, uint64_t seed,
                      long long global_row0, cudaStream_t st) {
  if (n <= 0) return 0;
  if (kind < 0 || kind > 3) return fail(1, "synthetic kind must be 0..3");
  synth_fill_kernel<<<148 * 8, 256, 0, st>>>(vec, row0, n, dp, d, kind, seed, global_row0);
  QG_CUDA_OK(cudaGetLastError());
  if (kind == 3) {
    long long blocks = (n + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_normalize_kernel<<<(int)blocks, 128, 0, st>>>(vec, row0, n, dp, d);
    QG_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// ---- K4: predicate mask ------------------------------------------------------------------------------
// Evaluates the normal form produced by the host-side compiler (quiver_b200/host/filter_compile.cpp)
// from facets.MatchesAllFilters (pkg/facets/facets.go:432-459) and core.matchesFilter
// (pkg/core/collection.go:532-575). One thread per row, one ballot per warp = one mask word.
__device__ __forceinline__ bool eval_clause(const qg_clause& c, const FacetColDev* cols, const FilterProgDev& prog,
                                            long long row) {
  bool r = false;
  if (c.op == QG_OP_FALSE) {
    r = false;
  } else if (c.op == QG_OP_TRUE) {
    r = true;
  } else {
    const FacetColDev col = cols[c.field];
    const uint8_t kb = col.kind[row];
    const int kind = kb & 0x7f;
    const bool nonempty = (kb & 0x80) != 0;
    switch (c.op) {
      case QG_OP_KIND_IN:
        r = ((c.ia >> kind) & 1) && (!c.ib || nonempty);
        break;
      case QG_OP_NUM_EQ_TOL:
        r = kind == QG_KIND_NUMBER && fabs(col.num[row] - c.fa) <= c.fb;
        break;
      case QG_OP_NUM_EQ:
        r = ((c.ia >> kind) & 1) && col.num[row] == c.fa;
        break;
      case QG_OP_NUM_CMP: {
        if (kind == QG_KIND_NUMBER) {
          const double v = col.num[row];
          r = c.ia == 0 ? v < c.fa : (c.ia == 1 ? v <= c.fa : (c.ia == 2 ? v > c.fa : v >= c.fa));
        }
        break;
      }
      case QG_OP_NUM_RANGE: {
        if (kind == QG_KIND_NUMBER) {
          const double v = col.num[row];
          bool lo_ok = true, hi_ok = true;
          if (c.ia & 1) lo_ok = (c.ia & 2) ? v >= c.fa : v > c.fa;
          if (c.ia & 4) hi_ok = (c.ia & 8) ? v <= c.fb : v < c.fb;
          r = lo_ok && hi_ok;
        }
        break;
      }
      case QG_OP_SCODE_EQ:
        r = ((c.ib >> kind) & 1) && col.scode[row] == c.ia;
        break;
      case QG_OP_SCODE_CMP:
        if ((c.ib >> kind) & 1) r = c.ic == 0 ? col.scode[row] < c.ia : col.scode[row] >= c.ia;
        break;
      case QG_OP_FCODE_EQ:
        r = kind == QG_KIND_STRING && col.fcode[row] == c.ia;
        break;
      case QG_OP_SCODE_IN:
        if ((c.ib >> kind) & 1) {
          const int v = col.scode[row];
          for (int j = 0; j < c.ic; ++j) r |= prog.iset[c.ia + j] == v;
        }
        break;
      case QG_OP_FCODE_IN:
        if (kind == QG_KIND_STRING) {
          const int v = col.fcode[row];
          for (int j = 0; j < c.ic; ++j) r |= prog.iset[c.ia + j] == v;
        }
        break;
      case QG_OP_NUM_IN_TOL:
        if (kind == QG_KIND_NUMBER) {
          const double v = col.num[row];
          for (int j = 0; j < c.ic; ++j) r |= fabs(v - prog.fset[c.ia + j]) <= c.fb;
        }
        break;
      case QG_OP_NUM_IN:
        if ((c.ib >> kind) & 1) {
          const double v = col.num[row];
          for (int j = 0; j < c.ic; ++j) r |= v == prog.fset[c.ia + j];
        }
        break;
      case QG_OP_NUM_BITS_EQ:
        r = kind == QG_KIND_NUMBER && __double_as_longlong(col.num[row]) == __double_as_longlong(c.fa);
        break;
      case QG_OP_ELEM_IN:
        if (kind == QG_KIND_OTHER && col.arr_off != nullptr) {
          const int e1 = col.arr_off[row + 1];
          for (int e = col.arr_off[row]; e < e1 && !r; ++e) {
            const int v = col.arr_code[e];
            for (int j = 0; j < c.ic; ++j) r |= prog.iset[c.ia + j] == v;
          }
        }
        break;
      case QG_OP_WHOLE_EQ:
        r = kind == QG_KIND_OTHER && col.fcode[row] == c.ia;
        break;
      default:
        r = false;
    }
  }
  return c.negate ? !r : r;
}

__global__ void __launch_bounds__(256) filter_eval_kernel(const FacetColDev* __restrict__ cols, FilterProgDev prog,
                                                          long long n, uint32_t* __restrict__ mask,
                                                          unsigned long long* matches) {
  const long long n_pad = (n + 31) & ~31ll;
  unsigned long long local = 0;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_pad;
       row += (long long)gridDim.x * blockDim.x) {
    bool ok = row < n;
    if (ok) {
      for (int pi = 0; pi < prog.n_preds && ok; ++pi) {
        const qg_pred pd = prog.preds[pi];
        bool any = false;
        bool norow = false;
        for (int ci = 0; ci < pd.n_clauses; ++ci) {
          const qg_clause c = prog.clauses[pd.first_clause + ci];
          if (pd.require_row && c.op >= QG_OP_KIND_IN) norow |= (cols[c.field].kind[row] & 0x7f) == QG_KIND_NOROW;
          any |= eval_clause(c, cols, prog, row);
        }
        bool res = pd.negate ? !any : any;
        if (norow) res = false;
        ok = res;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) {
      mask[row >> 5] = m;
      local += __popc(m);
    }
  }
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(matches, local);
}

int launch_filter_eval(const FacetColDev* cols, FilterProgDev prog, long long n, uint32_t* mask,
                       unsigned long long* matches, cudaStream_t st) {
  QG_CUDA_OK(cudaMemsetAsync(matches, 0, sizeof(unsigned long long), st));
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  filter_eval_kernel<<<(int)blocks, 256, 0, st>>>(cols, prog, n, mask, matches);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- mask AND / compaction -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_and_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                       long long n_words, uint32_t tail_mask,
                                                       uint32_t* __restrict__ out, unsigned long long* count) {
  unsigned long long local = 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
       w += (long long)gridDim.x * blockDim.x) {
    uint32_t v = a[w];
    if (b != nullptr) v &= b[w];
    if (w == n_words - 1) v &= tail_mask;
    out[w] = v;
    local += __popc(v);
  }
  for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

int launch_mask_and(const uint32_t* a, const uint32_t* b, long long n_rows, uint32_t* out, unsigned long long* count,
                    cudaStream_t st) {
  QG_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  const long long n_words = (n_rows + 31) / 32;
  if (n_words <= 0) return 0;
  const int rem = (int)(n_rows & 31);
  const uint32_t tail = rem ? ((1u << rem) - 1u) : 0xffffffffu;
  long long blocks = (n_words + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mask_and_kernel<<<(int)blocks, 256, 0, st>>>(a, b, n_words, tail, out, count);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) mask_compact_kernel(const uint32_t* __restrict__ mask, long long n_words,
                                                           uint32_t* __restrict__ list, unsigned long long* n_out) {
  const int lane = threadIdx.x & 31;
  const long long n_pad = (n_words + 31) & ~31ll;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_pad;
       w += (long long)gridDim.x * blockDim.x) {
    const uint32_t v = w < n_words ? mask[w] : 0u;
    const int c = __popc(v);
    // warp-inclusive scan of the counts, one atomic per warp
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(n_out, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned long long pos = base + (unsigned long long)(incl - c);
    uint32_t bits = v;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      list[pos++] = (uint32_t)(w * 32 + b);
    }
  }
}

int launch_mask_compact(const uint32_t* mask, long long n_rows, uint32_t* list, unsigned long long* n_out,
                        cudaStream_t st) {
  QG_CUDA_OK(cudaMemsetAsync(n_out, 0, sizeof(unsigned long long), st));
  const long long n_words = (n_rows + 31) / 32;
  if (n_words <= 0) return 0;
  long long blocks = (n_words + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mask_compact_kernel<<<(int)blocks, 256, 0, st>>>(mask, n_words, list, n_out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- tombstones / live bits -----------------------------------------------------------------------
__global__ void tombstone_kernel(uint32_t* live, const long long* __restrict__ rows, long long n, long long n_rows,
                                 unsigned long long* n_cleared) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = rows[i];
    if (r < 0 || r >= n_rows) continue;
    const uint32_t bit = 1u << (r & 31);
    const uint32_t old = atomicAnd(&live[r >> 5], ~bit);
    if (old & bit) atomicAdd(n_cleared, 1ull);
  }
}

int launch_tombstone(uint32_t* live, const long long* rows, long long n, long long n_rows,
                     unsigned long long* n_cleared, cudaStream_t st) {
  QG_CUDA_OK(cudaMemsetAsync(n_cleared, 0, sizeof(unsigned long long), st));
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  tombstone_kernel<<<(int)blocks, 256, 0, st>>>(live, rows, n, n_rows, n_cleared);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void set_live_kernel(uint32_t* live, long long row0, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = row0 + i;
    atomicOr(&live[r >> 5], 1u << (r & 31));
  }
}

int launch_set_live(uint32_t* live, long long row0, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  set_live_kernel<<<(int)blocks, 256, 0, st>>>(live, row0, n);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void fetch_rows_kernel(const float* __restrict__ vec, int dp, int d, long long n_rows,
                                  const long long* __restrict__ rows, long long n, float* __restrict__ out) {
  const long long total = n * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = i / d;
    const int c = (int)(i - j * d);
    const long long r = rows[j];
    out[i] = (r >= 0 && r < n_rows) ? vec[(size_t)r * dp + c] : 0.f;
  }
}

int launch_fetch_rows(const float* vec, int dp, int d, long long n_rows, const long long* rows, long long n,
                      float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = (n * d + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fetch_rows_kernel<<<(int)blocks, 256, 0, st>>>(vec, dp, d, n_rows, rows, n, out);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace qg
