// exact.cuh — warp-cooperative distance in the REFERENCE's arithmetic and operation order.
//
// The distances Quiver returns are float64-accumulated sums rounded once to float32
// (reference pkg/vectortypes/distances.go:12-104) or, after a reload, sequential float32
// sums (pkg/hnsw/adapter.go:105-167). Floating-point addition is not associative, so to
// return the SAME float32 a candidate's terms are produced in parallel (each term is exact:
// the product of two float32 values fits a float64 mantissa, so an FMA contraction cannot
// change it) and then summed by ONE lane in index order, like the Go loop. Only the
// handful of candidates that survive the fp32 scan pay this.
#pragma once
#include "common.cuh"

namespace qg {

enum { METRIC_COSINE = 0, METRIC_L2 = 1, METRIC_DOT = 2, METRIC_SQL2 = 3, METRIC_L1 = 4 };
enum { ARITH_VECTORTYPES = 0, ARITH_HNSW_F32 = 1 };

constexpr int EXACT_CHUNK = 128;                                     // elements per staging chunk
constexpr int EXACT_SCRATCH_BYTES = 3 * EXACT_CHUNK * sizeof(double);  // per warp

// All 32 lanes of a warp call this with the same arguments; the result is returned on every
// lane. `a` is the first argument of the reference's distFunc (the query), `b` the stored row.
// `scratch` is EXACT_SCRATCH_BYTES of shared memory private to the warp.
__device__ __forceinline__ float exact_distance_warp(int metric, int arith, const float* __restrict__ a,
                                                     const float* __restrict__ b, int d, double* scratch) {
  const int lane = threadIdx.x & 31;
  if (arith == ARITH_HNSW_F32 && (metric == METRIC_COSINE || metric == METRIC_L2 || metric == METRIC_DOT)) {
    // pkg/hnsw/adapter.go:105-167 — everything float32, sequential, no fused multiply-add.
    float* t0 = reinterpret_cast<float*>(scratch);
    float* t1 = t0 + EXACT_CHUNK;
    float* t2 = t1 + EXACT_CHUNK;
    float s = 0.f;  // lane 0: dot or sum; lane 1: normA; lane 2: normB
    for (int base = 0; base < d; base += EXACT_CHUNK) {
      const int n = min(EXACT_CHUNK, d - base);
      for (int i = lane; i < n; i += 32) {
        const float x = a[base + i], y = b[base + i];
        if (metric == METRIC_L2) {
          const float df = __fsub_rn(x, y);
          t0[i] = __fmul_rn(df, df);
        } else {
          t0[i] = __fmul_rn(x, y);
          if (metric == METRIC_COSINE) {
            t1[i] = __fmul_rn(x, x);
            t2[i] = __fmul_rn(y, y);
          }
        }
      }
      __syncwarp();
      if (lane < 3) {
        const float* t = lane == 0 ? t0 : (lane == 1 ? t1 : t2);
        if (lane == 0 || metric == METRIC_COSINE) {
          for (int i = 0; i < n; ++i) s = __fadd_rn(s, t[i]);
        }
      }
      __syncwarp();
    }
    const float s0 = __shfl_sync(0xffffffffu, s, 0);
    const float s1 = __shfl_sync(0xffffffffu, s, 1);
    const float s2 = __shfl_sync(0xffffffffu, s, 2);
    if (metric == METRIC_L2) return (float)sqrt((double)s0);  // adapter.go:150
    if (metric == METRIC_DOT) return __fsub_rn(1.0f, s0);      // adapter.go:165
    if (s1 == 0.f || s2 == 0.f) return 1.0f;                   // adapter.go:121-123
    const float sa = (float)sqrt((double)s1), sb = (float)sqrt((double)s2);
    float sim = __fdiv_rn(s0, __fmul_rn(sa, sb));              // adapter.go:127
    if (sim > 1.0f) sim = 1.0f;
    else if (sim < -1.0f) sim = -1.0f;
    return __fsub_rn(1.0f, sim);
  }

  if (metric == METRIC_SQL2) {
    // distances.go:60-72 — float32 subtraction, float32 square, sequential float32 sum.
    float* t0 = reinterpret_cast<float*>(scratch);
    float s = 0.f;
    for (int base = 0; base < d; base += EXACT_CHUNK) {
      const int n = min(EXACT_CHUNK, d - base);
      for (int i = lane; i < n; i += 32) {
        const float df = __fsub_rn(a[base + i], b[base + i]);
        t0[i] = __fmul_rn(df, df);
      }
      __syncwarp();
      if (lane == 0)
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, t0[i]);
      __syncwarp();
    }
    return __shfl_sync(0xffffffffu, s, 0);
  }

  // float64 accumulators (distances.go:17-22, 48-52, 82-85, 99-101).
  double* t0 = scratch;
  double* t1 = t0 + EXACT_CHUNK;
  double* t2 = t1 + EXACT_CHUNK;
  double s = 0.0;  // lane 0: dot / sum; lane 1: magnitudeA; lane 2: magnitudeB
  for (int base = 0; base < d; base += EXACT_CHUNK) {
    const int n = min(EXACT_CHUNK, d - base);
    for (int i = lane; i < n; i += 32) {
      const float x = a[base + i], y = b[base + i];
      if (metric == METRIC_L2) {
        const double df = (double)__fsub_rn(x, y);  // float32 subtraction first, distances.go:50
        t0[i] = df * df;
      } else if (metric == METRIC_L1) {
        t0[i] = fabs((double)__fsub_rn(x, y));  // distances.go:100
      } else {
        t0[i] = (double)x * (double)y;
        if (metric == METRIC_COSINE) {
          t1[i] = (double)x * (double)x;
          t2[i] = (double)y * (double)y;
        }
      }
    }
    __syncwarp();
    if (lane < 3) {
      const double* t = lane == 0 ? t0 : (lane == 1 ? t1 : t2);
      if (lane == 0 || metric == METRIC_COSINE) {
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, t[i]);
      }
    }
    __syncwarp();
  }
  const double s0 = __shfl_sync(0xffffffffu, s, 0);
  const double s1 = __shfl_sync(0xffffffffu, s, 1);
  const double s2 = __shfl_sync(0xffffffffu, s, 2);
  if (metric == METRIC_L2) return (float)sqrt(s0);            // distances.go:54
  if (metric == METRIC_L1) return (float)s0;                  // distances.go:103
  if (metric == METRIC_DOT) return (float)(1.0 - s0);         // distances.go:89
  if (s1 == 0.0 || s2 == 0.0) return 1.0f;                    // distances.go:25-27
  double sim = s0 / (sqrt(s1) * sqrt(s2));                    // distances.go:30
  if (sim > 1.0) sim = 1.0;
  else if (sim < -1.0) sim = -1.0;
  return (float)(1.0 - sim);                                  // distances.go:39
}

// Up to three stored rows against one query in a single pass of the warp: the reference's float64
// sums are strictly sequential (one lane per sum, ~20 cycles per addition), so the three staging arrays
// that cosine needs for one row carry one row each for L2 / L1 / dot and lanes 0..2 run the three
// chains side by side. Same operation order per row as exact_distance_warp, hence the same bits.
// out[j] is valid on every lane for j < nb. Metrics / arithmetic modes that need more than one array per
// row (cosine, the float32 variants) fall back to one row at a time.
__device__ __forceinline__ void exact_distance_warp3(int metric, int arith, const float* __restrict__ a,
                                                     const float* const* b, int nb, int d, double* scratch,
                                                     float* out) {
  const bool one_array = arith != ARITH_HNSW_F32 && (metric == METRIC_L2 || metric == METRIC_L1 || metric == METRIC_DOT);
  if (!one_array || nb == 1) {
    for (int j = 0; j < nb; ++j) out[j] = exact_distance_warp(metric, arith, a, b[j], d, scratch);
    return;
  }
  const int lane = threadIdx.x & 31;
  double s = 0.0;  // lane j < nb: the sum of row j
  for (int base = 0; base < d; base += EXACT_CHUNK) {
    const int n = min(EXACT_CHUNK, d - base);
    for (int i = lane; i < n; i += 32) {
      const float x = a[base + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j >= nb) break;
        const float y = b[j][base + i];
        double t;
        if (metric == METRIC_L2) {
          const double df = (double)__fsub_rn(x, y);
          t = df * df;
        } else if (metric == METRIC_L1) {
          t = fabs((double)__fsub_rn(x, y));
        } else {
          t = (double)x * (double)y;
        }
        scratch[j * EXACT_CHUNK + i] = t;
      }
    }
    __syncwarp();
    if (lane < nb) {
      const double* t = scratch + lane * EXACT_CHUNK;
      for (int i = 0; i < n; ++i) s = __dadd_rn(s, t[i]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= nb) break;
    const double sj = __shfl_sync(0xffffffffu, s, j);
    out[j] = metric == METRIC_L2 ? (float)sqrt(sj) : (metric == METRIC_L1 ? (float)sj : (float)(1.0 - sj));
  }
}

// ---- one (query, row) distance by ONE thread, reference arithmetic and order ----------------------
// q: the query (shared memory, all lanes read the same element: broadcast), x: the stored row.
// Same value as exact_distance_warp(metric, arith, q, x, d): the terms are formed with the same
// roundings and added in index order.
__device__ __forceinline__ float exact_distance_lane(int metric, int arith, const float* __restrict__ q,
                                                     const float* __restrict__ x, int d, int dp) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int n4 = dp >> 2;  // rows are zero padded to a multiple of 4 floats; zero terms do not change a sum
  if (arith == ARITH_HNSW_F32 && (metric == METRIC_COSINE || metric == METRIC_L2 || metric == METRIC_DOT)) {
    // pkg/hnsw/adapter.go:105-167 — everything float32, sequential, no fused multiply-add
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (int i = 0; i < n4; ++i) {
      const float4 v = __ldg(x4 + i);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i * 4 + j >= d) break;
        const float a = q[i * 4 + j], b = xs[j];
        if (metric == METRIC_L2) {
          const float df = __fsub_rn(a, b);
          s0 = __fadd_rn(s0, __fmul_rn(df, df));
        } else {
          s0 = __fadd_rn(s0, __fmul_rn(a, b));
          if (metric == METRIC_COSINE) {
            s1 = __fadd_rn(s1, __fmul_rn(a, a));
            s2 = __fadd_rn(s2, __fmul_rn(b, b));
          }
        }
      }
    }
    if (metric == METRIC_L2) return (float)sqrt((double)s0);
    if (metric == METRIC_DOT) return __fsub_rn(1.0f, s0);
    if (s1 == 0.f || s2 == 0.f) return 1.0f;
    const float sa = (float)sqrt((double)s1), sb = (float)sqrt((double)s2);
    float sim = __fdiv_rn(s0, __fmul_rn(sa, sb));
    if (sim > 1.0f) sim = 1.0f;
    else if (sim < -1.0f) sim = -1.0f;
    return __fsub_rn(1.0f, sim);
  }
  if (metric == METRIC_SQL2) {  // distances.go:60-72
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < n4; ++i) {
      const float4 v = __ldg(x4 + i);
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i * 4 + j >= d) break;
        const float df = __fsub_rn(q[i * 4 + j], xs[j]);
        s = __fadd_rn(s, __fmul_rn(df, df));
      }
    }
    return s;
  }
  // float64 accumulators (distances.go:17-22, 48-52, 82-85, 99-101); a float32 product is exact in float64
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 4
  for (int i = 0; i < n4; ++i) {
    const float4 v = __ldg(x4 + i);
    const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i * 4 + j >= d) break;
      const float a = q[i * 4 + j], b = xs[j];
      if (metric == METRIC_L2) {
        const double df = (double)__fsub_rn(a, b);
        s0 = __dadd_rn(s0, __dmul_rn(df, df));
      } else if (metric == METRIC_L1) {
        s0 = __dadd_rn(s0, fabs((double)__fsub_rn(a, b)));
      } else {
        s0 = __dadd_rn(s0, __dmul_rn((double)a, (double)b));
        if (metric == METRIC_COSINE) {
          s1 = __dadd_rn(s1, __dmul_rn((double)a, (double)a));
          s2 = __dadd_rn(s2, __dmul_rn((double)b, (double)b));
        }
      }
    }
  }
  if (metric == METRIC_L2) return (float)sqrt(s0);
  if (metric == METRIC_L1) return (float)s0;
  if (metric == METRIC_DOT) return (float)(1.0 - s0);
  if (s1 == 0.0 || s2 == 0.0) return 1.0f;
  double sim = s0 / (sqrt(s1) * sqrt(s2));
  if (sim > 1.0) sim = 1.0;
  else if (sim < -1.0) sim = -1.0;
  return (float)(1.0 - sim);
}


}  // namespace qg
