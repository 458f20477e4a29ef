/*
 * oracle/synth.h — counter-based synthetic data generator (CPU side).
 *
 * TEST INFRASTRUCTURE. Element (row, col) of a synthetic matrix depends only on
 * (kind, seed, row, col), so numpy, this C file and the device generator in
 * quiver_b200/csrc/synth.cuh produce identical float32 bits without a shared
 * stream ("counter-based RNG ... so Python, C++ and CUDA produce
 * identical bits"). Only IEEE single operations with a fixed order are used.
 *
 *   kind 0  uniform [0,1)                      (hybrid_property_test.go:464-471 shape)
 *   kind 1  SIFT-like: floor(u * 218) as fp32  (config 2)
 *   kind 2  approx N(0,1): ((u0+u1)+(u2+u3) - 2) * sqrt(3)   (config 3)
 *   kind 3  kind 2, each row divided by its L2 norm (fp64 sequential)  (config 4)
 */
#ifndef QUIVER_ORACLE_SYNTH_H
#define QUIVER_ORACLE_SYNTH_H
#include <math.h>
#include <stdint.h>

static inline uint64_t qo_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* 24-bit uniform in [0,1): exactly representable in float32. */
static inline float qo_uniform(uint64_t seed, uint64_t row, uint32_t col, uint32_t lane) {
  uint64_t h = qo_mix64(seed ^ qo_mix64(row * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull));
  h = qo_mix64(h ^ (((uint64_t)col << 8) | lane));
  return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

static inline float qo_synth_raw(int kind, uint64_t seed, uint64_t row, uint32_t col) {
  if (kind == 0) return qo_uniform(seed, row, col, 0);
  if (kind == 1) return floorf(qo_uniform(seed, row, col, 0) * 218.0f);
  {
    float u0 = qo_uniform(seed, row, col, 0), u1 = qo_uniform(seed, row, col, 1);
    float u2 = qo_uniform(seed, row, col, 2), u3 = qo_uniform(seed, row, col, 3);
    float s = (u0 + u1) + (u2 + u3);
    s = s - 2.0f;
    return s * 1.7320508f;
  }
}

/* Fill one row of `dim` floats. */
static inline void qo_synth_row(int kind, uint64_t seed, uint64_t row, int dim, float* out) {
  int k = kind == 3 ? 2 : kind;
  for (int c = 0; c < dim; ++c) out[c] = qo_synth_raw(k, seed, row, (uint32_t)c);
  if (kind == 3) {
    double n2 = 0.0;
    for (int c = 0; c < dim; ++c) n2 += (double)out[c] * (double)out[c];
    if (n2 > 0.0) {
      double inv = sqrt(n2);
      for (int c = 0; c < dim; ++c) out[c] = (float)((double)out[c] / inv);
    }
  }
}
#endif
