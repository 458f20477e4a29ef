// synth.cuh — device side of the counter-based synthetic generator (benchmarks and tests).
// Same specification as oracle/synth.h (an independent restatement on purpose: the parity
// tests check that both produce identical float32 bits). Only single IEEE operations in a
// fixed order are used, so host and device agree bit for bit.
#pragma once
#include <stdint.h>

namespace qg {

__host__ __device__ __forceinline__ uint64_t synth_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ float synth_uniform(uint64_t seed, uint64_t row, uint32_t col, uint32_t lane) {
  uint64_t h = synth_mix64(seed ^ synth_mix64(row * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull));
  h = synth_mix64(h ^ (((uint64_t)col << 8) | lane));
  return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

// kinds 0..2 (kind 3 = kind 2 followed by a per-row normalisation, done by the caller)
__device__ __forceinline__ float synth_raw(int kind, uint64_t seed, uint64_t row, uint32_t col) {
  if (kind == 0) return synth_uniform(seed, row, col, 0);
  if (kind == 1) return floorf(__fmul_rn(synth_uniform(seed, row, col, 0), 218.0f));
  const float u0 = synth_uniform(seed, row, col, 0), u1 = synth_uniform(seed, row, col, 1);
  const float u2 = synth_uniform(seed, row, col, 2), u3 = synth_uniform(seed, row, col, 3);
  float s = __fadd_rn(__fadd_rn(u0, u1), __fadd_rn(u2, u3));
  s = __fsub_rn(s, 2.0f);
  return __fmul_rn(s, 1.7320508f);
}

}  // namespace qg
