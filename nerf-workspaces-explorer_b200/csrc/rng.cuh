// Counter-based random numbers (Philox4x32-10) for the training path's three draws: stratified
// jitter (training handler:560), importance-sampling uniforms (rays.py:98) and sigma noise
// (model_utils.py:65).  The reference draws them with torch.rand/randn (on the CPU, then copies);
// here each kernel derives its value from (seed, offset, stream, element index), so nothing is
// stored in HBM and the compositing backward regenerates exactly the noise its forward used.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nwx {

struct RngSpec {
  uint64_t seed = 0, offset = 0;   // offset: e.g. the optimiser step, so every step draws fresh numbers
  uint32_t stream = 0;             // 0 jitter, 1 importance u, 2 noise (coarse), 3 noise (fine)
  float scale = 1.0f;              // raw_noise_std for the normal streams
  int on = 0;
};

__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}

// 4 x 32 random bits for element `idx` of (seed, offset, stream)
__host__ __device__ inline void philox4x32_10(uint64_t seed, uint64_t offset, uint32_t stream, uint64_t idx, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream, (uint32_t)offset};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// U[0,1) with 24-bit resolution, like torch.rand for float32
__device__ __forceinline__ float rng_uniform(const RngSpec& r, uint64_t idx) {
  uint32_t x[4];
  philox4x32_10(r.seed, r.offset, r.stream, idx, x);
  return (float)(x[0] >> 8) * (1.0f / 16777216.0f);
}
// N(0,1) * scale by Box-Muller
__device__ __forceinline__ float rng_normal(const RngSpec& r, uint64_t idx) {
  uint32_t x[4];
  philox4x32_10(r.seed, r.offset, r.stream, idx, x);
  const float u1 = (float)((x[0] >> 8) + 1u) * (1.0f / 16777216.0f);      // (0,1]
  const float u2 = (float)(x[1] >> 8) * (1.0f / 16777216.0f);
  return r.scale * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace nwx
