// Counter-based random numbers: Philox4x32-10 keyed by (seed, frame, site, stream) and a
// Box-Muller transform.  Every value is a pure function of its key, so any frame range produced
// by any rank, slab or chunking is identical.
#pragma once
#include "common.cuh"

namespace agf {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t frame, uint32_t site, uint32_t stream,
                                           uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)frame, (uint32_t)(frame >> 32), site, stream};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
  out[3] = c[3];
}

// three standard normals from one Philox block (two Box-Muller pairs, one value unused)
__device__ __forceinline__ void normal3(uint64_t seed, uint64_t frame, uint32_t site, uint32_t stream, float (&z)[3]) {
  uint32_t r[4];
  philox4x32(seed, frame, site, stream, r);
  const float u0 = ((float)r[0] + 0.5f) * 2.3283064365386963e-10f;
  const float u1 = ((float)r[1] + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r[2] + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = ((float)r[3] + 0.5f) * 2.3283064365386963e-10f;
  const float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  z[0] = m0 * c0;
  z[1] = m0 * s0;
  z[2] = m1 * c1;
  (void)s1;
}

}  // namespace agf
