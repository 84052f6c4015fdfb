// Philox4x32-10 counter-based generator (Salmon et al., SC'11) + Box-Muller normals.
// Replaces np.random.normal of the reference sampler (mcz.py:207-217, 517) and of the
// coefficient-noise draws (metscales.py:679-680, 948-953).  Counter layout:
//   c0 = sample index, c1 = call index, (c2, c3) = global galaxy index;  key = 64-bit seed.
// so results are independent of how galaxies are sharded over blocks / GPUs.
#pragma once
#include "mcz_common.cuh"

namespace mcz {

MCZ_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MCZ_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one 32x32->64 multiply per word pair (IMAD.WIDE), not separate high and low products
    const uint64_t p0 = (uint64_t)M0 * (uint64_t)c0, p1 = (uint64_t)M1 * (uint64_t)c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#if defined(__CUDACC__)
// Two standard normals from two 32-bit words.  u1 in [2^-24, 1-2^-24] (23 mantissa bits + half
// a step, exact in float), angle in [-pi, pi).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = __uint_as_float((a >> 9) | 0x3f800000u) - 0.99999994f;
  const float u2 = __uint_as_float((b >> 9) | 0x3f800000u) - 1.5f;        // [-0.5, 0.5)
  float r;                                                                       // sqrt(-2 ln u1)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(-1.3862943611198906f * __log2f(u1), 0.0f)));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  z0 = r * cs;
  z1 = r * sn;
}
#endif

}  // namespace mcz
