// Vocabulary shared by every kernel of the pyMCZ hot path: output keys, line order, scale tokens,
// per-galaxy line flags and the float/double math traits.
//
// Reference interfaces mirrored here (paths under /root/reference/pyMCZ):
//   * output key order            metallicity.py:29-56  (37 keys)
//   * line order ("alllines")     mcz.py:45             (11 lines)
//   * --md tokens                 metallicity.py:155-257
//   * has* flags                  metscales.py:284-384
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define MCZ_HD __host__ __device__ __forceinline__
#define MCZ_D __device__ __forceinline__
#else
#define MCZ_HD inline
#define MCZ_D inline
#endif

namespace mcz {

// ---- output keys (metallicity.py:29-56) -------------------------------------------------------
enum Key : int {
  K_EBV = 0, K_LOGR23, K_D02, K_Z94, K_M91, K_C01_N2S2, K_C01_R23, K_P05, K_P01, K_PP04_N2HA,
  K_PP04_O3N2, K_DP00, K_P10_ONS, K_P10_ON, K_M08_R23, K_M08_N2HA, K_M08_O3HB, K_M08_O2HB,
  K_M08_O3O2, K_M08_O3N2, K_M13_O3N2, K_M13_N2,
  K_D13_0, K_D13_1, K_D13_2, K_D13_3, K_D13_4, K_D13_5, K_D13_6, K_D13_7,
  K_KD02_N2O2, K_KD02_N2S2, K_KK04_N2HA, K_KK04_R23, K_KD02COMB, K_PM14, K_D16,
  NKEYS = 37
};

// Intermediate per-sample quantities the reference keeps as attributes of the diagnostics object
// (metscales.py:393-440, 468-489, 1112): handed to put() under indices >= NKEYS.  Only the replay
// kernel stores them (optional output); the production kernels' put() drops them at compile time.
enum AuxKey : int {
  A_LOGN2O2 = NKEYS,   // self.logN2O2 (metscales.py:404)
  A_LOGN2S2,           // self.logN2S2 (:396)
  A_N2S2,              // self.N2S2 (:395)
  A_R2, A_R3, A_R23,   // self.R2, self.R3, self.R23 (:431-433)
  A_ZINIT,             // self.Z_init_guess (:471-489)
  A_LOGQ,              // self.logq left by calcKK04_N2Ha (:1112, :1124)
  A_N2O2_ROOT_HI,      // the two real roots of the [NII]/[OII] quartic (of self.N2O2_roots, :423), NaN if complex
  A_N2O2_ROOT_LO,
  A_END,
  NAUX = A_END - NKEYS
};

// keys that are present but NaN for every sample (the reference's `highZ is True` tests never hold
// for a numpy.bool_, metscales.py:1013-1018, 1030-1047): never stored, reported as n = 0
MCZ_HD bool const_nan_key(int k) { return k == K_M08_R23 || k == K_M08_O3HB || k == K_M08_O2HB; }

// ---- input lines (mcz.py:45) ------------------------------------------------------------------
enum Line : int {
  L_O2 = 0, L_HB, L_O34959, L_O3, L_O1, L_HA, L_N2, L_S2A, L_S2B, L_S39069, L_S39532, NLINES = 11
};
// Lines the arithmetic consumes, in the order the sampler draws them.  [OIII]4959 is replaced by
// [OIII]5007/3 (metallicity.py:108), [OI]6300 is stored but never read (metscales.py:295) and
// [SIII]9532 only sets a flag (metscales.py:376-378), so their perturbed samples are dead values.
enum Used : int { U_HA = 0, U_HB, U_O2, U_O3, U_N2, U_S2A, U_S2B, U_S39069, NUSED = 8 };

// ---- per-galaxy line flags: any(sample > 0) (metscales.py:287-300, 346, 362-378) --------------
enum Flag : uint32_t {
  F_HA = 1u << 0, F_HB = 1u << 1, F_O2 = 1u << 2, F_O3 = 1u << 3, F_N2 = 1u << 4,
  F_S2 = 1u << 5,       // [SII]6717 > 0
  F_S2B = 1u << 6,      // [SII]6731 > 1e-9
  F_S39069 = 1u << 7,   // [SIII]9069 > 1e-9
  F_ALL = 0xffu
};

// ---- --md tokens (metallicity.py:161-257) -----------------------------------------------------
enum Token : uint32_t {
  T_D02 = 1u << 0, T_Z94 = 1u << 1, T_M91 = 1u << 2, T_PP04 = 1u << 3, T_P10 = 1u << 4,
  T_M08 = 1u << 5, T_M08ALL = 1u << 6, T_M13 = 1u << 7, T_KD02 = 1u << 8, T_P05 = 1u << 9,
  T_C01 = 1u << 10, T_P01 = 1u << 11, T_DP00 = 1u << 12, T_D16 = 1u << 13,
  T_ALL = T_D02 | T_Z94 | T_M91 | T_PP04 | T_P10 | T_M08 | T_M13 | T_KD02,  // metallicity.py:161-188
  // not a scale: the sanitise mode of metallicity.py:23-26, 98-102 rides in the token mask.  Set: the
  // reference's DROPNEGATIVES (a negative flux sample becomes NaN); clear: FIXNEGATIVES (it becomes 0).
  T_DROPNEG = 1u << 30
};

// dust request passed by the caller (metallicity.py:113-132)
enum Dust : int {
  DUST_NONE = 0,        // --nodust: corrections exactly 0 (metallicity.py:130-132)
  DUST_ON = 1,          // correct; a galaxy without Ha/Hb falls back to E=1e-5 *applied* (:116-129)
  DUST_ON_IGNORING = 2  // correct; sticky IGNOREDUST already set: fallback is "exactly 0" (:127,130)
};

// status codes of the percentile table (mcz.py:273-301)
enum Status : int8_t { ST_OK = 0, ST_ABSENT = 1, ST_EMPTY = 2, ST_NONPOS = 3 };

// CCM Rv=3.1 extinction constants (metscales.py:10-22)
#define MCZ_K_HA 2.535
#define MCZ_K_HB 3.609
#define MCZ_K_O2 4.771
#define MCZ_K_O35007 3.341
#define MCZ_K_O34959 3.384
#define MCZ_K_O3 ((MCZ_K_O35007 + MCZ_K_O34959) / 2.)
#define MCZ_K_N2 2.443
#define MCZ_K_S2 2.381
#define MCZ_K_S3 1.0

// ---- math traits ------------------------------------------------------------------------------
template <typename R> struct Math;

template <> struct Math<double> {
  static MCZ_HD double log10(double x) { return ::log10(x); }
  static MCZ_HD double exp10(double x) { return ::pow(10.0, x); }
  static MCZ_HD double sqrt(double x) { return ::sqrt(x); }
  static MCZ_HD double div(double a, double b) { return a / b; }
  static MCZ_HD double abs(double x) { return ::fabs(x); }
  static MCZ_HD double pow(double a, double b) { return ::pow(a, b); }
  static MCZ_HD double nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
  }
  static MCZ_HD bool isnan(double x) { return x != x; }
  static MCZ_HD bool isfinite(double x) { return ::fabs(x) <= 1.7976931348623157e308; }
};

template <> struct Math<float> {
  // float32 is the production arithmetic: MUFU-backed log2/exp2 scaled to base 10.
  static MCZ_HD float log10(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // MUFU.LG2 alone: no denormal rescaling (fluxes are O(1))
    return r * 0.30102999566398120f;
#else
    return ::log10f(x);
#endif
  }
  static MCZ_HD float exp10(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 3.3219280948873623f));   // MUFU.EX2, 2 ulp
    return r;
#else
    return ::powf(10.0f, x);
#endif
  }
  static MCZ_HD float sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));                        // MUFU.SQRT, 1 ulp
    return r;
#else
    return ::sqrtf(x);
#endif
  }
  static MCZ_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
  }
  static MCZ_HD float abs(float x) { return ::fabsf(x); }
  static MCZ_HD float pow(float a, float b) {
#if defined(__CUDA_ARCH__)
    return exp2f(b * __log2f(a));
#else
    return ::powf(a, b);
#endif
  }
  static MCZ_HD float nan() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7fc00000);
#else
    return NAN;
#endif
  }
  static MCZ_HD bool isnan(float x) { return x != x; }
  static MCZ_HD bool isfinite(float x) { return ::fabsf(x) <= 3.402823466e38f; }
};

// Horner evaluation of an ascending-coefficient polynomial (numpy.polynomial.polynomial.polyval).
template <typename R> MCZ_HD R poly2(R x, double c0, double c1, double c2) {
  return R(c0) + x * (R(c1) + x * R(c2));
}
template <typename R> MCZ_HD R poly3(R x, double c0, double c1, double c2, double c3) {
  return R(c0) + x * (R(c1) + x * (R(c2) + x * R(c3)));
}
template <typename R> MCZ_HD R poly4(R x, double c0, double c1, double c2, double c3, double c4) {
  return R(c0) + x * (R(c1) + x * (R(c2) + x * (R(c3) + x * R(c4))));
}

}  // namespace mcz
