// Per-sample physics of the pyMCZ hot path, written once for float (production) and double
// (replay/parity) arithmetic.  Everything here is a pure function of ONE Monte-Carlo sample plus
// per-galaxy flags; the per-galaxy collectives (has* flags, KK04 convergence means, percentile
// selection) live in the kernels that call these functions.
//
// Reference (paths under /root/reference/pyMCZ): metallicity.py:86-257 (dispatch, sanitising),
// metscales.py:273-489 (dust, ratios, R23, initial guess), :674-1264 (scales).  The arithmetic is
// restated in log space: every dust-corrected log ratio of the reference is
//   log10(F1/F2 * 10^(0.4 E (k1-k2))) = log10 F1 - log10 F2 + 0.4 E (k1-k2),
// so one log10 per line (+ one for [SII] sum, one for R23) replaces the reference's ~13 log10 and
// 7 pow calls per sample.  IEEE special values (0 -> -inf, 0/0 -> NaN) propagate identically.
#pragma once
#include "mcz_common.cuh"
#include "mcz_root_tables.h"

namespace mcz {

// ---- per-galaxy logic -------------------------------------------------------------------------

// metscales.py:233-246 with the dust branch of metallicity.py:113-132 folded in: by the time
// checkminimumreq runs, red_corr is only still True when Ha and Hb are both present.
MCZ_HD bool minimum_ok(uint32_t f) {
  return (f & F_N2) || ((f & F_O2) && (f & F_O3));
}

// 0: E(B-V) measured per sample; 1: E=1e-5 and corrections applied; 2: E=1e-5 reported, no correction
MCZ_HD int effective_dust(int dust, uint32_t f) {
  if (dust != DUST_NONE && (f & F_HA) && (f & F_HB)) return 0;   // metallicity.py:113-115
  if (dust == DUST_ON) return 1;                                  // metallicity.py:116-129
  return 2;                                                       // metallicity.py:130-132
}

// Which keys are not None for a galaxy (the `present` table).  One bit per Key.
MCZ_HD uint64_t present_mask(uint32_t f, uint32_t tok) {
  uint64_t m = 1ull << K_EBV;
  if (!minimum_ok(f)) return m;                                   // metallicity.py:145-146
  const bool ha = f & F_HA, hb = f & F_HB, o2 = f & F_O2, o3 = f & F_O3, n2 = f & F_N2;
  const bool s2 = f & F_S2, s39 = f & F_S39069;
  const bool r23 = o3 && o2 && hb;                                // metscales.py:430
  const bool o3o2 = o2 && o3;                                     // metscales.py:334-341
  const bool o3hb = hb && o3;                                     // metscales.py:326-331
  auto set = [&](int k, bool c) { if (c) m |= 1ull << k; };
  set(K_LOGR23, r23);
  set(K_D02, (tok & T_D02) && n2 && ha);                          // metscales.py:681
  set(K_Z94, (tok & T_Z94) && r23);                               // metscales.py:718-724
  set(K_M91, (tok & (T_M91 | T_KD02)) != 0);                      // metscales.py:915 (set before checks), :1208
  set(K_PP04_N2HA, (tok & T_PP04) && n2 && ha);                   // metscales.py:694
  set(K_PP04_O3N2, (tok & T_PP04) && n2 && ha && o3hb);           // metscales.py:700
  set(K_P10_ONS, (tok & T_P10) && hb);                            // metscales.py:774-778
  set(K_P10_ON, (tok & T_P10) && hb);
  const bool m08 = tok & (T_M08 | T_M08ALL);
  set(K_M08_O3O2, m08 && o3o2);                                   // metscales.py:975
  set(K_M08_N2HA, m08 && n2 && ha);                               // metscales.py:993
  set(K_M08_R23, m08 && r23);                                     // metscales.py:1003-1009
  const bool m08all = (tok & T_M08ALL) != 0;
  set(K_M08_O3HB, m08all && o3hb);                                // metscales.py:1024 (logO3Hb is not None)
  set(K_M08_O2HB, m08all && o2 && hb);                            // metscales.py:1036 (logO2Hb is not None)
  set(K_M08_O3N2, m08all && o3 && n2);                            // metscales.py:1051
  set(K_M13_N2, (tok & T_M13) && ha && n2);                       // metscales.py:943-950
  set(K_M13_O3N2, (tok & T_M13) && ha && n2 && hb && o3);         // metscales.py:951
  set(K_KD02_N2O2, (tok & T_KD02) && n2 && o2 && ha && hb);       // metscales.py:1076
  set(K_KK04_N2HA, (tok & T_KD02) && n2 && ha);                   // metscales.py:1105
  set(K_KK04_R23, (tok & T_KD02) && o3o2 && r23);                 // metscales.py:1151-1162
  set(K_KD02COMB, (tok & T_KD02) != 0);                           // metscales.py:1250
  set(K_P05, (tok & T_P05) && r23);                               // metscales.py:750
  set(K_P01, (tok & T_P01) && o3o2 && r23);                       // metscales.py:852-859
  set(K_C01_R23, (tok & T_C01) && o3 && o2 && o3hb);              // metscales.py:874
  set(K_C01_N2S2, (tok & T_C01) && s2 && n2 && o3 && o2 && o3hb); // metscales.py:890
  set(K_DP00, (tok & T_DP00) && s2 && s39 && hb);                 // metscales.py:448-449, 666-671
  set(K_D16, (tok & T_D16) && ha && n2 && s2);                    // metscales.py:1347
  return m;
}

// ---- closed-form roots (replacing fz_roots/np.roots, metscales.py:248-263) ---------------------
#define MCZ_RT_X(v) v,

template <typename R, int N> MCZ_HD R horner_desc(const double (&c)[N], R u) {
  R acc = R(c[N - 1]);
#pragma unroll
  for (int i = N - 2; i >= 0; --i) acc = acc * u + R(c[i]);
  return acc;
}

template <typename R> struct IsDouble { static const bool value = false; };
template <> struct IsDouble<double> { static const bool value = true; };

// two Newton steps on phi(t) = t*sqrt(h0+h1 t+h2 t^2) - s  (double only: 1e-8 -> < 1e-15)
template <typename R> MCZ_HD R rt_polish(R t, R s, double h0, double h1, double h2) {
  if (IsDouble<R>::value) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      R h = R(h0) + t * (R(h1) + t * R(h2));
      R q = Math<R>::sqrt(h);
      R dphi = q + t * (R(h1) + R(2.0 * h2) * t) / (R(2.0) * q);
      t = t - (t * q - s) / dphi;
    }
  }
  return t;
}

// t(s) for the KD02 [NII]/[OII] quartic (metscales.py:419-423)
template <typename R> MCZ_HD R rt_n2o2_t(R s) {
  R t;
  if (s < R(RT_N2O2_SEG0_HI)) {
    const double c[] = {RT_N2O2_SEG0_POLY(MCZ_RT_X)};
    t = horner_desc<R>(c, (s - R(RT_N2O2_SEG0_MID)) * R(RT_N2O2_SEG0_RHALF));
  } else {
    const double c[] = {RT_N2O2_SEG1_POLY(MCZ_RT_X)};
    t = horner_desc<R>(c, (s - R(RT_N2O2_SEG1_MID)) * R(RT_N2O2_SEG1_RHALF));
  }
  return rt_polish<R>(t, s, RT_N2O2_H0, RT_N2O2_H1, RT_N2O2_H2);
}

// t(s) for the M08 [NII]/Ha quartic (metscales.py:995-997)
template <typename R> MCZ_HD R rt_m08n2_t(R s) {
  R t;
  if (s < R(RT_M08N2_SEG1_HI)) {
    if (s < R(RT_M08N2_SEG0_HI)) {
      const double c[] = {RT_M08N2_SEG0_POLY(MCZ_RT_X)};
      t = horner_desc<R>(c, (s - R(RT_M08N2_SEG0_MID)) * R(RT_M08N2_SEG0_RHALF));
    } else {
      const double c[] = {RT_M08N2_SEG1_POLY(MCZ_RT_X)};
      t = horner_desc<R>(c, (s - R(RT_M08N2_SEG1_MID)) * R(RT_M08N2_SEG1_RHALF));
    }
  } else {
    if (s < R(RT_M08N2_SEG2_HI)) {
      const double c[] = {RT_M08N2_SEG2_POLY(MCZ_RT_X)};
      t = horner_desc<R>(c, (s - R(RT_M08N2_SEG2_MID)) * R(RT_M08N2_SEG2_RHALF));
    } else {
      const double c[] = {RT_M08N2_SEG3_POLY(MCZ_RT_X)};
      t = horner_desc<R>(c, (s - R(RT_M08N2_SEG3_MID)) * R(RT_M08N2_SEG3_RHALF));
    }
  }
  return rt_polish<R>(t, s, RT_M08N2_H0, RT_M08N2_H1, RT_M08N2_H2);
}

// KD02 [NII]/[OII] abundance (metscales.py:1085-1087): real roots r with |r| in [7.5, 9.4],
// last match in np.roots order wins.  The quartic has one real minimum f0 at x0; for L >= f0 it
// has two real roots x0 + t(-s) < x0 < x0 + t(+s), s = sqrt(L - f0).  np.roots lists the pair in
// descending order, so when both are inside the window (f0 <= L <= f(7.5)) the smaller one wins;
// above that only the larger is inside until L > f(9.4).  Non-finite L: the constant coefficient
// is zeroed (metscales.py:259) and no root of x*(cubic) is inside the window -> NaN.
template <typename R> MCZ_HD R kd02_n2o2(R L) {
  if (!Math<R>::isfinite(L)) return Math<R>::nan();
  const R d = L - R(RT_N2O2_F0);
  if (!(d >= R(0)) || L > R(RT_N2O2_F_HI)) return Math<R>::nan();
  const R s = Math<R>::sqrt(d);
  const R t = rt_n2o2_t<R>(L <= R(RT_N2O2_F_LO) ? -s : s);
  return R(RT_N2O2_X0) + t;
}
// the real root pair of that quartic (both NaN when it has none): part of self.N2O2_roots
template <typename R> MCZ_HD void kd02_n2o2_real_roots(R L, R& hi, R& lo) {
  hi = lo = Math<R>::nan();
  if (!Math<R>::isfinite(L)) return;
  const R d = L - R(RT_N2O2_F0);
  if (!(d >= R(0))) return;
  const R s = Math<R>::sqrt(d);
  // (the fitted t(s) covers the window of KD02_N2O2 only: outside it the roots are not reported)
  if (L > R(RT_N2O2_F_HI)) return;
  hi = R(RT_N2O2_X0) + rt_n2o2_t<R>(s);
  // The partner root lies below the window for most L (t(-s) is fitted for -s >= -0.39 only): Newton
  // on t sqrt(h(t)) = -s from the fitted value inside the fit's domain, from -s / sqrt(h(0)) outside.
  R t = (-s >= R(RT_N2O2_SEG0_MID) - R(1.0) / R(RT_N2O2_SEG0_RHALF)) ? rt_n2o2_t<R>(-s) : -s / Math<R>::sqrt(R(RT_N2O2_H0));
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const R h = R(RT_N2O2_H0) + t * (R(RT_N2O2_H1) + t * R(RT_N2O2_H2));
    const R q = Math<R>::sqrt(h);
    const R dphi = q + t * (R(RT_N2O2_H1) + R(2.0 * RT_N2O2_H2) * t) / (R(2.0) * q);
    t = t - (t * q + s) / dphi;
  }
  lo = R(RT_N2O2_X0) + t;
}

// M08 [NII]/Ha (metscales.py:994-999): first root (np.roots order) with root+8.69 in [7.1, 9.4].
// One real maximum f0 at x0; for y <= f0 two real roots, listed in descending order, so the
// larger wins while it is inside the window (y >= f(0.71)), else the smaller (y >= f(-1.59)).
// Non-finite y: constant coefficient zeroed -> root x=0 -> exactly 8.69.
template <typename R> MCZ_HD R m08_n2ha(R y) {
  if (!Math<R>::isfinite(y)) return R(0.0) + R(8.69);
  const R d = R(RT_M08N2_F0) - y;
  if (!(d >= R(0)) || y < R(RT_M08N2_F_LO)) return Math<R>::nan();
  const R s = Math<R>::sqrt(d);
  const R t = rt_m08n2_t<R>(y >= R(RT_M08N2_F_HI) ? s : -s);
  return (R(RT_M08N2_X0) + t) + R(8.69);
}

// M08 [OIII]/[OII] (metscales.py:976-991): quadratic -0.3172 x^2 - 1.3881 x - 0.2839 - y = 0;
// only the larger root can fall in the window; written in the cancellation-free form.
template <typename R> MCZ_HD R m08_o3o2(R y) {
  if (!Math<R>::isfinite(y)) return R(0.0) + R(8.69);
  const R cp = R(-0.2839) - y;
  const R D = R(1.3881 * 1.3881) + R(4.0 * 0.3172) * cp;
  if (!(D >= R(0))) return Math<R>::nan();
  const R x = Math<R>::div(R(2.0) * cp, Math<R>::sqrt(D) + R(1.3881));
  const R Z = x + R(8.69);
  return (Z >= R(7.1) && Z <= R(9.4)) ? Z : Math<R>::nan();
}

// t(s) for the M08 [OIII]/[NII] cubic (metscales.py:1051-1057)
template <typename R> MCZ_HD R rt_m08o3n2_t(R s) {
  R t;
  if (s < R(RT_M08O3N2_SEG0_HI)) {
    const double c[] = {RT_M08O3N2_SEG0_POLY(MCZ_RT_X)};
    t = horner_desc<R>(c, (s - R(RT_M08O3N2_SEG0_MID)) * R(RT_M08O3N2_SEG0_RHALF));
  } else if (s < R(RT_M08O3N2_SEG1_HI)) {
    const double c[] = {RT_M08O3N2_SEG1_POLY(MCZ_RT_X)};
    t = horner_desc<R>(c, (s - R(RT_M08O3N2_SEG1_MID)) * R(RT_M08O3N2_SEG1_RHALF));
  } else {
    const double c[] = {RT_M08O3N2_SEG2_POLY(MCZ_RT_X)};
    t = horner_desc<R>(c, (s - R(RT_M08O3N2_SEG2_MID)) * R(RT_M08O3N2_SEG2_RHALF));
  }
  return rt_polish<R>(t, s, RT_M08O3N2_H0, RT_M08O3N2_H1, RT_M08O3N2_H2);
}

// M08 [OIII]/[NII] of --md M08all (metscales.py:1051-1059): cubic 0.1347 x^3 - 0.7170 x^2 -
// 2.6096 x + 0.4520 - y = 0, first root (np.roots order) with root+8.69 in [7.1, 9.4].  Inside the
// window the cubic has one local maximum f0 at x0 = -1.325; its other extremum (x = 4.87) and the
// largest root are far outside.  np.roots lists [largest, smallest, middle], so the smallest root
// x0 + t(-s) wins once it is inside the window (y >= f(-1.59)), else the middle one x0 + t(+s)
// (inside while y >= f(0.71)).  Non-finite y: constant coefficient zeroed -> root 0 -> 8.69.
template <typename R> MCZ_HD R m08_o3n2(R y) {
  if (!Math<R>::isfinite(y)) return R(0.0) + R(8.69);
  const R d = R(RT_M08O3N2_F0) - y;
  if (!(d >= R(0)) || y < R(RT_M08O3N2_F_HI)) return Math<R>::nan();
  const R s = Math<R>::sqrt(d);
  const R t = rt_m08o3n2_t<R>(y >= R(RT_M08O3N2_F_LO) ? -s : s);
  return (R(RT_M08O3N2_X0) + t) + R(8.69);
}

// ---- per-sample state carried into the KK04 iterations ----------------------------------------
template <typename R> struct IterState {
  R num0, num1, den0, den1;   // calclogq (metscales.py:464-465): logq = (num0+Z num1)/(den0+Z den1)
  R A, B;                     // KK04_N2Ha update (metscales.py:1114-1116): Z = A - logq*B
  R U0, U1, L0, L1;           // KK04_R23 branches (metscales.py:1172-1176)
  R Z1, Z2;                   // running Z of the two iterations
  R lq1, lq2;                 // last logq of the two iterations
  uint32_t aux;               // bits 0-1: Z_init_guess class (0: 8.2, 1: 8.4, 2: 8.7); bit 2: logN2Ha > 0.8
};

enum : uint32_t { AUX_ZI_MASK = 3u, AUX_N2HA_GT08 = 4u, AUX_KD_VALID = 8u /* KD02_N2O2 is not NaN */ };

template <typename R> MCZ_HD R calclogq(const IterState<R>& st, R Z) {
  return Math<R>::div(st.num0 + Z * st.num1, st.den0 + Z * st.den1);
}

// one trip of the KK04_N2Ha loop body (metscales.py:1112-1118); returns |logq - logq_save|
template <typename R> MCZ_HD R kk04_n2ha_trip(IterState<R>& st) {
  const R lq = calclogq<R>(st, st.Z1);
  st.Z1 = st.A - lq * st.B;
  const R d = Math<R>::abs(lq - st.lq1);
  st.lq1 = lq;
  return d;
}

// one trip of the KK04_R23 loop body (metscales.py:1167-1178); returns logqold - logq
template <typename R> MCZ_HD R kk04_r23_trip(IterState<R>& st) {
  const R lq = calclogq<R>(st, st.Z2);
  // Zmax = 8.4 where 6.7 <= logq < 8.3 else 0; lower branch where Z_init_guess <= Zmax
  const bool lower = ((st.aux & AUX_ZI_MASK) <= 1u) && (lq >= R(6.7)) && (lq < R(8.3));
  st.Z2 = lower ? (st.L0 - lq * st.L1) : (st.U0 - lq * st.U1);
  const R d = st.lq2 - lq;
  st.lq2 = lq;
  return d;
}

// What phase A hands to the iteration phase for one sample (16 bytes in float).
template <typename R> struct Carry {
  R y;         // logO3O2 (metscales.py:338-339)
  R n2ha;      // logN2Ha (metscales.py:350)
  R x;         // logR23  (metscales.py:434)
  uint32_t aux;
};

// The loop constants in three groups, for callers that recompute them per batch and need only the
// groups of the loops still running (same expressions as make_iter below, which calls them).
template <typename R> MCZ_HD void iter_logq(R y, IterState<R>& st) {       // calclogq, metscales.py:464-465
  const R y2 = y * y;
  st.num0 = R(32.81) - R(1.153) * y2;
  st.num1 = R(-3.396) - R(0.025) * y + R(0.1444) * y2;
  st.den0 = R(4.603) - R(0.3119) * y - R(0.163) * y2;
  st.den1 = R(-0.48) + R(0.0271) * y + R(0.02037) * y2;
}
template <typename R> MCZ_HD void iter_n2ha(R n, IterState<R>& st) {       // metscales.py:1114-1116
  st.A = poly3<R>(n, 7.04, 5.28, 6.28, 2.37);
  st.B = poly3<R>(n, -2.44, -2.01, -0.325, 0.128) - Math<R>::exp10(n - R(0.2)) * (R(-3.16) + R(4.65) * n);
}
template <typename R> MCZ_HD void iter_r23(R x, IterState<R>& st) {        // metscales.py:1172-1176
  st.U0 = poly4<R>(x, 9.72, -0.777, -0.951, -0.072, -0.811);
  st.U1 = poly4<R>(x, 0.0737, -0.0713, -0.141, 0.0373, -0.058);
  st.L0 = poly2<R>(x, 9.40, 4.65, -3.17);
  st.L1 = poly2<R>(x, 0.272, 0.547, -0.513);
}

// Build the loop constants from the carry; kd = this sample's KD02_N2O2 (NaN when absent).
template <typename R> MCZ_HD void make_iter(const Carry<R>& c, R kd, IterState<R>& st) {
  iter_logq<R>(c.y, st);
  iter_n2ha<R>(c.n2ha, st);
  iter_r23<R>(c.x, st);
  const uint32_t zi = c.aux & AUX_ZI_MASK;
  st.Z1 = kd;        // the kernel overrides with 8.6 when KD02_N2O2 is absent / all-NaN (:1099-1103)
  st.Z2 = zi == 0u ? R(8.2) : (zi == 1u ? R(8.4) : R(8.7));               // metscales.py:1156
  st.lq1 = R(0);                                                          // logq_save (:1106)
  st.lq2 = R(100);                                                        // logqold (:1163)
  st.aux = c.aux;
}

// ---- phase A: everything that needs no cross-sample information -------------------------------
// f[NUSED]: sanitised fluxes (non-finite -> 0, negative -> 0; metallicity.py:96-99) in `Used` order
// e[5]:     scaled coefficient noise: D02 e1,e2 (metscales.py:679-680), M13_N2 e1,e2 (:948-949),
//           M13_O3N2 e1 (:952; its e2 is drawn but never used, :953-955)
// put(key, value) is called for every key that is present for this galaxy except the three KK04
// iteration outputs, which phase C writes; with Put::WANT_AUX also for the AuxKey intermediates.  `carry` is filled when the KD02 group is requested.
template <typename R, class Put>
MCZ_HD void phase_a(const R* f, const R* e, uint32_t flags, uint32_t tok, int dust_eff,
                    Put& put, Carry<R>& carry) {
  typedef Math<R> M;
  const R nan = M::nan();
  const bool ha = flags & F_HA, hb = flags & F_HB, o2 = flags & F_O2, o3 = flags & F_O3;
  const bool n2 = flags & F_N2, s2 = flags & F_S2, s2b = flags & F_S2B, s39 = flags & F_S39069;
  const bool r23ok = o3 && o2 && hb, o3o2 = o2 && o3, o3hb = hb && o3;

  const R lHa = M::log10(f[U_HA]), lHb = M::log10(f[U_HB]);
  // E(B-V) (metscales.py:389-390) and a = 0.4 E (metscales.py:277)
  R E, a;
  if (dust_eff == 0) {
    E = M::div(R(0.45636603312904175) + lHb - lHa, R(0.4 * (MCZ_K_HA - MCZ_K_HB)));   // log10(2.86)
    if (E <= R(0)) E = R(1e-5);
    a = R(0.4) * E;
  } else {
    E = R(1e-5);
    a = dust_eff == 1 ? R(0.4 * 1e-5) : R(0);
  }
  put(K_EBV, E);
  if (!minimum_ok(flags)) return;

  const R lO2 = M::log10(f[U_O2]), lO3 = M::log10(f[U_O3]), lN2 = M::log10(f[U_N2]);
  const R L43 = R(0.12493873660829995);                       // log10(4/3): [OIII]4959+5007 = 5007*(1+1/3)
  const R logN2Ha = lN2 - lHa;                                 // metscales.py:350 (no dust)
  const R logO35007O2 = lO3 - lO2 + a * R(MCZ_K_O3 - MCZ_K_O2);          // metscales.py:303-306
  const R y = logO35007O2 + L43;                               // logO3O2, metscales.py:338-339
  const R logO3Hb = lO3 - lHb + a * R(MCZ_K_O35007 - MCZ_K_HB);          // metscales.py:327-330
  const R logN2O2 = lN2 - lO2 + a * R(MCZ_K_N2 - MCZ_K_O2);              // metscales.py:404
  // R23 (metscales.py:431-435)
  const R lr2 = lO2 - lHb + a * R(MCZ_K_O2 - MCZ_K_HB);
  const R lr3 = lO3 + L43 - lHb + a * R(MCZ_K_O3 - MCZ_K_HB);
  R R2 = nan, R3 = nan, R23 = nan, x = nan;
  if (r23ok) {
    R2 = M::exp10(lr2);
    R3 = M::exp10(lr3);
    R23 = R2 + R3;
    x = M::log10(R23);
    put(K_LOGR23, x);
    if (Put::WANT_AUX) { put(A_R2, R2); put(A_R3, R3); put(A_R23, R23); }
  }
  if (Put::WANT_AUX && n2 && o2) {                                        // calcNIIOII, metscales.py:402-423
    put(A_LOGN2O2, logN2O2);
    R rhi, rlo;
    kd02_n2o2_real_roots<R>(logN2O2, rhi, rlo);
    put(A_N2O2_ROOT_HI, rhi); put(A_N2O2_ROOT_LO, rlo);
  }
  if (Put::WANT_AUX && s2 && n2) {                                        // calcNIISII, metscales.py:393-396
    const R lns = lN2 - M::log10(f[U_S2A]) + a * R(MCZ_K_N2 - MCZ_K_S2);
    put(A_LOGN2S2, lns);
    put(A_N2S2, M::exp10(lns));
  }

  // Z_init_guess (metscales.py:471-489)
  uint32_t zi = 2u;
  if (ha && n2 && f[U_N2] != R(0)) {
    if (logN2Ha < R(-1.3)) zi = 0u;
    else if (logN2Ha < R(-1.1)) zi = 1u;
    else if (logN2Ha >= R(-1.1)) zi = 2u;
  }
  if (n2 && o2) {
    const bool nz = hb ? (f[U_N2] * f[U_HA] * f[U_HB] * f[U_O2] != R(0)) : true;
    if (nz) {
      if (logN2O2 < R(-1.2)) zi = 0u;
      else if (logN2O2 >= R(-1.2)) zi = 2u;
    }
  }
  const bool zi_lt84 = zi == 0u;      // Z_init_guess < 8.4
  if (Put::WANT_AUX) put(A_ZINIT, zi == 0u ? R(8.2) : (zi == 1u ? R(8.4) : R(8.7)));

  if ((tok & T_D02) && n2 && ha)                                          // metscales.py:682
    put(K_D02, R(9.12) + e[0] + (R(0.73) + e[1]) * logN2Ha);

  if ((tok & T_Z94) && r23ok) {                                           // metscales.py:724-725
    R v = poly4<R>(x, 9.265, -0.33, -0.202, -0.207, -0.333);
    if (x > R(0.9)) v = nan;
    put(K_Z94, v);
  }

  R m91 = nan;
  if (tok & (T_M91 | T_KD02)) {                                           // metscales.py:915-936
    if (r23ok) {
      const R low = poly2<R>(x, 12.0 - 4.944, 0.767, 0.602) - y * poly2<R>(x, 0.29, 0.332, -0.331);
      const R up = poly4<R>(x, 12.0 - 2.939, -0.2, -0.237, -0.305, -0.0283) -
                   y * poly4<R>(x, 0.0047, -0.0221, -0.102, -0.0817, -0.00717);
      if (M::abs(y) > R(0) && M::abs(x) > R(0)) m91 = zi_lt84 ? low : up;
      if (up < low) m91 = nan;
    }
    put(K_M91, m91);
  }

  if ((tok & T_PP04) && n2 && ha) {                                       // metscales.py:695-703
    R v = poly3<R>(logN2Ha, 9.37, 2.03, 1.26, 0.32);
    if (!(logN2Ha > R(-2.5) && logN2Ha < R(-0.3))) v = nan;
    put(K_PP04_N2HA, v);
    if (o3hb) {
      R w = R(8.73) - R(0.32) * (logO3Hb - logN2Ha);
      if (logO3Hb > R(2)) w = nan;
      put(K_PP04_O3N2, w);
    }
  }

  if ((tok & T_P10) && hb) {                                              // metscales.py:777-841
    const R LN10 = R(2.302585092994046);
    const R P = r23ok ? M::div(R3, R23) : nan;                            // metscales.py:740
    const R lR2 = r23ok ? LN10 * lr2 : nan;                               // np.log (sic) :788
    const R lR3 = r23ok ? LN10 * lr3 : nan;                               // np.log (sic) :791
    const R lN2p = n2 ? LN10 * (lN2 + R(0.12385164096708581) - lHb) + a * R(MCZ_K_N2 - MCZ_K_HB) : nan;  // :798, log10(1.33)
    R lS2 = nan;
    if (s2 && s2b) lS2 = M::log10(f[U_S2A] + f[U_S2B]) - lHb + a * R(MCZ_K_S2 - MCZ_K_HB);              // :802-805
    const R lN2S2 = lN2p - lS2, lN2R2 = lN2p - lR2, lS2R2 = lS2 - lR2;
    R ons = nan, on = nan;
    if (lN2p > R(-0.1)) {
      ons = R(8.277) + R(0.657) * P + R(-0.399) * lR3 + R(-0.061) * lN2R2 + R(0.005) * lS2R2;
      on = R(8.606) + R(-0.105) * lR3 + R(-0.410) * lR2 + R(-0.150) * lN2R2;
    } else if (lN2p < R(-0.1) && lN2S2 > R(-0.25)) {
      ons = R(8.816) + R(-0.733) * P + R(0.454) * lR3 + R(0.710) * lN2R2 + R(-0.337) * lS2R2;
      on = R(8.642) + R(0.077) * lR3 + R(0.411) * lR2 + R(0.601) * lN2R2;
    } else if (lN2p < R(-0.1) && lN2S2 < R(-0.25)) {
      ons = R(8.774) + R(-1.855) * P + R(1.517) * lR3 + R(0.304) * lN2R2 + R(0.328) * lS2R2;
      on = R(8.013) + R(0.905) * lR3 + R(0.602) * lR2 + R(0.751) * lN2R2;
    }
    if (!(ons > R(7.1) && on > R(7.1) && ons < R(9.4) && on < R(9.4))) { ons = nan; on = nan; }
    put(K_P10_ONS, ons);
    put(K_P10_ON, on);
  }

  if (tok & (T_M08 | T_M08ALL)) {                                         // metscales.py:975-1018
    if (o3o2) put(K_M08_O3O2, m08_o3o2<R>(logO35007O2));
    if (n2 && ha) put(K_M08_N2HA, m08_n2ha<R>(logN2Ha));
    if (r23ok) put(K_M08_R23, nan);            // `highZ is True` never holds for numpy.bool_ (:1013-1018)
    if (tok & T_M08ALL) {                                                 // metscales.py:1024-1059
      if (o3hb) put(K_M08_O3HB, nan);          // same dead `highZ is True / is False` tests (:1030, :1033)
      if (o2 && hb) put(K_M08_O2HB, nan);      // (:1042, :1045)
      // log10(O3/N2) MINUS the dust term: the reference adds dustcorrect() to the coefficient (:1054-1055)
      if (o3 && n2) put(K_M08_O3N2, m08_o3n2<R>(lO3 - lN2 - a * R(MCZ_K_O35007 - MCZ_K_N2)));
    }
  }

  if ((tok & T_M13) && ha && n2) {                                        // metscales.py:948-959
    put(K_M13_N2, R(8.743) + e[2] + (R(0.462) + e[3]) * logN2Ha);
    if (hb && o3) {
      const R O3N2 = logO3Hb - logN2Ha;
      R v = R(8.533) + e[4] - (R(0.214) + e[4]) * O3N2;
      if (O3N2 > R(1.7) || O3N2 < R(-1.1)) v = nan;
      put(K_M13_O3N2, v);
    }
  }

  if ((tok & T_P05) && r23ok) {                                           // metscales.py:755-761
    const R P = M::div(R3, R23), Psq = P * P;
    const R up = M::div(R23 + R(726.1) + R(842.2) * P + R(337.5) * Psq,
                        R(85.96) + R(82.76) * P + R(43.98) * Psq + R(1.793) * R23);
    const R low = M::div(R23 + R(106.4) + R(106.8) * P - R(3.40) * Psq,
                         R(17.72) + R(6.60) * P + R(6.95) * Psq - R(0.302) * R23);
    put(K_P05, zi_lt84 ? low : up);
  }

  if ((tok & T_P01) && o3o2 && r23ok) {                                   // metscales.py:853-863
    const R t = M::exp10(y);
    const R P = M::div(t, R(1) + t), Psq = P * P;
    const R old = M::div(R23 + R(54.2) + R(59.45) * P + R(7.31) * Psq,
                         R(6.07) + R(6.71) * P + R(0.371) * Psq + R(0.243) * R23);
    put(K_P01, zi_lt84 ? nan : old);
  }

  if ((tok & T_C01) && o3 && o2 && o3hb) {                                // metscales.py:874-891
    const R lx2 = -logO35007O2 - R(0.17609125905568124);                  // log10(O2O35007/1.5)
    const R lx3 = logO3Hb + R(-0.3010299956639812);                       // log10(10^logO3Hb*0.5)
    R v = nan;
    if (-logO35007O2 < R(-0.09691001300805639))                           // O2O35007 < 0.8
      v = R(-3.4225082001627745) + R(0.17) * lx2 + R(-0.44) * lx3 + R(12.0);   // log10(3.78e-4)
    else if (-logO35007O2 >= R(-0.09691001300805639))
      v = R(-3.4023048140744878) + R(-0.46) * lx3 + R(12.0);                   // log10(3.96e-4)
    put(K_C01_R23, v);
    if (s2 && n2) {
      const R logN2S2 = lN2 - M::log10(f[U_S2A]) + a * R(MCZ_K_N2 - MCZ_K_S2);  // metscales.py:395-396
      put(K_C01_N2S2, R(-3.2932822176632413) + R(0.17) * lx2 +
                          R(1.17) * (logN2S2 - R(-0.07058107428570727)) + R(12));  // log10(5.09e-4), log10(0.85)
    }
  }

  if ((tok & T_DP00) && s2 && s39 && hb) {                                // metscales.py:450-453, 671
    const R t1 = M::exp10(M::log10(f[U_S2A]) - lHb + a * R(MCZ_K_S2 - MCZ_K_HB));
    const R t2 = M::exp10(M::log10(f[U_S39069]) - lHb + a * R(MCZ_K_S3 - MCZ_K_HB));
    const R ls = M::log10(t1 + t2);
    put(K_DP00, R(1.53) * ls + R(8.27) + M::div(R(1.0), R(2.0) - R(9.0) * ls * ls * ls));
  }

  if ((tok & T_D16) && ha && n2 && s2) {                                  // metscales.py:1351-1356
    const R logN2S2 = lN2 - M::log10(f[U_S2A]) + a * R(MCZ_K_N2 - MCZ_K_S2);
    const R yy = logN2S2 + R(0.264) * logN2Ha;
    const R q = yy + R(0.3);
    R v = R(8.77) + yy - R(0.45) * (q * q * q * q * q);
    if (yy < R(-1.) || yy > R(0.5)) v = nan;
    put(K_D16, v);
  }

  if (tok & T_KD02) {
    // KD02_N2O2 (metscales.py:1076-1087)
    R kd = nan;
    const bool kd_present = n2 && o2 && ha && hb;
    if (kd_present) {
      kd = kd02_n2o2<R>(logN2O2);
      put(K_KD02_N2O2, kd);
    }
    carry.y = y;
    carry.n2ha = logN2Ha;
    carry.x = x;
    carry.aux = zi | ((logN2Ha > R(0.8)) ? AUX_N2HA_GT08 : 0u) | (M::isnan(kd) ? 0u : AUX_KD_VALID);
  }
}

// KK04_N2Ha without [OIII]/[OII] (metscales.py:1122-1126): one evaluation with logq = 7.37177
template <typename R> MCZ_HD R kk04_n2ha_noq(const IterState<R>& st) {
  return st.A - R(7.37177) * st.B;
}

// ---- phase C: finalise the KK04 outputs and KD02comb once the trip counts are known ------------
// kd, m91: this sample's KD02_N2O2 / M91 values (NaN when absent).
template <typename R, class Put>
MCZ_HD void phase_c(const IterState<R>& st, uint64_t present, bool n2ha_capped, bool r23_capped,
                    R kd, R m91, Put& put) {
  typedef Math<R> M;
  const R nan = M::nan();
  R kn = nan, kr = nan;
  if (present & (1ull << K_KK04_N2HA)) {                                  // metscales.py:1119-1129
    kn = n2ha_capped ? nan : st.Z1;
    if (st.aux & AUX_N2HA_GT08) kn = nan;
    put(K_KK04_N2HA, kn);
  }
  if (present & (1ull << K_KK04_R23)) {                                   // metscales.py:1179-1187
    kr = r23_capped ? nan : st.Z2;
    const R lo = st.L0 - st.lq2 * st.L1, up = st.U0 - st.lq2 * st.U1;
    if (lo > up) kr = nan;
    put(K_KK04_R23, kr);
  }
  if (present & (1ull << K_KD02COMB)) {                                   // metscales.py:1250-1261
    const bool ig = (st.aux & AUX_ZI_MASK) == 2u;                         // Z_init_guess > 8.4
    R v = ig ? kd : kn;
    if ((present & (1ull << K_KK04_R23)) && !ig && !M::isnan(kr) && !M::isnan(m91))
      v = R(0.5) * (kr + m91);
    put(K_KD02COMB, v);
  }
}

}  // namespace mcz
