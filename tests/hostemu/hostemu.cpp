// TEST-ONLY serial emulation of one galaxy using the SAME per-sample device functions
// (pymcz_b200/csrc/mcz_physics.cuh) compiled for the host with g++.  It exists so that the
// physics can be checked against the oracle in the GPU-less build container; it is not part of
// the product (nothing under pymcz_b200/ builds, loads or calls it).
#include <vector>
#include <cstring>
#include "../../pymcz_b200/csrc/mcz_physics.cuh"

using namespace mcz;

template <typename R> struct ArrPut {
  static constexpr bool WANT_AUX = false;
  double* Z; int N; int s;
  void operator()(int key, R v) { if (key < NKEYS) Z[(size_t)key * N + s] = (double)v; }
};

template <typename R>
static void run(const double* flux, const double* noise, int N, uint32_t tok, int dust,
                double* Z, uint64_t* present_out, int* trips, int* ret) {
  static const int used2line[NUSED] = {L_HA, L_HB, L_O2, L_O3, L_N2, L_S2A, L_S2B, L_S39069};
  std::vector<R> f((size_t)N * NUSED);
  uint32_t flags = 0;
  for (int s = 0; s < N; ++s)
    for (int u = 0; u < NUSED; ++u) {
      double v = flux[(size_t)used2line[u] * N + s];
      if (!(std::fabs(v) <= 1.7976931348623157e308)) v = 0.0;   // metallicity.py:96
      if (v < 0) v = 0.0;                                        // metallicity.py:99
      R r = (R)v;
      f[(size_t)s * NUSED + u] = r;
      const bool pos = (u >= U_S2B) ? (r > R(1e-9)) : (r > R(0));
      if (pos) flags |= 1u << u;
    }
  const int de = effective_dust(dust, flags);
  const uint64_t present = present_mask(flags, tok);
  *present_out = present;
  *ret = minimum_ok(flags) ? 0 : -1;
  std::vector<IterState<R>> st(N);
  for (int s = 0; s < N; ++s) {
    R e[5];
    for (int j = 0; j < 5; ++j) e[j] = (R)noise[(size_t)j * N + s];
    ArrPut<R> put{Z, N, s};
    std::memset(&st[s], 0, sizeof(IterState<R>));
    Carry<R> c;
    std::memset(&c, 0, sizeof(c));
    phase_a<R>(&f[(size_t)s * NUSED], e, flags, tok, de, put, c);
    if ((tok & T_KD02) && minimum_ok(flags)) {
      const R kd = (present & (1ull << K_KD02_N2O2)) ? (R)Z[(size_t)K_KD02_N2O2 * N + s] : Math<R>::nan();
      make_iter<R>(c, kd, st[s]);
    }
  }
  trips[0] = trips[1] = 0;
  if (!(tok & T_KD02) || !minimum_ok(flags)) return;
  const bool has_kn = present & (1ull << K_KK04_N2HA), has_kr = present & (1ull << K_KK04_R23);
  const bool o3o2 = (flags & F_O2) && (flags & F_O3);
  // start value of the N2Ha loop (metscales.py:1097-1103)
  bool allnan = true;
  if (present & (1ull << K_KD02_N2O2))
    for (int s = 0; s < N; ++s) if (!Math<R>::isnan(st[s].Z1)) { allnan = false; break; }
  if (allnan) for (int s = 0; s < N; ++s) st[s].Z1 = R(8.6);
  int ii1 = 0, ii2 = 0;
  if (has_kn) {
    if (o3o2) {
      double conv = 100;
      while (conv > 1.0e-3 && ii1 < 100) {
        ++ii1;
        double sum = 0;
        for (int s = 0; s < N; ++s) sum += (double)kk04_n2ha_trip<R>(st[s]);
        conv = sum / N;
      }
    } else {
      for (int s = 0; s < N; ++s) st[s].Z1 = kk04_n2ha_noq<R>(st[s]);
    }
  }
  if (has_kr) {
    double conv = 100;
    while (conv > 1e-4 && ii2 < 100) {
      ++ii2;
      double sum = 0;
      for (int s = 0; s < N; ++s) sum += (double)kk04_r23_trip<R>(st[s]);
      conv = std::fabs(sum / N);
    }
  }
  trips[0] = ii1; trips[1] = ii2;
  for (int s = 0; s < N; ++s) {
    ArrPut<R> put{Z, N, s};
    const R kd = (present & (1ull << K_KD02_N2O2)) ? (R)Z[(size_t)K_KD02_N2O2 * N + s] : Math<R>::nan();
    const R m91 = (R)Z[(size_t)K_M91 * N + s];
    phase_c<R>(st[s], present, ii1 >= 100, ii2 >= 100, kd, m91, put);
  }
}

extern "C" int hostemu_galaxy(int use_float, const double* flux, const double* noise, int N,
                              unsigned tok, int dust, double* Z, unsigned long long* present,
                              int* trips, int* ret) {
  uint64_t p = 0;
  if (use_float) run<float>(flux, noise, N, tok, dust, Z, &p, trips, ret);
  else run<double>(flux, noise, N, tok, dust, Z, &p, trips, ret);
  *present = p;
  return 0;
}

#include "../../pymcz_b200/csrc/mcz_philox.cuh"
extern "C" void hostemu_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                               unsigned k1, unsigned* out) {
  uint32_t o[4];
  mcz::philox4x32_10(c0, c1, c2, c3, k0, k1, o);
  for (int i = 0; i < 4; ++i) out[i] = o[i];
}

// The closed-form root selections on their own: which = 0 KD02 [NII]/[OII], 1 M08 [NII]/Ha,
// 2 M08 [OIII]/[OII], 3 M08 [OIII]/[NII] (M08all); y = the log ratio the reference subtracts from
// the constant coefficient.
extern "C" void hostemu_roots(int which, int use_float, const double* y, int n, double* out) {
  using namespace mcz;
  for (int i = 0; i < n; ++i) {
    if (use_float) {
      const float v = (float)y[i];
      out[i] = which == 0 ? kd02_n2o2<float>(v) : which == 1 ? m08_n2ha<float>(v)
               : which == 2 ? m08_o3o2<float>(v) : m08_o3n2<float>(v);
    } else {
      out[i] = which == 0 ? kd02_n2o2<double>(y[i]) : which == 1 ? m08_n2ha<double>(y[i])
               : which == 2 ? m08_o3o2<double>(y[i]) : m08_o3n2<double>(y[i]);
    }
  }
}
