// Percentile selection for one stored column: exact order statistics at the six ranks that
// np.percentile(data, [50, 16, 84]) interpolates between (reference mcz.py:273-301), one warp per
// column, float32 keys.  Included by mcz_production.cu after PT/PW/HB/CAP are defined.
//
// Every pass streams the column with 128-bit shared-memory loads and is branch-free: counters are
// per-lane PRIVATE 8-bit (or 16-bit) cells laid out so that lane l only touches bank l (no atomics,
// no bank conflicts), and non-members update a dummy cell.
//   pass 1  n, sum and a 64-bin histogram of f>>4, f = 10-bit "fine index" over a robust range
//           estimated from the first 64 samples;
//   pass 2  the 16 sub-bins (f & 15) of the (at most four) bins holding the six target ranks;
//   pass 3  gather of the few members of the target sub-bins, then a 64-key register sort.
// A "crowded" target sub-bin (ties such as the clamped E(B-V) = 1e-5, or an extreme fine index that
// also collects everything outside the range) is resolved by select_in_interval(): the same kind of
// pass restricted to the exact [min, max] of its members, narrowing until <= 64 members remain.
#pragma once

namespace mcz {

__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

// diagnostics: how often each selection path is taken (build with -DMCZ_SEL_STATS)
__device__ unsigned long long g_sel_counters[8];
enum { SEL_TOTAL = 0, SEL_FAST, SEL_NO_SUBSAMPLE, SEL_GENERAL, SEL_CROWDED, SEL_TRIVIAL, SEL_NARROW_ROUNDS };
__device__ __forceinline__ void sel_count(int which) {
#ifdef MCZ_SEL_STATS
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_sel_counters[which], 1ull);
#else
  (void)which;
#endif
}

// Per-lane private counters: 64 entries per lane.
//   8-bit : entry e -> byte offset g(e) = (e>>2)*128 + (e&3)   (+ lane*4)   rows 0..15, dummy rows 16..19
//   16-bit: entry e -> byte offset g(e) = (e>>1)*128 + (e&1)*2 (+ lane*4)   rows 0..31, dummy rows 32..39
template <typename HistT> struct PrivHist;

template <> struct PrivHist<unsigned char> {
  static constexpr int ROWS = 20, DUMMY = 16 * 128;
  static __device__ __forceinline__ int g(int e) { return ((e >> 2) << 7) + (e & 3); }
  static __device__ __forceinline__ void bump(unsigned char* hl, int off) {
    unsigned char* q = hl + off;
    *q = (unsigned char)(*q + 1);
  }
  static __device__ __forceinline__ unsigned total(const unsigned char* h, int e) {
    const uint4* row = reinterpret_cast<const uint4*>(h + ((e >> 2) << 7));
    const unsigned sel = 1u << (8 * (e & 3));
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 w = row[i];
      acc = __dp4a(w.x, sel, acc);
      acc = __dp4a(w.y, sel, acc);
      acc = __dp4a(w.z, sel, acc);
      acc = __dp4a(w.w, sel, acc);
    }
    return acc;
  }
};

template <> struct PrivHist<unsigned short> {
  static constexpr int ROWS = 40, DUMMY = 32 * 128;
  static __device__ __forceinline__ int g(int e) { return ((e >> 1) << 7) + ((e & 1) << 1); }
  static __device__ __forceinline__ void bump(unsigned char* hl, int off) {
    unsigned short* q = reinterpret_cast<unsigned short*>(hl + off);
    *q = (unsigned short)(*q + 1);
  }
  static __device__ __forceinline__ unsigned total(const unsigned char* h, int e) {
    const uint4* row = reinterpret_cast<const uint4*>(h + ((e >> 1) << 7));
    const int sh = 16 * (e & 1);
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 w = row[i];
      acc += ((w.x >> sh) & 0xffffu) + ((w.y >> sh) & 0xffffu) + ((w.z >> sh) & 0xffffu) + ((w.w >> sh) & 0xffffu);
    }
    return acc;
  }
};

template <typename HistT> __device__ __forceinline__ void priv_zero(unsigned char* h, int lane) {
  uint32_t* hw = reinterpret_cast<uint32_t*>(h);
#pragma unroll
  for (int i = 0; i < PrivHist<HistT>::DUMMY / 128; ++i) hw[i * 32 + lane] = 0u;    // dummy rows need no reset
  __syncwarp();
}

// lane l <- totals of entries l and l+32 and their exclusive prefixes (entry order 0..63)
template <typename HistT>
__device__ __forceinline__ void priv_readout(const unsigned char* h, int lane, unsigned& c0, unsigned& c1,
                                             unsigned& e0, unsigned& e1) {
  __syncwarp();
  c0 = PrivHist<HistT>::total(h, lane);
  c1 = PrivHist<HistT>::total(h, lane + 32);
  unsigned i0 = c0, i1 = c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { i0 += t0; i1 += t1; }
  }
  const unsigned tot0 = __shfl_sync(0xffffffffu, i0, 31);
  e0 = i0 - c0;
  e1 = tot0 + i1 - c1;
  __syncwarp();
}

// entry holding 0-based rank r (relative to the histogram) + its exclusive prefix and count
__device__ __forceinline__ void locate_rank(unsigned r, unsigned c0, unsigned c1, unsigned e0,
                                            unsigned e1, int& bin, unsigned& excl, unsigned& cnt) {
  const unsigned m0 = __ballot_sync(0xffffffffu, c0 > 0 && r >= e0 && r < e0 + c0);
  const unsigned m1 = __ballot_sync(0xffffffffu, c1 > 0 && r >= e1 && r < e1 + c1);
  if (m0) {
    const int l = __ffs(m0) - 1;
    bin = l;
    excl = __shfl_sync(0xffffffffu, e0, l);
    cnt = __shfl_sync(0xffffffffu, c0, l);
  } else {
    const int l = m1 ? __ffs(m1) - 1 : 31;
    bin = 32 + l;
    excl = __shfl_sync(0xffffffffu, e1, l);
    cnt = __shfl_sync(0xffffffffu, c1, l);
  }
}

// ascending bitonic sort of 64 keys held two per lane: key i<32 is `a` of lane i, key i>=32 is `b`
// of lane i-32
__device__ __forceinline__ void sort64_regs(float& a, float& b, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        const float lo = fminf(a, b), hi = fmaxf(a, b);
        a = lo; b = hi;
      } else {
        const float pa = __shfl_xor_sync(0xffffffffu, a, j), pb = __shfl_xor_sync(0xffffffffu, b, j);
        const bool lower = (lane & j) == 0;
        const bool up_a = (lane & k) == 0;
        const bool up_b = ((lane + 32) & k) == 0;
        a = (lower == up_a) ? fminf(a, pa) : fmaxf(a, pa);
        b = (lower == up_b) ? fminf(b, pb) : fmaxf(b, pb);
      }
    }
  }
}
__device__ __forceinline__ float sorted64_get(float a, float b, unsigned i) {
  const float va = __shfl_sync(0xffffffffu, a, i & 31), vb = __shfl_sync(0xffffffffu, b, i & 31);
  return i < 32u ? va : vb;
}

// numpy.lib._function_base_impl._lerp in float64 on the float32 order statistics
__device__ __forceinline__ float np_lerp_f(float a, float b, double t) {
  const double da = a, db = b, d = db - da;
  return (float)((t >= 0.5) ? (db - d * (1.0 - t)) : (da + d * t));
}

// Work areas of one warp.
struct SelWarp {
  const float* v;          // the column, Npad readable floats, [N, Npad) NaN
  int Npad;
  unsigned char* hbase;    // private-histogram area (PrivHist<HistT>::ROWS * 128 bytes)
  float* cand;             // 256 floats: tables + candidate list
  int lane;
};

// min / max of the finite values of the column; n_eq_min = how many equal the minimum
__device__ __forceinline__ void column_minmax(const SelWarp& w, float& mn, float& mx) {
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  mn = pinf; mx = ninf;
  for (int s0 = 4 * w.lane; s0 < w.Npad; s0 += 128) {
    const float4 x4 = *reinterpret_cast<const float4*>(w.v + s0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool fin = finite_f(xs[j]);
      mn = fminf(mn, fin ? xs[j] : pinf);
      mx = fmaxf(mx, fin ? xs[j] : ninf);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
}

// Exact order statistic inside an interval: 0-based rank r among the finite values x with
// lo <= x <= hi (there must be more than r of them).  Each round histograms the members into 64
// linear bins; a bin with <= 64 members is gathered and sorted, a larger one is narrowed to the
// exact [min, max] of its members (every round leaves at least two non-empty bins, so the member
// count strictly decreases; a tie at the minimum returns at once).
template <typename HistT>
__device__ float select_in_interval(const SelWarp& w, float lo, float hi, unsigned r) {
  typedef PrivHist<HistT> PH;
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  unsigned char* const hl = w.hbase + (w.lane << 2);
  unsigned* const ccount = reinterpret_cast<unsigned*>(w.cand);
  float* const clist = w.cand + 32;
  for (;;) {
    sel_count(SEL_NARROW_ROUNDS);
    if (!(hi > lo)) return lo;
    const float scale = 64.0f / (hi - lo), nLs = -lo * scale;
    if (!finite_f(scale) || !finite_f(nLs)) return lo;        // values differing by denormal amounts
    priv_zero<HistT>(w.hbase, w.lane);
    for (int s0 = 4 * w.lane; s0 < w.Npad; s0 += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(w.v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        const int b = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 63.0f));
        PH::bump(hl, (x >= lo && x <= hi) ? PH::g(b) : PH::DUMMY);       // NaN fails both compares
      }
    }
    unsigned c0, c1, e0, e1, excl, cnt;
    int b;
    priv_readout<HistT>(w.hbase, w.lane, c0, c1, e0, e1);
    locate_rank(r, c0, c1, e0, e1, b, excl, cnt);
    if (cnt <= 64u) {
      if (w.lane == 0) *ccount = 0u;
      __syncwarp();
      for (int s0 = 4 * w.lane; s0 < w.Npad; s0 += 128) {
        const float4 x4 = *reinterpret_cast<const float4*>(w.v + s0);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = xs[j];
          const int bb = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 63.0f));
          if (x >= lo && x <= hi && bb == b) clist[atomicAdd(ccount, 1u) & 63u] = x;
        }
      }
      __syncwarp();
      float ca = (unsigned)w.lane < cnt ? clist[w.lane] : pinf;
      float cb = (unsigned)(w.lane + 32) < cnt ? clist[w.lane + 32] : pinf;
      sort64_regs(ca, cb, w.lane);
      const float out = sorted64_get(ca, cb, (r - excl) & 63u);
      __syncwarp();
      return out;
    }
    // narrow to the exact range of the crowded bin
    float mn = pinf, mx = ninf;
    int neq = 0;
    for (int s0 = 4 * w.lane; s0 < w.Npad; s0 += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(w.v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        const int bb = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 63.0f));
        if (x >= lo && x <= hi && bb == b) {
          if (x < mn) { mn = x; neq = 1; }
          else if (x == mn) ++neq;
          mx = fmaxf(mx, x);
        }
      }
    }
    float gm = mn;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const unsigned nmin = (unsigned)__reduce_add_sync(0xffffffffu, mn == gm ? neq : 0);
    r -= excl;
    if (r < nmin) return gm;
    lo = gm;
    hi = mx;
  }
}

// percentile extraction for one stored column (mcz.py:273-301); all 32 lanes participate.
// v has Npad (multiple of 128) readable entries; entries in [N, Npad) are NaN.
template <typename HistT>
__device__ void warp_percentiles(const float* __restrict__ v, int N, int Npad, HistT* hist_, float* cand,
                                 float pct[3], int& nvalid, int& status) {
  typedef PrivHist<HistT> PH;
  (void)N;
  const int lane = threadIdx.x & 31;
  SelWarp w;
  w.v = v; w.Npad = Npad; w.hbase = reinterpret_cast<unsigned char*>(hist_); w.cand = cand; w.lane = lane;
  unsigned char* const hbase = w.hbase;
  unsigned char* const hl = hbase + (lane << 2);
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  // per-warp tables in the candidate buffer (1 KB): candidate counter and list first (shared with
  // select_in_interval), then bin -> byte offset of the target bin's first sub-entry (DUMMY for
  // the others), bin -> 16-bit mask of the sub-bins to visit in pass 3, bin -> first entry index
  unsigned* const ccount = reinterpret_cast<unsigned*>(cand);
  float* const clist = cand + 32;                                                 // [64]
  unsigned short* const G2 = reinterpret_cast<unsigned short*>(cand + 96);        // [64]
  unsigned short* const W3 = G2 + 64;                                             // [64]
  unsigned char* const tmap = reinterpret_cast<unsigned char*>(W3 + 64);          // [64]
  const float nanf_ = Math<float>::nan();
  pct[0] = pct[1] = pct[2] = nanf_;
  sel_count(SEL_TOTAL);
  // ---- robust range from the first 64 samples (i.i.d. draws: any fixed subset is a fair
  // subsample): [q06 - 1.5 w, q94 + 1.5 w] with w = q94 - q06, so that a few wild values do not
  // stretch the fine index; whatever falls outside is clamped into the two extreme entries.
  float L, H;
  {
    float a = v[lane], b = v[lane + 32];
    a = finite_f(a) ? a : pinf;
    b = finite_f(b) ? b : pinf;
    const int m = __popc(__ballot_sync(0xffffffffu, a < pinf)) + __popc(__ballot_sync(0xffffffffu, b < pinf));
    sort64_regs(a, b, lane);
    float qlo = pinf, qhi = ninf;
    if (m >= 8) {
      qlo = sorted64_get(a, b, (unsigned)(0.06f * (float)(m - 1)));
      qhi = sorted64_get(a, b, (unsigned)(m - 1) - (unsigned)(0.06f * (float)(m - 1)));
      if (!(qhi > qlo)) { qlo = sorted64_get(a, b, 0u); qhi = sorted64_get(a, b, (unsigned)(m - 1)); }
    }
    if (!(qhi > qlo)) {
      // too few finite values up front, or a constant prefix: use the exact range of the column
      sel_count(SEL_NO_SUBSAMPLE);
      column_minmax(w, qlo, qhi);
      if (!(qhi >= qlo)) { nvalid = 0; status = ST_EMPTY; return; }                 // mcz.py:277-280
      if (!(qhi > qlo)) { qlo -= 1.0f; qhi += 1.0f; }       // constant column: any range will do
      L = qlo; H = qhi;
    } else {
      const float ww = 1.5f * (qhi - qlo);
      L = qlo - ww;
      H = qhi + ww;
    }
  }
  float scale = 1024.0f / (H - L), nLs = -L * scale;
  if (!finite_f(scale) || !finite_f(nLs)) { L = -FLT_MAX * 0.25f; H = FLT_MAX * 0.25f; scale = 1024.0f / (H - L); nLs = -L * scale; }

  // ---- pass 1: sum and the 64-bin histogram of f >> 4 (non-finite values hit the dummy rows:
  // f = 1024 -> g(64) is the first dummy row)
  priv_zero<HistT>(hbase, lane);
  float sum = 0.0f;
#pragma unroll 2
  for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
    const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = xs[j];
      const bool fin = finite_f(x);                                                 // mcz.py:273
      const int f = fin ? __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 1023.0f)) : 1024;
      PH::bump(hl, PH::g(f >> 4));
      sum += fin ? x : 0.0f;
    }
  }
  unsigned c0, c1, e0, e1;
  priv_readout<HistT>(hbase, lane, c0, c1, e0, e1);
  const int n = (int)(__shfl_sync(0xffffffffu, e1, 31) + __shfl_sync(0xffffffffu, c1, 31));
  const double total = warp_sum((double)sum);
  nvalid = n;
  if (n == 0) { sel_count(SEL_TRIVIAL); status = ST_EMPTY; return; }                // mcz.py:277-280
  if (!(total > 0.0)) { sel_count(SEL_TRIVIAL); status = ST_NONPOS; return; }       // mcz.py:282-285
  status = ST_OK;
  // ranks in ascending order: p16 pair, median pair, p84 pair
  const double qs[3] = {16.0 / 100.0, 50.0 / 100.0, 84.0 / 100.0};                  // mcz.py:288
  unsigned rk[6];
  double gam[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double vi = (double)(n - 1) * qs[q];
    const double lo = floor(vi);
    gam[q] = vi - lo;
    rk[2 * q] = (unsigned)lo;
    rk[2 * q + 1] = min((unsigned)lo + 1u, (unsigned)(n - 1));
  }
  float res[6];
  int bins[6], ids[6];
  unsigned excl[6], cnts[6];
  int nd = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    locate_rank(rk[i], c0, c1, e0, e1, bins[i], excl[i], cnts[i]);
    if (i == 0 || bins[i] != bins[i - 1]) ++nd;
    ids[i] = nd - 1;
  }
  bool general = nd > 4;       // the six ranks straddle more than four bins (two pairs split): rare
  int ent[6];
  unsigned off[6], ecnt[6];
  unsigned long long want = 0ull;
  int bigs[3] = {-1, -1, -1};
  int nbig = 0;
  unsigned ncand = 0;
  if (!general) {
    G2[lane] = (unsigned short)PH::DUMMY;
    G2[lane + 32] = (unsigned short)PH::DUMMY;
    W3[lane] = 0;
    W3[lane + 32] = 0;
    tmap[lane] = 64;
    tmap[lane + 32] = 64;
    if (lane == 0) *ccount = 0u;
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        G2[bins[i]] = (unsigned short)PH::g(ids[i] << 4);
        tmap[bins[i]] = (unsigned char)(ids[i] << 4);
      }
    }
    // ---- pass 2: the 16 sub-bins (f & 15) of each target bin
    priv_zero<HistT>(hbase, lane);
#pragma unroll 2
    for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        const int f = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 1023.0f));
        const int base = finite_f(x) ? (int)G2[f >> 4] : PH::DUMMY;
        PH::bump(hl, base + PH::g(f & 15));      // g(eb + sub) = g(eb) + g(sub) since eb % 16 == 0
      }
    }
    priv_readout<HistT>(hbase, lane, c0, c1, e0, e1);
    // locate each rank among the entries (id*16 + sub); entries are in ascending value order
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      // exclusive prefix at the first entry of this target bin = pass-2 elements in lower target bins
      const int eb = ids[i] << 4;
      const unsigned base = eb < 32 ? __shfl_sync(0xffffffffu, e0, eb) : __shfl_sync(0xffffffffu, e1, eb - 32);
      unsigned ex;
      locate_rank(base + (rk[i] - excl[i]), c0, c1, e0, e1, ent[i], ex, ecnt[i]);
      off[i] = base + (rk[i] - excl[i]) - ex;
    }
    // Entries with few members are gathered and sorted (<= 64 candidates in all); crowded ones
    // (and the two extreme fine indices) are resolved from the exact range of their members.
    for (unsigned thr = 32u;; thr >>= 1) {
      want = 0ull; nbig = 0; ncand = 0;
      bigs[0] = bigs[1] = bigs[2] = -1;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i > 0 && ent[i] == ent[i - 1]) continue;
        const int sub = ent[i] & 15;
        const bool extreme = (bins[i] == 0 && sub == 0) || (bins[i] == HB - 1 && sub == 15);
        if (ecnt[i] > thr || extreme) {
          if (nbig < 3) bigs[nbig] = ent[i];
          ++nbig;
        } else {
          want |= 1ull << ent[i];
          ncand += ecnt[i];
        }
      }
      if (ncand <= 64u || thr <= 8u) break;
    }
    general = nbig > 3 || ncand > 64u;
  }
  if (general) {
    // order statistics one by one over the exact range of the column (a handful of columns per
    // million): same machinery, more passes
    sel_count(SEL_GENERAL);
    float mn, mx;
    column_minmax(w, mn, mx);
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
      if (i > 0 && rk[i] == rk[i - 1]) { res[i] = res[i - 1]; continue; }
      res[i] = select_in_interval<HistT>(w, mn, mx, rk[i]);
    }
  } else {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) W3[bins[i]] = (unsigned short)(W3[bins[i]] | (1u << (ent[i] & 15)));
    }
    __syncwarp();
    // ---- pass 3: visit the members of the marked sub-bins (a few per cent of the column)
    float bmn[3] = {pinf, pinf, pinf}, bmx[3] = {ninf, ninf, ninf};
    int bcm[3] = {0, 0, 0};          // how many members equal the running minimum (tie detection)
    if (nbig == 0) {
#pragma unroll 2
      for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
        const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = xs[j];
          const int f = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 1023.0f));
          const unsigned m = W3[f >> 4];
          if (((m >> (f & 15)) & 1u) && finite_f(x)) clist[atomicAdd(ccount, 1u) & 63u] = x;
        }
      }
    } else {
      sel_count(SEL_CROWDED);
      for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
        const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = xs[j];
          const int f = __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 1023.0f));
          const unsigned m = W3[f >> 4];
          if (((m >> (f & 15)) & 1u) && finite_f(x)) {
            const int e = (int)tmap[f >> 4] | (f & 15);
            if ((want >> e) & 1ull) {
              clist[atomicAdd(ccount, 1u) & 63u] = x;
            } else {
#pragma unroll
              for (int bb = 0; bb < 3; ++bb)
                if (e == bigs[bb]) {
                  if (x < bmn[bb]) { bmn[bb] = x; bcm[bb] = 1; }
                  else if (x == bmn[bb]) ++bcm[bb];
                  bmx[bb] = fmaxf(bmx[bb], x);
                }
            }
          }
        }
      }
    }
    sel_count(SEL_FAST);
    __syncwarp();
    float ca = (unsigned)lane < ncand ? clist[lane] : pinf;
    float cb = (unsigned)(lane + 32) < ncand ? clist[lane + 32] : pinf;
    sort64_regs(ca, cb, lane);
    // position of each rank in the sorted candidate list: candidates of lower wanted entries first
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      unsigned below = 0;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        // count each distinct lower gathered entry once (entries are non-decreasing in j)
        if (ent[j] < ent[i] && ((want >> ent[j]) & 1ull) && (j == 0 || ent[j] != ent[j - 1])) below += ecnt[j];
      }
      res[i] = sorted64_get(ca, cb, (below + off[i]) & 63u);
    }
    if (nbig) {
      // A crowded entry is exactly the set of finite values in [bmn, bmx] (the fine index is
      // monotone), so rank off[i] inside it is an order statistic of that interval.
      unsigned nmin[3];
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        float gm = bmn[bb];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
          bmx[bb] = fmaxf(bmx[bb], __shfl_xor_sync(0xffffffffu, bmx[bb], o));
        }
        nmin[bb] = (unsigned)__reduce_add_sync(0xffffffffu, bmn[bb] == gm ? bcm[bb] : 0);
        bmn[bb] = gm;
      }
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        if ((want >> ent[i]) & 1ull) continue;
        if (i > 0 && ent[i - 1] == ent[i] && off[i] == off[i - 1]) { res[i] = res[i - 1]; continue; }
        const int bb = ent[i] == bigs[0] ? 0 : (ent[i] == bigs[1] ? 1 : 2);
        const float lo = bb == 0 ? bmn[0] : (bb == 1 ? bmn[1] : bmn[2]);
        const float hi = bb == 0 ? bmx[0] : (bb == 1 ? bmx[1] : bmx[2]);
        const unsigned nm = bb == 0 ? nmin[0] : (bb == 1 ? nmin[1] : nmin[2]);
        res[i] = (lo == hi || off[i] < nm) ? lo : select_in_interval<HistT>(w, lo, hi, off[i]);
      }
    }
  }
  pct[0] = np_lerp_f(res[2], res[3], gam[1]);       // median
  pct[1] = np_lerp_f(res[0], res[1], gam[0]);       // 16th
  pct[2] = np_lerp_f(res[4], res[5], gam[2]);       // 84th
}

}  // namespace mcz
