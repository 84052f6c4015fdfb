// Percentile selection for one stored column: exact order statistics at the six ranks that
// np.percentile(data, [50, 16, 84]) interpolates between (reference mcz.py:273-301), one warp per
// column, float32 keys.  Included by mcz_production.cu after PT/PW/HB/CAP are defined.
//
// Fast path (warp_percentiles): three passes over the column with per-lane PRIVATE, bank-conflict
// free 8-bit counters (no atomics): a 10-bit "fine index" f(x) over a range taken from the first
// 128 samples; pass 1 histograms f>>4 (64 bins) and accumulates n and sum; pass 2 histograms the
// 16 sub-bins of the (at most four) bins that hold the six target ranks; pass 3 gathers the few
// members of the target sub-bins, which are then sorted.  Anything unusual (targets in an edge
// bin, heavy ties, more than 64 candidates) falls back to the generic narrowing select below.
#pragma once
namespace mcz {

// ---- selection (one warp per stored scale) ----------------------------------------------------
__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

__device__ __forceinline__ int bin_of(float x, float L, float scale) {
  int b = __float2int_rz((x - L) * scale);
  return min(max(b, 0), HB - 1);
}

template <typename HistT> __device__ __forceinline__ unsigned hist_row_sum(const HistT* row);
template <> __device__ __forceinline__ unsigned hist_row_sum<unsigned char>(const unsigned char* row) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row);
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = __dp4a(w[i], 0x01010101u, acc);
  return acc;
}
template <> __device__ __forceinline__ unsigned hist_row_sum<unsigned short>(const unsigned short* row) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row);
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += (w[i] & 0xffffu) + (w[i] >> 16);
  return acc;
}

// Histogram of the members of [L, H] into HB linear bins; per-lane private counters
// hist[bin][lane] (no atomics, no bank conflicts).  On return lane l holds the counts of bins l
// and l+32 and their exclusive prefix sums.
template <typename HistT>
__device__ __forceinline__ void warp_histogram(const float* __restrict__ v, int N, float L, float H,
                                               float scale, HistT* hist, unsigned& c0, unsigned& c1,
                                               unsigned& e0, unsigned& e1) {
  const int lane = threadIdx.x & 31;
  uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
#pragma unroll
  for (int i = 0; i < (int)(HB * 32 * sizeof(HistT) / 4 / 32); ++i) hw[i * 32 + lane] = 0u;
  __syncwarp();
  for (int s = lane; s < N; s += 32) {
    const float x = v[s];
    if (x >= L && x <= H) {          // NaN fails both; +-inf never inside a finite [L, H]
      HistT* h = hist + bin_of(x, L, scale) * 32 + lane;
      *h = (HistT)(*h + 1);
    }
  }
  __syncwarp();
  c0 = hist_row_sum<HistT>(hist + lane * 32);
  c1 = hist_row_sum<HistT>(hist + (lane + 32) * 32);
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

// bin holding 0-based rank r (relative to the histogram) + its exclusive prefix and count
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

// gather the members of [L,H] whose bin is `b` into cand[] (cnt <= CAP) and sort them ascending
__device__ __forceinline__ void gather_sort(const float* __restrict__ v, int N, float L, float H,
                                            float scale, int b, unsigned cnt, float* cand) {
  const int lane = threadIdx.x & 31;
  unsigned filled = 0;
  for (int s0 = 0; s0 < N; s0 += 32) {
    const int s = s0 + lane;
    const float x = s < N ? v[s] : 0.0f;
    const bool hit = s < N && x >= L && x <= H && bin_of(x, L, scale) == b;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) cand[filled + __popc(m & ((1u << lane) - 1u))] = x;
    filled += __popc(m);
  }
  unsigned M = 32;
  while (M < cnt) M <<= 1;
  for (unsigned i = cnt + lane; i < M; i += 32) cand[i] = __int_as_float(0x7f800000);   // +inf pad
  __syncwarp();
  for (unsigned k = 2; k <= M; k <<= 1) {
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned i = lane; i < M; i += 32) {
        const unsigned ixj = i ^ j;
        if (ixj > i) {
          const float a = cand[i], c = cand[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { cand[i] = c; cand[ixj] = a; }
        }
      }
      __syncwarp();
    }
  }
}

// min / max of the members of [L,H] falling in bin b
__device__ __forceinline__ void bin_minmax(const float* __restrict__ v, int N, float L, float H,
                                           float scale, int b, float& mn, float& mx) {
  const int lane = threadIdx.x & 31;
  mn = __int_as_float(0x7f800000);
  mx = __int_as_float(0xff800000);
  for (int s = lane; s < N; s += 32) {
    const float x = v[s];
    if (x >= L && x <= H && bin_of(x, L, scale) == b) { mn = fminf(mn, x); mx = fmaxf(mx, x); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
}

// generic exact order statistic: rank r (0-based among the finite values) by repeated narrowing
template <typename HistT>
__device__ float select_one(const float* __restrict__ v, int N, unsigned r, float L, float H,
                            HistT* hist, float* cand) {
  unsigned base = 0;
  for (;;) {
    if (!(H > L)) return L;
    const float scale = (float)HB / (H - L);
    unsigned c0, c1, e0, e1, excl, cnt;
    int b;
    warp_histogram<HistT>(v, N, L, H, scale, hist, c0, c1, e0, e1);
    locate_rank(r - base, c0, c1, e0, e1, b, excl, cnt);
    if (cnt <= (unsigned)CAP) {
      gather_sort(v, N, L, H, scale, b, cnt, cand);
      const float out = cand[r - base - excl];
      __syncwarp();
      return out;
    }
    float mn, mx;
    bin_minmax(v, N, L, H, scale, b, mn, mx);
    base += excl;
    L = mn;
    H = mx;
  }
}

// numpy.lib._function_base_impl._lerp in float64 on the float32 order statistics
__device__ __forceinline__ float np_lerp_f(float a, float b, double t) {
  const double da = a, db = b, d = db - da;
  return (float)((t >= 0.5) ? (db - d * (1.0 - t)) : (da + d * t));
}

// generic percentile extraction for one stored column (mcz.py:273-301): own statistics pass,
// 64-bin histogram over [min, max], gather + sort of the target bins, narrowing when needed
template <typename HistT>
__device__ void warp_percentiles_generic(const float* __restrict__ v, int N, HistT* hist, float* cand,
                                 float pct[3], int& nvalid, int& status) {
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  float sum = 0.0f, mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (int s = lane; s < N; s += 32) {
    const float x = v[s];
    if (finite_f(x)) { ++cnt; sum += x; mn = fminf(mn, x); mx = fmaxf(mx, x); }   // mcz.py:273
  }
  const int n = __reduce_add_sync(0xffffffffu, cnt);
  const double total = warp_sum((double)sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const float nanf_ = Math<float>::nan();
  pct[0] = pct[1] = pct[2] = nanf_;
  nvalid = n;
  if (n == 0) { status = ST_EMPTY; return; }                                      // mcz.py:277-280
  if (!(total > 0.0)) { status = ST_NONPOS; return; }                             // mcz.py:282-285
  status = ST_OK;
  const double qs[3] = {50.0 / 100.0, 16.0 / 100.0, 84.0 / 100.0};                // mcz.py:288
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
  if (!(mx > mn)) {
#pragma unroll
    for (int i = 0; i < 6; ++i) res[i] = mn;
  } else {
    // level-1 histogram shared by all six ranks
    const float scale = (float)HB / (mx - mn);
    unsigned c0, c1, e0, e1;
    warp_histogram<HistT>(v, N, mn, mx, scale, hist, c0, c1, e0, e1);
    int bins[6];
    unsigned excl[6], cnts[6];
    bool fast = true;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      locate_rank(rk[i], c0, c1, e0, e1, bins[i], excl[i], cnts[i]);
      fast = fast && cnts[i] <= (unsigned)CAP;
    }
    if (fast) {
      unsigned done = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (done & (1u << i)) continue;
        gather_sort(v, N, mn, mx, scale, bins[i], cnts[i], cand);
#pragma unroll
        for (int j = i; j < 6; ++j)
          if (!(done & (1u << j)) && bins[j] == bins[i]) { res[j] = cand[rk[j] - excl[j]]; done |= 1u << j; }
        __syncwarp();
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        if (i > 0 && rk[i] == rk[i - 1]) { res[i] = res[i - 1]; continue; }
        res[i] = select_one<HistT>(v, N, rk[i], mn, mx, hist, cand);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) pct[q] = np_lerp_f(res[2 * q], res[2 * q + 1], gam[q]);
  (void)lane;
}

}  // namespace mcz

// =================================================================================================
// Fast path
// =================================================================================================
namespace mcz {

// diagnostics: how often each selection path is taken (read back with mcz_debug_counters)
__device__ unsigned long long g_sel_counters[8];
enum { SEL_TOTAL = 0, SEL_FAST, SEL_FB_RANGE, SEL_FB_EDGE, SEL_FB_CAND, SEL_TRIVIAL };
enum { SEL_FB_2BIG = 6, SEL_BIG_RESOLVED = 7 };
__device__ __forceinline__ void sel_count(int which) {
#ifdef MCZ_SEL_STATS
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_sel_counters[which], 1ull);
#else
  (void)which;
#endif
}

// 10-bit fine index, monotone non-decreasing in x (one FFMA, two FMNMX, one F2I); NaN -> 0
__device__ __forceinline__ int fine_index(float x, float scale, float nLs) {
  return __float2int_rz(fminf(fmaxf(fmaf(x, scale, nLs), 0.0f), 1023.0f));
}

// Per-lane private counters: 64 entries per lane, laid out so that lane l only ever touches bank l
// (no atomics, no bank conflicts), plus a dummy region that absorbs the updates of non-members so
// that the inner loops are branch-free.
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

// percentile extraction for one stored column (mcz.py:273-301); all 32 lanes participate.
// v has Npad (multiple of 128) readable entries; entries in [N, Npad) are NaN.
template <typename HistT>
__device__ void warp_percentiles(const float* __restrict__ v, int N, int Npad, HistT* hist_, float* cand,
                                 float pct[3], int& nvalid, int& status) {
  typedef PrivHist<HistT> PH;
  const int lane = threadIdx.x & 31;
  unsigned char* const hbase = reinterpret_cast<unsigned char*>(hist_);
  unsigned char* const hl = hbase + (lane << 2);
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  // per-warp tables in the candidate buffer (1 KB): bin -> byte offset of the target bin's first
  // sub-entry (DUMMY for the others); bin -> 16-bit mask of the sub-bins to visit in pass 3;
  // bin -> first entry index; candidate counter and list
  unsigned short* const G2 = reinterpret_cast<unsigned short*>(cand);            // [64]
  unsigned short* const W3 = G2 + 64;                                             // [64]
  unsigned char* const tmap = reinterpret_cast<unsigned char*>(W3 + 64);          // [64]
  unsigned* const ccount = reinterpret_cast<unsigned*>(tmap + 64);
  float* const clist = cand + 96;                                                 // [64]
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
    if (m < 4) { sel_count(SEL_FB_RANGE); warp_percentiles_generic<HistT>(v, N, hist_, cand, pct, nvalid, status); return; }
    float qlo = sorted64_get(a, b, (unsigned)(0.06f * (float)(m - 1))),
          qhi = sorted64_get(a, b, (unsigned)(m - 1) - (unsigned)(0.06f * (float)(m - 1)));
    if (!(qhi > qlo)) { qlo = sorted64_get(a, b, 0u); qhi = sorted64_get(a, b, (unsigned)(m - 1)); }
    if (!(qhi > qlo)) { sel_count(SEL_FB_RANGE); warp_percentiles_generic<HistT>(v, N, hist_, cand, pct, nvalid, status); return; }
    const float w = 1.5f * (qhi - qlo);
    L = qlo - w;
    H = qhi + w;
  }
  const float scale = 1024.0f / (H - L), nLs = -L * scale;
  if (!finite_f(scale) || !finite_f(nLs)) { sel_count(SEL_FB_RANGE); warp_percentiles_generic<HistT>(v, N, hist_, cand, pct, nvalid, status); return; }

  // ---- pass 1: n, sum, 64-bin histogram of f >> 4 (branch-free: non-finite values hit the dummy rows)
  priv_zero<HistT>(hbase, lane);
  int cnt = 0;
  float sum = 0.0f;
#pragma unroll 2
  for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
    const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = xs[j];
      const bool fin = finite_f(x);                                                 // mcz.py:273
      const int f = fine_index(x, scale, nLs);
      PH::bump(hl, fin ? PH::g(f >> 4) : PH::DUMMY);
      cnt += fin ? 1 : 0;
      sum += fin ? x : 0.0f;
    }
  }
  const int n = __reduce_add_sync(0xffffffffu, cnt);
  const double total = warp_sum((double)sum);
  const float nanf_ = Math<float>::nan();
  pct[0] = pct[1] = pct[2] = nanf_;
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
  unsigned c0, c1, e0, e1;
  priv_readout<HistT>(hbase, lane, c0, c1, e0, e1);
  int bins[6], ids[6];
  unsigned excl[6], cnts[6];
  int nd = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    locate_rank(rk[i], c0, c1, e0, e1, bins[i], excl[i], cnts[i]);
    if (i == 0 || bins[i] != bins[i - 1]) ++nd;
    ids[i] = nd - 1;
  }
  if (nd > 4) {      // the six ranks straddle more than four bins (two pairs split): rare
    sel_count(SEL_FB_EDGE);
    warp_percentiles_generic<HistT>(v, N, hist_, cand, pct, nvalid, status);
    return;
  }
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
  // ---- pass 2: the 16 sub-bins (f & 15) of each target bin (branch-free)
  priv_zero<HistT>(hbase, lane);
#pragma unroll 2
  for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
    const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = xs[j];
      const int f = fine_index(x, scale, nLs);
      const int base = finite_f(x) ? (int)G2[f >> 4] : PH::DUMMY;
      PH::bump(hl, base + PH::g(f & 15));        // g(eb + sub) = g(eb) + g(sub) since eb % 16 == 0
    }
  }
  priv_readout<HistT>(hbase, lane, c0, c1, e0, e1);
  // locate each rank among the entries (id*16 + sub); entries are in ascending value order
  int ent[6];
  unsigned off[6], ecnt[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    // exclusive prefix at the first entry of this target bin = pass-2 elements in lower target bins
    const int eb = ids[i] << 4;
    const unsigned base = eb < 32 ? __shfl_sync(0xffffffffu, e0, eb) : __shfl_sync(0xffffffffu, e1, eb - 32);
    unsigned ex;
    locate_rank(base + (rk[i] - excl[i]), c0, c1, e0, e1, ent[i], ex, ecnt[i]);
    off[i] = base + (rk[i] - excl[i]) - ex;
  }
  // Entries with few members are gathered and sorted (<= 64 candidates in all).  "Crowded" entries
  // (ties, or the two extreme fine indices that also collect everything outside [L, H]) are
  // resolved separately from the min/max of their members; at most three of them.
  unsigned long long want = 0ull;
  int bigs[3] = {-1, -1, -1};
  int nbig = 0;
  unsigned ncand = 0;
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
  if (nbig > 3 || ncand > 64u) {
    sel_count(nbig > 3 ? SEL_FB_2BIG : SEL_FB_CAND);
    warp_percentiles_generic<HistT>(v, N, hist_, cand, pct, nvalid, status);
    return;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) W3[bins[i]] = (unsigned short)(W3[bins[i]] | (1u << (ent[i] & 15)));
  }
  __syncwarp();
  // ---- pass 3: visit the members of the wanted / crowded entries (a few per cent of the column)
  float bmn[3] = {pinf, pinf, pinf}, bmx[3] = {ninf, ninf, ninf};
  int bcm[3] = {0, 0, 0};          // how many members equal the running minimum (tie detection)
#pragma unroll 2
  for (int s0 = 4 * lane; s0 < Npad; s0 += 128) {
    const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    unsigned hits = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = fine_index(xs[j], scale, nLs);
      const unsigned m = W3[f >> 4];
      const bool hit = ((m >> (f & 15)) & 1u) && finite_f(xs[j]);
      hits |= hit ? (1u << j) : 0u;
    }
    if (hits) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (hits & (1u << j)) {
          const float x = xs[j];
          const int f = fine_index(x, scale, nLs);
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
  float res[6];
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
    // monotone), so rank off[i] inside it is an order statistic of that interval.  Ties at the
    // minimum (clamped values such as E(B-V) = 1e-5) are recognised from the count of members equal
    // to the minimum; anything else is narrowed by the generic select.
    sel_count(SEL_BIG_RESOLVED);
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
      res[i] = (lo == hi || off[i] < nm) ? lo : select_one<HistT>(v, N, off[i], lo, hi, hist_, cand);
    }
  }
  pct[0] = np_lerp_f(res[2], res[3], gam[1]);       // median
  pct[1] = np_lerp_f(res[0], res[1], gam[0]);       // 16th
  pct[2] = np_lerp_f(res[4], res[5], gam[2]);       // 84th
}

}  // namespace mcz
