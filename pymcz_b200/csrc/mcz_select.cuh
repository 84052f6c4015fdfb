// Percentile selection for one stored column: exact order statistics at the six ranks that
// np.percentile(data, [50, 16, 84]) interpolates between (reference mcz.py:273-301), one warp per
// column, float32 keys.  Included by mcz_production.cu.
//
// One "round" = two streaming passes over the column (128-bit loads, no branches in the hot loop):
//   pass H  1024-bin histogram of a monotone "fine index" f(x) over a value range [L, H], counted
//           with shared-memory atomics on 16-bit counters packed two per word (2 KB per warp; one
//           branch-free ATOMS per sample -- the pass is bound by the shared-atomic pipe, measured
//           4.9 clk per warp-wide ATOMS with 17 warps at it, tools/microbench/atoms2.cu); the sum
//           of the column rides along;
//   pass G  gather of the members of the (at most three) fine-index ranges that hold the target
//           ranks -- a rank pair (k, k+1) always lies in one bin or in two bins with nothing in
//           between -- recorded as one bit per sample and appended afterwards, followed by a
//           32- or 64-key bitonic sort in registers.
// The first round uses a robust range from the first 64 samples, [q06 - 1.5w, q94 + 1.5w] cut to
// that prefix's own [min, max]; values outside are clamped into the end bins.  Each of the (at
// most three) target ranges may hold 64 candidates: one shared list and one sort when they fit in
// 64 together, one list and one sort per range otherwise.  A value that repeats heavily in the
// prefix (the clamped E(B-V) = 1e-5) is counted by vote, left out of the lists and merged back by
// rank.  A range that still holds more than 64 samples is "crowded": pass G records the exact
// min / max of its members and how many equal the min (a tie is then answered at once), and
// select_in_interval() runs further rounds restricted to [min, max] of the one bin that holds
// both ranks (two ranks in different bins are the max of one and the min of the other); those
// rounds index by ordered bit patterns, so the interval shrinks by at least 512x per round
// whatever the spacing of the values, and the loop terminates.
#pragma once

namespace mcz {

__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

// diagnostics: how often each selection path is taken (build with -DMCZ_SEL_STATS)
__device__ unsigned long long g_sel_counters[8];
enum { SEL_TOTAL = 0, SEL_FAST, SEL_NO_SUBSAMPLE, SEL_CROWDED, SEL_TRIVIAL, SEL_ROUNDS };
__device__ __forceinline__ void sel_count(int which) {
#ifdef MCZ_SEL_STATS
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_sel_counters[which], 1ull);
#else
  (void)which;
#endif
}

#ifndef MCZ_GEN_UBG
#define MCZ_GEN_UBG 4
#endif
static constexpr int FB = 1024;     // fine-index bins per round

// 1024 counters per warp: PACK16 -> 512 words, two 16-bit counters each (columns of < 65536
// samples); otherwise 1024 32-bit words.
template <bool PACK16> struct FineHist {
  static constexpr int WORDS = PACK16 ? FB / 2 : FB;
  // + 32 spare words, one per lane: the streaming pass sends the increment of a sample that must not
  // be counted (NaN, inf, padding) there instead of making the atomic conditional
  // + 4 words for the cluster kernel's exchange: [WORDS + 32] the block's count of the repeated value,
  // [WORDS + 33] the block's column sum (float bits)
  static constexpr int BYTES = WORDS * 4 + 128 + 16;
  static __device__ __forceinline__ void zero(unsigned* h, int lane) {
#pragma unroll
    for (int i = 0; i < WORDS / 32; ++i) h[i * 32 + lane] = 0u;
    __syncwarp();
  }
  static __device__ __forceinline__ void add(unsigned* h, int f) {
    if (PACK16) atomicAdd(h + (f >> 1), 1u << ((f & 1) << 4));
    else atomicAdd(h + f, 1u);
  }
  // branch-free form for the streaming pass: a sample that is not counted adds 0 to a word of
  // its lane's own bank, so the four samples of a 128-bit load keep independent instruction chains
  static __device__ __forceinline__ void add_if(unsigned* h, int f, bool ok, int lane) {
#ifdef MCZ_PRED_ATOMS
    // a PREDICATED shared reduction (no branch, no redirected address): the byte address of the word
    // is 2 * (f & ~1), the increment 1 or 65536 by the parity of the bin
    (void)lane;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(h) + (PACK16 ? 2u * ((unsigned)f & ~1u) : 4u * (unsigned)f);
    const unsigned val = PACK16 ? 1u + ((unsigned)f & 1u) * 65535u : 1u;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p red.shared.add.u32 [%0], %1; }"
                 :: "r"(addr), "r"(val), "r"((unsigned)ok) : "memory");
#else
    if (PACK16) atomicAdd(h + (ok ? (f >> 1) : lane), ok ? (1u << ((f & 1) << 4)) : 0u);
    else atomicAdd(h + (ok ? f : lane), ok ? 1u : 0u);
#endif
  }
  // the same on the raw index bits (fine_tbits) and 32-bit shared addresses: `hs` = address of the
  // table, `ls` = address of the lane's own word (a sample that is not counted adds 0 there)
  static __device__ __forceinline__ void add_t(unsigned hs, unsigned ls, int tb, bool ok) {
    unsigned addr, val;
    if (PACK16) {
      addr = hs + 2u * ((unsigned)tb & (unsigned)(FB - 2));        // word (bin >> 1): byte offset 2 * (bin & ~1)
      // low or high 16-bit counter: 1 + parity * 65535, as a multiply-add (FMA pipe, not compare + select)
      asm("mad.lo.u32 %0, %1, 65535, 1;" : "=r"(val) : "r"((unsigned)tb & 1u));
    } else {
      addr = hs + 4u * ((unsigned)tb & (unsigned)(FB - 1));
      val = 1u;
    }
    addr = ok ? addr : ls;                                         // ls: the lane's spare word
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(val) : "memory");
  }
  static __device__ __forceinline__ void add_n(unsigned* h, int f, unsigned n) {      // n < 65536 when PACK16
    if (PACK16) atomicAdd(h + (f >> 1), n << ((f & 1) << 4));
    else atomicAdd(h + f, n);
  }
  static __device__ __forceinline__ unsigned get(const unsigned* h, int f) {
    if (PACK16) return (h[f >> 1] >> ((f & 1) << 4)) & 0xffffu;
    return h[f];
  }
  // sum of the 32 bins [32*lane, 32*lane+32)
  static __device__ __forceinline__ unsigned lane_total(const unsigned* h, int lane) {
    // each lane starts at a different one of its 128-bit pieces so that the eight lanes of a
    // shared-memory phase touch eight different bank groups
    const uint4* q = reinterpret_cast<const uint4*>(h + lane * (WORDS / 32));
    const int rot = PACK16 ? (lane >> 1) : lane;
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < WORDS / 128; ++i) {
      const uint4 w = q[(i + rot) & (WORDS / 128 - 1)];
      if (PACK16) acc += (w.x & 0xffffu) + (w.x >> 16) + (w.y & 0xffffu) + (w.y >> 16) + (w.z & 0xffffu) + (w.z >> 16) + (w.w & 0xffffu) + (w.w >> 16);
      else acc += w.x + w.y + w.z + w.w;
    }
    return acc;
  }
};

// ascending bitonic sort of 64 keys held two per lane: key i is `a` (i even) or `b` (i odd) of
// lane i / 2, so the six distance-1 stages stay inside a lane and 15 stages need a shuffle
__device__ __forceinline__ void sort64_regs(float& a, float& b, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
    const bool up = (lane & (k >> 1)) == 0 || k == 64;
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 1) {
        const float lo = fminf(a, b), hi = fmaxf(a, b);
        a = up ? lo : hi;
        b = up ? hi : lo;
      } else {
        const float pa = __shfl_xor_sync(0xffffffffu, a, j >> 1), pb = __shfl_xor_sync(0xffffffffu, b, j >> 1);
        const bool lower = (lane & (j >> 1)) == 0;
        a = (lower == up) ? fminf(a, pa) : fmaxf(a, pa);
        b = (lower == up) ? fminf(b, pb) : fmaxf(b, pb);
      }
    }
  }
}
// ascending bitonic sort of 32 keys, one per lane
__device__ __forceinline__ void sort32_regs(float& a, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float pa = __shfl_xor_sync(0xffffffffu, a, j);
      const bool lower = (lane & j) == 0;
      const bool up = (lane & k) == 0 || k == 32;
      a = (lower == up) ? fminf(a, pa) : fmaxf(a, pa);
    }
  }
}
__device__ __forceinline__ float sorted64_get(float a, float b, unsigned i) {
  const float va = __shfl_sync(0xffffffffu, a, (i >> 1) & 31), vb = __shfl_sync(0xffffffffu, b, (i >> 1) & 31);
  return (i & 1u) ? vb : va;
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
  int ngroups;             // 128-sample groups select_in_interval scans: Npad/128, or fewer once the
  int nmine;               // members of the crowded ranges were compacted (nmine = this lane's count)
  unsigned* hist;          // FineHist<>::WORDS words
  unsigned* ccount;        // candidate counter
  float* clist;            // 64 candidates (select_in_interval)
  float* glist;            // 192 candidates of pass G; aliases hist, which the plan has consumed by then
  int lane;
  float scale, nLs;        // fine index of the first round
  // cluster kernel, second round: the (at most three) target ranges of the first round are
  // subdivided into RSUB bins each (see fine_tbits_w); rnr = 0: plain first-round index
  int rnr, rflo[3], rfhi[3];
  float rmul[3];
#ifdef MCZ_COLTIME_DUMP
  mutable int dbg;
#endif
};

// Monotone map of a value to a bin of [0, FB): u = saturate(x * scale + nLs) in [0, 1] (one FFMA.SAT:
// it clamps AND sends NaN to 0), then round(u * (FB - 1)) read from the mantissa of a 2^23-biased
// FMA.  Two instructions on the FMA pipe -- no FMNMX clamps, no float-to-int conversion: the
// streaming passes of the selection are bound by the ALU pipe (integer, compare, min/max, select
// instructions issue at half the FMA rate), so work is moved off it wherever it can be.
// fine_tbits() returns the raw bits 0x4B000000 + bin; the callers mask or compare them as they are.
static constexpr int FINE_BIAS = 0x4B000000;
__device__ __forceinline__ int fine_tbits(float x, float scale, float nLs) {
  float u;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(x), "f"(scale), "f"(nLs));
  return __float_as_int(fmaf(u, (float)(FB - 1), 8388608.0f));
}
__device__ __forceinline__ int fine_index(float x, float scale, float nLs) {
  return fine_tbits(x, scale, nLs) - FINE_BIAS;
}

// Second-round index of the cluster kernel.  Layout of the 1024 bins, RSUB = 339 per range:
//   [0] below range 0 | [1, RSUB] range 0 | [RSUB+1] between | [RSUB+2, 2 RSUB+1] range 1 | [2 RSUB+2]
//   between | [2 RSUB+3, 3 RSUB+2] range 2 | [3 RSUB+3] above.  Inside a range the position is
//   u * 1023 - (lo - 0.5) in [0, width), u the saturated first-round coordinate, cut into RSUB
//   pieces: monotone in x like the first-round index, ~340 / width times finer.
static constexpr int RSUB = 339;
__device__ __forceinline__ int fine_tbits_w(const SelWarp& w, float x) {
  const int tb = fine_tbits(x, w.scale, w.nLs);
  if (w.rnr == 0) return tb;
  const int f1 = tb - FINE_BIAS;
  float u;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(x), "f"(w.scale), "f"(w.nLs));
  int idx = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (r < w.rnr) {
      const int base = r * (RSUB + 1);
      if (f1 > w.rfhi[r]) idx = base + RSUB + 1;
      else if (f1 >= w.rflo[r]) {
        const float v = fmaf(u, (float)(FB - 1), 0.5f - (float)w.rflo[r]);
        int sub = __float2int_rd(v * w.rmul[r]);
        sub = sub < 0 ? 0 : (sub > RSUB - 1 ? RSUB - 1 : sub);
        idx = base + 1 + sub;
      }
    }
  }
  return idx + FINE_BIAS;
}

// After pass H: every lane gets its exclusive prefix over lanes (32 bins per lane) and its total.
template <bool PACK16>
__device__ __forceinline__ void fine_prefix(const SelWarp& w, unsigned& lane_excl, unsigned& lane_tot, unsigned& n) {
  __syncwarp();
  lane_tot = FineHist<PACK16>::lane_total(w.hist, w.lane);
  unsigned inc = lane_tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (w.lane >= o) inc += t;
  }
  lane_excl = inc - lane_tot;
  n = __shfl_sync(0xffffffffu, inc, 31);
}

// fine bin holding 0-based rank r, its exclusive prefix and its count
template <bool PACK16>
__device__ __forceinline__ void fine_locate(const SelWarp& w, unsigned lane_excl, unsigned lane_tot, unsigned r,
                                            int& f, unsigned& excl, unsigned& cnt) {
  const unsigned own = __ballot_sync(0xffffffffu, lane_tot > 0 && r >= lane_excl && r < lane_excl + lane_tot);
  const int ol = own ? __ffs(own) - 1 : 31;
  const unsigned base = __shfl_sync(0xffffffffu, lane_excl, ol);
  const unsigned c = FineHist<PACK16>::get(w.hist, 32 * ol + w.lane);
  unsigned inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (w.lane >= o) inc += t;
  }
  const unsigned ex = base + inc - c;
  const unsigned hit = __ballot_sync(0xffffffffu, c > 0 && r >= ex && r < ex + c);
  const int hl = hit ? __ffs(hit) - 1 : 31;
  f = 32 * ol + hl;
  excl = __shfl_sync(0xffffffffu, ex, hl);
  cnt = __shfl_sync(0xffffffffu, c, hl);
}

// Same for two consecutive ranks rA <= rB <= rA + 1: they share one scan when the same lane's 32
// bins hold both (nearly always).  Returns the bins, the samples below fA and the samples up to
// and including fB.
template <bool PACK16>
__device__ __forceinline__ void fine_locate_pair(const SelWarp& w, unsigned lane_excl, unsigned lane_tot,
                                                 unsigned rA, unsigned rB, int& fA, unsigned& exA, int& fB,
                                                 unsigned& endB, bool& redoB) {
  const unsigned ownA = __ballot_sync(0xffffffffu, lane_tot > 0 && rA >= lane_excl && rA < lane_excl + lane_tot);
  const unsigned ownB = __ballot_sync(0xffffffffu, lane_tot > 0 && rB >= lane_excl && rB < lane_excl + lane_tot);
  const int olA = ownA ? __ffs(ownA) - 1 : 31;
  const int olB = ownB ? __ffs(ownB) - 1 : 31;
  const unsigned base = __shfl_sync(0xffffffffu, lane_excl, olA);
  const unsigned c = FineHist<PACK16>::get(w.hist, 32 * olA + w.lane);
  unsigned inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (w.lane >= o) inc += t;
  }
  const unsigned ex = base + inc - c;
  const unsigned hitA = __ballot_sync(0xffffffffu, c > 0 && rA >= ex && rA < ex + c);
  const unsigned hitB = __ballot_sync(0xffffffffu, c > 0 && rB >= ex && rB < ex + c);
  const int hlA = hitA ? __ffs(hitA) - 1 : 31;
  const int hlB = hitB ? __ffs(hitB) - 1 : 31;
  fA = 32 * olA + hlA;
  exA = __shfl_sync(0xffffffffu, ex, hlA);
  fB = 32 * olA + hlB;
  endB = __shfl_sync(0xffffffffu, ex + c, hlB);
  redoB = olB != olA;         // rB starts the next non-empty group of bins: the caller locates it alone
}

// min / max of the finite values of the column
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

// Order-preserving image of a float in the unsigned integers (-0 and +0 coincide).
__device__ __forceinline__ unsigned ordered_bits(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? (unsigned)b + 0x80000000u : 0x80000000u - ((unsigned)b & 0x7fffffffu);
}

// Exact order statistics inside an interval: 0-based ranks rA <= rB <= rA+1 among the finite
// values x with lo <= x <= hi (there are more than rB of them).
template <bool PACK16>
__device__ void select_in_interval(const SelWarp& w, float lo, float hi, unsigned rA, unsigned rB,
                                   float& outA, float& outB) {
  typedef FineHist<PACK16> FH;
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  bool haveA = false;
  // (each round divides the bit-pattern span by >= 512, so 4 rounds always suffice; the cap only
  // guards the GPU against a hang should that reasoning ever be wrong)
  for (int round = 0; round < 16; ++round) {
    sel_count(SEL_ROUNDS);
#ifdef MCZ_COLTIME_DUMP
    w.dbg += 256;
#endif
    // The index of these rounds is taken from the ordered bit patterns, (u(x) - u(lo)) >> shift with
    // the shift that maps [lo, hi] onto < 1024 bins: monotone like the float index of the first
    // round, but free of overflow whatever the spacing (denormals, values near FLT_MAX).
    const unsigned ulo = ordered_bits(lo), uspan = ordered_bits(hi) - ulo;
    if (!(hi > lo) || uspan == 0u) {                    // all equal
      if (!haveA) outA = lo;
      outB = lo;
      return;
    }
    const int shift = max(0, 22 - __clz(uspan));       // uspan >> shift in [1, 1023]
    FH::zero(w.hist, w.lane);
    for (int g = 0; g < w.ngroups; ++g) {
      const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * w.lane + 128 * g);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        if (4 * g + j < w.nmine && x >= lo && x <= hi) FH::add(w.hist, (int)((ordered_bits(x) - ulo) >> shift));   // NaN fails
      }
    }
    unsigned lex, ltot, n;
    fine_prefix<PACK16>(w, lex, ltot, n);
    int fa, fb;
    unsigned ea, ca, eb, cb;
    fine_locate<PACK16>(w, lex, ltot, rA, fa, ea, ca);
    fine_locate<PACK16>(w, lex, ltot, rB, fb, eb, cb);
    if (fa != fb) {
      // Two consecutive ranks in different bins: nothing lies between them, so rA is the largest
      // member of bin fa and rB the smallest member of bin fb.  (Narrowing to [min, max] of the
      // two bins would not shrink the interval when they are its two ends.)
      float mxA = ninf, mnB = pinf;
      for (int g = 0; g < w.ngroups; ++g) {
        const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * w.lane + 128 * g);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = xs[j];
          if (4 * g + j < w.nmine && x >= lo && x <= hi) {
            const int f = (int)((ordered_bits(x) - ulo) >> shift);
            if (f == fa) mxA = fmaxf(mxA, x);
            if (f == fb) mnB = fminf(mnB, x);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, o));
        mnB = fminf(mnB, __shfl_xor_sync(0xffffffffu, mnB, o));
      }
      if (!haveA) outA = mxA;
      outB = mnB;
      return;
    }
    const unsigned m = ca;                               // both ranks in bin fa
    if (w.lane == 0) *w.ccount = 0u;
    __syncwarp();
    const unsigned span = 0u;
    if (m <= 64u) {
      for (int g = 0; g < w.ngroups; ++g) {
        const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * w.lane + 128 * g);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = xs[j];
          const unsigned d = ((ordered_bits(x) - ulo) >> shift) - (unsigned)fa;
          if (4 * g + j < w.nmine && x >= lo && x <= hi && d <= span) w.clist[atomicAdd(w.ccount, 1u) & 63u] = x;
        }
      }
      __syncwarp();
      float ca_ = (unsigned)w.lane < m ? w.clist[w.lane] : pinf;
      float cb_ = (unsigned)(w.lane + 32) < m ? w.clist[w.lane + 32] : pinf;
      sort64_regs(ca_, cb_, w.lane);
      if (!haveA) outA = sorted64_get(ca_, cb_, (rA - ea) & 63u);
      outB = sorted64_get(ca_, cb_, (rB - ea) & 63u);
      __syncwarp();
      return;
    }
    // crowded: narrow to the exact range of the members of bin fa (at most 1/1024 of the interval,
    // so the rounds terminate: the interval shrinks until it is a single value)
    float mn = pinf, mx = ninf;
    int neq = 0;
    for (int g = 0; g < w.ngroups; ++g) {
      const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * w.lane + 128 * g);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        const unsigned d = ((ordered_bits(x) - ulo) >> shift) - (unsigned)fa;
        if (4 * g + j < w.nmine && x >= lo && x <= hi && d <= span) {
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
    rA -= ea;
    rB -= ea;
    if (rB < nmin) { if (!haveA) outA = gm; outB = gm; return; }
    if (!haveA && rA < nmin) { outA = gm; haveA = true; rA = rB; }
    lo = gm;
    hi = mx;
  }
  if (!haveA) outA = lo;
  outB = lo;
}

// Robust fine-index range of a column from its first 64 samples (i.i.d. draws: any fixed subset is
// a fair subsample): [q06 - 1.5 w, q94 + 1.5 w], w = q94 - q06, so that a few wild values do not
// stretch the index.  Returns false when that prefix has fewer than 8 finite values or is constant.
// Also reports a heavily repeated value of the prefix (>= 5 equal samples out of 64), e.g. the
// E(B-V) = 1e-5 clamp: the histogram pass counts such a value with a vote instead of one shared
// atomic per sample (same-address atomics serialise).
// (a, b: samples `lane` and `lane + 32` of the column)
__device__ __forceinline__ bool subsample_range_ab(float a, float b, int lane, float& scale, float& nLs,
                                                   bool& has_tie, float& tie) {
  const float pinf = __int_as_float(0x7f800000);
  a = finite_f(a) ? a : pinf;
  b = finite_f(b) ? b : pinf;
  const int m = __popc(__ballot_sync(0xffffffffu, a < pinf)) + __popc(__ballot_sync(0xffffffffu, b < pinf));
  sort64_regs(a, b, lane);
  {
    // sorted[i] == sorted[i + 4]  <=>  at least five copies of that value
    // (key i + 4 is the same register two lanes up)
    const float a4 = __shfl_down_sync(0xffffffffu, a, 2);
    const float b4 = __shfl_down_sync(0xffffffffu, b, 2);
    const unsigned ma = __ballot_sync(0xffffffffu, lane < 30 && a < pinf && a == a4);
    const unsigned mb = __ballot_sync(0xffffffffu, lane < 30 && b < pinf && b == b4);
    has_tie = (ma | mb) != 0u;
    tie = ma ? __shfl_sync(0xffffffffu, a, __ffs(ma) - 1) : __shfl_sync(0xffffffffu, b, mb ? __ffs(mb) - 1 : 0);
  }
  if (m < 8) return false;
  float qlo = sorted64_get(a, b, (unsigned)(0.06f * (float)(m - 1)));
  float qhi = sorted64_get(a, b, (unsigned)(m - 1) - (unsigned)(0.06f * (float)(m - 1)));
  if (!(qhi > qlo)) { qlo = sorted64_get(a, b, 0u); qhi = sorted64_get(a, b, (unsigned)(m - 1)); }
  if (!(qhi > qlo)) return false;
  const float ww = 1.5f * (qhi - qlo);
  float L = qlo - ww, H = qhi + ww;
  if (m >= 48) {
    // A target rank lies outside the prefix's own [min, max] with probability 0.84^m per column,
    // so the index need not reach further: bimodal columns (two branches of a diagnostic) then
    // keep bins fine enough for the 64-candidate lists.
    // (two bins of margin: the end bins then hold only values outside the prefix's range, and
    // NaN / inf, which index into them, are never members of a target range in practice)
    const float pmin = sorted64_get(a, b, 0u), pmax = sorted64_get(a, b, (unsigned)(m - 1));
    const float d = (pmax - pmin) * (2.0f / (float)(FB - 4));
    L = fmaxf(L, pmin - d);
    H = fminf(H, pmax + d);
  }
  scale = 1.0f / (H - L);               // u = saturate((x - L) / (H - L)), see fine_tbits
  nLs = -L * scale;
  return finite_f(scale) && finite_f(nLs) && scale > 0.0f;
}

__device__ __forceinline__ bool subsample_range(const float* __restrict__ v, int lane, float& scale, float& nLs,
                                                bool& has_tie, float& tie) {
  return subsample_range_ab(v[lane], v[lane + 32], lane, scale, nLs, has_tie, tie);
}

// ---- the selection of one column, in stages --------------------------------------------------
// Stage "plan" (after the histogram is complete): n, the six ranks, the fine-index ranges that
// hold them and which of those are gathered (<= 64 candidates in all) or crowded.
struct SelPlan {
  int n;
  unsigned rk[6];          // ascending: p16 pair, median pair, p84 pair
  double gam[3];
  int nr;                  // number of merged ranges (<= 3)
  int rlo[3], rhi[3];      // fine-index range
  unsigned rE[3], rM[3];   // samples below the range / inside it
  int rq[3];               // range of pair q
  bool crowded[3];         // too many members to gather: resolved by further rounds over the column
  bool big[3];             // 65 .. BIG_MAX members: gathered, but selected by bisection instead of a 64-key sort
  unsigned need[3];        // members gathered for the range (0 if crowded)
  unsigned loff[3];        // where its candidates start in the sorted list (when not split)
  bool split;              // false: <= 64 candidates in all, one sort; true: one sort per range
  unsigned ncand;          // candidates gathered in all (<= 192)
  // a heavily repeated value T (found in the 64-sample prefix, counted by vote in pass H): its ntie
  // copies are not gathered; the range that holds T's fine bin merges them back by rank
  int tie_range;           // -1: none
  unsigned ntie;
  float tie;
};

// BIG (the long-column instantiations, N' > 2304): a range with 65 .. BIG_MAX members is still gathered,
// as long as all candidates fit the 256-entry list, and its ranks are found by bisection on the list
// (list_select) -- bins of a 1024-bin index hold five times more samples at N' = 11000 than at 2200, and
// a range that is "crowded" costs further passes over the whole column.
static constexpr unsigned BIG_MAX = 192u, LIST_CAP = 256u;
template <bool PACK16, bool BIG = false>
__device__ __forceinline__ void sel_make_plan(const SelWarp& w, SelPlan& P, bool has_tie, float tie, int ftie,
                                              unsigned ntie) {
  unsigned lex, ltot, un;
  fine_prefix<PACK16>(w, lex, ltot, un);
  P.n = (int)un;
  P.nr = 0;
  P.ncand = 0;
  P.tie_range = -1;
  P.ntie = 0;
  P.tie = Math<float>::nan();
  if (P.n == 0) return;
  const double qs[3] = {16.0 / 100.0, 50.0 / 100.0, 84.0 / 100.0};                  // mcz.py:288
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double vi = (double)(P.n - 1) * qs[q];
    const double lo = floor(vi);
    P.gam[q] = vi - lo;
    P.rk[2 * q] = (unsigned)lo;
    P.rk[2 * q + 1] = min((unsigned)lo + 1u, (unsigned)(P.n - 1));
  }
  int fa[3], fb[3];
  unsigned exa[3], endb[3];
  bool redo[3];
#pragma unroll
  for (int q = 0; q < 3; ++q)      // three independent chains of shuffles, no branch between them
    fine_locate_pair<PACK16>(w, lex, ltot, P.rk[2 * q], P.rk[2 * q + 1], fa[q], exa[q], fb[q], endb[q], redo[q]);
  if (redo[0] || redo[1] || redo[2]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (redo[q]) {
        unsigned exB, cnB;
        fine_locate<PACK16>(w, lex, ltot, P.rk[2 * q + 1], fb[q], exB, cnB);
        endb[q] = exB + cnB;
      }
    }
  }
  // Merge the three pair ranges [fa, fb] where they touch or overlap (small n, ties).  Written out
  // case by case with constant indices: a loop over a run-time range count would put the whole
  // plan into local memory.
  {
    // range 0 starts as pair 0
    int lo0 = fa[0], hi0 = fb[0];
    unsigned E0 = exa[0], M0 = endb[0] - exa[0];
    int lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
    unsigned E1 = 0u, M1 = 0u, E2 = 0u, M2 = 0u;
    // pair 1: joins range 0 or opens range 1
    const bool j1 = fa[1] <= hi0;
    if (j1) {
      if (fb[1] > hi0) { hi0 = fb[1]; M0 = endb[1] - E0; }
    } else {
      lo1 = fa[1]; hi1 = fb[1]; E1 = exa[1]; M1 = endb[1] - exa[1];
    }
    // pair 2: joins the last range or opens the next one
    const int last_hi = j1 ? hi0 : hi1;
    const bool j2 = fa[2] <= last_hi;
    int q2;
    if (j2) {
      if (fb[2] > last_hi) {
        if (j1) { hi0 = fb[2]; M0 = endb[2] - E0; } else { hi1 = fb[2]; M1 = endb[2] - E1; }
      }
      q2 = j1 ? 0 : 1;
    } else if (j1) {
      lo1 = fa[2]; hi1 = fb[2]; E1 = exa[2]; M1 = endb[2] - exa[2];
      q2 = 1;
    } else {
      lo2 = fa[2]; hi2 = fb[2]; E2 = exa[2]; M2 = endb[2] - exa[2];
      q2 = 2;
    }
    P.rlo[0] = lo0; P.rhi[0] = hi0; P.rE[0] = E0; P.rM[0] = M0;
    P.rlo[1] = lo1; P.rhi[1] = hi1; P.rE[1] = E1; P.rM[1] = M1;
    P.rlo[2] = lo2; P.rhi[2] = hi2; P.rE[2] = E2; P.rM[2] = M2;
    P.rq[0] = 0; P.rq[1] = j1 ? 0 : 1; P.rq[2] = q2;
    P.nr = q2 + 1;
  }
  const int nr = P.nr;
  if (has_tie && ntie > 0u) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (r < nr && ftie >= P.rlo[r] && ftie <= P.rhi[r]) { P.tie_range = r; P.ntie = ntie; P.tie = tie; }
  }
  // a range is gathered if it has at most 64 members (not counting the copies of the tie value);
  // the others are "crowded"
  unsigned ncand = 0;
  bool anybig = false;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    P.crowded[r] = false;
    P.big[r] = false;
    P.need[r] = 0u;
    if (r < nr) {
      const unsigned need = P.rM[r] - (r == P.tie_range ? P.ntie : 0u);
      if (need <= 64u) { P.need[r] = need; ncand += need; }
      else if (BIG && need <= BIG_MAX) { P.need[r] = need; ncand += need; P.big[r] = true; anybig = true; }
      else P.crowded[r] = true;
    }
  }
  if (BIG && anybig && ncand > LIST_CAP) {            // the list cannot take them all: the big ranges are crowded after all
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (P.big[r]) { P.big[r] = false; P.crowded[r] = true; ncand -= P.need[r]; P.need[r] = 0u; anybig = false; }
  }
  if (P.tie_range >= 0) {
    const bool cr = P.tie_range == 0 ? P.crowded[0] : (P.tie_range == 1 ? P.crowded[1] : P.crowded[2]);
    if (cr) { P.tie_range = -1; P.ntie = 0; P.tie = Math<float>::nan(); }   // still crowded: ties stay members
  }
  P.split = ncand > 64u || (BIG && anybig);
  P.loff[0] = 0u;
  P.loff[1] = P.need[0];
  P.loff[2] = P.need[0] + P.need[1];
  P.ncand = ncand;
}

// range tests of pass G: (unsigned)(f - tlo) <= tsp; `gathered` selects the ranges that are
// gathered (true) or the crowded ones (false); the others get a test that never matches
__device__ __forceinline__ void sel_tests(const SelPlan& P, bool gathered, int tlo[3], unsigned tsp[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const bool use = r < P.nr && (P.crowded[r] != gathered);
    tlo[r] = use ? P.rlo[r] : 0x40000000;
    tsp[r] = use ? (unsigned)(P.rhi[r] - P.rlo[r]) : 0u;
  }
}

// Exact order statistics of the members of one range among the gathered candidates: x[i] = entry
// i * 32 + lane of the list, bit i of `in` says whether it belongs to the range.  Returns the values of
// 0-based ranks ra <= rb <= ra + 1 by MSB-first bisection on the order-preserving bit patterns,
// starting below the common prefix of the members: ~20 steps of eight compares and one REDUX.
__device__ __forceinline__ void list_select(const float x[8], unsigned in, unsigned ra, unsigned rb, float& outA, float& outB) {
  unsigned u[8];
  unsigned umin = 0xffffffffu, umax = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool m = (in >> i) & 1u;
    u[i] = m ? ordered_bits(x[i]) : 0xffffffffu;        // never <= a bisection threshold (those have a zero bit)
    umin = min(umin, u[i]);
    umax = max(umax, m ? u[i] : 0u);
  }
  umin = __reduce_min_sync(0xffffffffu, umin);
  umax = __reduce_max_sync(0xffffffffu, umax);
  const unsigned diff = umin ^ umax;
  const int top = diff ? 31 - __clz(diff) : -1;          // highest bit in which two members differ
  unsigned K = top >= 31 ? 0u : (umax >> (top + 1)) << (top + 1);   // the common prefix
  for (int b = top; b >= 0; --b) {
    const unsigned T = K | ((1u << b) - 1u);             // keys with bit b clear (given the prefix so far)
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += (u[i] <= T) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((unsigned)c <= ra) K |= 1u << b;
  }
  // K = key of rank ra; rank rb is the same value if enough members are <= K, else the next larger key
  int cle = 0;
  unsigned nxt = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cle += (u[i] <= K) ? 1 : 0;
    nxt = min(nxt, u[i] > K ? u[i] : 0xffffffffu);
  }
  cle = __reduce_add_sync(0xffffffffu, cle);
  nxt = __reduce_min_sync(0xffffffffu, nxt);
  const unsigned KB = ((unsigned)cle > rb || nxt == 0xffffffffu) ? K : nxt;
  outA = __uint_as_float(K >= 0x80000000u ? K - 0x80000000u : ((0x80000000u - K) | 0x80000000u));
  outB = __uint_as_float(KB >= 0x80000000u ? KB - 0x80000000u : ((0x80000000u - KB) | 0x80000000u));
}

// Stage "finish": the gathered candidates are in w.glist (range r at P.loff[r], P.need[r] of them,
// any order); sort them, read the order statistics, resolve crowded ranges, interpolate.
// For crowded ranges the caller supplies the min / max of their members and how many equal the
// min, and has compacted those members to the front of each lane's slots (w.nmine, w.ngroups).
template <bool PACK16, bool BIG = false>
__device__ __forceinline__ void sel_finish(const SelWarp& w, const SelPlan& P, const float bmn[3],
                                           const float bmx[3], const unsigned nmin[3], float pct[3]) {
  const float pinf = __int_as_float(0x7f800000);
  const int lane = w.lane;
  __syncwarp();
  float res[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) res[i] = pinf;
  // one sort of the whole list (ranges are disjoint and ascending, so their candidates come out
  // range after range); when the list is split, the members of each range are first picked out
  // of the (unordered) list into w.clist and sorted range by range
  const int nsort = P.split ? P.nr : 1;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
  for (int t = 0; t < nsort; ++t) {
    const float* src = w.glist;
    unsigned cnt = P.ncand;
    if (P.split) {
      const int lo_t = t == 0 ? P.rlo[0] : (t == 1 ? P.rlo[1] : P.rlo[2]);
      const unsigned sp_t = (unsigned)((t == 0 ? P.rhi[0] : (t == 1 ? P.rhi[1] : P.rhi[2])) - lo_t);
      const bool cr_t = t == 0 ? P.crowded[0] : (t == 1 ? P.crowded[1] : P.crowded[2]);
      if (BIG && (t == 0 ? P.big[0] : (t == 1 ? P.big[1] : P.big[2]))) {
        // a big range: its members stay where they are in the list; ranks by bisection
        float x[8];
        unsigned in = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const unsigned idx = (unsigned)(i * 32 + lane);
          x[i] = w.glist[idx];
          const bool m = idx < P.ncand && (unsigned)(fine_tbits_w(w, x[i]) - FINE_BIAS - lo_t) <= sp_t;
          in |= m ? (1u << i) : 0u;
        }
        unsigned tless_b = 0;
        if (P.tie_range == t) {
          int c = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) c += (((in >> i) & 1u) && x[i] < P.tie) ? 1 : 0;
          tless_b = (unsigned)__reduce_add_sync(0xffffffffu, c);
        }
        const unsigned E = t == 0 ? P.rE[0] : (t == 1 ? P.rE[1] : P.rE[2]);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (P.rq[q] != t) continue;
          unsigned rho[2];
          bool is_tie[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            rho[h] = P.rk[2 * q + h] - E;
            is_tie[h] = false;
            if (P.tie_range == t) {
              if (rho[h] >= tless_b && rho[h] < tless_b + P.ntie) is_tie[h] = true;
              else if (rho[h] >= tless_b + P.ntie) rho[h] -= P.ntie;
            }
          }
          // (a rank that falls on the repeated value is not among the gathered members; the bisection
          // then runs for the other rank of the pair)
          const unsigned ra = is_tie[0] ? rho[1] : rho[0], rb = is_tie[1] ? ra : rho[1];
          float va, vb;
          list_select(x, in, ra, rb >= ra ? rb : ra, va, vb);
          res[2 * q] = is_tie[0] ? P.tie : va;
          res[2 * q + 1] = is_tie[1] ? P.tie : (is_tie[0] ? va : vb);
        }
        continue;
      }
      unsigned c = 0;
      if (!cr_t) {
        for (unsigned i0 = 0; i0 < P.ncand; i0 += 32u) {
          const unsigned i = i0 + (unsigned)lane;
          const float x = w.glist[i & 255u];
          const bool in = i < P.ncand && (unsigned)(fine_tbits_w(w, x) - FINE_BIAS - lo_t) <= sp_t;
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) w.clist[(c + __popc(m & lt)) & 63u] = x;
          c += __popc(m);
        }
      }
      __syncwarp();
      src = w.clist;
      cnt = c;
    }
    float ca = (unsigned)lane < cnt ? src[lane] : pinf;
    float cb = (unsigned)(lane + 32) < cnt ? src[lane + 32] : pinf;
    __syncwarp();
    const bool small = cnt <= 32u;                  // the usual case: a handful of candidates (cb = +inf)
    if (small) sort32_regs(ca, lane);
    else sort64_regs(ca, cb, lane);
    const unsigned ia = small ? (unsigned)lane : 2u * (unsigned)lane;        // sorted position of ca, cb
    const unsigned ib = small ? 64u : 2u * (unsigned)lane + 1u;
    // tie range: candidates below the tie value (the sorted list holds the range at [tb, tb + tm))
    unsigned tless = 0;
    if (P.tie_range >= 0 && (!P.split || P.tie_range == t)) {
      const unsigned tb = P.split ? 0u : (P.tie_range == 0 ? P.loff[0] : (P.tie_range == 1 ? P.loff[1] : P.loff[2]));
      const unsigned tm = P.tie_range == 0 ? P.need[0] : (P.tie_range == 1 ? P.need[1] : P.need[2]);
      const bool ina = ia - tb < tm && ca < P.tie;
      const bool inb = ib - tb < tm && cb < P.tie;
      tless = __popc(__ballot_sync(0xffffffffu, ina)) + __popc(__ballot_sync(0xffffffffu, inb));
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int r = P.rq[i >> 1];
      if (P.split && r != t) continue;
      const unsigned bl = P.split ? 0u : (r == 0 ? P.loff[0] : (r == 1 ? P.loff[1] : P.loff[2]));
      const unsigned E = r == 0 ? P.rE[0] : (r == 1 ? P.rE[1] : P.rE[2]);
      unsigned rho = P.rk[i] - E;                   // rank inside the range
      bool is_tie = false;
      if (r == P.tie_range) {
        if (rho >= tless && rho < tless + P.ntie) is_tie = true;
        else if (rho >= tless + P.ntie) rho -= P.ntie;
      }
      const unsigned at = bl + rho;
      const float val = small ? __shfl_sync(0xffffffffu, ca, at & 31u) : sorted64_get(ca, cb, at & 63u);
      res[i] = is_tie ? P.tie : val;
    }
  }
  if (P.crowded[0] || P.crowded[1] || P.crowded[2]) {
    // A crowded range is exactly the set of finite values in [bmn, bmx] (the fine index is
    // monotone), so a rank inside it is an order statistic of that interval.
    sel_count(SEL_CROWDED);
#pragma unroll
    for (int q = 0; q < 3; ++q) {          // (unrolled: constant indices keep the plan in registers)
      const int r = P.rq[q];
      const bool cr = r == 0 ? P.crowded[0] : (r == 1 ? P.crowded[1] : P.crowded[2]);
      if (!cr) continue;
      const float lo = r == 0 ? bmn[0] : (r == 1 ? bmn[1] : bmn[2]);
      const float hi = r == 0 ? bmx[0] : (r == 1 ? bmx[1] : bmx[2]);
      const unsigned nm = r == 0 ? nmin[0] : (r == 1 ? nmin[1] : nmin[2]);
      const unsigned E = r == 0 ? P.rE[0] : (r == 1 ? P.rE[1] : P.rE[2]);
      const unsigned ra = P.rk[2 * q] - E, rb = P.rk[2 * q + 1] - E;
      if (lo == hi || rb < nm) { res[2 * q] = lo; res[2 * q + 1] = lo; continue; }
      select_in_interval<PACK16>(w, lo, hi, ra, rb, res[2 * q], res[2 * q + 1]);
    }
  }
  pct[0] = np_lerp_f(res[2], res[3], P.gam[1]);       // median
  pct[1] = np_lerp_f(res[0], res[1], P.gam[0]);       // 16th
  pct[2] = np_lerp_f(res[4], res[5], P.gam[2]);       // 84th
}

// Fine-index range of a column when the 64-sample estimate is not usable: the exact range.
// Returns false for a column without finite values.
__device__ __forceinline__ bool exact_range(const SelWarp& w, float& scale, float& nLs) {
  float qlo, qhi;
  column_minmax(w, qlo, qhi);
  if (!(qhi >= qlo)) return false;
  if (!(qhi > qlo)) { qlo -= 1.0f; qhi += 1.0f; }       // constant column: any range will do
  else {                                                // keep the end bins for NaN / inf alone
    const float d = (qhi - qlo) * (2.0f / (float)(FB - 4));
    qlo -= d;
    qhi += d;
  }
  scale = 1.0f / (qhi - qlo);
  nLs = -qlo * scale;
  // a span beyond the float range (or a flushed scale): one bin for everything, the rounds on the
  // ordered bit patterns (select_in_interval) take it from there
  if (!finite_f(scale) || !finite_f(nLs) || !(scale > 0.0f)) { scale = 0.0f; nLs = 0.5f; }
  return true;
}

// Whole selection of one column by one warp (mcz.py:273-301); all 32 lanes participate.
// v has Npad (multiple of 128) readable entries; entries in [N, Npad) are NaN.
// hist: FineHist<PACK16>::WORDS words; cand: 128 floats.
// NG: number of 128-sample groups when it is known at compile time (Npad == 128 * NG), else 0.
// ring: optional staging area of the warp for long columns in global memory (RING_FLOATS floats, 16-byte
// aligned): the gather pass then prefetches the column two chunks ahead with cp.async.
// PAIR: TWO warps share the column (launch shapes with at most half as many columns as warps, e.g.
// `--md KD02` on long columns): warp `part` (0 / 1) streams every other 128-sample group in pass H and
// every other chunk in the gather pass, into the SAME histogram and candidate list (`hist`, `cand`:
// the pair's; `ring`: the warp's own); both make the (identical) plan; they meet at the named barrier
// `barid` (64 threads).  Warp 0 finishes and returns the result; a column that needs the general path
// (crowded ranges, end bins) is done by warp 0 alone from the gather pass on.
static constexpr int RING_CH = 6, RING_SLOTS = 3, RING_FLOATS = RING_SLOTS * RING_CH * 128;
template <bool PACK16, int NG, bool PAIR = false>
__device__ void warp_percentiles(const float* __restrict__ v, int Npad, unsigned* hist, float* cand,
                                 float pct[3], int& nvalid, int& status, float* ring = nullptr, int part = 0, int barid = 0) {
  typedef FineHist<PACK16> FH;
  static_assert(!PAIR || NG == 0, "two warps per column: long columns only");
  const int lane = threadIdx.x & 31;
  constexpr int np = PAIR ? 2 : 1;
  auto pair_sync = [&]() { if (PAIR) asm volatile("bar.sync %0, 64;" :: "r"(barid) : "memory"); };
  constexpr int HSTEP = 128 * np;                           // samples between two 128-bit loads of a lane in pass H
  const int hstart = 4 * lane + (PAIR ? 128 * part : 0);
  SelWarp w;
  w.v = v; w.Npad = Npad; w.hist = hist; w.lane = lane;
  w.rnr = 0;
  w.ccount = reinterpret_cast<unsigned*>(cand);
  w.clist = cand + 32;
  w.glist = reinterpret_cast<float*>(hist);
  const float nanf_ = Math<float>::nan();
  pct[0] = pct[1] = pct[2] = nanf_;
  nvalid = 0;
  sel_count(SEL_TOTAL);
  float scale, nLs, tie;
  bool has_tie;
  unsigned ntie_total = 0;
#if defined(MCZ_COLTIME_DUMP) && MCZ_COLTIME_DUMP >= 2
  const long long ts0 = clock64();
#define MCZ_TS(v) const long long v = clock64()
#else
#define MCZ_TS(v)
#endif
  if (!subsample_range(v, lane, scale, nLs, has_tie, tie)) {
    sel_count(SEL_NO_SUBSAMPLE);
    if (!exact_range(w, scale, nLs)) { status = ST_EMPTY; return; }                 // mcz.py:277-280
  }
  w.scale = scale;
  w.nLs = nLs;
  MCZ_TS(ts1);
  // ---- pass H: the 1024-bin histogram (out-of-range finite values land in the end bins) and the
  // sum of the column (the atomics bound this pass, so the extra arithmetic is free here)
  float sum = 0.0f;
  if (!PAIR || part == 0) FH::zero(hist, lane);
  pair_sync();
  const int nH = NG > 0 ? NG * 128 : Npad;      // (compile-time trip count when the length is known)
  const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist);
  const unsigned lane_sa = hist_sa + 4u * (unsigned)(FH::WORDS + lane);
  if (!has_tie) {
    int s0 = hstart;
    if constexpr (NG == 0) {
      // Long column in global memory: the shared-memory reductions are `asm volatile` with a memory
      // clobber, so the compiler keeps every load behind the reductions of the iteration before --
      // one 128-bit load per lane in flight against an L2 / HBM latency of ~1 k cycles.  Four loads
      // are issued before the first of their samples is counted.
      constexpr int UB = 4;
      for (; s0 + (UB - 1) * HSTEP < nH; s0 += UB * HSTEP) {
        float4 xb[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) xb[u] = *reinterpret_cast<const float4*>(v + s0 + u * HSTEP);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const float xs[4] = {xb[u].x, xb[u].y, xb[u].z, xb[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool fin = finite_f(xs[j]);
            sum += fin ? xs[j] : 0.0f;
            FH::add_t(hist_sa, lane_sa, fine_tbits(xs[j], scale, nLs), fin);
          }
        }
      }
    }
#pragma unroll 3
    for (; s0 < nH; s0 += HSTEP) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool fin = finite_f(xs[j]);
        sum += fin ? xs[j] : 0.0f;
        FH::add_t(hist_sa, lane_sa, fine_tbits(xs[j], scale, nLs), fin);            // mcz.py:273
      }
    }
  } else {
    int ntie = 0;
    int s0 = hstart;
    if constexpr (NG == 0) {
      // (long column: four loads ahead, as above -- this variant runs on every E(B-V) column that holds
      // the 1e-5 clamp, a third of the galaxies, and was the slowest column of those galaxies)
      constexpr int UB = 4;
      for (; s0 + (UB - 1) * HSTEP < nH; s0 += UB * HSTEP) {
        float4 xb[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) xb[u] = *reinterpret_cast<const float4*>(v + s0 + u * HSTEP);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const float xs[4] = {xb[u].x, xb[u].y, xb[u].z, xb[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool is_tie = xs[j] == tie;
            const bool fin = finite_f(xs[j]);
            ntie += is_tie ? 1 : 0;
            sum += fin ? xs[j] : 0.0f;
            FH::add_t(hist_sa, lane_sa, fine_tbits(xs[j], scale, nLs), fin && !is_tie);
          }
        }
      }
    }
#pragma unroll 3
    for (; s0 < nH; s0 += HSTEP) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool is_tie = xs[j] == tie;
        const bool fin = finite_f(xs[j]);
        ntie += is_tie ? 1 : 0;
        sum += fin ? xs[j] : 0.0f;
        FH::add_t(hist_sa, lane_sa, fine_tbits(xs[j], scale, nLs), fin && !is_tie);
      }
    }
    ntie = __reduce_add_sync(0xffffffffu, ntie);
    if (lane == 0 && ntie > 0) FH::add_n(hist, fine_index(tie, scale, nLs), (unsigned)ntie);
    ntie_total = (unsigned)ntie;
  }
  double dsum = 0.0;
  if (PAIR) {
    // the two halves of the column sum and of the tie count meet in the spare words of the candidate area
    dsum = warp_sum((double)sum);
    if (lane == 0) {
      reinterpret_cast<double*>(cand + 2)[part] = dsum;
      reinterpret_cast<unsigned*>(cand)[6 + part] = ntie_total;
    }
    pair_sync();                                  // the histogram is complete
    dsum = reinterpret_cast<const double*>(cand + 2)[0] + reinterpret_cast<const double*>(cand + 2)[1];
    ntie_total = reinterpret_cast<const unsigned*>(cand)[6] + reinterpret_cast<const unsigned*>(cand)[7];
  }
  MCZ_TS(ts2);
  SelPlan P;
  constexpr bool BIG = NG == 0;         // the long-column instantiations
  sel_make_plan<PACK16, BIG>(w, P, has_tie, tie, has_tie ? fine_index(tie, scale, nLs) : 0, ntie_total);
  const float skip = P.tie;           // NaN unless a tie value is merged back by rank (never equal then)
  nvalid = P.n;
  if (P.n == 0) { sel_count(SEL_TRIVIAL); status = ST_EMPTY; return; }              // mcz.py:277-280
  // ---- pass G: gather the members of the uncrowded ranges, sum everything
  MCZ_TS(ts3);
  int tlo[3];
  unsigned tsp[3];
  sel_tests(P, true, tlo, tsp);
#pragma unroll
  for (int r = 0; r < 3; ++r) tlo[r] += FINE_BIAS;        // the tests below compare the raw index bits
  __syncwarp();
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
  float bmn[3] = {pinf, pinf, pinf}, bmx[3] = {ninf, ninf, ninf};
  unsigned nmin[3] = {0u, 0u, 0u};
  w.ngroups = Npad / 128;
  w.nmine = Npad;
  bool general = P.crowded[0] || P.crowded[1] || P.crowded[2];
#pragma unroll
  for (int r = 0; r < 3; ++r)                    // NaN / inf index into the end bins
    if (r < P.nr && (P.rlo[r] == 0 || P.rhi[r] == FB - 1)) general = true;
  if (PAIR) {
    if (part == 0 && lane == 0) *w.ccount = 0u;
    pair_sync();                                 // both plans are made: the candidate list may overwrite the histogram
    if (general && part != 0) return;            // (the general path compacts the column in place: one warp)
    __syncwarp();
  }
  if (!general) {
    // The streaming loop only records which samples are members (one bit each, six 128-sample
    // groups per 24-bit mask); the few members are appended after each chunk with one shared
    // atomic per lane that has any.
    constexpr int CH = 6;
    constexpr int NCH = (NG + CH - 1) / CH;
    const int ngroups = Npad / 128;
    if (!PAIR && lane == 0) *w.ccount = 0u;
    __syncwarp();
    bool done = false;
    if constexpr (NG > 0 && NCH <= 4) {
      if (!(skip == skip)) {
        // known column length and no value to leave out: straight-line code, all masks kept, members
        // appended once at the end.  The membership bits are shifted in (mask = 2 * mask + hit: one
        // select and one multiply-add per sample), so sample k of a chunk ends up in bit (nbits - 1 - k).
        unsigned m[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          unsigned mask = 0u;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            if (c * CH + i < NG) {
              const float4 x4 = *reinterpret_cast<const float4*>(v + 4 * lane + 128 * (c * CH + i));
              const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int f = fine_tbits(xs[j], scale, nLs);
                const bool hit = ((unsigned)(f - tlo[0]) <= tsp[0]) || ((unsigned)(f - tlo[1]) <= tsp[1]) ||
                                 ((unsigned)(f - tlo[2]) <= tsp[2]);
                mask = 2u * mask + (hit ? 1u : 0u);
              }
            }
          }
          m[c] = mask;
        }
        unsigned mine = 0u;
#pragma unroll
        for (int c = 0; c < NCH; ++c) mine += (unsigned)__popc(m[c]);
        if (mine) {
          unsigned pos = atomicAdd(w.ccount, mine);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            constexpr int dummy = 0; (void)dummy;
            const int nbits = 4 * ((NG - c * CH) < CH ? (NG - c * CH) : CH);
            unsigned mask = m[c];
            while (mask) {
              const int k = nbits - __ffs(mask);          // sample k of the chunk (group k >> 2, element k & 3)
              mask &= mask - 1u;
              w.glist[pos & 255u] = v[4 * lane + 128 * (c * CH + (k >> 2)) + (k & 3)];
              ++pos;
            }
          }
        }
        done = true;
      }
    }
    if (NG == 0 && ring != nullptr) {
      // Long column in global memory: the lane's own 16-byte pieces of the next two chunks are in flight
      // (cp.async into the lane's slots of the ring) while it tests the current one; nobody else reads
      // those slots, so no synchronisation beyond the lane's own wait_group is needed.
      static_assert(CH == RING_CH, "ring layout");
      const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring) + 16u * (unsigned)lane;
      // (PAIR: this warp takes chunks part, part + 2, ...; c counts the warp's own chunks)
      const int nchunks = PAIR ? ((ngroups + CH - 1) / CH + np - 1 - part) / np : (ngroups + CH - 1) / CH;
      auto chunk0 = [&](int c) { return (PAIR ? np * c + part : c) * CH; };     // first 128-sample group of chunk c
      auto issue = [&](int c, int slot) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int g = chunk0(c) + i;
          if (g < ngroups)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(ring_sa + (unsigned)((slot * CH + i) * 512)),
                         "l"(v + 4 * lane + 128 * g) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      issue(0, 0);
      issue(1, 1);
      int slot = 0;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int nslot = slot == 0 ? 2 : slot - 1;          // the slot of chunk c + 2 = the one chunk c - 1 used
        issue(c + 2, nslot);                                 // (past the end: an empty group keeps the count uniform)
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        unsigned mask = 0u;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          if (chunk0(c) + i < ngroups) {
            float4 x4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x4.x), "=f"(x4.y), "=f"(x4.z), "=f"(x4.w)
                         : "r"(ring_sa + (unsigned)((slot * CH + i) * 512)) : "memory");
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x = xs[j];
              const int f = fine_tbits(x, scale, nLs);            // non-finite x: an end bin, no member here
              const bool hit = ((unsigned)(f - tlo[0]) <= tsp[0]) || ((unsigned)(f - tlo[1]) <= tsp[1]) ||
                               ((unsigned)(f - tlo[2]) <= tsp[2]);
              mask |= (hit && x != skip) ? (1u << (4 * i + j)) : 0u;
            }
          }
        }
        if (mask) {
          unsigned pos = atomicAdd(w.ccount, (unsigned)__popc(mask));
          do {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1u;
            w.glist[pos & 255u] = v[4 * lane + 128 * (chunk0(c) + (b >> 2)) + (b & 3)];
            ++pos;
          } while (mask);
        }
        slot = slot == 2 ? 0 : slot + 1;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      done = true;
    }
    if (!done)
#pragma unroll 1
    for (int g0 = PAIR ? CH * part : 0; g0 < ngroups; g0 += PAIR ? np * CH : CH) {
      unsigned mask = 0u;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (g0 + i < ngroups) {
          const float4 x4 = *reinterpret_cast<const float4*>(v + 4 * lane + 128 * (g0 + i));
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = xs[j];
            const int f = fine_tbits(x, scale, nLs);            // non-finite x: an end bin, no member here
            const bool hit = ((unsigned)(f - tlo[0]) <= tsp[0]) || ((unsigned)(f - tlo[1]) <= tsp[1]) ||
                             ((unsigned)(f - tlo[2]) <= tsp[2]);
            mask |= (hit && x != skip) ? (1u << (4 * i + j)) : 0u;
          }
        }
      }
      if (mask) {
        unsigned pos = atomicAdd(w.ccount, (unsigned)__popc(mask));
        do {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1u;
          w.glist[pos & 255u] = v[4 * lane + 128 * (g0 + (b >> 2)) + (b & 3)];
          ++pos;
        } while (mask);
      }
    }
  } else {
    // Some range is crowded (ties) or reaches an end bin.  Same pass, but with the test for
    // non-finite values, candidates appended by vote, and the members of the crowded ranges
    // are also (a) reduced to min / max / number equal to the min and (b) compacted, in place, to
    // the front of this lane's own slots: from here on the column is only needed for them, so
    // the follow-up rounds of select_in_interval scan a few groups instead of the whole column.
    int clo[3];
    unsigned csp[3];
    sel_tests(P, false, clo, csp);
#pragma unroll
    for (int r = 0; r < 3; ++r) clo[r] += FINE_BIAS;
    int bcm[3] = {0, 0, 0};
    int cw = 0;                                  // members compacted by this lane so far
    float* const vm = const_cast<float*>(v);
    {
      // The warp is usually alone on the SM by now (the other columns are done), so the
      // pass runs at the latency of ONE instruction stream (~7 cycles per instruction): no vote in the
      // loop (a candidate takes its list position with a shared atomic), no divergent branches, loads
      // issued ahead on long columns.  The in-place compaction writes at or before the group that is being handled,
      // never into one whose load is still pending.  (Per crowded column of 11000 samples: 275 k -> 175 k
      // cycles; the statistics alone were 100 k as branches.)
      constexpr int UBG = NG == 0 ? MCZ_GEN_UBG : 1;
      if (!PAIR && lane == 0) *w.ccount = 0u;
      __syncwarp();
      for (int s00 = 4 * lane; s00 < Npad; s00 += 128 * UBG) {
        float4 xg[UBG];
#pragma unroll
        for (int u = 0; u < UBG; ++u)
          xg[u] = s00 + 128 * u < Npad ? *reinterpret_cast<const float4*>(v + s00 + 128 * u)
                                       : make_float4(nanf_, nanf_, nanf_, nanf_);
        // branch-free but for the rare candidate: the statistics of the crowded ranges are updated under
        // selects (ranges that are not crowded have a test that never matches), the compaction store
        // and its counter are predicated.  (Storing EVERY sample at the compaction position and advancing
        // only past members is also correct and saves the predicate, but measured slower: more spills.)
#pragma unroll
        for (int u = 0; u < UBG; ++u) {
          const float xs[4] = {xg[u].x, xg[u].y, xg[u].z, xg[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = xs[j];
            const bool fin = finite_f(x);
            const int f = fine_tbits(x, scale, nLs);
            const bool take = fin && x != skip &&
                              (((unsigned)(f - tlo[0]) <= tsp[0]) || ((unsigned)(f - tlo[1]) <= tsp[1]) ||
                               ((unsigned)(f - tlo[2]) <= tsp[2]));
            bool memb = false;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const bool in = fin && ((unsigned)(f - clo[r]) <= csp[r]);
              const float xm = in ? x : pinf;
              const bool lt = xm < bmn[r];
              const bool eq = in && x == bmn[r];
              bcm[r] = lt ? 1 : (bcm[r] + (eq ? 1 : 0));
              bmn[r] = fminf(bmn[r], xm);
              bmx[r] = fmaxf(bmx[r], in ? x : ninf);
              memb = memb || in;
            }
            if (take) w.glist[atomicAdd(w.ccount, 1u) & 255u] = x;
            if (memb) vm[4 * lane + 128 * (cw >> 2) + (cw & 3)] = x;
            cw += memb ? 1 : 0;
          }
        }
      }
      __syncwarp();
    }
    int mxc = cw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mxc = max(mxc, __shfl_xor_sync(0xffffffffu, mxc, o));
    w.ngroups = (mxc + 3) >> 2;
    w.nmine = cw;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float gm = bmn[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        bmx[r] = fmaxf(bmx[r], __shfl_xor_sync(0xffffffffu, bmx[r], o));
      }
      nmin[r] = (unsigned)__reduce_add_sync(0xffffffffu, bmn[r] == gm ? bcm[r] : 0);
      bmn[r] = gm;
    }
  }
  if (PAIR && !general) {
    pair_sync();                                 // every candidate is in the list
    if (part != 0) return;
  }
  if (!PAIR) dsum = warp_sum((double)sum);
  if (!(dsum > 0.0)) { sel_count(SEL_TRIVIAL); status = ST_NONPOS; return; }   // mcz.py:282-285
  status = ST_OK;
  MCZ_TS(ts4);
  sel_count(SEL_FAST);
#ifdef MCZ_COLTIME_DUMP
  w.dbg = (P.crowded[0] ? 1 : 0) | (P.crowded[1] ? 2 : 0) | (P.crowded[2] ? 4 : 0) | (P.tie_range >= 0 ? 8 : 0) |
          (has_tie ? 16 : 0) | (P.nr << 5);
#endif
  sel_finish<PACK16, BIG>(w, P, bmn, bmx, nmin, pct);
#ifdef MCZ_COLTIME_DUMP
  nvalid |= w.dbg << 16;
#endif
#if defined(MCZ_COLTIME_DUMP) && MCZ_COLTIME_DUMP == 2
  pct[0] = (float)(ts2 - ts0); pct[1] = (float)(ts3 - ts2); pct[2] = (float)(ts4 - ts3);   // range + pass H, plan, pass G
#elif defined(MCZ_COLTIME_DUMP) && MCZ_COLTIME_DUMP == 3
  pct[0] = (float)(ts1 - ts0); pct[1] = (float)(clock64() - ts4); pct[2] = (float)(clock64() - ts0);   // range, finish, total
#endif
}

// ---- cluster kernel: the selection of one column whose samples are spread over the blocks ------
// Block r of the cluster holds piece r of every column (`ng` groups of 128 samples, NaN beyond the
// block's share).  Warp w of EVERY block works on column slot w; one of the blocks is the column's
// OWNER (slot mod cluster size).  The five or six warps of a column talk to each other through the
// owner's shared memory with asynchronous DSMEM stores / reductions that signal mbarriers -- there is
// no cluster-wide barrier inside the selection, so the columns of a galaxy proceed independently,
// as in the one-block kernel:
//   all     range from the column's first 64 samples (block 0's piece, read through DSMEM);
//           pass H over the own piece into the own 1024-bin table;
//   peers   push the table into the owner's (red.async.add, 2 KB) plus count of the repeated value, sum,
//           extremes (st.async); wait for the owner's command;
//   owner   waits until every peer's 2064 bytes have landed; plan from the merged table; command to
//           the peers (64 bytes each): GATHER (ranges), REFINE (a second round with a finer index),
//           FALLBACK (whole column through global memory) or NONE;
//   all     pass G over the own piece: members appended to the owner's candidate list (remote atomic
//           for the position, st.async for the values, counted in bytes by the owner's third barrier);
//   owner   sorts the candidates and interpolates, exactly as the one-block kernel does.
// A column with a crowded range gets ONE second round (the target ranges subdivided 339-fold, or the
// exact range for a column whose first 64 samples gave none); if that is still not enough, every
// block copies its piece to a global scratch column and the owner runs the one-block selection
// (warp_percentiles) on it: rare, and every special case keeps the code the adversarial tests cover.
enum { CLM_NONE = 0, CLM_GATHER = 1, CLM_REFINE_RANGE = 2, CLM_REFINE_SUB = 3, CLM_FALLBACK = 4 };
static constexpr unsigned CL_CMD_BYTES = 64u;
static constexpr unsigned CL_PUSH_BYTES = FineHist<true>::WORDS * 4u + 16u;

// stage timers of the cluster selection (-DMCZ_PHASE_TIMING): cycles summed over columns into
// g_col_cycles[stage] (owner) / g_col_cycles[16 + stage] (peers); [15] / [31] count the columns
#ifdef MCZ_PHASE_TIMING
#define CLT_MARK(stage) do { const long long now_ = clock64(); if (lane == 0) atomicAdd(&g_col_cycles[(own ? 0 : 16) + (stage)], (unsigned long long)(now_ - tmark_)); tmark_ = now_; } while (0)
#define CLT_START() long long tmark_ = clock64(); if (lane == 0) atomicAdd(&g_col_cycles[(own ? 15 : 31)], 1ull)
#else
#define CLT_MARK(stage)
#define CLT_START()
#endif

struct ClusterCol {
  typedef FineHist<true> FH;
  unsigned crank, csize, owner;
  int ng;                       // groups of 128 samples every block streams
  unsigned cand_sa, hist_sa, glist_sa;
  unsigned long long* poison_word;
  unsigned par;                 // bit 0: merge barrier, bit 1: command barrier, bit 2: gather barrier
  // diagnostics
  bool did_refine, did_fallback;

  // pass H over the own piece with the current index of `w`; the table must be zero (or hold the
  // peers' contributions: the owner's table is zeroed before anything can arrive)
  __device__ __forceinline__ void histogram(const SelWarp& w, bool has_tie, float tie, bool track,
                                            float& sum, unsigned& ntie, float& mn, float& mx) const {
    const int lane = w.lane;
    const unsigned lane_sa = hist_sa + 4u * (unsigned)(FH::WORDS + lane);
    const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
    float sm = 0.0f;
    int nt = 0;
    mn = pinf; mx = ninf;
    if (w.rnr == 0 && !track) {
#pragma unroll 3
      for (int g = 0; g < ng; ++g) {
        const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * lane + 128 * g);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool is_tie = has_tie && xs[j] == tie;
          const bool fin = finite_f(xs[j]);
          nt += is_tie ? 1 : 0;
          sm += fin ? xs[j] : 0.0f;
          FH::add_t(hist_sa, lane_sa, fine_tbits(xs[j], w.scale, w.nLs), fin && !is_tie);
        }
      }
    } else {
#pragma unroll 2
      for (int g = 0; g < ng; ++g) {
        const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * lane + 128 * g);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool is_tie = has_tie && xs[j] == tie;
          const bool fin = finite_f(xs[j]);
          nt += is_tie ? 1 : 0;
          sm += fin ? xs[j] : 0.0f;
          mn = fminf(mn, fin ? xs[j] : pinf);
          mx = fmaxf(mx, fin ? xs[j] : ninf);
          FH::add_t(hist_sa, lane_sa, fine_tbits_w(w, xs[j]), fin && !is_tie);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
    }
    nt = __reduce_add_sync(0xffffffffu, nt);
    sum = warp_sum(sm);
    ntie = (unsigned)nt;
    __syncwarp();
    if (lane == 0 && nt > 0) FH::add_n(w.hist, fine_tbits_w(w, tie) - FINE_BIAS, (unsigned)nt);
    __syncwarp();
  }

  // pass G over the own piece: members of the (uncrowded) target ranges -> the owner's candidate list.
  // Returns how many members this WARP appended.
  __device__ __forceinline__ unsigned gather(const SelWarp& w, const int tlo_[3], const unsigned tsp[3], float skip) const {
    const int lane = w.lane;
    const bool own = crank == owner;
    int tlo[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) tlo[r] = tlo_[r] + FINE_BIAS;
    constexpr int CH = 6, NCH = 3;                    // <= 18 groups
    unsigned m[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      unsigned mask = 0u;
#pragma unroll 2
      for (int i = 0; i < CH; ++i) {
        if (c * CH + i < ng) {
          const float4 x4 = *reinterpret_cast<const float4*>(w.v + 4 * lane + 128 * (c * CH + i));
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int f = w.rnr == 0 ? fine_tbits(xs[j], w.scale, w.nLs) : fine_tbits_w(w, xs[j]);
            const bool hit = ((unsigned)(f - tlo[0]) <= tsp[0]) || ((unsigned)(f - tlo[1]) <= tsp[1]) ||
                             ((unsigned)(f - tlo[2]) <= tsp[2]);
            mask |= (hit && finite_f(xs[j]) && xs[j] != skip) ? (1u << (4 * i + j)) : 0u;
          }
        }
      }
      m[c] = mask;
    }
    unsigned mine = 0u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) mine += (unsigned)__popc(m[c]);
    if (mine) {
      const unsigned cap = (unsigned)(CL_GLIST_FLOATS - 1);
      if (own) {
        unsigned pos = atomicAdd(w.ccount + 1, mine);               // (gcount)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          unsigned mask = m[c];
          while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1u;
            w.glist[min(pos, cap)] = w.v[4 * lane + 128 * (c * CH + (b >> 2)) + (b & 3)];
            ++pos;
          }
        }
      } else {
        const unsigned o_gcount = cl_map(cand_sa + ClMail::GCOUNT, owner), o_glist = cl_map(glist_sa, owner);
        const unsigned o_gather = cl_map(cand_sa + ClMail::MB_GATHER, owner);
        unsigned pos = cl_atom_add(o_gcount, mine);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          unsigned mask = m[c];
          while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1u;
            st_async_u32(o_glist + 4u * min(pos, cap), __float_as_uint(w.v[4 * lane + 128 * (c * CH + (b >> 2)) + (b & 3)]), o_gather);
            ++pos;
          }
        }
      }
    }
    return (unsigned)__reduce_add_sync(0xffffffffu, (int)mine);
  }

  // The whole protocol for one column; every lane of the column's warp in every block calls it.
  // col: the block's piece; hist / cand / glist: the warp's areas; gcol: the cluster's global scratch
  // column (csize * ng * 128 floats).  The owner returns the statistics (is_owner = true).
  __device__ void run(const float* col, unsigned* hist, float* cand, float* glist, float* gcol,
                      float pct[3], int& nvalid, int& status) {
    const int lane = threadIdx.x & 31;
    const bool own = crank == owner;
    const float nanf_ = Math<float>::nan();
    pct[0] = pct[1] = pct[2] = nanf_;
    nvalid = 0;
    status = ST_EMPTY;
    did_refine = false; did_fallback = false;
    SelWarp w;
    w.v = col; w.Npad = ng * 128; w.hist = hist; w.lane = lane;
    w.ccount = reinterpret_cast<unsigned*>(cand); w.clist = cand + 32; w.glist = glist;
    w.scale = 0.0f; w.nLs = 0.5f; w.ngroups = ng; w.nmine = ng * 128; w.rnr = 0;
    const unsigned mb_merge = cand_sa + ClMail::MB_MERGE, mb_cmd = cand_sa + ClMail::MB_CMD, mb_gather = cand_sa + ClMail::MB_GATHER;
    unsigned* const cmdw = reinterpret_cast<unsigned*>(cand) + ClMail::CMD / 4;
    unsigned* const statw = reinterpret_cast<unsigned*>(cand) + ClMail::STAT / 4;
    if (lane == 0) w.ccount[1] = 0u;                   // gcount (appends come after two message flights)
    CLT_START();
    // ---- range of the first round from the column's first 64 samples (block 0's piece)
    float a, b;
    if (crank == 0u) { a = col[lane]; b = col[lane + 32]; }
    else { const unsigned ra = cl_map(col, 0u); a = cl_ldf(ra + 4u * (unsigned)lane); b = cl_ldf(ra + 4u * (unsigned)(lane + 32)); }
    bool has_tie;
    float tie;
    // No usable range (usually: no finite value at all, or one repeated value, which `has_tie` covers):
    // one bin for everything in the first round, which also reduces the exact extremes of the column;
    // a column that does hold a spread of finite values gets its histogram in the second round.
    const bool norange = !subsample_range_ab(a, b, lane, w.scale, w.nLs, has_tie, tie);
    if (norange) { w.scale = 0.0f; w.nLs = 0.5f; }
    CLT_MARK(0);                                       // range
    SelPlan P;
    float sum_total = 0.0f;
    for (int round = 0;; ++round) {
      float sum, mn, mx;
      unsigned ntie;
      if (!own && round > 0) FH::zero(hist, lane);     // (first round: zeroed before the barrier that precedes the selection)
      histogram(w, has_tie, tie, norange && round == 0, sum, ntie, mn, mx);
      CLT_MARK(1);                                     // pass H
      if (!own) {
        // ---- peer: table and statistics to the owner, then wait for the command
        const unsigned o_hist = cl_map(hist_sa, owner), o_merge = cl_map(mb_merge, owner);
#pragma unroll
        for (int i = 0; i < FH::WORDS / 32; ++i) red_async_add(o_hist + 4u * (unsigned)(i * 32 + lane), hist[i * 32 + lane], o_merge);
        if (lane == 0) {
          st_async_v4(cl_map(cand_sa + ClMail::STAT + 16u * crank, owner), ntie, __float_as_uint(sum), __float_as_uint(mn),
                      __float_as_uint(mx), o_merge);
          mbar_expect_tx_arrive(mb_cmd, CL_CMD_BYTES);
        }
        CLT_MARK(2);                                   // push
        mbar_wait(mb_cmd, (par >> 1) & 1u, poison_word, 3u);
        par ^= 2u;
        CLT_MARK(3);                                   // wait for the command
        const unsigned mode = cmdw[0];
        if (mode == CLM_GATHER) {
          int tlo[3];
          unsigned tsp[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) { tlo[r] = (int)cmdw[1 + r]; tsp[r] = cmdw[4 + r]; }
          gather(w, tlo, tsp, __uint_as_float(cmdw[7]));
          CLT_MARK(5);                                 // pass G
          return;
        }
        if (mode == CLM_REFINE_RANGE) {
          w.scale = __uint_as_float(cmdw[11]); w.nLs = __uint_as_float(cmdw[12]);
          did_refine = true;
          continue;
        }
        if (mode == CLM_REFINE_SUB) {
          w.rnr = (int)cmdw[7];
#pragma unroll
          for (int r = 0; r < 3; ++r) { w.rflo[r] = (int)cmdw[1 + r]; w.rfhi[r] = (int)cmdw[4 + r]; w.rmul[r] = __uint_as_float(cmdw[8 + r]); }
          did_refine = true;
          continue;
        }
        if (mode == CLM_FALLBACK) {
          for (int s = lane; s < ng * 128; s += 32) gcol[(int)crank * ng * 128 + s] = col[s];
          __threadfence();
          __syncwarp();
          if (lane == 0) st_async_u32(cl_map(cand_sa + ClMail::STAT + 16u * crank, owner), 0u, cl_map(mb_gather, owner));
          did_fallback = true;
        }
        return;                                        // CLM_NONE: nothing to gather
      }
      // ---- owner: merged table, plan, command
      if (lane == 0) {
        statw[4 * crank] = ntie; statw[4 * crank + 1] = __float_as_uint(sum);
        statw[4 * crank + 2] = __float_as_uint(mn); statw[4 * crank + 3] = __float_as_uint(mx);
        mbar_expect_tx_arrive(mb_merge, (csize - 1u) * CL_PUSH_BYTES);
      }
      mbar_wait(mb_merge, par & 1u, poison_word, 2u);
      par ^= 1u;
      CLT_MARK(3);                                     // wait for the peers' tables
      __syncwarp();
      unsigned ntie_total = 0u;
      float gmin = __int_as_float(0x7f800000), gmax = __int_as_float(0xff800000);
      sum_total = 0.0f;
      for (unsigned r = 0; r < csize; ++r) {           // rank order: the same float total whoever owns the column
        const uint4 sv = *reinterpret_cast<const uint4*>(statw + 4 * r);
        ntie_total += sv.x;
        sum_total += __uint_as_float(sv.y);
        gmin = fminf(gmin, __uint_as_float(sv.z));
        gmax = fmaxf(gmax, __uint_as_float(sv.w));
      }
      sel_make_plan<true>(w, P, has_tie, tie, has_tie ? fine_tbits_w(w, tie) - FINE_BIAS : 0, ntie_total);
      unsigned mode = CLM_GATHER;
      unsigned cw[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) cw[i] = 0u;
      if (P.n == 0) {
        mode = CLM_NONE;
      } else {
        const bool crowded = P.crowded[0] || P.crowded[1] || P.crowded[2];
        bool edge = false;
#pragma unroll
        for (int r = 0; r < 3; ++r)
          if (r < P.nr && (P.rlo[r] == 0 || P.rhi[r] == FB - 1)) edge = true;      // NaN / inf index into the end bins
        if (round == 0 && norange && crowded && gmax > gmin) {
          // second round over the exact range of the column (as exact_range() of the one-block selection)
          mode = CLM_REFINE_RANGE;
          const float d = (gmax - gmin) * (2.0f / (float)(FB - 4));
          const float qlo = gmin - d, qhi = gmax + d;
          w.scale = 1.0f / (qhi - qlo);
          w.nLs = -qlo * w.scale;
          if (!finite_f(w.scale) || !finite_f(w.nLs) || !(w.scale > 0.0f)) { w.scale = 0.0f; w.nLs = 0.5f; }
          cw[11] = __float_as_uint(w.scale); cw[12] = __float_as_uint(w.nLs);
        } else if (round == 0 && crowded && !edge) {
          // second round: the target ranges subdivided
          mode = CLM_REFINE_SUB;
          w.rnr = P.nr;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            w.rflo[r] = r < P.nr ? P.rlo[r] : 0;
            w.rfhi[r] = r < P.nr ? P.rhi[r] : 0;
            w.rmul[r] = r < P.nr ? (float)RSUB / (float)(P.rhi[r] - P.rlo[r] + 1) : 0.0f;
            cw[1 + r] = (unsigned)w.rflo[r]; cw[4 + r] = (unsigned)w.rfhi[r]; cw[8 + r] = __float_as_uint(w.rmul[r]);
          }
          cw[7] = (unsigned)w.rnr;
        } else if (crowded || edge) {
          mode = CLM_FALLBACK;
        } else {
          int tlo[3];
          unsigned tsp[3];
          sel_tests(P, true, tlo, tsp);
#pragma unroll
          for (int r = 0; r < 3; ++r) { cw[1 + r] = (unsigned)tlo[r]; cw[4 + r] = tsp[r]; }
          cw[7] = __float_as_uint(P.tie);              // NaN unless the repeated value is merged back by rank
        }
      }
      CLT_MARK(4);                                     // plan
      cw[0] = mode;
      if (mode == CLM_REFINE_RANGE || mode == CLM_REFINE_SUB) {
        FH::zero(hist, lane);                          // before any peer can push its second-round table
        did_refine = true;
      }
      __syncwarp();
      if ((unsigned)lane < csize && (unsigned)lane != crank) {
        const unsigned pc = cl_map(cand_sa + ClMail::CMD, (unsigned)lane), pm = cl_map(mb_cmd, (unsigned)lane);
        st_async_v4(pc, cw[0], cw[1], cw[2], cw[3], pm);
        st_async_v4(pc + 16u, cw[4], cw[5], cw[6], cw[7], pm);
        st_async_v4(pc + 32u, cw[8], cw[9], cw[10], cw[11], pm);
        st_async_v4(pc + 48u, cw[12], cw[13], cw[14], cw[15], pm);
      }
      if (mode == CLM_REFINE_RANGE || mode == CLM_REFINE_SUB) continue;
      if (mode == CLM_NONE) return;                                                     // mcz.py:277-280
      if (mode == CLM_FALLBACK) {
        did_fallback = true;
        for (int s = lane; s < ng * 128; s += 32) gcol[(int)crank * ng * 128 + s] = col[s];
        if (lane == 0) mbar_expect_tx_arrive(mb_gather, 4u * (csize - 1u));
        mbar_wait(mb_gather, (par >> 2) & 1u, poison_word, 5u);
        par ^= 4u;
        __threadfence();                               // (acquire side: drops stale L1 lines of the scratch column)
        __syncwarp();
        warp_percentiles<true, 0>(gcol, (int)csize * ng * 128, hist, cand, pct, nvalid, status);
        CLT_MARK(8);                                   // whole-column path
        return;
      }
      // ---- gather: own members, then wait for the peers' (counted in bytes)
      {
        int tlo[3];
        unsigned tsp[3];
        sel_tests(P, true, tlo, tsp);
        const unsigned ownc = gather(w, tlo, tsp, P.tie);
        CLT_MARK(5);                                   // pass G
        if (lane == 0) mbar_expect_tx_arrive(mb_gather, 4u * (P.ncand - min(ownc, P.ncand)));
        mbar_wait(mb_gather, (par >> 2) & 1u, poison_word, 4u);
        par ^= 4u;
        __syncwarp();
        CLT_MARK(6);                                   // wait for the peers' candidates
      }
      nvalid = P.n;
      if (!(sum_total > 0.0f)) { status = ST_NONPOS; return; }                          // mcz.py:282-285
      status = ST_OK;
      const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);
      const float bmn[3] = {pinf, pinf, pinf}, bmx[3] = {ninf, ninf, ninf};
      const unsigned nmin[3] = {0u, 0u, 0u};
      sel_finish<true>(w, P, bmn, bmx, nmin, pct);
      CLT_MARK(7);                                     // finish
      return;
    }
  }
};

}  // namespace mcz
