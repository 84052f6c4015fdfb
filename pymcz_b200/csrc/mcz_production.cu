// Fused production kernel of the pyMCZ hot path: for every galaxy, in ONE persistent kernel,
//   A. Philox4x32-10 + Box-Muller perturbation of the line fluxes (reference mcz.py:509-519,
//      436-442, 580-589), sanitising (metallicity.py:96-99), has* flags (metscales.py:287-378),
//      E(B-V) + dust, every closed-form scale and the root-based ones (metscales.py:387-1087);
//   B. the two KK04 fixed-point loops whose trip count is a per-galaxy mean over all samples
//      (metscales.py:1111-1121, 1166-1181), state held in registers;
//   C. KK04 validity masks and KD02comb (metscales.py:1127-1129, 1186, 1250-1261);
//   D. per (galaxy, scale) finite filter, n, sum<=0 test and the 50/16/84th percentiles with
//      numpy's linear interpolation (mcz.py:273-301), one warp per scale.
// Layout: one thread block per galaxy (dynamic galaxy queue; the next galaxy's ticket and inputs
// are fetched by the 18th warp during phase D), per-scale sample columns staged in shared memory
// [slot][sample] (conflict-free: consecutive threads own consecutive samples); nothing but the
// 2x11 input floats and the percentile table touches HBM.  production_kernel<0, WIDE> keeps the
// columns in an L2-resident global scratch for N' too large for 227 KB (<0, false, false, true>: two
// warps per column in phase D, for launches that store at most nine columns).  With several GPUs the
// result rows are written into every rank's full table (ProdParams::ndst): the gather of the
// path is fused into the kernel.
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <cooperative_groups.h>
#include <atomic>
#include <type_traits>
#include "mcz_physics.cuh"
#include "mcz_philox.cuh"
#include "mcz_device_utils.cuh"
#include "mcz_internal.h"

namespace mcz {

static constexpr int PT = 576;            // threads per block: 18 warps
static constexpr int PW = PT / 32;
static constexpr int CAND_FLOATS = 128;   // per-warp candidate area (counter + 64-entry list)
static constexpr int CL_GLIST_FLOATS = 192;   // cluster kernel: per-warp candidate list of the owner block (3 ranges x 64)
static constexpr int CTRL_BYTES = 1024;
static constexpr int MAX_DST = 8;                 // ranks of one node
static constexpr uint32_t SEEN_KD = 1u << 8;     // ctrl->flags bits above the eight line flags: some sample has a KD02_N2O2 value
static constexpr int SCRATCH_HEADER = 256;       // scratch_d starts with: galaxy queue counter (u64) at 0, launch statistics at 64
static constexpr int SCRATCH_STATS_OFF = 64;
#ifndef MCZ_DIV_TRIP
#define MCZ_DIV_TRIP 6
#endif
static constexpr int DIVERGENCE_CHECK_TRIP = MCZ_DIV_TRIP;   // when to test a still-running N2Ha loop for provable divergence

struct ProdParams {
  const float* meas;
  const float* err;
  long long G;
  int N, Npad;
  uint32_t tok;
  int dust;
  uint32_t seed_lo, seed_hi;
  unsigned long long goff;
  int correlated, deterministic;
  uint32_t calls;                          // bit c set: Philox call c is needed
  // result tables: ndst destinations (1 = the caller's local tables; > 1 = the gathered tables of
  // every rank, written over NVLink peer mappings), rows start at row0 in each of them
  int ndst;
  long long row0;
  float* pct[MAX_DST];
  int* nvalid[MAX_DST];
  signed char* status[MAX_DST];
  int* trips;
  int* ret;
  float* Z;
  int nstore;
  int scap;                                // global variant: samples whose KK04 loop state lives in shared memory
  int bsm;                                 // global variant: the inputs and the state of the KK04 loops fit in shared memory
  int ring;                                // global variant: per-warp prefetch rings for the gather pass
  int pair;                                // global variant: two warps per column in phase D (nstore <= PW / 2)
  int cper, cng;                           // cluster kernel: samples per block, 128-sample groups streamed per block
  signed char slot_of_key[NKEYS];
  signed char key_of_slot[NKEYS];
  int col_off[NKEYS];                      // slot * Npad (float index of the column), -1 if not stored
  // wide kernel: the galaxy's samples may also be sharded over the GPUs of a node; wshared[q] is rank
  // q's control block + merged histograms + barrier counter as mapped into this process
  int wrank, wworld;
  void* wshared[MAX_DST];
  unsigned long long* counter;
  unsigned long long* stats;               // scratch header: [0] galaxies, [1] N2Ha trips executed, [2] R23 trips executed, [3] N2Ha early exits, [4] R23 early exits
  float* gscratch;                         // global variant: per-block columns + stash + loop state
  long long gscratch_per_block;            // floats
};

struct Ctrl {
  // cluster kernel: exchange of the per-block KK04 sums.  Two barriers / slot sets, used alternately;
  // xslot[p][r] = what block r sent: four fixed-point sums, "changed" flag, NaN flag (8 words, 32 bytes)
  unsigned long long xbar[2];
  int xslot[2][8][8];
  long long g;
  uint32_t flags;
  int any_valid_kd;
  float line[2 * NLINES];
  float2 redf[2 * PW];
  int acc[3][4];           // fixed-point block sums: two KK04 convergence measures x two trips per barrier
  unsigned chg[3];         // some sample's R23 loop state changed over the batch (rotates with acc)
  unsigned nanw[3];        // cluster kernel: some logq of the batch is NaN in this block
  unsigned nan4;           // which loop / trip saw a NaN (rare path)
  unsigned long long* poison_word;
  int cl_soff, cl_n;       // cluster kernel: this block's first sample and sample count
};
static_assert(sizeof(Ctrl) <= CTRL_BYTES, "control block too large");

// ---- sampling ---------------------------------------------------------------------------------
// Normal index j: 0..4 coefficient noise, 5..12 the consumed lines in `Used` order.
__device__ __forceinline__ void draw_normals(uint32_t s, unsigned long long gg, const ProdParams& p,
                                             float z[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (p.calls & (1u << c)) {
      uint32_t o[4];
      philox4x32_10(s, (uint32_t)c, (uint32_t)gg, (uint32_t)(gg >> 32), p.seed_lo, p.seed_hi, o);
      box_muller(o[0], o[1], z[4 * c + 0], z[4 * c + 1]);
      box_muller(o[2], o[3], z[4 * c + 2], z[4 * c + 3]);
    } else {
      z[4 * c + 0] = z[4 * c + 1] = z[4 * c + 2] = z[4 * c + 3] = 0.0f;
    }
  }
}

// the same with the set of Philox calls known at compile time
template <uint32_t CALLS>
__device__ __forceinline__ void draw_normals_c(uint32_t s, unsigned long long gg, const ProdParams& p, float z[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (CALLS & (1u << c)) {
      uint32_t o[4];
      philox4x32_10(s, (uint32_t)c, (uint32_t)gg, (uint32_t)(gg >> 32), p.seed_lo, p.seed_hi, o);
      box_muller(o[0], o[1], z[4 * c + 0], z[4 * c + 1]);
      box_muller(o[2], o[3], z[4 * c + 2], z[4 * c + 3]);
    } else {
      z[4 * c + 0] = z[4 * c + 1] = z[4 * c + 2] = z[4 * c + 3] = 0.0f;
    }
  }
}

// Block-wide sums of two floats, every thread gets both totals (NaN propagates); one barrier per
// call thanks to the alternating buffer.
__device__ __forceinline__ void block_sum2f(float& a, float& c, float2* buf, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  c = warp_sum(c);
  float2* b = buf + phase * PW;
  if (lane == 0) b[warp] = make_float2(a, c);
  __syncthreads();
  const float2 pr = lane < PW ? b[lane] : make_float2(0.0f, 0.0f);
  a = warp_sum(pr.x);
  c = warp_sum(pr.y);
  phase ^= 1;
}

// The same for four floats behind ONE barrier; every total is formed in the same order as
// block_sum2f forms it, so the two are interchangeable bit for bit.  bufA, bufB: 2 * PW float2 each.
__device__ __forceinline__ void block_sum4f(float (&v)[4], float2* bufA, float2* bufB, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = warp_sum(v[i]);
  float2* a = bufA + phase * PW;
  float2* b = bufB + phase * PW;
  if (lane == 0) { a[warp] = make_float2(v[0], v[1]); b[warp] = make_float2(v[2], v[3]); }
  __syncthreads();
  const float2 pa = lane < PW ? a[lane] : make_float2(0.0f, 0.0f);
  const float2 pb = lane < PW ? b[lane] : make_float2(0.0f, 0.0f);
  v[0] = warp_sum(pa.x); v[1] = warp_sum(pa.y);
  v[2] = warp_sum(pb.x); v[3] = warp_sum(pb.y);
  phase ^= 1;
}

// ---- thread-block clusters: distributed shared memory ----------------------------------------
// For N' > 2304 a galaxy is given to a CLUSTER of ceil(N' / 2304) thread blocks: block r holds samples
// [2304 r, 2304 (r + 1)) in its own shared memory, exactly like the one-block kernel, and what the
// path needs from the whole sample vector (the has* flags, the KK04 loop means per trip, the
// histograms and candidates of the selection) is exchanged through the blocks' shared memories
// (mapa / st|ld|red|atom.shared::cluster) between cluster barriers.
__device__ __forceinline__ unsigned cl_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// shared-space address of the same variable in block `rank` of the cluster
__device__ __forceinline__ unsigned cl_map(unsigned saddr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ unsigned cl_map(const void* p, unsigned rank) { return cl_map((unsigned)__cvta_generic_to_shared(p), rank); }
__device__ __forceinline__ void cl_st(unsigned addr, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_stf(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ unsigned cl_ld(unsigned addr) { unsigned v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ float cl_ldf(unsigned addr) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void cl_red_add(unsigned addr, int v) { asm volatile("red.shared::cluster.add.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_red_or(unsigned addr, unsigned v) { asm volatile("red.shared::cluster.or.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned cl_atom_add(unsigned addr, unsigned v) {
  unsigned o;
  asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory");
  return o;
}
// mbarriers (shared memory, 8 bytes) with transaction counts: the receiving end of the asynchronous
// remote stores / reductions below.  A barrier is initialised with ONE expected arrival -- the
// receiver's own arrive.expect_tx(bytes) -- and its phase completes when that arrival has happened
// and `bytes` of remote data have landed (data that lands before the expect_tx just drives the
// transaction count negative for a while).
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_arrive(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0u;
}
// Wait for phase `parity` of a local barrier.  A protocol error (the blocks of a cluster disagreeing
// on what comes next) would otherwise hang the GPU: after ~1 s the wait gives up, raises the poison
// flag in the launch statistics (checked by the host) and every later wait returns at once.
__device__ __noinline__ void mbar_wait_slow(unsigned a, unsigned parity, unsigned long long* poison_word, unsigned where) {
  const long long t0 = clock64();
  for (;;) {
    if ((*reinterpret_cast<volatile unsigned long long*>(poison_word) >> 63) != 0ull) return;
    for (int it = 0; it < 16; ++it)
      if (mbar_try_wait(a, parity)) return;
    if (clock64() - t0 > 1000000000ll) {
      const unsigned long long old = atomicOr(poison_word, 1ull << 63);
      if (!(old >> 63)) atomicOr(poison_word, (unsigned long long)where << 56);      // who gave up first
      return;
    }
  }
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity, unsigned long long* poison_word, unsigned where) {
#pragma unroll 1
  for (int it = 0; it < 64; ++it)
    if (mbar_try_wait(a, parity)) return;
  mbar_wait_slow(a, parity, poison_word, where);
}
// asynchronous remote stores / reductions (shared::cluster addresses) that signal `mbar` (a
// shared::cluster address in the SAME block as the data) with the number of bytes written
__device__ __forceinline__ void st_async_v4(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_u32(unsigned addr, unsigned a, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" :: "r"(addr), "r"(a), "r"(mbar) : "memory");
}
__device__ __forceinline__ void red_async_add(unsigned addr, unsigned a, unsigned mbar) {
  asm volatile("red.async.relaxed.cluster.mbarrier::complete_tx::bytes.shared::cluster.add.u32 [%0], %1, [%2];"
               :: "r"(addr), "r"(a), "r"(mbar) : "memory");
}
// byte offsets inside a warp's 512-byte candidate area: counters, the three barriers of the selection
// protocol (ClusterCol), the command from the owner, the 64-entry list of select_in_interval, and
// what each block reports with its table (count of the repeated value, sum, min, max)
struct ClMail {
  static constexpr unsigned CCOUNT = 0, GCOUNT = 4, MB_MERGE = 8, MB_CMD = 16, MB_GATHER = 24, CMD = 32, PAR = 96, CLIST = 128, STAT = 384;
};
// barrier over every thread of every block of the cluster; orders shared AND global memory accesses
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool CL> __device__ __forceinline__ void block_or_cluster_sync() {
  if (CL) cl_sync(); else __syncthreads();
}

// Block-wide sums of the convergence measures in fixed point with FX_BITS fractional bits: the
// thread's partial sum (<= 4 samples, each clamped to +-8 by the caller, which cannot change the
// outcome of "mean > tol" since tol * N < 8) is rounded to 2^-FX_BITS, one REDUX per warp
// (|warp total| <= 128 * 8 * 2^20 = 2^30), the warp total clamped to +-64 (2^26: only a warp whose
// samples are nowhere near convergence gets there, and 18 such totals still fit 32 bits), one
// 32-bit shared atomic per warp (a 64-bit shared atomicAdd is a CAS loop), and a single barrier
// that also ORs the NaN flags (BAR.RED.OR).  The rounding error of the block total is ~7 quanta
// (7e-6 at 20 bits, against thresholds tol * N of 0.22 and 2.2): below the float32 rounding of the
// per-sample terms themselves, so the integer total is an order-independent, deterministic
// stand-in for the reference's float64 mean (metscales.py:1117, 1177).
#ifndef MCZ_CLUSTER_DEFAULT
#define MCZ_CLUSTER_DEFAULT 0
#endif
#ifndef MCZ_A_UNROLL
#define MCZ_A_UNROLL 1
#endif
static constexpr int A_UNROLL = MCZ_A_UNROLL;     // samples of a thread evaluated side by side in phase A
#ifndef MCZ_BSM_UNROLL
#define MCZ_BSM_UNROLL 2
#endif
#ifndef MCZ_BSM_TRIPS
#define MCZ_BSM_TRIPS 2
#endif
static constexpr int BSM_TRIPS = MCZ_BSM_TRIPS;     // trips per block-wide reduction in the shared-memory KK04 loop (2 or 3)
static constexpr int BSM_UNROLL = MCZ_BSM_UNROLL;   // samples of a thread evaluated side by side in the shared-memory KK04 loop
#ifndef MCZ_FX_BITS
#define MCZ_FX_BITS 20
#endif
static constexpr float FX_SCALE = (float)(1 << MCZ_FX_BITS);
static constexpr int FX_WARP_CLAMP = 64 << MCZ_FX_BITS;
typedef int fx_t;
__device__ __forceinline__ int fx_clamp(int q) { return max(-FX_WARP_CLAMP, min(FX_WARP_CLAMP, q)); }
static constexpr int FX_BLOCK_CLAMP = 128 << MCZ_FX_BITS;     // cluster kernel: per-block totals, <= 8 of them in 32 bits
// Cluster kernel, after the block-local sums: thread r of every block sends the block's (clamped)
// totals and flags to block r of the cluster with two asynchronous 16-byte stores that signal the
// receiver's barrier; every thread then waits on its own block's barrier and adds up what the
// blocks sent.  No cluster barrier: ~one DSMEM flight per exchange, and the blocks only wait for data.
// `xq` counts the exchanges (the same in every thread of the cluster): barrier q & 1, phase (q >> 1) & 1.
struct ClusterX {
  unsigned xq;
};
__device__ __forceinline__ void cluster_exchange(Ctrl* ctrl, int ph, ClusterX& cx, int tot[4], bool& anychg, bool& anynan) {
  const int t = threadIdx.x;
  const unsigned q = cx.xq++, pb = q & 1u, par = (q >> 1) & 1u;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&ctrl->xbar[pb]);
  const unsigned csize = cl_size();
  __syncthreads();                       // the block's own totals are complete
  if (t < (int)csize) {
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = max(-FX_BLOCK_CLAMP, min(FX_BLOCK_CLAMP, ctrl->acc[ph][k]));
    const unsigned dst = cl_map((unsigned)__cvta_generic_to_shared(&ctrl->xslot[pb][cl_rank()][0]), (unsigned)t);
    const unsigned rbar = cl_map(bar, (unsigned)t);
    st_async_v4(dst, (unsigned)v[0], (unsigned)v[1], (unsigned)v[2], (unsigned)v[3], rbar);
    st_async_v4(dst + 16u, ctrl->chg[ph], ctrl->nanw[ph], 0u, 0u, rbar);
  } else if (t == 32) {
    mbar_expect_tx_arrive(bar, csize * 32u);
  }
  mbar_wait(bar, par, ctrl->poison_word, 1u);
  int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  unsigned c = 0u, n = 0u;
  for (unsigned r = 0; r < csize; ++r) {
    const int4 a = *reinterpret_cast<const int4*>(&ctrl->xslot[pb][r][0]);
    const int2 b = *reinterpret_cast<const int2*>(&ctrl->xslot[pb][r][4]);
    s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
    c |= (unsigned)b.x; n |= (unsigned)b.y;
  }
  tot[0] = s0; tot[1] = s1; tot[2] = s2; tot[3] = s3;
  anychg = c != 0u;
  anynan = n != 0u;
}
template <bool CL>
__device__ __forceinline__ void block_sum2_fx(float s1, float s2, bool nan, Ctrl* ctrl, int& ph,
                                              fx_t& t1, fx_t& t2, bool& anynan, ClusterX& cx) {
  const int lane = threadIdx.x & 31;
  int q1 = __reduce_add_sync(0xffffffffu, __float2int_rn(s1 * FX_SCALE));
  int q2 = __reduce_add_sync(0xffffffffu, __float2int_rn(s2 * FX_SCALE));
  const bool wnan = CL ? __any_sync(0xffffffffu, nan) : false;
  if (lane == 0) {
    atomicAdd(&ctrl->acc[ph][0], fx_clamp(q1));
    atomicAdd(&ctrl->acc[ph][1], fx_clamp(q2));
    if (CL && wnan) ctrl->nanw[ph] = 1u;
  }
  const int nx = ph == 2 ? 0 : ph + 1;
  if (threadIdx.x < 4) ctrl->acc[nx][threadIdx.x] = 0;
  if (threadIdx.x == 4) { ctrl->chg[nx] = 0u; ctrl->nanw[nx] = 0u; }
  if (CL) {
    int tot[4];
    bool chg;
    cluster_exchange(ctrl, ph, cx, tot, chg, anynan);
    t1 = tot[0];
    t2 = tot[1];
  } else {
    anynan = __syncthreads_or(nan);
    t1 = ctrl->acc[ph][0];
    t2 = ctrl->acc[ph][1];
  }
  ph = nx;
}

// Same for four sums (two loops x two trips per barrier); `nanbits` is OR-ed over the block.
// `changed`: this thread's R23 loop state differs from the state before the batch; `anychanged`
// is its OR over the block (cluster).
template <bool CL>
__device__ __forceinline__ void block_sum4_fx(const float s[4], unsigned nanbits, bool changed, Ctrl* ctrl, int& ph,
                                              fx_t t[4], bool& anynan, bool& anychanged, ClusterX& cx) {
  const int lane = threadIdx.x & 31;
  int q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = __reduce_add_sync(0xffffffffu, __float2int_rn(s[k] * FX_SCALE));
  const bool wchg = __any_sync(0xffffffffu, changed);
  const bool wnan = CL ? __any_sync(0xffffffffu, nanbits != 0u) : false;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&ctrl->acc[ph][k], fx_clamp(q[k]));
    if (wchg) ctrl->chg[ph] = 1u;
    if (CL && wnan) ctrl->nanw[ph] = 1u;
  }
  const int nx = ph == 2 ? 0 : ph + 1;
  if (threadIdx.x < 4) ctrl->acc[nx][threadIdx.x] = 0;
  if (threadIdx.x == 4) { ctrl->chg[nx] = 0u; ctrl->nanw[nx] = 0u; }
  if (CL) {
    cluster_exchange(ctrl, ph, cx, t, anychanged, anynan);
  } else {
    anynan = __syncthreads_or(nanbits != 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = ctrl->acc[ph][k];
    anychanged = ctrl->chg[ph] != 0u;
  }
  ph = nx;
}

// MUFU.RCP without the range checks of __fdividef (the KK04 denominators are O(1))
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Column store.  In the shared-memory variant the 32-bit shared address of the thread's sample
// is computed once and the store is issued as st.shared: through a generic pointer the compiler
// re-derives the shared window base (S2UR / ULEA / IMAD) at every one of the 17 stores.
template <bool SHARED> struct SlotPut {
  static constexpr bool WANT_AUX = false;
  float* col;                 // &vals[0][s]
  const int* col_off;         // per key: float offset of its column (kernel parameter space)
  uint32_t saddr;             // shared-space byte address of col (SHARED only)
  __device__ __forceinline__ void operator()(int key, float v) const {
    // every key the physics can emit under the launch's tokens owns a column, except the
    // constant-NaN ones (M08_R23, M08_O3Hb, M08_O2Hb: const_nan_key), which are never stored
    if (const_nan_key(key) || key >= NKEYS) return;       // (aux keys: replay only)
    if (SHARED) asm volatile("st.shared.f32 [%0], %1;" :: "r"(saddr + 4u * (unsigned)col_off[key]), "f"(v) : "memory");
    else col[col_off[key]] = v;
  }
};
template <bool SHARED> __device__ __forceinline__ void store_f32(float* gptr, uint32_t saddr, float v) {
  if (SHARED) asm volatile("st.shared.f32 [%0], %1;" :: "r"(saddr), "f"(v) : "memory");
  else *gptr = v;
}

__device__ unsigned long long g_phase_cycles[8];
__device__ unsigned long long g_col_cycles[2 * NKEYS];   // per key: cycles, max-cycles accumulators (cluster kernel: per stage)

}  // namespace mcz
#include "mcz_select.cuh"
namespace mcz {

// optional phase timing (build with -DMCZ_PHASE_TIMING): cycles summed over blocks, read back with
// mcz_debug_phase_cycles: [0] galaxies, [1] phase A, [2] phases B+C, [3] phase D own work (mean
// over the 18 warps), [4] phase D wall (start of D to the block barrier), [5] other
#ifdef MCZ_PHASE_TIMING
#define MCZ_T(var) const long long var = clock64()
#define MCZ_TADD(idx, val) do { if (tid == 0) atomicAdd(&g_phase_cycles[idx], (unsigned long long)(val)); } while (0)
#else
#define MCZ_T(var)
#define MCZ_TADD(idx, val)
#endif

// ---- the kernel -------------------------------------------------------------------------------
// SLOTS > 0: at most SLOTS samples per thread, loop state in registers, columns in shared memory.
// SLOTS == 0: any N', columns / stash / loop state in the per-block global scratch.
// WIDE: 32-bit histogram counters (columns of 65536 samples or more); global variant only.
// CL: launched in clusters of ceil(N' / (SLOTS * PT)) blocks, one galaxy per CLUSTER: block r of the
//     cluster holds samples [r SLOTS PT, (r + 1) SLOTS PT) of every column in its shared memory.
// PAIRK (global variant): two warps per column in phase D, for launches with at most PW / 2 columns.
template <int SLOTS, bool WIDE, bool CL, bool PAIRK = false>
__global__ void __launch_bounds__(PT, 1)
production_kernel(const __grid_constant__ ProdParams p) {
  static_assert(!PAIRK || (SLOTS == 0 && !WIDE && !CL), "two warps per column: global variant, 16-bit counters");
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr bool GLOBAL = (SLOTS == 0);
  static_assert(!CL || SLOTS > 0, "the cluster kernel keeps its columns in shared memory");
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Npad = p.Npad;                // samples a block holds per column (a multiple of 128)
  // cluster kernel: the samples are split evenly, block r holds [r cper, (r + 1) cper) of every column
  // (cper <= SLOTS * PT) and streams cng = ceil(cper / 128) groups of 128 in the selection passes.
  // The block's rank, first sample (SOFF) and sample count (NLOC) are re-read where they are needed
  // (special registers / control block) instead of living in registers across the KK04 loops.
  // NLOC: samples of the galaxy that THIS block holds (all of them without clusters); the loop
  // tolerances and the percentile ranks use the galaxy's p.N
#define CRANK (CL ? cl_rank() : 0u)
#define CSIZE (CL ? cl_size() : 1u)
#define SOFF (CL ? ctrl->cl_soff : 0)
#define NLOC (CL ? ctrl->cl_n : p.N)
  if (CL && threadIdx.x == 0) {
    const int so = (int)cl_rank() * p.cper;
    ctrl->cl_soff = so;
    ctrl->cl_n = max(0, min(p.N - so, p.cper));
  }
  const int Nfill = CL ? p.cng * 128 : Npad;                // [NLOC, Nfill) of every column is NaN
  ClusterX cx;
  cx.xq = 0u;

  constexpr bool PACK16 = !WIDE;          // 16-bit packed histogram counters: columns of < 65536 samples
  constexpr size_t HIST_BYTES = FineHist<PACK16>::BYTES;
  float* vals;              // columns [nstore][Npad]
  float* st0; float* st1; float* st2; float* st3;   // four per-sample scratch arrays (phase-A carry, then R23 constants)
  float* giter = nullptr;   // global variant: loop state, fields 10..15 of [..][Npad]
  float* sstate = nullptr;  // global variant: Z1, Z2, logq1, logq2 of the first p.scap samples, in shared memory
  unsigned char* work;      // selection scratch: PW histograms, then PW candidate areas
  if (GLOBAL) {
    float* base = p.gscratch + (long long)blockIdx.x * p.gscratch_per_block;
    vals = base;
    st0 = base + (long long)p.nstore * Npad;
    st1 = st0 + Npad; st2 = st1 + Npad; st3 = st2 + Npad;
    giter = st3 + Npad;
    work = smem + CTRL_BYTES;
    sstate = reinterpret_cast<float*>(work + (size_t)PW * (HIST_BYTES + CAND_FLOATS * sizeof(float)));
    if (p.bsm) {
      // the three carry values of every sample live in shared memory (over the selection work area,
      // which is dead between two phases D); the flag word stays in the global scratch
      const int Nq = (p.N + 3) & ~3;
      st0 = reinterpret_cast<float*>(smem + CTRL_BYTES);
      st1 = st0 + Nq;
      st2 = st1 + Nq;
    }
  } else {
    // Shared-memory variant (Npad == SLOTS*PT always).  The histograms are live from phase B on
    // (they are filled while the KK04 loops wait on their barriers), so the per-sample scratch
    // arrays must not alias them: three of them are the columns that only phase C writes
    // (KK04_N2Ha, KK04_R23, KD02comb), the fourth is an extra array after the candidate areas.
    vals = reinterpret_cast<float*>(smem + CTRL_BYTES);
    work = smem + CTRL_BYTES + (size_t)p.nstore * Npad * sizeof(float);
    st3 = reinterpret_cast<float*>(work + (size_t)PW * (HIST_BYTES + CAND_FLOATS * sizeof(float)));
    const bool kd = p.tok & T_KD02;
    st0 = kd ? vals + p.col_off[K_KK04_N2HA] : st3;
    st1 = kd ? vals + p.col_off[K_KK04_R23] : st3;
    st2 = kd ? vals + p.col_off[K_KD02COMB] : st3;
  }
  // (the mov hides where the address comes from, or the compiler re-derives it at every store)
  uint32_t vals_sa = GLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(vals);
  asm volatile("mov.u32 %0, %0;" : "+r"(vals_sa));
  const uint32_t st0_sa = GLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(st0);
  const uint32_t st1_sa = GLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(st1);
  const uint32_t st2_sa = GLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(st2);
  const uint32_t st3_sa = GLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(st3);
  unsigned* hist = reinterpret_cast<unsigned*>(work + (size_t)warp * HIST_BYTES);
  float* cand = reinterpret_cast<float*>(work + (size_t)PW * HIST_BYTES) + warp * CAND_FLOATS;

  const int used2line[NUSED] = {L_HA, L_HB, L_O2, L_O3, L_N2, L_S2A, L_S2B, L_S39069};
  const int nsl = (Npad + PT - 1) / PT;     // Npad is a multiple of 128; [N, Npad) is NaN-filled
  int phase = 0;

  // The next galaxy (queue ticket, its 22 input numbers, cleared flags) is fetched by the last warp
  // while the block is busy with phase D, so the loop top costs one barrier and no global latency.
  auto fetch_next = [&]() {
    if (CL && CRANK != 0u) return;        // the cluster's first block takes the ticket and hands it to all
    long long gn = 0;
    if (lane == 0) gn = (long long)atomicAdd(p.counter, 1ull);
    gn = __shfl_sync(0xffffffffu, gn, 0);
    float lv = 0.0f;
    if (gn < p.G) {
      if (lane < NLINES) lv = p.meas[gn * NLINES + lane];
      else if (lane < 2 * NLINES) lv = p.err[gn * NLINES + (lane - NLINES)];
    }
    if (CL) {
      const unsigned csize = CSIZE;
      for (unsigned r = 0; r < csize; ++r) {
        if (lane < 2 * NLINES) cl_stf(cl_map(&ctrl->line[lane], r), lv);
        if (lane == 0) {
          const unsigned ga = cl_map(&ctrl->g, r);
          cl_st(ga, (unsigned)(unsigned long long)gn);
          cl_st(ga + 4u, (unsigned)((unsigned long long)gn >> 32));
          cl_st(cl_map(&ctrl->flags, r), 0u);
          cl_st(cl_map(&ctrl->nan4, r), 0u);
          for (int k = 0; k < 4; ++k) cl_st(cl_map(&ctrl->acc[0][k], r), 0u);
          cl_st(cl_map(&ctrl->chg[0], r), 0u);
          cl_st(cl_map(&ctrl->nanw[0], r), 0u);
        }
      }
    } else {
      if (lane < 2 * NLINES) ctrl->line[lane] = lv;
      if (lane == 0) {
        ctrl->g = gn;
        ctrl->flags = 0u; ctrl->acc[0][0] = 0; ctrl->acc[0][1] = 0; ctrl->acc[0][2] = 0; ctrl->acc[0][3] = 0;
        ctrl->chg[0] = 0u;
      }
    }
  };
  if constexpr (CL) {
    // barriers of the KK04 exchange (control block) and of the selection protocol (every warp's
    // candidate area): one expected arrival each, the receiver's own arrive.expect_tx
    if (tid == 0) {
      mbar_init((unsigned)__cvta_generic_to_shared(&ctrl->xbar[0]), 1u);
      mbar_init((unsigned)__cvta_generic_to_shared(&ctrl->xbar[1]), 1u);
      ctrl->poison_word = p.stats + 7;
    }
    if (lane == 0) {
      const unsigned ca = (unsigned)__cvta_generic_to_shared(cand);
      mbar_init(ca + ClMail::MB_MERGE, 1u);
      mbar_init(ca + ClMail::MB_CMD, 1u);
      mbar_init(ca + ClMail::MB_GATHER, 1u);
      reinterpret_cast<unsigned*>(cand)[ClMail::PAR / 4] = 0u;
    }
    mbar_fence_init();
    cl_sync();                  // every block of the cluster runs and has its barriers before any remote access
  }
  if (warp == PW - 1) fetch_next();

  for (;;) {
    block_or_cluster_sync<CL>();
    const long long g = ctrl->g;
    if (g >= p.G) {
      if (p.ndst > 1) __threadfence_system();    // result rows written into other GPUs' tables
      break;
    }
    float lm[NUSED], le[NUSED];
    uint32_t assumed = 0u;
#pragma unroll
    for (int u = 0; u < NUSED; ++u) {
      lm[u] = ctrl->line[used2line[u]];
      le[u] = p.deterministic ? 0.0f : ctrl->line[NLINES + used2line[u]];
      const float thr = (u >= U_S2B) ? 1e-9f : 0.0f;
      // any(sample > thr): certain when err == 0, assumed true (and verified below) otherwise
      bool a = finite_f(lm[u]) && finite_f(le[u]);
      if (a && le[u] == 0.0f) a = lm[u] > thr;
      if (a) assumed |= 1u << u;
    }
    const unsigned long long gg = p.goff + (unsigned long long)g;

    MCZ_T(tA0);
    // ---- phase A (repeated once in the rare case the assumed flags were wrong) ----------------
    uint32_t flags = assumed;
    bool anykd = false;
    for (;;) {
      const int de = effective_dust(p.dust, flags);
      uint32_t seen = 0u;
      // The line flags, the token mask, the dust mode and the sampling mode are the same for every
      // sample of the galaxy, but phase_a tests them per sample (about forty uniform branches).  The
      // two launch shapes that matter for throughput -- every line measured, dust on, independent
      // sampling, `--md all` or `--md KD02` -- get a copy of the loop with those values as
      // compile-time constants: straight-line code, identical arithmetic.
      const uint32_t need = F_ALL & ~F_S39069;      // [SIII]9069 only matters for DP00
      // (|z| <= sqrt(-2 ln 2^-24) = 5.8 for the Box-Muller normals: with |meas| + 6 |err| inside the float
      // range no perturbed flux can overflow, and the +inf -> 0 clamp of metallicity.py:96 is dead code)
      bool bounded = true;
#pragma unroll
      for (int u = 0; u < NUSED; ++u) bounded = bounded && (fabsf(lm[u]) + 6.0f * fabsf(le[u]) < FLT_MAX);
      const bool full = (flags & need) == need && de == 0 && !p.correlated && !p.deterministic && bounded;
      auto pass = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;      // 0 generic, 1 --md all, 2 --md KD02
        const uint32_t fl = MODE ? (uint32_t)F_ALL : flags;
        const uint32_t tk = MODE == 1 ? (uint32_t)T_ALL : (MODE == 2 ? (uint32_t)T_KD02 : p.tok);
        const int dm = MODE ? 0 : de;
        const int N = NLOC, soff = SOFF;
#pragma unroll A_UNROLL
        for (int i = 0; i < nsl; ++i) {
          const int s = i * PT + tid;
          if (s < N) {
            float z[16];
            if (MODE == 1) draw_normals_c<7u>((uint32_t)(soff + s), gg, p, z);
            else if (MODE == 2) draw_normals_c<6u>((uint32_t)(soff + s), gg, p, z);
            else draw_normals((uint32_t)(soff + s), gg, p, z);
            float f[NUSED], e[5];
#pragma unroll
            for (int u = 0; u < NUSED; ++u) {
              const float raw = fmaf(le[u], (!MODE && p.correlated) ? z[5] : z[5 + u], lm[u]);   // mcz.py:442 / :589
              float v = fmaxf(raw, 0.0f);               // NaN -> 0 and negative -> 0 (metallicity.py:96, 99)
              if (!MODE && !(v <= FLT_MAX)) v = 0.0f;   // +inf -> 0 (metallicity.py:96); cannot happen when `bounded`
              if (!MODE && (p.tok & T_DROPNEG) && raw < 0.0f && raw >= -FLT_MAX) v = Math<float>::nan();   // DROPNEGATIVES (metallicity.py:101-102)
              f[u] = v;
              const bool pos = (u >= U_S2B) ? (v > 1e-9f) : (v > 0.0f);
              if (pos) seen |= 1u << u;
            }
            e[0] = 0.05f * z[0];            // metscales.py:679
            e[1] = 0.1f * z[1];             // metscales.py:680
            e[2] = 0.027f * z[2];           // metscales.py:948
            e[3] = 0.024f * z[3];           // metscales.py:949
            e[4] = 0.012f * z[4];           // metscales.py:952
            SlotPut<!GLOBAL> put{vals + s, p.col_off, vals_sa + 4u * (unsigned)s};
            Carry<float> c;
            c.y = c.n2ha = c.x = 0.0f;
            c.aux = 0u;
            phase_a<float>(f, e, fl, tk, dm, put, c);
            if (c.aux & AUX_KD_VALID) seen |= SEEN_KD;
            store_f32<!GLOBAL>(st0 + s, st0_sa + 4u * (unsigned)s, c.y);
            store_f32<!GLOBAL>(st1 + s, st1_sa + 4u * (unsigned)s, c.n2ha);
            store_f32<!GLOBAL>(st2 + s, st2_sa + 4u * (unsigned)s, c.x);
            store_f32<!GLOBAL>(st3 + s, st3_sa + 4u * (unsigned)s, __uint_as_float(c.aux));
          } else if (s < Nfill) {
            for (int sl = 0; sl < p.nstore; ++sl) vals[(long long)sl * Npad + s] = Math<float>::nan();
          }
        }
      };
      if (full && p.tok == T_ALL && p.calls == 7u) pass(std::integral_constant<int, 1>());
      else if (full && p.tok == T_KD02 && p.calls == 6u) pass(std::integral_constant<int, 2>());
      else pass(std::integral_constant<int, 0>());
      seen = __reduce_or_sync(0xffffffffu, seen);
      if (lane == 0 && seen) {
        if (CL) {
          for (unsigned r = 0, csize = CSIZE; r < csize; ++r) cl_red_or(cl_map(&ctrl->flags, r), seen);
        } else {
          atomicOr(&ctrl->flags, seen);
        }
      }
      block_or_cluster_sync<CL>();
      const uint32_t all = ctrl->flags;
      const uint32_t actual = all & F_ALL;
      anykd = (all & SEEN_KD) != 0u;          // some sample has a KD02_N2O2 value (start of the N2Ha loop)
      if (actual == flags) break;
      flags = actual;               // deterministic samples: the second pass reproduces `actual`
      block_or_cluster_sync<CL>();  // (rare path) drop the KD bit of the discarded pass
      if (tid == 0) ctrl->flags = actual;
      block_or_cluster_sync<CL>();
    }
    const uint64_t pm = present_mask(flags, p.tok);
    const bool ok = minimum_ok(flags);

    MCZ_T(tB0);
#ifdef MCZ_PHASE_TIMING
    long long tBL0 = tB0, tBL1 = tB0;
#endif
    // ---- phases B and C ----------------------------------------------------------------------
    int ii1 = 0, ii2 = 0;
    int executed1 = -1;       // N2Ha trips actually executed (differs from ii1 after a proven early exit)
    int executed2 = -1;       // same for the R23 loop (periodic-state exit)
    if ((p.tok & T_KD02) && ok) {
      const bool has_kd = (pm >> K_KD02_N2O2) & 1ull;
      const bool has_kn = (pm >> K_KK04_N2HA) & 1ull, has_kr = (pm >> K_KK04_R23) & 1ull;
      const bool o3o2 = (flags & F_O2) && (flags & F_O3);
      const int sl_kd = p.slot_of_key[K_KD02_N2O2], sl_m91 = p.slot_of_key[K_M91];
      bool act1 = has_kn && o3o2, act2 = has_kr;
      // mean > tol  <=>  sum > tol * N  (NaN sum compares false: the loop stops, as in numpy)
      const float tol1 = 1.0e-3f * (float)p.N, tol2 = 1.0e-4f * (float)p.N;
      if constexpr (!GLOBAL) {
        // Loop state: 11 registers per sample; the four R23 branch constants of each sample live
        // in shared memory, in the space of the (now dead) stash.  Slots without a sample get
        // constants that make every trip contribute exactly 0.
        float num0[SLOTS], num1[SLOTS], den0[SLOTS], den1[SLOTS], cA[SLOTS], cB[SLOTS];
        float Z1[SLOTS], lq1[SLOTS], Z2[SLOTS], lq2[SLOTS];
        uint32_t aux[SLOTS];
        constexpr int CS = SLOTS * PT;       // stride of the shared-memory constant arrays
        // U0, U1, L0, L1 of every sample replace the carry in st0..st3 (Npad == CS in this variant)
        float tU0[SLOTS], tU1[SLOTS], tL0[SLOTS], tL1[SLOTS];
        bool any = false;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
          const int s = i * PT + tid;
          num0[i] = 0.0f; num1[i] = 0.0f; den0[i] = 1.0f; den1[i] = 0.0f; cA[i] = 0.0f; cB[i] = 0.0f;
          Z1[i] = 0.0f; lq1[i] = 0.0f; Z2[i] = 0.0f; lq2[i] = 0.0f; aux[i] = 2u;
          tU0[i] = 0.0f; tU1[i] = 0.0f; tL0[i] = 0.0f; tL1[i] = 0.0f;
          if (s < NLOC) {
            Carry<float> c;
            c.y = st0[s]; c.n2ha = st1[s]; c.x = st2[s];
            c.aux = __float_as_uint(st3[s]);
            const float kd = has_kd ? vals[sl_kd * Npad + s] : Math<float>::nan();
            IterState<float> st;
            make_iter<float>(c, kd, st);
            any = any || !Math<float>::isnan(kd);
            num0[i] = st.num0; num1[i] = st.num1; den0[i] = st.den0; den1[i] = st.den1;
            cA[i] = st.A; cB[i] = st.B; Z1[i] = st.Z1; lq1[i] = st.lq1; Z2[i] = st.Z2; lq2[i] = st.lq2;
            aux[i] = st.aux;
            tU0[i] = st.U0; tU1[i] = st.U1; tL0[i] = st.L0; tL1[i] = st.L1;
          }
        }
        // (no barrier: a thread overwrites only the stash entries it has just read itself; whether
        // any sample has a KD02_N2O2 value was collected with the line flags in phase A)
        (void)any;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
          const int s = i * PT + tid;
          st0[s] = tU0[i]; st1[s] = tU1[i]; st2[s] = tL0[i]; st3[s] = tL1[i];
        }
        // start of the N2Ha loop: 8.6 if KD02_N2O2 is None or all-NaN (metscales.py:1097-1103)
        if (!anykd) {
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) if (i * PT + tid < NLOC) Z1[i] = 8.6f;
        }
        if (has_kn && !o3o2) {                                          // metscales.py:1122-1126
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) Z1[i] = cA[i] - 7.37177f * cB[i];
        }
        // thresholds tol * N in the same fixed point (exact: N * 2^FX_BITS / 1000 in float64)
        const fx_t itol1 = (fx_t)(1.0e-3 * (double)p.N * (double)FX_SCALE), itol2 = (fx_t)(1.0e-4 * (double)p.N * (double)FX_SCALE);
        int ph = 0;
        // Two trips per barrier: trips a and b are computed back to back, their four sums are
        // reduced together, then the reference's loop conditions are replayed trip by trip.  If a
        // loop turns out to have ended after trip a, its state is put back to that point (logq of
        // trip a is kept; Z follows from it).
#ifdef MCZ_PHASE_TIMING
        tBL0 = clock64();
#endif
        while (act1 || act2) {
          float sm4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          bool bad = false;           // some logq of this batch is NaN (it stays NaN: tested once per loop and batch)
          bool changed = false;       // R23 loop: some sample's logq after this batch differs from before it
          float lq1a[SLOTS], lq2a[SLOTS];
          if (act1) {
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {                            // metscales.py:1112-1118
              float lq = (num0[i] + Z1[i] * num1[i]) * fast_rcp(den0[i] + Z1[i] * den1[i]);
              float z = cA[i] - lq * cB[i];
              float d = fabsf(lq - lq1[i]);
              sm4[0] += fminf(d, 8.0f);
              lq1a[i] = lq;
              const float lqb = (num0[i] + z * num1[i]) * fast_rcp(den0[i] + z * den1[i]);
              Z1[i] = cA[i] - lqb * cB[i];
              d = fabsf(lqb - lq);
              sm4[1] += fminf(d, 8.0f);
              bad = bad || (lqb != lqb);
              lq1[i] = lqb;
            }
          }
          if (act2) {
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {                            // metscales.py:1167-1178
              const int s = i * PT + tid;
              const float u0 = st0[s], u1 = st1[s], l0 = st2[s], l1 = st3[s];
              const bool lowz = (aux[i] & AUX_ZI_MASK) <= 1u;
              float lq = (num0[i] + Z2[i] * num1[i]) * fast_rcp(den0[i] + Z2[i] * den1[i]);
              bool lower = lowz && (lq >= 6.7f) && (lq < 8.3f);
              float z = (lower ? l0 : u0) - lq * (lower ? l1 : u1);
              float d = lq2[i] - lq;
              sm4[2] += fminf(fmaxf(d, -8.0f), 8.0f);
              lq2a[i] = lq;
              const float lqb = (num0[i] + z * num1[i]) * fast_rcp(den0[i] + z * den1[i]);
              lower = lowz && (lqb >= 6.7f) && (lqb < 8.3f);
              Z2[i] = (lower ? l0 : u0) - lqb * (lower ? l1 : u1);
              d = lq - lqb;
              sm4[3] += fminf(fmaxf(d, -8.0f), 8.0f);
              bad = bad || (lqb != lqb);
              changed = changed || (lqb != lq2[i]);
              lq2[i] = lqb;
            }
          }
          fx_t t4[4];
          bool anynan, anychanged;
          block_sum4_fx<CL>(sm4, bad ? 1u : 0u, changed, ctrl, ph, t4, anynan, anychanged, cx);
          unsigned nb = 0u;
          if (anynan) {
            // rare: which loop / trip saw the NaN first (a NaN mean stops that loop, as in numpy).  A NaN
            // logq makes every later one NaN, so trip a's mean is NaN iff some logq after trip a is.
            unsigned nanb = 0u;
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {
              if (act1) nanb |= ((lq1a[i] != lq1a[i]) ? 1u : 0u) | ((lq1[i] != lq1[i]) ? 2u : 0u);
              if (act2) nanb |= ((lq2a[i] != lq2a[i]) ? 4u : 0u) | ((lq2[i] != lq2[i]) ? 8u : 0u);
            }
            if (CL) {
              nanb = __reduce_or_sync(0xffffffffu, nanb);
              if (lane == 0 && nanb)
                for (unsigned r = 0, csize = CSIZE; r < csize; ++r) cl_red_or(cl_map(&ctrl->nan4, r), nanb);
              cl_sync();
              nb = ctrl->nan4;
              cl_sync();                    // everyone has read it: clear for the next use
              if (tid == 0) ctrl->nan4 = 0u;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) nb |= __syncthreads_or((nanb >> k) & 1u) ? (1u << k) : 0u;
            }
          }
          if (act1) {                                                     // metscales.py:1111,1117
            ++ii1;
            if ((nb & 1u) || !(t4[0] > itol1) || ii1 >= 100) {
              act1 = false;       // ended after trip a: undo trip b
#pragma unroll
              for (int i = 0; i < SLOTS; ++i) { lq1[i] = lq1a[i]; Z1[i] = cA[i] - lq1a[i] * cB[i]; }
            } else {
              ++ii1;
              act1 = !(nb & 2u) && (t4[1] > itol1) && ii1 < 100;
            }
          }
          if (act2) {                                                     // metscales.py:1166,1177
            ++ii2;
            if ((nb & 4u) || !((t4[2] < 0 ? -t4[2] : t4[2]) > itol2) || ii2 >= 100) {
              act2 = false;
#pragma unroll
              for (int i = 0; i < SLOTS; ++i) {
                const int s = i * PT + tid;
                const float lq = lq2a[i];
                const bool lower = ((aux[i] & AUX_ZI_MASK) <= 1u) && (lq >= 6.7f) && (lq < 8.3f);
                lq2[i] = lq;
                Z2[i] = (lower ? st2[s] : st0[s]) - lq * (lower ? st3[s] : st1[s]);
              }
            } else {
              ++ii2;
              act2 = !(nb & 8u) && ((t4[3] < 0 ? -t4[3] : t4[3]) > itol2) && ii2 < 100;
              // Periodic-state exit.  logq of EVERY sample is bit-identical to what it was two trips
              // ago, and Z is a function of logq: every later batch repeats this one, sums included.
              // Both of its trips passed the loop test with a factor 2 to spare, so the loop runs to
              // the 100-trip cap and the galaxy is blanked (metscales.py:1179-1181) -- report that now.
              // (What the reference's capped R23 loops do: samples caught in a 2-cycle between the
              // lower and the upper branch keep |mean(dlogq)| at a constant far above the tolerance.)
              if (act2 && !anychanged && (t4[2] < 0 ? -t4[2] : t4[2]) > 2 * itol2 && (t4[3] < 0 ? -t4[3] : t4[3]) > 2 * itol2) {
                executed2 = ii2; ii2 = 100; act2 = false;
              }
            }
          }
          if (act1 && ii1 == DIVERGENCE_CHECK_TRIP) {
            // Early exit for galaxies whose N2Ha loop provably never converges (SURVEY A.3 item 4:
            // log N2Ha >~ -0.35).  Per sample the trip is the Moebius map
            //   lq' = (al - be lq) / (ga - de lq),  al = num0 + A num1, be = B num1, ga = den0 + A den1, de = B den1.
            // When it has no real fixed point ((be+ga)^2 < 4 al de) the step |lq' - lq| is bounded
            // below, for EVERY lq, by m = min |q(u)| / r over the two critical points
            // u = (ga +- r)/de, r = sqrt(al de - be ga), q(u) = de u^2 - (be+ga) u + al.
            // If sum_s m_s > 4 tol N, mean|dlogq| > tol at every trip: the reference runs to its
            // 100-trip cap and blanks the galaxy (metscales.py:1119-1121) -- so do that now.
            float sm = 0.0f;
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {
              const float al = num0[i] + cA[i] * num1[i], be = cB[i] * num1[i];
              const float ga = den0[i] + cA[i] * den1[i], de = cB[i] * den1[i];
              const float bg = be + ga, disc = bg * bg - 4.0f * al * de;
              float m = 0.0f;
              if (disc < 0.0f) {
                const float r = __fsqrt_rn(al * de - be * ga), ide = fast_rcp(de);
                const float u1 = (ga + r) * ide, u2 = (ga - r) * ide;
                const float q1 = (de * u1 - bg) * u1 + al, q2 = (de * u2 - bg) * u2 + al;
                m = fminf(fabsf(q1), fabsf(q2)) * fast_rcp(r);
              }
              sm += (m == m) ? fminf(m, 8.0f) : 0.0f;
            }
            fx_t tm, tdummy;
            bool nandummy;
            block_sum2_fx<CL>(sm, 0.0f, false, ctrl, ph, tm, tdummy, nandummy, cx);
            if (tm > 4 * itol1) { executed1 = ii1; ii1 = 100; act1 = false; }
          }
        }
        if (executed1 < 0) executed1 = ii1;
#ifdef MCZ_PHASE_TIMING
        tBL1 = clock64();
#endif
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
          const int s = i * PT + tid;
          if (s < NLOC) {
            SlotPut<!GLOBAL> put{vals + s, p.col_off, vals_sa + 4u * (unsigned)s};
            const float kd = has_kd ? vals[sl_kd * Npad + s] : Math<float>::nan();
            const float m91 = vals[sl_m91 * Npad + s];
            IterState<float> st;
            st.U0 = st0[s]; st.U1 = st1[s]; st.L0 = st2[s]; st.L1 = st3[s];
            st.Z1 = Z1[i]; st.Z2 = Z2[i]; st.lq2 = lq2[i]; st.aux = aux[i];
            phase_c<float>(st, pm, ii1 >= 100, ii2 >= 100, kd, m91, put);
          } else {
            // the scratch use of the three phase-C columns ends here: restore their NaN padding
            st0[s] = Math<float>::nan(); st1[s] = Math<float>::nan(); st2[s] = Math<float>::nan();
          }
        }
      } else if (p.bsm) {
        // Loop inputs and state in shared memory, 20 bytes + 1 bit per sample: logO3O2, logN2Ha, logR23
        // (written there by phase A), the last logq of the two loops, and "Z_init_guess <= 8.4".  The ten
        // constants of the two maps are recomputed from the three inputs in every two-trip batch
        // (~35 FMAs) and the running Z is recomputed from logq (Z = A - logq B, or the R23 branch
        // polynomial: exactly the expression that produced it), so a batch reads and writes shared
        // memory only -- except logq after the first trip of the pair, which goes to the global
        // scratch in case the loop turns out to have ended there (write-only until then).
        // Until a loop has run its first batch, its logq array holds the INITIAL Z instead
        // (metscales.py:1097-1106, 1156-1163) and the previous logq is the constant 0 / 100.
        const int N = p.N;
        const int Nq = (N + 3) & ~3;
        float* sQ1 = st2 + Nq;
        float* sQ2 = sQ1 + Nq;
        unsigned* sLow = reinterpret_cast<unsigned*>(sQ2 + Nq);
        auto consts = [&](int s, IterState<float>& st) {
          Carry<float> c;
          c.y = st0[s]; c.n2ha = st1[s]; c.x = st2[s];
          c.aux = __float_as_uint(st3[s]);
          make_iter<float>(c, 0.0f, st);
        };
        bool any = false;
        for (int s0 = warp * 32; s0 < N; s0 += PT) {
          const int s = s0 + lane;
          const bool in = s < N;
          const uint32_t aux = in ? __float_as_uint(st3[s]) : 2u;
          const float kd = (in && has_kd) ? vals[(long long)sl_kd * Npad + s] : Math<float>::nan();
          any = any || (in && !Math<float>::isnan(kd));
          const unsigned low = __ballot_sync(0xffffffffu, in && (aux & AUX_ZI_MASK) <= 1u);
          if (lane == 0) sLow[s0 >> 5] = low;
          if (in) {
            const uint32_t zi = aux & AUX_ZI_MASK;
            sQ1[s] = kd;
            sQ2[s] = zi == 0u ? 8.2f : (zi == 1u ? 8.4f : 8.7f);                          // metscales.py:1156
          }
        }
        const bool allnan = !__syncthreads_or(any);
        if (allnan || (has_kn && !o3o2)) {
          for (int s = tid; s < N; s += PT) {
            float z1 = allnan ? 8.6f : sQ1[s];                                            // metscales.py:1097-1103
            if (has_kn && !o3o2) {                                                        // metscales.py:1122-1126
              IterState<float> st;
              iter_n2ha<float>(st1[s], st);
              z1 = st.A - 7.37177f * st.B;
            }
            sQ1[s] = z1;
          }
        }
        const bool ran1 = act1, ran2 = act2;
        bool first = true;
        // BT trips per batch: the ~45 instructions that load a sample and rebuild its constants are paid
        // once per batch, so three trips per batch cost less per trip than two; logq after each but
        // the last trip of the batch goes to the global scratch (fields 10.. of the loop-state area,
        // write-only unless a loop turns out to have ended there), and the reference's loop tests are
        // replayed trip by trip on the BT sums.
        constexpr int BT = BSM_TRIPS;
        static_assert(BT >= 2 && BT <= 3, "two global arrays per loop are reserved");
        float* Fl1 = giter;                        // [BT - 1][Npad]: N2Ha loop, logq after trips 1 .. BT-1 of the batch
        float* Fl2 = giter + 2ll * Npad;           // same for the R23 loop
        while (act1 || act2) {
          float sm1[BT], sm2[BT];
#pragma unroll
          for (int t = 0; t < BT; ++t) { sm1[t] = 0.0f; sm2[t] = 0.0f; }
          // (a NaN step makes the float sum of its trip NaN, and a NaN sum fails the loop test: the loop stops
          // there, as numpy's mean does -- no separate NaN flags are needed with float sums)
#pragma unroll BSM_UNROLL
          for (int s = tid; s < N; s += PT) {
            IterState<float> k;                       // only the constants of the loops still running
            iter_logq<float>(st0[s], k);
            const float num0 = k.num0, num1 = k.num1, den0 = k.den0, den1 = k.den1;
            if (act1) {
              iter_n2ha<float>(st1[s], k);
              const float cA = k.A, cB = k.B;
              const float q = sQ1[s];
              float Z = first ? q : cA - q * cB;
              float lqp = first ? 0.0f : q;                                               // logq_save (:1106)
#pragma unroll
              for (int t = 0; t < BT; ++t) {
                const float lq = (num0 + Z * num1) * fast_rcp(den0 + Z * den1);          // metscales.py:1112-1118
                const float d = fabsf(lq - lqp);
                sm1[t] += d;
                if (t < BT - 1) { Fl1[(long long)t * Npad + s] = lq; Z = cA - lq * cB; }
                lqp = lq;
              }
              sQ1[s] = lqp;
            }
            if (act2) {
              iter_r23<float>(st2[s], k);
              const float u0 = k.U0, u1 = k.U1, l0 = k.L0, l1 = k.L1;
              const bool lowz = (sLow[s >> 5] >> (s & 31)) & 1u;
              const float q = sQ2[s];
              bool lower = lowz && (q >= 6.7f) && (q < 8.3f);
              float Z = first ? q : (lower ? l0 : u0) - q * (lower ? l1 : u1);
              float lqp = first ? 100.0f : q;                                             // logqold (:1163)
#pragma unroll
              for (int t = 0; t < BT; ++t) {
                const float lq = (num0 + Z * num1) * fast_rcp(den0 + Z * den1);          // metscales.py:1167-1178
                const float d = lqp - lq;
                sm2[t] += d;
                if (t < BT - 1) {
                  Fl2[(long long)t * Npad + s] = lq;
                  lower = lowz && (lq >= 6.7f) && (lq < 8.3f);
                  Z = (lower ? l0 : u0) - lq * (lower ? l1 : u1);
                }
                lqp = lq;
              }
              sQ2[s] = lqp;
            }
          }
          first = false;
          // block sums of all the trips of the batch behind one barrier
          {
            float flat[2 * BT];
#pragma unroll
            for (int t = 0; t < BT; ++t) { flat[t] = sm1[t]; flat[BT + t] = sm2[t]; }
            // (second buffer: the exchange slots of the cluster kernel, unused here)
            static_assert(sizeof(ctrl->xslot) >= 2 * PW * sizeof(float2), "second reduction buffer");
            if constexpr (BT == 2) {
              block_sum4f(flat, ctrl->redf, reinterpret_cast<float2*>(ctrl->xslot), phase);
            } else {
#pragma unroll
              for (int i = 0; i < 2 * BT; i += 2) block_sum2f(flat[i], flat[i + 1], ctrl->redf, phase);
            }
#pragma unroll
            for (int t = 0; t < BT; ++t) { sm1[t] = flat[t]; sm2[t] = flat[BT + t]; }
          }
          if (act1) {                                                     // metscales.py:1111,1117
#pragma unroll
            for (int t = 0; t < BT; ++t) {
              if (act1) {
                ++ii1;
                if (!(sm1[t] > tol1) || ii1 >= 100) {
                  act1 = false;
                  if (t < BT - 1)     // ended before the last trip of the batch: put that logq back
                    for (int s = tid; s < N; s += PT) sQ1[s] = Fl1[(long long)t * Npad + s];
                }
              }
            }
          }
          if (act2) {                                                     // metscales.py:1166,1177
#pragma unroll
            for (int t = 0; t < BT; ++t) {
              if (act2) {
                ++ii2;
                if (!(fabsf(sm2[t]) > tol2) || ii2 >= 100) {
                  act2 = false;
                  if (t < BT - 1)
                    for (int s = tid; s < N; s += PT) sQ2[s] = Fl2[(long long)t * Npad + s];
                }
              }
            }
          }
          if (act1 && ii1 == DIVERGENCE_CHECK_TRIP) {
            // proven-divergence exit, as in the register variant above (same bound, same threshold)
            float sm = 0.0f, unused = 0.0f;
            for (int s = tid; s < N; s += PT) {
              IterState<float> k;
              iter_logq<float>(st0[s], k);
              iter_n2ha<float>(st1[s], k);
              const float n0 = k.num0, n1 = k.num1, d0 = k.den0, d1 = k.den1, A = k.A, B = k.B;
              const float al = n0 + A * n1, be = B * n1, ga = d0 + A * d1, de = B * d1;
              const float bg = be + ga, disc = bg * bg - 4.0f * al * de;
              float m = 0.0f;
              if (disc < 0.0f) {
                const float r = __fsqrt_rn(al * de - be * ga), ide = fast_rcp(de);
                const float u1 = (ga + r) * ide, u2 = (ga - r) * ide;
                const float q1 = (de * u1 - bg) * u1 + al, q2 = (de * u2 - bg) * u2 + al;
                m = fminf(fabsf(q1), fabsf(q2)) * fast_rcp(r);
              }
              sm += (m == m) ? fminf(m, 8.0f) : 0.0f;
            }
            block_sum2f(sm, unused, ctrl->redf, phase);
            if (sm > 4.0f * tol1) { executed1 = ii1; ii1 = 100; act1 = false; }
          }
        }
        if (executed1 < 0) executed1 = ii1;
        for (int s = tid; s < N; s += PT) {
          IterState<float> st;
          consts(s, st);
          const float q1 = sQ1[s], q2 = sQ2[s];
          st.Z1 = ran1 ? st.A - q1 * st.B : q1;
          if (ran2) {
            const bool lower = ((st.aux & AUX_ZI_MASK) <= 1u) && (q2 >= 6.7f) && (q2 < 8.3f);
            st.Z2 = (lower ? st.L0 : st.U0) - q2 * (lower ? st.L1 : st.U1);
            st.lq2 = q2;
          }
          SlotPut<!GLOBAL> put{vals + s, p.col_off, vals_sa + 4u * (unsigned)s};
          const float kd = has_kd ? vals[(long long)sl_kd * Npad + s] : Math<float>::nan();
          const float m91 = vals[(long long)sl_m91 * Npad + s];
          phase_c<float>(st, pm, ii1 >= 100, ii2 >= 100, kd, m91, put);
        }
      } else {
        // Memory-backed loop state.  The ten per-sample constants of the two maps are recomputed
        // from the four carry values (st0..st3: logO3O2, logN2Ha, logR23, flags) in every batch:
        // ~35 FMAs against 40 bytes of L2 traffic per sample, and this loop is bound by L2
        // bandwidth.  The four values that change every trip (Z1, Z2, logq1, logq2 = fields 10..13)
        // live in shared memory for the first p.scap samples (a multiple of the block size, so
        // each pass of the sample loop is on one side), in the global scratch beyond.
        const int N = p.N;
        float* F = giter - 10ll * Npad;          // only fields 10..15 exist (the constants are recomputed)
        const int scap = p.scap;
#define LST(k, s) (*((s) < scap ? sstate + ((k) - 10) * scap + (s) : F + (long long)(k) * Npad + (s)))
        auto consts = [&](int s, IterState<float>& st) {
          Carry<float> c;
          c.y = st0[s]; c.n2ha = st1[s]; c.x = st2[s];
          c.aux = __float_as_uint(st3[s]);
          make_iter<float>(c, 0.0f, st);
        };
        bool any = false;
        for (int s = tid; s < N; s += PT) {
          const float kd = has_kd ? vals[(long long)sl_kd * Npad + s] : Math<float>::nan();
          IterState<float> st;
          consts(s, st);
          any = any || !Math<float>::isnan(kd);
          LST(10, s) = kd; LST(11, s) = st.Z2;
          LST(12, s) = st.lq1; LST(13, s) = st.lq2;
        }
        const bool allnan = !__syncthreads_or(any);
        if (allnan || (has_kn && !o3o2)) {
          for (int s = tid; s < N; s += PT) {
            float z1 = allnan ? 8.6f : LST(10, s);
            if (has_kn && !o3o2) {
              IterState<float> st;
              consts(s, st);
              z1 = st.A - 7.37177f * st.B;
            }
            LST(10, s) = z1;
          }
        }
        // Two trips per barrier, as in the register variant; logq after the first trip is kept in
        // F[14], F[15] so that a loop that turns out to have ended there can be put back.
        while (act1 || act2) {
          float sm4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          unsigned nanb = 0u;
#pragma unroll 2
          for (int s = tid; s < N; s += PT) {
            IterState<float> k;                       // only the constants of the loops still running
            iter_logq<float>(st0[s], k);
            if (act1) iter_n2ha<float>(st1[s], k);
            if (act2) { iter_r23<float>(st2[s], k); k.aux = __float_as_uint(st3[s]); }
            const float num0 = k.num0, num1 = k.num1, den0 = k.den0, den1 = k.den1;
            if (act1) {
              const float cA = k.A, cB = k.B;
              const float Z1 = LST(10, s), lq1 = LST(12, s);
              const float lq = (num0 + Z1 * num1) * fast_rcp(den0 + Z1 * den1);        // metscales.py:1112-1118
              const float z = cA - lq * cB;
              float d = fabsf(lq - lq1);
              nanb |= (d != d) ? 1u : 0u;
              sm4[0] += d;
              const float lqb = (num0 + z * num1) * fast_rcp(den0 + z * den1);
              d = fabsf(lqb - lq);
              nanb |= (d != d) ? 2u : 0u;
              sm4[1] += d;
              LST(10, s) = cA - lqb * cB;
              LST(12, s) = lqb;
              F[14ll * Npad + s] = lq;
            }
            if (act2) {
              const float u0 = k.U0, u1 = k.U1, l0 = k.L0, l1 = k.L1;
              const float Z2 = LST(11, s), lq2 = LST(13, s);
              const bool lowz = (k.aux & AUX_ZI_MASK) <= 1u;
              const float lq = (num0 + Z2 * num1) * fast_rcp(den0 + Z2 * den1);        // metscales.py:1167-1178
              bool lower = lowz && (lq >= 6.7f) && (lq < 8.3f);
              const float z = (lower ? l0 : u0) - lq * (lower ? l1 : u1);
              float d = lq2 - lq;
              nanb |= (d != d) ? 4u : 0u;
              sm4[2] += d;
              const float lqb = (num0 + z * num1) * fast_rcp(den0 + z * den1);
              lower = lowz && (lqb >= 6.7f) && (lqb < 8.3f);
              d = lq - lqb;
              nanb |= (d != d) ? 8u : 0u;
              sm4[3] += d;
              LST(11, s) = (lower ? l0 : u0) - lqb * (lower ? l1 : u1);
              LST(13, s) = lqb;
              F[15ll * Npad + s] = lq;
            }
          }
          // float block sums of the unclamped terms (columns of any length: the 16.16 fixed point of
          // the register variant would overflow); NaN flags by vote as there
          block_sum2f(sm4[0], sm4[1], ctrl->redf, phase);
          block_sum2f(sm4[2], sm4[3], ctrl->redf, phase);
          unsigned nb = 0u;
          if (__syncthreads_or(nanb != 0u)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) nb |= __syncthreads_or((nanb >> k) & 1u) ? (1u << k) : 0u;
          }
          if (act1) {                                                     // metscales.py:1111,1117
            ++ii1;
            if ((nb & 1u) || !(sm4[0] > tol1) || ii1 >= 100) {
              act1 = false;       // ended after the first trip of the pair: put that state back
              for (int s = tid; s < N; s += PT) {
                IterState<float> k;
                consts(s, k);
                const float lq = F[14ll * Npad + s];
                LST(12, s) = lq;
                LST(10, s) = k.A - lq * k.B;
              }
            } else {
              ++ii1;
              act1 = !(nb & 2u) && (sm4[1] > tol1) && ii1 < 100;
            }
          }
          if (act2) {                                                     // metscales.py:1166,1177
            ++ii2;
            if ((nb & 4u) || !(fabsf(sm4[2]) > tol2) || ii2 >= 100) {
              act2 = false;
              for (int s = tid; s < N; s += PT) {
                IterState<float> k;
                consts(s, k);
                const float lq = F[15ll * Npad + s];
                const bool lower = ((k.aux & AUX_ZI_MASK) <= 1u) && (lq >= 6.7f) && (lq < 8.3f);
                LST(13, s) = lq;
                LST(11, s) = (lower ? k.L0 : k.U0) - lq * (lower ? k.L1 : k.U1);
              }
            } else {
              ++ii2;
              act2 = !(nb & 8u) && (fabsf(sm4[3]) > tol2) && ii2 < 100;
            }
          }
          if (act1 && ii1 == DIVERGENCE_CHECK_TRIP) {
            // proven-divergence exit, as in the register variant above (same bound, same threshold)
            float sm = 0.0f, unused = 0.0f;
            for (int s = tid; s < N; s += PT) {
              IterState<float> k;
              consts(s, k);
              const float n0 = k.num0, n1 = k.num1, d0 = k.den0, d1 = k.den1, A = k.A, B = k.B;
              const float al = n0 + A * n1, be = B * n1, ga = d0 + A * d1, de = B * d1;
              const float bg = be + ga, disc = bg * bg - 4.0f * al * de;
              float m = 0.0f;
              if (disc < 0.0f) {
                const float r = __fsqrt_rn(al * de - be * ga), ide = fast_rcp(de);
                const float u1 = (ga + r) * ide, u2 = (ga - r) * ide;
                const float q1 = (de * u1 - bg) * u1 + al, q2 = (de * u2 - bg) * u2 + al;
                m = fminf(fabsf(q1), fabsf(q2)) * fast_rcp(r);
              }
              sm += (m == m) ? fminf(m, 8.0f) : 0.0f;
            }
            block_sum2f(sm, unused, ctrl->redf, phase);
            if (sm > 4.0f * tol1) { executed1 = ii1; ii1 = 100; act1 = false; }
          }
        }
        if (executed1 < 0) executed1 = ii1;
        for (int s = tid; s < N; s += PT) {
          IterState<float> st;
          consts(s, st);
          st.Z1 = LST(10, s); st.Z2 = LST(11, s);
          st.lq2 = LST(13, s);
          SlotPut<!GLOBAL> put{vals + s, p.col_off, vals_sa + 4u * (unsigned)s};
          const float kd = has_kd ? vals[(long long)sl_kd * Npad + s] : Math<float>::nan();
          const float m91 = vals[(long long)sl_m91 * Npad + s];
          phase_c<float>(st, pm, ii1 >= 100, ii2 >= 100, kd, m91, put);
        }
      }
    }
    if (CL) FineHist<true>::zero(hist, lane);      // (the owner's table must be empty before any peer pushes into it)
    block_or_cluster_sync<CL>();     // all columns final; the stash is dead from here (selection reuses it)

    MCZ_T(tD0);
    // ---- optional per-sample output (pickle / --asciidistrib path) ----------------------------
    if (p.Z) {
      for (int k = 0; k < NKEYS; ++k) {
        const int sl = p.slot_of_key[k];
        const bool live = ((pm >> k) & 1ull) && sl >= 0;
        float* dst = p.Z + ((long long)g * NKEYS + k) * p.N + SOFF;
        for (int s = tid, nloc = NLOC; s < nloc; s += PT) dst[s] = live ? vals[(long long)sl * Npad + s] : Math<float>::nan();
      }
    }

    // ---- phase D: percentiles, one warp per stored column -------------------------------------
    if (CRANK == 0u && tid < NKEYS) {
      const int k = tid, sl = p.slot_of_key[k];
      const bool pres = (pm >> k) & 1ull;
      if (!(pres && sl >= 0)) {
        const long long o = (p.row0 + g) * NKEYS + k;
        for (int q = 0; q < p.ndst; ++q) {
          p.status[q][o] = pres ? ST_EMPTY : ST_ABSENT;     // present but constant-NaN (M08_R23): n = 0
          p.nvalid[q][o] = 0;
          p.pct[q][3 * o] = p.pct[q][3 * o + 1] = p.pct[q][3 * o + 2] = Math<float>::nan();
        }
      }
    }
    if (CRANK == 0u && tid == 0) {
      atomicAdd(&p.stats[0], 1ull);
      atomicAdd(&p.stats[1], (unsigned long long)(executed1 < 0 ? ii1 : executed1));
      atomicAdd(&p.stats[2], (unsigned long long)(executed2 < 0 ? ii2 : executed2));
      if (executed1 >= 0 && executed1 != ii1) atomicAdd(&p.stats[3], 1ull);
      if (executed2 >= 0 && executed2 != ii2) atomicAdd(&p.stats[4], 1ull);
      p.trips[2 * g] = ii1;
      p.trips[2 * g + 1] = ii2;
      if (p.ret) p.ret[g] = ok ? 0 : -1;
    }
    if (warp == PW - 1) fetch_next();    // (the lines / ticket of this galaxy were read into registers in phase A)
    if constexpr (CL) {
      // one warp of EVERY block per column; the warps of a column exchange tables, commands and
      // candidates through the owner block's shared memory (ClusterCol), no cluster barrier
      float* glist = st3 + SLOTS * PT + warp * CL_GLIST_FLOATS;
      const int niter = (p.nstore + PW - 1) / PW;
      ClusterCol cc;
      cc.crank = cl_rank(); cc.csize = cl_size(); cc.ng = p.cng;
      cc.cand_sa = (unsigned)__cvta_generic_to_shared(cand);
      cc.hist_sa = (unsigned)__cvta_generic_to_shared(hist);
      cc.glist_sa = (unsigned)__cvta_generic_to_shared(glist);
      cc.poison_word = p.stats + 7;
      cc.par = reinterpret_cast<unsigned*>(cand)[ClMail::PAR / 4];     // phase parities of this warp's three barriers
      for (int it = 0; it < niter; ++it) {
        if (it > 0) {            // more than 18 stored columns: the tables are reused
          cl_sync();
          FineHist<true>::zero(hist, lane);
          cl_sync();
        }
        const int sl = it * PW + warp;
        const bool has = sl < p.nstore;
        const int k = has ? p.key_of_slot[sl] : 0;
        const bool live = has && ((pm >> k) & 1ull);
        if (live) {
          cc.owner = (unsigned)sl % cc.csize;
          float* gcol = p.gscratch + ((long long)cl_id() * p.nstore + sl) * ((long long)cc.csize * p.cng * 128);
          float pc[3];
          int nv, stt;
          cc.run(vals + (long long)sl * Npad, hist, cand, glist, gcol, pc, nv, stt);
          if (cc.crank == cc.owner && lane == 0) {
            atomicAdd(&p.stats[5], 1ull);
            if (cc.did_fallback) atomicAdd(&p.stats[6], 1ull);
            if (cc.did_refine) atomicAdd(&p.stats[7], 1ull);
            const long long o = (p.row0 + g) * NKEYS + k;
            for (int q = 0; q < p.ndst; ++q) {
              p.status[q][o] = (signed char)stt;
              p.nvalid[q][o] = nv;
              p.pct[q][3 * o] = pc[0];
              p.pct[q][3 * o + 1] = pc[1];
              p.pct[q][3 * o + 2] = pc[2];
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) reinterpret_cast<unsigned*>(cand)[ClMail::PAR / 4] = cc.par;
    } else
    for (int sl = PAIRK ? (warp >> 1) : warp; sl < p.nstore; sl += PW) {
      // (PAIRK: at most half as many columns as warps -- `--md KD02`: 7 --, warps 2c and 2c + 1 share
      // column c; its histogram / candidate areas are warp 2c's)
      const int part = PAIRK ? (warp & 1) : 0;
      const int k = p.key_of_slot[sl];
      if (!((pm >> k) & 1ull)) continue;
      float pc[3];
      int nv, stt;
#ifdef MCZ_PHASE_TIMING
      const long long tc0 = clock64();
#endif
      // (global variant: the gather pass prefetches through a per-warp ring behind the candidate areas --
      // the space the KK04 loops used)
      float* ring = (GLOBAL && p.ring) ? reinterpret_cast<float*>(work + (size_t)PW * (HIST_BYTES + CAND_FLOATS * sizeof(float))) + warp * RING_FLOATS
                           : nullptr;
      if constexpr (PAIRK) {
        unsigned* phist = reinterpret_cast<unsigned*>(work + (size_t)(warp & ~1) * HIST_BYTES);
        float* pcand = cand - (warp & 1) * CAND_FLOATS;
        warp_percentiles<PACK16, 0, true>(vals + (long long)sl * Npad, Npad, phist, pcand, pc, nv, stt, ring, part, 1 + sl);
      } else {
        warp_percentiles<PACK16, GLOBAL ? 0 : SLOTS * PT / 128>(vals + (long long)sl * Npad, Npad, hist, cand, pc, nv, stt, ring);
      }
#ifdef MCZ_PHASE_TIMING
      if (lane == 0) {
        const unsigned long long dt = (unsigned long long)(clock64() - tc0);
        atomicAdd(&g_col_cycles[k], dt);
        atomicMax(&g_col_cycles[NKEYS + k], dt);
      }
#endif
      if (lane == 0 && part == 0) {
        const long long o = (p.row0 + g) * NKEYS + k;
        for (int q = 0; q < p.ndst; ++q) {
          p.status[q][o] = (signed char)stt;
          p.nvalid[q][o] = nv;
          p.pct[q][3 * o] = pc[0];
          p.pct[q][3 * o + 1] = pc[1];
          p.pct[q][3 * o + 2] = pc[2];
        }
      }
    }
#ifdef MCZ_PHASE_TIMING
    {
      const long long tD1 = clock64();
      if (lane == 0) atomicAdd(&g_phase_cycles[3], (unsigned long long)((tD1 - tD0) / PW));
      __syncthreads();
      const long long tD2 = clock64();
      MCZ_TADD(0, 1);
      MCZ_TADD(1, tB0 - tA0);
      MCZ_TADD(2, tD0 - tB0);
      MCZ_TADD(4, tD2 - tD0);
      MCZ_TADD(5, tBL0 - tB0);        // B prologue (loop state, constants)
      MCZ_TADD(6, tBL1 - tBL0);       // KK04 loops
      MCZ_TADD(7, tD0 - tBL1);        // phase C
    }
#endif
  }
}

#undef CRANK
#undef CSIZE
#undef SOFF
#undef NLOC

// ---- host side --------------------------------------------------------------------------------
static void build_slots(uint32_t tok, ProdParams& p) {
  // every key that can be present under `tok` with all lines measured, minus constant-NaN ones
  const uint64_t pm = present_mask(F_ALL, tok);
  int n = 0;
  for (int k = 0; k < NKEYS; ++k) {
    p.slot_of_key[k] = -1;
    p.key_of_slot[k] = -1;
  }
  int keys[NKEYS];
  for (int k = 0; k < NKEYS; ++k)
    if (((pm >> k) & 1ull) && !const_nan_key(k)) keys[n++] = k;
  p.nstore = n;
  // Slot s is selected by warp s % 18, and warps 0, 4, 8, 12, 16 share one SM sub-partition: with 17
  // columns it is the only one that runs five selections at once.  Give it the columns that are
  // most often empty (a diagnostic outside its validity range yields an all-NaN column, which
  // costs almost nothing), so that it usually has four like the others.
  static const int often_empty[] = {K_M13_O3N2, K_M08_O3O2, K_P10_ONS, K_P10_ON, K_KD02_N2O2, K_Z94, K_KK04_N2HA};
  bool placed[NKEYS] = {false};
  int pos = 0;
  for (int k : often_empty) {
    if (pos >= n || pos >= 18) break;
    if (!((pm >> k) & 1ull) || const_nan_key(k)) continue;
    p.slot_of_key[k] = (signed char)pos;
    p.key_of_slot[pos] = (signed char)k;
    placed[k] = true;
    pos += 4;
  }
  int sl = 0;
  for (int i = 0; i < n; ++i) {
    const int k = keys[i];
    if (placed[k]) continue;
    while (p.key_of_slot[sl] >= 0) ++sl;
    p.slot_of_key[k] = (signed char)sl;
    p.key_of_slot[sl] = (signed char)k;
  }
}
static void finish_slots(ProdParams& p) {
  for (int k = 0; k < NKEYS; ++k) p.col_off[k] = p.slot_of_key[k] >= 0 ? p.slot_of_key[k] * p.Npad : -1;
}

static uint32_t philox_calls(uint32_t tok, int correlated, int deterministic) {
  uint32_t calls = 0;
  if (tok & (T_D02 | T_M13)) calls |= 1u | 2u;          // normals 0..4
  if (!deterministic) {
    calls |= 2u;                                         // normals 5..7 (Ha, Hb, [OII])
    if (!correlated) {
      calls |= 4u;                                       // normals 8..11 ([OIII], [NII], [SII] x2)
      if (tok & T_DP00) calls |= 8u;                     // normal 12 ([SIII]9069)
    }
  }
  return calls;
}

struct ProdPlan {
  int cluster;    // > 1: one galaxy per cluster of this many blocks (columns in distributed shared memory)
  bool global;
  bool wide;      // global variant with 32-bit histogram counters (Npad >= 65536)
  int scap;       // global variant: samples with their loop state in shared memory
  int bsm;        // global variant: loop inputs + state of all samples in shared memory (20 bytes + 1 bit each)
  int ring;       // global variant: room for the per-warp prefetch rings of the gather pass
  int pair;       // global variant: two warps per column in phase D
  int cper, cng;  // cluster kernel: samples per block, 128-sample groups per block
  int Npad, grid;
  size_t smem;
  long long per_block_floats;
};

// Per-device state (one process may drive several GPUs through mcz_ctx_create(device)): the SM
// count, the "227 KB of dynamic shared memory" opt-in of every kernel (a function attribute is
// per device) and the cooperative grid limit of the wide kernel are cached per device ordinal.
static constexpr int MAX_DEVICES = 64;
enum KernelId { KID_PROD_SHARED = 0, KID_PROD_GLOBAL, KID_PROD_GLOBAL_WIDE, KID_PROD_CLUSTER, KID_WIDE, KID_SELECT, KID_PROD_GLOBAL_PAIR, KID_COUNT };
struct DeviceState {
  std::atomic<int> sms{0};
  std::atomic<int> attr_done[KID_COUNT];
  std::atomic<int> wide_max_blocks{0};
  DeviceState() { for (auto& a : attr_done) a.store(0); }
};
static DeviceState g_dev[MAX_DEVICES];
static DeviceState& device_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev[(dev >= 0 && dev < MAX_DEVICES) ? dev : 0];
}
static int num_sms() {
  DeviceState& d = device_state();
  int n = d.sms.load();
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    d.sms.store(n);
  }
  return n;
}
template <class K> static int ensure_smem_attr(K kernel, KernelId id, int bytes) {
  DeviceState& d = device_state();
  if (d.attr_done[id].load() >= bytes) return 0;
  const int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "smem attribute");
  if (rc) return rc;
  d.attr_done[id].store(bytes);
  return 0;
}

// how many clusters of `csz` blocks of the cluster kernel can be resident at once on this device
// (0: the device / driver refuses the shape); cached per device and cluster size
static int max_active_clusters(int csz, size_t smem) {
  static std::atomic<int> cache[MAX_DEVICES][9];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= MAX_DEVICES) dev = 0;
  const int c = cache[dev][csz].load();
  if (c != 0) return c < 0 ? 0 : c;
  int n = 0;
  if (ensure_smem_attr(production_kernel<4, false, true>, KID_PROD_CLUSTER, 227 * 1024) == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csz * 64u);
    cfg.blockDim = dim3(PT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, production_kernel<4, false, true>, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
  } else {
    cudaGetLastError();
  }
  cache[dev][csz].store(n > 0 ? n : -1);
  return n > 0 ? n : 0;
}

// MCZ_CLUSTER=1 / 0 forces the cluster kernel on / off for the shapes it covers; default: see DESIGN.md
static bool cluster_wanted(const char* env) {
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return true;
  return MCZ_CLUSTER_DEFAULT != 0;
}

static ProdPlan make_plan(long long G, long long N, int nstore) {
  ProdPlan pl;
  // shared-memory variant: fixed column stride of 4*PT samples; layout = control block, columns,
  // PW histograms, PW candidate areas, one extra per-sample array
  const size_t smem_local = CTRL_BYTES + (size_t)nstore * 4 * PT * sizeof(float) +
                            (size_t)PW * (FineHist<true>::BYTES + CAND_FLOATS * sizeof(float)) +
                            (size_t)4 * PT * sizeof(float);
  pl.global = !(N <= 4ll * PT && smem_local <= 227ull * 1024);
  pl.wide = false;
  pl.scap = 0;
  pl.bsm = 0;
  pl.ring = 0;
  pl.pair = 0;
  pl.cluster = 1;
  pl.cper = 0; pl.cng = 0;
  // Longer columns, up to 8 x 2304 samples: a cluster of blocks per galaxy, each block laid out like the
  // one-block kernel plus the owner's candidate lists (MCZ_CLUSTER=0: never).  The cluster size is at
  // least ceil(N' / 2304); a larger one can keep more SMs busy (clusters do not span GPCs: with 18 or
  // 20 SMs per GPC, clusters of 5 use 130 of the 148 SMs, clusters of 6 use 144), so the size with the
  // best modelled throughput is taken: resident clusters / (fixed cost + per-sample cost x samples per
  // block), the two costs from the phase timings of the kernel.  MCZ_CLUSTER_SIZE=n forces a size.
  const size_t smem_cluster = smem_local + (size_t)PW * CL_GLIST_FLOATS * sizeof(float);
  const long long cmin = (N + 4ll * PT - 1) / (4ll * PT);
  const char* envc = getenv("MCZ_CLUSTER");
  if (pl.global && cmin >= 2 && cmin <= 8 && smem_cluster <= 227ull * 1024 && cluster_wanted(envc)) {
    const char* envs = getenv("MCZ_CLUSTER_SIZE");
    const int forced = envs ? atoi(envs) : 0;
    int best = 0, best_n = 0;
    double best_rate = 0.0;
    for (int c = (int)cmin; c <= 8; ++c) {
      if (forced >= cmin && forced <= 8 && c != forced) continue;
      const int n = max_active_clusters(c, smem_cluster);
      if (n <= 0) continue;
      const double per = (double)((N + c - 1) / c);
      const double rate = (double)(G < n ? (G > 0 ? G : 1) : n) / (14000.0 + 17.0 * per);
      if (rate > best_rate) { best_rate = rate; best = c; best_n = n; }
    }
    if (best > 0) {
      pl.global = false;
      pl.cluster = best;
      pl.Npad = 4 * PT;
      pl.smem = smem_cluster;
      pl.per_block_floats = 0;
      pl.cper = (int)((N + best - 1) / best);
      pl.cng = (pl.cper + 127) / 128;
      const long long nc = G < best_n ? (G > 0 ? G : 1) : best_n;
      pl.grid = (int)(nc * best);
      return pl;
    }
  }
  if (!pl.global) {
    pl.Npad = 4 * PT;
    pl.smem = smem_local;
    pl.per_block_floats = 0;
  } else {
    pl.Npad = (int)((N + 127) / 128 * 128);
    pl.wide = pl.Npad >= 65536;
    const size_t base = CTRL_BYTES + (size_t)PW * ((pl.wide ? FineHist<false>::BYTES : FineHist<true>::BYTES) +
                                                    CAND_FLOATS * sizeof(float));
    // the rest of the 227 KB holds the per-trip loop state (16 B per sample) of as many samples as fit
    const long long fit = (long long)((227ull * 1024 - base) / 16 / PT) * PT;
    const long long need = (N + PT - 1) / PT * PT;
    pl.scap = (int)(fit < need ? fit : need);
    pl.smem = base + (size_t)pl.scap * 16;
    // N' <= ~11400 (configs 3 and 5): logO3O2, logN2Ha, logR23 and the two logq of EVERY sample (20 bytes)
    // plus one flag bit fit in the 226 KB next to the control block -- the selection work area is dead
    // while the loops run -- so the loops read nothing from global memory (MCZ_BSM=0: never)
    // (phase D: a 9 KB prefetch ring per warp in the same space)
    {
      const char* envp = getenv("MCZ_PAIR");           // MCZ_PAIR=0: one warp per column whatever their number
      pl.pair = (!pl.wide && 2 * nstore <= PW && !(envp && envp[0] == '0')) ? 1 : 0;
    }
    const size_t ring_bytes = base + (size_t)PW * RING_FLOATS * sizeof(float);
    pl.ring = ring_bytes <= 227ull * 1024;             // (not next to the 32-bit histogram counters of >= 65536-sample columns)
    if (pl.ring && pl.smem < ring_bytes) pl.smem = ring_bytes;
    const size_t nq = (size_t)((N + 3) / 4 * 4);
    const size_t bsm_bytes = CTRL_BYTES + 20 * nq + 4 * ((nq + 31) / 32) + 16;
    const char* envb = getenv("MCZ_BSM");
    if (bsm_bytes <= 227ull * 1024 && !(envb && envb[0] == '0')) {
      pl.bsm = 1;
      pl.scap = 0;
      pl.smem = (!pl.ring || bsm_bytes > ring_bytes) ? bsm_bytes : ring_bytes;
      if (pl.smem < base) pl.smem = base;
    }
    pl.per_block_floats = ((long long)nstore + 4 + 6) * pl.Npad;
  }
  const long long sms = num_sms();
  pl.grid = (int)(G < sms ? (G > 0 ? G : 1) : sms);
  return pl;
}

// The launch statistics live in the header of the caller's scratch buffer (no process-wide state):
// bytes [64, 128) = galaxies, N2Ha trips executed, R23 trips executed, N2Ha proven-divergence exits,
// R23 periodic-state exits, cluster kernel: columns selected, of which by the whole-column path; 1 spare.
int read_launch_stats(const void* scratch_d, unsigned long long* out8, cudaStream_t stream) {
  int rc = check_cuda(cudaMemcpyAsync(out8, reinterpret_cast<const unsigned char*>(scratch_d) + SCRATCH_STATS_OFF,
                                      8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream), "read launch stats");
  if (rc) return rc;
  return check_cuda(cudaStreamSynchronize(stream), "read launch stats");
}

int read_col_cycles(unsigned long long* out, int reset) {
  int rc = check_cuda(cudaMemcpyFromSymbol(out, g_col_cycles, 2 * NKEYS * sizeof(unsigned long long)), "read col cycles");
  if (rc) return rc;
  if (reset) {
    unsigned long long z[2 * NKEYS] = {0};
    rc = check_cuda(cudaMemcpyToSymbol(g_col_cycles, z, sizeof(z)), "reset col cycles");
  }
  return rc;
}

int read_phase_cycles(unsigned long long* out8, int reset) {
  int rc = check_cuda(cudaMemcpyFromSymbol(out8, g_phase_cycles, 8 * sizeof(unsigned long long)), "read cycles");
  if (rc) return rc;
  if (reset) {
    unsigned long long z[8] = {0};
    rc = check_cuda(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)), "reset cycles");
  }
  return rc;
}

int read_sel_counters(unsigned long long* out8, int reset) {
  int rc = check_cuda(cudaMemcpyFromSymbol(out8, g_sel_counters, 8 * sizeof(unsigned long long)), "read counters");
  if (rc) return rc;
  if (reset) {
    unsigned long long z[8] = {0};
    rc = check_cuda(cudaMemcpyToSymbol(g_sel_counters, z, sizeof(z)), "reset counters");
  }
  return rc;
}

static size_t wide_scratch_bytes(long long N, int nstore);
static bool use_wide(long long G, long long N, const float* Z);
static int launch_wide(ProdParams& p, void* scratch, cudaStream_t stream, const WideShare* ws);

size_t production_scratch_bytes(long long G, long long n_draw, uint32_t tok) {
  ProdParams p;
  build_slots(tok, p);
  const ProdPlan pl = make_plan(G, n_draw, p.nstore);
  size_t need = SCRATCH_HEADER + (size_t)pl.grid * (size_t)pl.per_block_floats * sizeof(float);
  if (pl.cluster > 1)     // the fallback columns of every cluster: nstore x (cluster x 2304) floats at most
    need = SCRATCH_HEADER + (size_t)(pl.grid / pl.cluster) * (size_t)p.nstore * (size_t)pl.cluster * (size_t)pl.Npad * sizeof(float);
  if (use_wide(G, n_draw, nullptr)) {
    const size_t w = wide_scratch_bytes(n_draw, p.nstore);
    if (w > need) need = w;
  }
  return need;
}

// name of the kernel launch_production() picks for this launch shape on the current device
const char* production_kernel_name(long long G, long long n_draw, uint32_t tok) {
  ProdParams p;
  build_slots(tok, p);
  if (use_wide(G, n_draw, nullptr)) return "wide_kernel";
  const ProdPlan pl = make_plan(G, n_draw, p.nstore);
  if (pl.cluster > 1) return "production_kernel<4,0,cluster>";
  if (!pl.global) return "production_kernel<4,0>";
  return pl.wide ? "production_kernel<0,1>" : (pl.pair ? "production_kernel<0,0,pair>" : "production_kernel<0,0>");
}

int launch_production(const float* meas, const float* err, long long G, long long n_draw, uint32_t tok,
                      int dust, unsigned long long seed, unsigned long long galaxy_offset,
                      int sample_mode, int deterministic, float* pct, int* nvalid,
                      signed char* status, int* trips, int* ret, float* Z, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream, const ProdDst* dst, const WideShare* ws) {
  if (G <= 0) return 0;
  if (n_draw <= 0 || n_draw > 0x7fffffffll / 64) return set_error("mcz_forward_production: bad n_draw");
  ProdParams p;
  build_slots(tok, p);
  const ProdPlan pl = make_plan(G, n_draw, p.nstore);
  if (scratch_bytes < ((ws && ws->world > 1) ? wide_private_bytes(n_draw, tok) : production_scratch_bytes(G, n_draw, tok)))
    return set_error("mcz_forward_production: scratch too small");
  p.meas = meas; p.err = err; p.G = G; p.N = (int)n_draw; p.Npad = pl.Npad; p.scap = pl.scap; p.bsm = pl.bsm; p.ring = pl.ring; p.pair = pl.pair; p.cper = pl.cper; p.cng = pl.cng; p.tok = tok; p.dust = dust;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.goff = galaxy_offset;
  p.correlated = sample_mode == 1; p.deterministic = deterministic != 0;
  p.calls = philox_calls(tok, p.correlated, p.deterministic);
  finish_slots(p);
  if (dst && dst->n > 0) {
    if (dst->n > MAX_DST) return set_error("mcz_forward_production: more than 8 destination tables");
    p.ndst = dst->n;
    p.row0 = dst->row0;
    for (int q = 0; q < dst->n; ++q) {
      p.pct[q] = dst->pct[q]; p.nvalid[q] = dst->nvalid[q]; p.status[q] = dst->status[q];
      if (!p.pct[q] || !p.nvalid[q] || !p.status[q]) return set_error("mcz_forward_production: null destination table");
    }
  } else {
    p.ndst = 1;
    p.row0 = 0;
    p.pct[0] = pct; p.nvalid[0] = nvalid; p.status[0] = status;
  }
  p.trips = trips; p.ret = ret; p.Z = Z;
  int rc = check_cuda(cudaMemsetAsync(scratch, 0, SCRATCH_HEADER, stream), "scratch header memset");
  if (rc) return rc;
  p.counter = reinterpret_cast<unsigned long long*>(scratch);
  p.stats = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(scratch) + SCRATCH_STATS_OFF);
  if ((ws && ws->world > 1) || use_wide(G, n_draw, Z)) return launch_wide(p, scratch, stream, ws);
  p.gscratch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + SCRATCH_HEADER);
  p.gscratch_per_block = pl.per_block_floats;
  if (getenv("MCZ_DEBUG_PLAN"))
    fprintf(stderr, "mcz plan: cluster %d global %d wide %d grid %d smem %zu Npad %d\n", pl.cluster, (int)pl.global, (int)pl.wide,
            pl.grid, pl.smem, pl.Npad);
  if (pl.cluster > 1) {
    if ((rc = ensure_smem_attr(production_kernel<4, false, true>, KID_PROD_CLUSTER, 227 * 1024))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3(PT);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)pl.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    rc = check_cuda(cudaLaunchKernelEx(&cfg, production_kernel<4, false, true>, p), "cluster kernel launch");
    if (rc) return rc;
    if (getenv("MCZ_DEBUG_POISON")) {      // debugging aid: wait for the kernel and report a timed-out protocol wait
      unsigned long long st[8];
      if (read_launch_stats(scratch, st, stream) == 0 && (st[7] >> 63))
        fprintf(stderr, "mcz cluster kernel: protocol wait timed out (where = %llu), results invalid\n", (st[7] >> 56) & 63ull);
    }
  } else if (!pl.global) {
    if ((rc = ensure_smem_attr(production_kernel<4, false, false>, KID_PROD_SHARED, 227 * 1024))) return rc;
    production_kernel<4, false, false><<<pl.grid, PT, pl.smem, stream>>>(p);
  } else if (!pl.wide && pl.pair) {
    if ((rc = ensure_smem_attr(production_kernel<0, false, false, true>, KID_PROD_GLOBAL_PAIR, 227 * 1024))) return rc;
    production_kernel<0, false, false, true><<<pl.grid, PT, pl.smem, stream>>>(p);
  } else if (!pl.wide) {
    if ((rc = ensure_smem_attr(production_kernel<0, false, false>, KID_PROD_GLOBAL, 227 * 1024))) return rc;
    production_kernel<0, false, false><<<pl.grid, PT, pl.smem, stream>>>(p);
  } else {
    if ((rc = ensure_smem_attr(production_kernel<0, true, false>, KID_PROD_GLOBAL_WIDE, 227 * 1024))) return rc;
    production_kernel<0, true, false><<<pl.grid, PT, pl.smem, stream>>>(p);
  }
  count_launch();
  return check_cuda(cudaGetLastError(), "production_kernel launch");
}

// ---- wide kernel: one galaxy at a time, samples tiled across all thread blocks ------------------
// For single-object runs (few galaxies, N' in the 10^5 .. 10^7 range) one thread block
// per galaxy would leave all but a few SMs idle.  This cooperative kernel gives every block a
// contiguous slice of the samples of the SAME galaxy; what the phases need from the whole sample
// set goes through global memory and grid-wide barriers:
//   A  has* flags: atomicOr;
//   B  the two KK04 means per trip: fixed-point 64-bit atomics, one grid barrier per two trips;
//   D  an MSB radix select on the order-preserving bit patterns of the floats, 8 bits per round, for
//      the six ranks at once: per-slice histograms (shared memory) merged into a global table, the
//      ranks located redundantly by every block after the barrier; four rounds give exact values.
// Per-sample results are identical to the per-galaxy kernels (same Philox indexing, same device
// functions).  With p.wworld > 1 the samples of the galaxy are also sharded over the GPUs of the
// node: every atomic goes to every rank's copy of the control block / radix tables (NVLink peer
// mappings) and each grid barrier is followed by a counter barrier across the ranks.
struct WideCtl {
  unsigned flags;                 // line flags | SEEN_KD
  unsigned nan[3];                // NaN bits of the KK04 sums, rotating with sums[]
  long long sums[3][4];           // fixed point, 18 fractional bits, per-sample rounding (wfx)
  double colsum[NKEYS];           // per slot: sum of the finite values (sum <= 0 test, mcz.py:282-285)
  unsigned umaxo[NKEYS];          // per slot: max of the ordered bit patterns of the finite values ...
  unsigned nmino[NKEYS];          // ... and max of their complements (= ~min): the radix select skips the common prefix
};
static_assert(sizeof(WideCtl) <= 4096, "wide control block too large");

// Wide kernel: the per-sample terms of the KK04 convergence sums are rounded to 2^-18 ONE BY ONE and
// added as integers (the 1.5 * 2^23 bias makes the FMA itself round to an integer held in the low
// mantissa bits; exact for |d| <= 8).  Integer sums do not depend on how the samples are grouped
// into threads, blocks or GPUs, so a run gives the same trip counts -- hence bit-identical tables --
// whatever the grid and however many ranks share the object.
static constexpr float WFX_SCALE = 262144.0f;     // 2^18
__device__ __forceinline__ int wfx(float d) {     // d already clamped to [-8, 8]; NaN is flagged by the caller
  return __float_as_int(fmaf(d, WFX_SCALE, 12582912.0f)) - 0x4B400000;
}

__global__ void __launch_bounds__(PT, 1)
wide_kernel(const __grid_constant__ ProdParams p) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem[];
  typedef FineHist<false> FH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = gridDim.x, b = blockIdx.x;
  const int N = p.N, Npad = p.Npad, ngroups = Npad / 128;
  // this rank's share of the 128-sample groups, and this block's slice of it
  const int rg_lo = (int)((long long)ngroups * p.wrank / p.wworld), rg_hi = (int)((long long)ngroups * (p.wrank + 1) / p.wworld);
  const int g_lo = rg_lo + (int)((long long)(rg_hi - rg_lo) * b / nb), g_hi = rg_lo + (int)((long long)(rg_hi - rg_lo) * (b + 1) / nb);
  const int s_lo = g_lo * 128, s_hi = g_hi * 128;
  unsigned* hist = reinterpret_cast<unsigned*>(smem) + (size_t)warp * FB;

  // shared part (this rank's copy; every rank's copy receives every rank's atomics)
  const int W = p.wworld, R = p.wrank;
  unsigned char* shl = reinterpret_cast<unsigned char*>(p.wshared[R]);
  WideCtl* ctl = reinterpret_cast<WideCtl*>(shl);
  unsigned* ghist = reinterpret_cast<unsigned*>(shl + 4096);             // [4 rounds][nstore][6 targets][256]
  const size_t ghist_bytes = (size_t)4 * p.nstore * 6 * 256 * sizeof(unsigned);
  unsigned* xbar = reinterpret_cast<unsigned*>(shl + 4096 + ghist_bytes);   // cross-rank barrier counter
  auto peer = [&](void* local, int q) -> void* {
    return reinterpret_cast<unsigned char*>(p.wshared[q]) + (reinterpret_cast<unsigned char*>(local) - shl);
  };
  unsigned xepoch = 0;
  // grid barrier; with several ranks also a barrier across them (every rank's atomics have landed)
  auto xsync = [&]() {
    if (W > 1) __threadfence_system();
    grid.sync();
    if (W > 1) {
      ++xepoch;
      if (b == 0 && tid == 0) {
        for (int q = 0; q < W; ++q) atomicAdd_system(reinterpret_cast<unsigned*>(peer(xbar, q)), 1u);
        while (*reinterpret_cast<volatile unsigned*>(xbar) < xepoch * (unsigned)W) { }
        __threadfence_system();
      }
      grid.sync();
    }
  };
  unsigned char* base = reinterpret_cast<unsigned char*>(p.gscratch);
  float* vals = reinterpret_cast<float*>(base);
  float* st0 = vals + (long long)p.nstore * Npad;
  float* st1 = st0 + Npad; float* st2 = st1 + Npad; float* st3 = st2 + Npad;
  float* F = st3 + Npad;                                   // fields 0..5: Z1, Z2, lq1, lq2, lq1a, lq2a
  // Several ranks: the blocks of a rank first merge their tables of a radix round in this LOCAL table
  // [nstore][6][256]; after a grid barrier its non-zero counters are sent to every rank once (the private
  // scratch of a sharded run has room behind the loop state: wide_scratch_bytes counts the shared
  // region, which then lives in the symmetric buffer).  148 blocks x 17 columns x W destinations of
  // system-scope atomics per round took ~0.14 ms per peer.
  unsigned* lhist = reinterpret_cast<unsigned*>(F + 6ll * Npad);
  const int used2line[NUSED] = {L_HA, L_HB, L_O2, L_O3, L_N2, L_S2A, L_S2B, L_S39069};

  // (-DMCZ_PHASE_TIMING: cycles of block 0 between the grid barriers, summed into g_phase_cycles[1..7])
#ifdef MCZ_PHASE_TIMING
  long long wt_ = clock64();
#define WT(i) do { const long long n_ = clock64(); if (b == 0 && tid == 0) atomicAdd(&g_phase_cycles[i], (unsigned long long)(n_ - wt_)); wt_ = n_; } while (0)
#else
#define WT(i)
#endif
  for (long long g = 0; g < p.G; ++g) {
    if (b == 0 && tid == 0) { MCZ_TADD(0, 1); }
    // ---- reset the control block and the merged histograms ------------------------------------
    if (b == 0) {
      for (int i = tid; i < (int)(sizeof(WideCtl) / 4); i += PT) reinterpret_cast<unsigned*>(ctl)[i] = 0u;
    }
    for (long long i = (long long)b * PT + tid; i < 4ll * p.nstore * 6 * 256; i += (long long)nb * PT) ghist[i] = 0u;
    if (W > 1)
      for (long long i = (long long)b * PT + tid; i < (long long)p.nstore * 6 * 256; i += (long long)nb * PT) lhist[i] = 0u;
    xsync();
    WT(1);   // reset
    float lm[NUSED], le[NUSED];
    uint32_t assumed = 0u;
#pragma unroll
    for (int u = 0; u < NUSED; ++u) {
      lm[u] = p.meas[g * NLINES + used2line[u]];
      le[u] = p.deterministic ? 0.0f : p.err[g * NLINES + used2line[u]];
      const float thr = (u >= U_S2B) ? 1e-9f : 0.0f;
      bool a = finite_f(lm[u]) && finite_f(le[u]);
      if (a && le[u] == 0.0f) a = lm[u] > thr;
      if (a) assumed |= 1u << u;
    }
    const unsigned long long gg = p.goff + (unsigned long long)g;
    // ---- phase A ------------------------------------------------------------------------------
    uint32_t flags = assumed;
    bool anykd = false;
    for (;;) {
      const int de = effective_dust(p.dust, flags);
      uint32_t seen = 0u;
      // (as in the per-galaxy kernels: the launch shape `--md all`, every line measured, dust on, independent
      // sampling gets a copy of the loop with flags / tokens / dust mode as compile-time constants)
      const uint32_t need = F_ALL & ~F_S39069;
      bool bounded = true;
#pragma unroll
      for (int u = 0; u < NUSED; ++u) bounded = bounded && (fabsf(lm[u]) + 6.0f * fabsf(le[u]) < FLT_MAX);
      const bool full = (flags & need) == need && de == 0 && !p.correlated && !p.deterministic && bounded &&
                        p.tok == T_ALL && p.calls == 7u;
      auto pass = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;      // 0 generic, 1 --md all
        const uint32_t fl = MODE ? (uint32_t)F_ALL : flags;
        const uint32_t tk = MODE ? (uint32_t)T_ALL : p.tok;
        const int dm = MODE ? 0 : de;
        for (int s = s_lo + tid; s < s_hi; s += PT) {
          if (s < N) {
            float z[16];
            if (MODE == 1) draw_normals_c<7u>((uint32_t)s, gg, p, z);
            else draw_normals((uint32_t)s, gg, p, z);
            float f[NUSED], e[5];
#pragma unroll
            for (int u = 0; u < NUSED; ++u) {
              const float raw = fmaf(le[u], (!MODE && p.correlated) ? z[5] : z[5 + u], lm[u]);
              float v = fmaxf(raw, 0.0f);
              if (!MODE && !(v <= FLT_MAX)) v = 0.0f;
              if (!MODE && (p.tok & T_DROPNEG) && raw < 0.0f && raw >= -FLT_MAX) v = Math<float>::nan();
              f[u] = v;
              const bool pos = (u >= U_S2B) ? (v > 1e-9f) : (v > 0.0f);
              if (pos) seen |= 1u << u;
            }
            e[0] = 0.05f * z[0]; e[1] = 0.1f * z[1]; e[2] = 0.027f * z[2]; e[3] = 0.024f * z[3]; e[4] = 0.012f * z[4];
            SlotPut<false> put{vals + s, p.col_off, 0u};
            Carry<float> c;
            c.y = c.n2ha = c.x = 0.0f;
            c.aux = 0u;
            phase_a<float>(f, e, fl, tk, dm, put, c);
            if (c.aux & AUX_KD_VALID) seen |= SEEN_KD;
            st0[s] = c.y; st1[s] = c.n2ha; st2[s] = c.x; st3[s] = __uint_as_float(c.aux);
          } else {
            for (int sl = 0; sl < p.nstore; ++sl) vals[(long long)sl * Npad + s] = Math<float>::nan();
          }
        }
      };
      if (full) pass(std::integral_constant<int, 1>());
      else pass(std::integral_constant<int, 0>());
      seen = __reduce_or_sync(0xffffffffu, seen);
      if (lane == 0 && seen)
        for (int q = 0; q < W; ++q) atomicOr_system(reinterpret_cast<unsigned*>(peer(&ctl->flags, q)), seen);
      xsync();
      WT(2);   // phase A
      const uint32_t all = *reinterpret_cast<volatile unsigned*>(&ctl->flags);
      const uint32_t actual = all & F_ALL;
      anykd = (all & SEEN_KD) != 0u;
      if (actual == flags) break;
      flags = actual;
      xsync();                                       // everyone has read the flags of the discarded pass
      if (b == 0 && tid == 0) ctl->flags = actual;
      xsync();
    }
    const uint64_t pm = present_mask(flags, p.tok);
    const bool ok = minimum_ok(flags);

    // ---- phases B and C -----------------------------------------------------------------------
    int ii1 = 0, ii2 = 0;
    if ((p.tok & T_KD02) && ok) {
      const bool has_kd = (pm >> K_KD02_N2O2) & 1ull;
      const bool has_kn = (pm >> K_KK04_N2HA) & 1ull, has_kr = (pm >> K_KK04_R23) & 1ull;
      const bool o3o2 = (flags & F_O2) && (flags & F_O3);
      const int sl_kd = p.slot_of_key[K_KD02_N2O2], sl_m91 = p.slot_of_key[K_M91];
      bool act1 = has_kn && o3o2, act2 = has_kr;
      const long long itol1 = (long long)(1.0e-3 * (double)N * (double)WFX_SCALE), itol2 = (long long)(1.0e-4 * (double)N * (double)WFX_SCALE);
      auto consts = [&](int s, IterState<float>& st) {
        Carry<float> c;
        c.y = st0[s]; c.n2ha = st1[s]; c.x = st2[s];
        c.aux = __float_as_uint(st3[s]);
        make_iter<float>(c, 0.0f, st);
      };
      for (int s = s_lo + tid; s < s_hi && s < N; s += PT) {
        IterState<float> st;
        consts(s, st);
        float z1 = has_kd ? vals[(long long)sl_kd * Npad + s] : Math<float>::nan();
        if (!anykd) z1 = 8.6f;                                             // metscales.py:1097-1103
        if (has_kn && !o3o2) z1 = st.A - 7.37177f * st.B;                  // metscales.py:1122-1126
        F[0ll * Npad + s] = z1; F[1ll * Npad + s] = st.Z2;
        F[2ll * Npad + s] = st.lq1; F[3ll * Npad + s] = st.lq2;
      }
      int ph = 0;
      // fixed-point block -> grid sums of (up to) four terms; returns the four totals and NaN bits
      auto grid_sums = [&](const long long sm4[4], unsigned nanb, long long t4[4], unsigned& nb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // (a thread holds at most a few hundred samples of <= 2^21 each; the warp total may pass 2^31)
          long long q = sm4[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          if (lane == 0 && q)
            for (int r = 0; r < W; ++r)
              atomicAdd_system(reinterpret_cast<unsigned long long*>(peer(&ctl->sums[ph][k], r)), (unsigned long long)q);
        }
        nanb = __reduce_or_sync(0xffffffffu, nanb);
        if (lane == 0 && nanb)
          for (int r = 0; r < W; ++r) atomicOr_system(reinterpret_cast<unsigned*>(peer(&ctl->nan[ph], r)), nanb);
        const int nx = ph == 2 ? 0 : ph + 1;
        if (b == 0 && tid < 4) ctl->sums[nx][tid] = 0;
        if (b == 0 && tid == 4) ctl->nan[nx] = 0u;
        xsync();
        WT(3);   // KK04 batches
#pragma unroll
        for (int k = 0; k < 4; ++k) t4[k] = *reinterpret_cast<volatile long long*>(&ctl->sums[ph][k]);
        nb = *reinterpret_cast<volatile unsigned*>(&ctl->nan[ph]);
        ph = nx;
      };
      while (act1 || act2) {
        long long sm4[4] = {0, 0, 0, 0};
        unsigned nanb = 0u;
        // two samples of the thread per iteration, all sixteen loads first: the state comes from global
        // memory (L2 at best) and one sample's loads in flight per thread leave the pass latency-bound
        for (int s = s_lo + tid; s < s_hi && s < N; s += 2 * PT) {
          const bool two = s + PT < s_hi && s + PT < N;
          const int sx[2] = {s, two ? s + PT : s};
          float c0[2], c1[2], c2[2], c3[2], vZ1[2], vq1[2], vZ2[2], vq2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            c0[u] = st0[sx[u]];
            c1[u] = act1 ? st1[sx[u]] : 0.0f;
            vZ1[u] = act1 ? F[0ll * Npad + sx[u]] : 0.0f;
            vq1[u] = act1 ? F[2ll * Npad + sx[u]] : 0.0f;
            c2[u] = act2 ? st2[sx[u]] : 0.0f;
            c3[u] = act2 ? st3[sx[u]] : 0.0f;
            vZ2[u] = act2 ? F[1ll * Npad + sx[u]] : 0.0f;
            vq2[u] = act2 ? F[3ll * Npad + sx[u]] : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) continue;
            const int sq = sx[u];
            IterState<float> k;                         // only the constants of the loops still running
            iter_logq<float>(c0[u], k);
            if (act1) iter_n2ha<float>(c1[u], k);
            if (act2) { iter_r23<float>(c2[u], k); k.aux = __float_as_uint(c3[u]); }
            const float num0 = k.num0, num1 = k.num1, den0 = k.den0, den1 = k.den1;
            if (act1) {
              const float Z1 = vZ1[u], lq1 = vq1[u];
              const float lq = (num0 + Z1 * num1) * fast_rcp(den0 + Z1 * den1);
              const float z = k.A - lq * k.B;
              float d = fabsf(lq - lq1);
              nanb |= (d != d) ? 1u : 0u;
              sm4[0] += wfx(fminf(d, 8.0f));
              const float lqb = (num0 + z * num1) * fast_rcp(den0 + z * den1);
              d = fabsf(lqb - lq);
              nanb |= (d != d) ? 2u : 0u;
              sm4[1] += wfx(fminf(d, 8.0f));
              F[0ll * Npad + sq] = k.A - lqb * k.B;
              F[2ll * Npad + sq] = lqb;
              F[4ll * Npad + sq] = lq;
            }
            if (act2) {
              const float Z2 = vZ2[u], lq2 = vq2[u];
              const bool lowz = (k.aux & AUX_ZI_MASK) <= 1u;
              const float lq = (num0 + Z2 * num1) * fast_rcp(den0 + Z2 * den1);
              bool lower = lowz && (lq >= 6.7f) && (lq < 8.3f);
              const float z = (lower ? k.L0 : k.U0) - lq * (lower ? k.L1 : k.U1);
              float d = lq2 - lq;
              nanb |= (d != d) ? 4u : 0u;
              sm4[2] += wfx(fminf(fmaxf(d, -8.0f), 8.0f));
              const float lqb = (num0 + z * num1) * fast_rcp(den0 + z * den1);
              lower = lowz && (lqb >= 6.7f) && (lqb < 8.3f);
              d = lq - lqb;
              nanb |= (d != d) ? 8u : 0u;
              sm4[3] += wfx(fminf(fmaxf(d, -8.0f), 8.0f));
              F[1ll * Npad + sq] = (lower ? k.L0 : k.U0) - lqb * (lower ? k.L1 : k.U1);
              F[3ll * Npad + sq] = lqb;
              F[5ll * Npad + sq] = lq;
            }
          }
        }
        long long t4[4];
        unsigned nb;
        grid_sums(sm4, nanb, t4, nb);
        if (act1) {
          ++ii1;
          if ((nb & 1u) || !(t4[0] > itol1) || ii1 >= 100) {
            act1 = false;
            for (int s = s_lo + tid; s < s_hi && s < N; s += PT) {
              IterState<float> k;
              consts(s, k);
              const float lq = F[4ll * Npad + s];
              F[2ll * Npad + s] = lq;
              F[0ll * Npad + s] = k.A - lq * k.B;
            }
          } else {
            ++ii1;
            act1 = !(nb & 2u) && (t4[1] > itol1) && ii1 < 100;
          }
        }
        if (act2) {
          ++ii2;
          if ((nb & 4u) || !((t4[2] < 0 ? -t4[2] : t4[2]) > itol2) || ii2 >= 100) {
            act2 = false;
            for (int s = s_lo + tid; s < s_hi && s < N; s += PT) {
              IterState<float> k;
              consts(s, k);
              const float lq = F[5ll * Npad + s];
              const bool lower = ((k.aux & AUX_ZI_MASK) <= 1u) && (lq >= 6.7f) && (lq < 8.3f);
              F[3ll * Npad + s] = lq;
              F[1ll * Npad + s] = (lower ? k.L0 : k.U0) - lq * (lower ? k.L1 : k.U1);
            }
          } else {
            ++ii2;
            act2 = !(nb & 8u) && ((t4[3] < 0 ? -t4[3] : t4[3]) > itol2) && ii2 < 100;
          }
        }
        if (act1 && ii1 == DIVERGENCE_CHECK_TRIP) {      // proven-divergence exit (see production_kernel)
          long long sm[4] = {0, 0, 0, 0};
          for (int s = s_lo + tid; s < s_hi && s < N; s += PT) {
            IterState<float> k;
            consts(s, k);
            const float al = k.num0 + k.A * k.num1, be = k.B * k.num1, ga = k.den0 + k.A * k.den1, de = k.B * k.den1;
            const float bg = be + ga, disc = bg * bg - 4.0f * al * de;
            float m = 0.0f;
            if (disc < 0.0f) {
              const float r = __fsqrt_rn(al * de - be * ga), ide = fast_rcp(de);
              const float u1 = (ga + r) * ide, u2 = (ga - r) * ide;
              const float q1 = (de * u1 - bg) * u1 + al, q2 = (de * u2 - bg) * u2 + al;
              m = fminf(fabsf(q1), fabsf(q2)) * fast_rcp(r);
            }
            sm[0] += wfx((m == m) ? fminf(m, 8.0f) : 0.0f);
          }
          long long t4b[4];
          unsigned nbb;
          grid_sums(sm, 0u, t4b, nbb);
          if (t4b[0] > 4 * itol1) { ii1 = 100; act1 = false; }
        }
      }
      // two samples per thread, all eighteen loads first (the column stores of one sample would keep the
      // loads of the next behind them: everything here is derived from one scratch pointer)
      for (int s = s_lo + tid; s < s_hi && s < N; s += 2 * PT) {
        const bool two = s + PT < s_hi && s + PT < N;
        const int sx[2] = {s, two ? s + PT : s};
        float cy[2], cn[2], cx[2], ca[2], z1[2], z2[2], l2[2], kdv[2], m9[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int t = sx[q];
          cy[q] = st0[t]; cn[q] = st1[t]; cx[q] = st2[t]; ca[q] = st3[t];
          z1[q] = F[0ll * Npad + t]; z2[q] = F[1ll * Npad + t]; l2[q] = F[3ll * Npad + t];
          kdv[q] = has_kd ? vals[(long long)sl_kd * Npad + t] : Math<float>::nan();
          m9[q] = vals[(long long)sl_m91 * Npad + t];
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          Carry<float> c;
          c.y = cy[q]; c.n2ha = cn[q]; c.x = cx[q];
          c.aux = __float_as_uint(ca[q]);
          IterState<float> st;
          make_iter<float>(c, 0.0f, st);
          st.Z1 = z1[q]; st.Z2 = z2[q]; st.lq2 = l2[q];
          SlotPut<false> put{vals + sx[q], p.col_off, 0u};
          phase_c<float>(st, pm, ii1 >= 100, ii2 >= 100, kdv[q], m9[q], put);
        }
      }
    }
    grid.sync();          // all columns final
    WT(4);   // phase C
    if (p.Z) {            // optional per-sample output (pickle / --asciidistrib path): this block's slice
      for (int k = 0; k < NKEYS; ++k) {
        const int sl = p.slot_of_key[k];
        const bool livek = ((pm >> k) & 1ull) && sl >= 0;
        float* dst = p.Z + ((long long)g * NKEYS + k) * N;
        for (int s = s_lo + tid; s < s_hi && s < N; s += PT) dst[s] = livek ? vals[(long long)sl * Npad + s] : Math<float>::nan();
      }
    }

    // ---- phase D: distributed selection -------------------------------------------------------
    if (b == 0 && tid < NKEYS) {
      const int k = tid, sl = p.slot_of_key[k];
      const bool pres = (pm >> k) & 1ull;
      if (!(pres && sl >= 0)) {
        const long long o = (p.row0 + g) * NKEYS + k;
        for (int q = 0; q < p.ndst; ++q) {
          p.status[q][o] = pres ? ST_EMPTY : ST_ABSENT;
          p.nvalid[q][o] = 0;
          p.pct[q][3 * o] = p.pct[q][3 * o + 1] = p.pct[q][3 * o + 2] = Math<float>::nan();
        }
      }
    }
    if (b == 0 && tid == 0) {
      p.trips[2 * g] = ii1;
      p.trips[2 * g + 1] = ii2;
      if (p.ret) p.ret[g] = ok ? 0 : -1;
    }
    // MSB radix select on the order-preserving bit patterns of the floats, 8 bits per round, for the
    // six ranks np.percentile interpolates between: exact whatever the distribution (ties end in
    // one bit pattern), no gather, no sort.  Per round every block histograms its slice (shared
    // memory, 16-bit counters), merges into the global table, and after the grid barrier every
    // block locates the six ranks in the merged table (redundantly, identically).
    unsigned tpre[2][6];        // per slot (this warp handles slots warp, warp + PW): prefix of each target
    unsigned trk[2][6];         // rank of each target inside its prefix
    double gam[2][3];
    int nfin[2] = {0, 0};
    bool live[2] = {false, false};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int sl = warp + h * PW;
      live[h] = sl < p.nstore && ((pm >> p.key_of_slot[sl < p.nstore ? sl : 0]) & 1ull);
#pragma unroll
      for (int t = 0; t < 6; ++t) { tpre[h][t] = 0u; trk[h][t] = 0u; }
    }
    unsigned short* rh = reinterpret_cast<unsigned short*>(hist);      // [6][256] 16-bit counters
    // The finite values of a column share their leading bits (sign, exponent, often the first
    // mantissa bits): a digit taken there puts every sample into one or two bins, and the shared-
    // memory atomics of a warp serialise on them.  So the extremes of the ordered bit patterns are
    // reduced first (one pass, one barrier), the common prefix is taken as known, and the 8-bit
    // rounds start at the first bit in which the extremes differ: usually three well-spread rounds
    // instead of four, the first two of which were all conflicts.
    int Tbits[2] = {0, 0};      // bits still to resolve per slot
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (!live[h]) continue;
      const int sl = warp + h * PW;
      const float* col = vals + (long long)sl * Npad;
      unsigned lmax = 0u, lnmin = 0u;
      // (four 128-bit loads per lane in flight: with one, the 17 warps of a block keep 8.7 KB in flight per
      // SM and the pass is bound by memory latency, not bandwidth)
      for (int g0 = g_lo; g0 < g_hi; g0 += 4) {
        float4 x4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          x4[q] = g0 + q < g_hi ? *reinterpret_cast<const float4*>(col + 4 * lane + 128 * (g0 + q))
                                : make_float4(Math<float>::nan(), Math<float>::nan(), Math<float>::nan(), Math<float>::nan());
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float xs[4] = {x4[q].x, x4[q].y, x4[q].z, x4[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const unsigned u = finite_f(xs[j]) ? ordered_bits(xs[j]) : 0u;
            lmax = max(lmax, u);
            lnmin = max(lnmin, u ? ~u : 0u);
          }
        }
      }
      lmax = __reduce_max_sync(0xffffffffu, lmax);
      lnmin = __reduce_max_sync(0xffffffffu, lnmin);
      if (lane == 0 && lmax)
        for (int r = 0; r < W; ++r) {
          atomicMax_system(reinterpret_cast<unsigned*>(peer(&ctl->umaxo[sl], r)), lmax);
          atomicMax_system(reinterpret_cast<unsigned*>(peer(&ctl->nmino[sl], r)), lnmin);
        }
    }
    xsync();
    WT(5);   // radix select: extremes pass
    int Rmax = 1;
    for (int sl = 0; sl < p.nstore; ++sl) {
      const unsigned umx = *reinterpret_cast<volatile unsigned*>(&ctl->umaxo[sl]);
      const unsigned umn = ~*reinterpret_cast<volatile unsigned*>(&ctl->nmino[sl]);
      const int T = umx ? 32 - __clz(umx ^ umn) : 0;
      Rmax = max(Rmax, (T + 7) / 8);
      if (sl == warp || sl == warp + PW) {
        const int h = sl == warp ? 0 : 1;
        if (!umx) live[h] = false;                       // no finite value in the column
        Tbits[h] = T;
        const unsigned pre = T >= 32 ? 0u : (umx >> T);
#pragma unroll
        for (int t = 0; t < 6; ++t) tpre[h][t] = pre;
      }
    }
#pragma unroll 1
    for (int round = 0; round < Rmax; ++round) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (!live[h] || (round > 0 && 8 * round >= Tbits[h])) continue;
        // this round resolves bits [lo, hi) of the patterns; everything above hi is in tpre
        const int hi = Tbits[h] - 8 * round, lo = max(hi - 8, 0);
        const int shift = lo;
        const unsigned dmask = (1u << (hi - lo)) - 1u;      // (a column of equal values: hi = lo = 0, one bin)
        const int sl = warp + h * PW;
        const float* col = vals + (long long)sl * Npad;
        // distinct prefixes among the six targets (ascending ranks -> non-decreasing prefixes)
        // (built with constant indices only: a run-time index would put uq[] into local memory, and the
        // match below would read it from there for every sample)
        unsigned tp[6], uq[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) { tp[t] = tpre[h][t]; uq[t] = 0xffffffffu; }
        int nu = 0;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const bool isnew = t == 0 || tp[t] != tp[t - 1];
#pragma unroll
          for (int k = 0; k < 6; ++k) if (isnew && nu == k) uq[k] = tp[t];
          nu += isnew ? 1 : 0;
        }
        float sum = 0.0f;
        unsigned* gh = ghist + (((long long)round * p.nstore + sl) * 6) * 256;
        // (16-bit counters: the slice is taken 480 groups = 61440 samples at a time)
#pragma unroll 1
        for (int gc = g_lo; gc < g_hi; gc += 480) {
        const int gce = gc + 480 < g_hi ? gc + 480 : g_hi;
        for (int i = lane; i < 6 * 128; i += 32) hist[i] = 0u;
        __syncwarp();
        // The prefix of a sample is matched against the DISTINCT prefixes of the six targets: one in the
        // first round, at most three as a rule (the two ranks of a percentile are neighbours) -- the
        // loop is specialised on that number (a fixed six-way match was a sixth of all the instructions
        // the kernel executes).
        auto scan = [&](auto nuq_tag) {
          constexpr int NUQ = decltype(nuq_tag)::value;
          for (int g0 = gc; g0 < gce; g0 += 4) {
            float4 x44[4];                       // four loads in flight (see the extremes pass)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              x44[q] = g0 + q < gce ? *reinterpret_cast<const float4*>(col + 4 * lane + 128 * (g0 + q))
                                    : make_float4(Math<float>::nan(), Math<float>::nan(), Math<float>::nan(), Math<float>::nan());
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float xs[4] = {x44[q].x, x44[q].y, x44[q].z, x44[q].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = xs[j];
                const bool fin = finite_f(x);
                if (round == 0) sum += fin ? x : 0.0f;
                const unsigned u = ordered_bits(x);
                const unsigned pre = hi >= 32 ? 0u : (u >> hi);
                int which = -1;
#pragma unroll
                for (int i = 0; i < NUQ; ++i) if ((NUQ == 1 || i < nu) && uq[i] == pre) which = i;
                if (fin && which >= 0) {
                  const unsigned bin = (u >> shift) & dmask;
                  const unsigned idx = (unsigned)which * 256u + bin;
                  atomicAdd(hist + (idx >> 1), 1u << ((idx & 1u) << 4));
                }
              }
            }
          }
        };
        if (nu == 1) scan(std::integral_constant<int, 1>());
        else if (nu <= 3) scan(std::integral_constant<int, 3>());
        else scan(std::integral_constant<int, 6>());
        __syncwarp();
        for (int i = lane; i < nu * 256; i += 32) {
          const unsigned c = rh[i];
          if (c) {
            if (W > 1) atomicAdd(lhist + (long long)sl * 6 * 256 + i, c);
            else atomicAdd(gh + i, c);
          }
        }
        __syncwarp();
        }
        if (round == 0) {
          const double tot = warp_sum((double)sum);
          if (lane == 0)
            for (int r = 0; r < W; ++r) atomicAdd_system(reinterpret_cast<double*>(peer(&ctl->colsum[sl], r)), tot);
        }
        __syncwarp();
      }
      if (W > 1) {
        grid.sync();
        unsigned* ghr = ghist + (long long)round * p.nstore * 6 * 256;
        for (long long i = (long long)b * PT + tid; i < (long long)p.nstore * 6 * 256; i += (long long)nb * PT) {
          const unsigned c = lhist[i];
          if (c) {
            lhist[i] = 0u;
            for (int r = 0; r < W; ++r) atomicAdd_system(reinterpret_cast<unsigned*>(peer(ghr + i, r)), c);
          }
        }
      }
      xsync();
      WT(6);   // radix rounds
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (!live[h] || (round > 0 && 8 * round >= Tbits[h])) continue;
        const int hi = Tbits[h] - 8 * round, lo = max(hi - 8, 0);
        const int sl = warp + h * PW;
        const unsigned* gh = ghist + (((long long)round * p.nstore + sl) * 6) * 256;
        unsigned uq[6];
        int nu = 0;
#pragma unroll
        for (int t = 0; t < 6; ++t)
          if (nu == 0 || tpre[h][t] != uq[nu - 1 < 0 ? 0 : nu - 1]) { uq[nu] = tpre[h][t]; ++nu; }
        if (round == 0) {
          unsigned c = 0;
          for (int i = lane; i < 256; i += 32) c += __ldcg(gh + i);
          nfin[h] = (int)__reduce_add_sync(0xffffffffu, c);
          if (nfin[h] == 0) { live[h] = false; continue; }
          const double qs[3] = {16.0 / 100.0, 50.0 / 100.0, 84.0 / 100.0};              // mcz.py:288
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double vi = (double)(nfin[h] - 1) * qs[q];
            const double lo = floor(vi);
            gam[h][q] = vi - lo;
            trk[h][2 * q] = (unsigned)lo;
            trk[h][2 * q + 1] = min((unsigned)lo + 1u, (unsigned)(nfin[h] - 1));
          }
        }
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          int which = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) if (i < nu && uq[i] == tpre[h][t]) which = i;
          // lane holds bins [8 lane, 8 lane + 8) of the target's table
          unsigned c8[8], tot = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) { c8[i] = __ldcg(gh + which * 256 + 8 * lane + i); tot += c8[i]; }
          unsigned inc = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
          }
          const unsigned ex = inc - tot, r = trk[h][t];
          const unsigned own = __ballot_sync(0xffffffffu, tot > 0 && r >= ex && r < ex + tot);
          const int ol = own ? __ffs(own) - 1 : 0;
          // inside the owning lane: which of its 8 bins
          unsigned run = ex, bin = 0, before = ex;
          bool found = false;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!found && r < run + c8[i]) { bin = (unsigned)i; before = run; found = true; }
            run += c8[i];
          }
          const unsigned gbin = __shfl_sync(0xffffffffu, 8u * (unsigned)lane + bin, ol);
          const unsigned gbefore = __shfl_sync(0xffffffffu, before, ol);
          tpre[h][t] = (hi - lo >= 32 ? 0u : (tpre[h][t] << (hi - lo))) | gbin;
          trk[h][t] = r - gbefore;
        }
      }
    }
    // ---- results: block 0, lane 0 of each warp --------------------------------------------------
    for (int h = 0; h < 2; ++h) {
      const int sl = warp + h * PW;
      if (sl >= p.nstore || b != 0) continue;
      const int k = p.key_of_slot[sl];
      if (!((pm >> k) & 1ull)) continue;
      float pc[3] = {Math<float>::nan(), Math<float>::nan(), Math<float>::nan()};
      int stt = ST_EMPTY;
      if (nfin[h] > 0) {
        const double tot = *reinterpret_cast<volatile double*>(&ctl->colsum[sl]);
        if (!(tot > 0.0)) {
          stt = ST_NONPOS;
        } else {
          float val[6];
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const unsigned u = tpre[h][t];
            val[t] = __uint_as_float(u >= 0x80000000u ? u - 0x80000000u : ((0x80000000u - u) | 0x80000000u));
          }
          pc[0] = np_lerp_f(val[2], val[3], gam[h][1]);       // median
          pc[1] = np_lerp_f(val[0], val[1], gam[h][0]);       // 16th
          pc[2] = np_lerp_f(val[4], val[5], gam[h][2]);       // 84th
          stt = ST_OK;
        }
      }
      if (lane == 0) {
        const long long o = (p.row0 + g) * NKEYS + k;
        for (int q = 0; q < p.ndst; ++q) {
          p.status[q][o] = (signed char)stt;
          p.nvalid[q][o] = nfin[h];
          p.pct[q][3 * o] = pc[0]; p.pct[q][3 * o + 1] = pc[1]; p.pct[q][3 * o + 2] = pc[2];
        }
      }
    }
    grid.sync();          // the scratch is reused by the next galaxy
    WT(7);   // results + end barrier
  }
  if (p.ndst > 1) __threadfence_system();
}


static size_t wide_shared_bytes_n(int nstore) {
  return 4096 + (size_t)4 * nstore * 6 * 256 * sizeof(unsigned) + 256;
}
static size_t wide_scratch_bytes(long long N, int nstore) {
  const long long Npad = (N + 127) / 128 * 128;
  return SCRATCH_HEADER + wide_shared_bytes_n(nstore) + (size_t)((long long)nstore + 4 + 6) * Npad * sizeof(float);
}
size_t wide_private_bytes(long long N, uint32_t tok) {
  ProdParams p;
  build_slots(tok, p);
  return wide_scratch_bytes(N, p.nstore);
}
size_t wide_shared_bytes(uint32_t tok) {
  ProdParams p;
  build_slots(tok, p);
  return wide_shared_bytes_n(p.nstore);
}

// Few galaxies with very many samples each (single-object runs): the cooperative wide kernel.
static bool use_wide(long long G, long long N, const float* Z) {
  const char* force = getenv("MCZ_WIDE");              // "1": always (tests), "0": never
  (void)Z;
  if (force && force[0] == '0') return false;
  if (force && force[0] == '1') return true;
  // Measured on B200 (tools/time_wide.py): the wide kernel takes ~0.12 ms + 0.27 ns per sample for
  // each galaxy, one after the other; a per-galaxy block ~29 ns per sample (columns of this length
  // stream through L2 / DRAM), up to one block per SM at a time.
  if (N < 32768 || G > num_sms()) return false;
  return (double)G * (0.12 + 2.7e-7 * (double)N) < 2.9e-5 * (double)N;
}

static int launch_wide(ProdParams& p, void* scratch, cudaStream_t stream, const WideShare* ws) {
  p.Npad = (int)((p.N + 127ll) / 128 * 128);
  finish_slots(p);
  const size_t shb = wide_shared_bytes_n(p.nstore);
  if (ws && ws->world > 1) {
    if (ws->world > MAX_DST || ws->rank < 0 || ws->rank >= ws->world) return set_error("wide kernel: bad rank / world");
    if (p.Z) return set_error("wide kernel: per-sample output is not available when one object is sharded over GPUs");
    p.wworld = ws->world; p.wrank = ws->rank;
    for (int q = 0; q < ws->world; ++q) {
      if (!ws->shared[q]) return set_error("wide kernel: null shared region");
      p.wshared[q] = ws->shared[q];
    }
    p.gscratch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + SCRATCH_HEADER);
    // the cross-rank barrier counter starts at 0 (the caller makes sure every rank has done this
    // before any rank launches: one host-side barrier)
  } else {
    p.wworld = 1; p.wrank = 0;
    unsigned char* body = reinterpret_cast<unsigned char*>(scratch) + SCRATCH_HEADER;
    p.wshared[0] = body;
    p.gscratch = reinterpret_cast<float*>(body + shb);
    const int rc = check_cuda(cudaMemsetAsync(body + shb - 256, 0, 256, stream), "barrier memset");
    if (rc) return rc;
  }
  const size_t smem = (size_t)PW * FineHist<false>::BYTES;
  DeviceState& dstate = device_state();
  int max_blocks = dstate.wide_max_blocks.load();
  if (!max_blocks) {
    int rc = ensure_smem_attr(wide_kernel, KID_WIDE, (int)smem);
    if (rc) return rc;
    int per_sm = 0;
    rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_kernel, PT, smem), "occupancy");
    if (rc) return rc;
    if (per_sm < 1) return set_error("wide_kernel does not fit on an SM");
    max_blocks = per_sm * num_sms();
    dstate.wide_max_blocks.store(max_blocks);
  }
  int grid = max_blocks < num_sms() ? max_blocks : num_sms();
  const int ngroups = p.Npad / 128;
  const int mine = (int)((long long)ngroups * (p.wrank + 1) / p.wworld - (long long)ngroups * p.wrank / p.wworld);
  if (grid > mine) grid = mine > 0 ? mine : 1;
  void* args[] = {&p};
  const int rc = check_cuda(cudaLaunchCooperativeKernel((void*)wide_kernel, dim3(grid), dim3(PT), args, smem, stream),
                            "wide_kernel launch");
  if (rc) return rc;
  count_launch();
  return 0;
}

// ---- selection on caller-supplied columns (test entry point) -----------------------------------
// One warp per column, exactly the code phase D runs: cols[ncols][N] f32 in global memory are
// staged into a shared-memory column of NG*128 floats (NaN padding) or, for longer columns, into a
// padded global scratch, then warp_percentiles() is called.  Lets the tests feed the selection
// with distributions the physics never produces (ties, outliers, two-point, sorted, denormals).
template <int NG, bool WIDE>
__global__ void __launch_bounds__(128, 1)
select_columns_kernel(const float* __restrict__ cols, int ncols, int N, int Npad, float* __restrict__ padded,
                      float* __restrict__ pct, int* __restrict__ nvalid, signed char* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr bool SH = NG > 0;
  typedef FineHist<!WIDE> FH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 4 + warp;
  if (c >= ncols) return;
  unsigned char* mine = smem + (size_t)warp * ((SH ? NG * 128 * 4 : 0) + FH::BYTES + 512);
  float* col = SH ? reinterpret_cast<float*>(mine) : padded + (long long)c * Npad;
  unsigned* hist = reinterpret_cast<unsigned*>(mine + (SH ? NG * 128 * 4 : 0));
  float* cand = reinterpret_cast<float*>(mine + (SH ? NG * 128 * 4 : 0) + FH::BYTES);
  for (int s = lane; s < Npad; s += 32) col[s] = s < N ? cols[(long long)c * N + s] : Math<float>::nan();
  __syncwarp();
  float pc[3];
  int nv, st;
  warp_percentiles<!WIDE, NG>(col, Npad, hist, cand, pc, nv, st);
  if (lane == 0) {
    pct[3 * c] = pc[0]; pct[3 * c + 1] = pc[1]; pct[3 * c + 2] = pc[2];
    nvalid[c] = nv;
    status[c] = (signed char)st;
  }
}

int launch_select_columns(const float* cols_d, long long ncols, long long N, float* pct_d, int* nvalid_d,
                          signed char* status_d, float* padded_d, cudaStream_t stream) {
  if (ncols <= 0 || N <= 0) return 0;
  const int grid = (int)((ncols + 3) / 4);
  if (N <= 18 * 128) {
    const size_t sm = 4 * (size_t)(18 * 128 * 4 + FineHist<true>::BYTES + 512);
    const int rc = ensure_smem_attr(select_columns_kernel<18, false>, KID_SELECT, (int)sm);
    if (rc) return rc;
    select_columns_kernel<18, false><<<grid, 128, sm, stream>>>(cols_d, (int)ncols, (int)N, 18 * 128, nullptr, pct_d, nvalid_d, status_d);
  } else {
    if (!padded_d) return set_error("mcz_debug_select_columns: columns longer than 2304 need the padded scratch");
    const int Npad = (int)((N + 127) / 128 * 128);
    if (Npad < 65536) {           // the instantiations the production kernel uses for these lengths
      const size_t sm = 4 * (size_t)(FineHist<true>::BYTES + 512);
      select_columns_kernel<0, false><<<grid, 128, sm, stream>>>(cols_d, (int)ncols, (int)N, Npad, padded_d, pct_d, nvalid_d, status_d);
    } else {
      const size_t sm = 4 * (size_t)(FineHist<false>::BYTES + 512);
      select_columns_kernel<0, true><<<grid, 128, sm, stream>>>(cols_d, (int)ncols, (int)N, Npad, padded_d, pct_d, nvalid_d, status_d);
    }
  }
  count_launch();
  return check_cuda(cudaGetLastError(), "select_columns_kernel launch");
}

// ---- the sampler on its own (test entry point) --------------------------------------------------
// Writes the perturbed, sanitised fluxes and the scaled coefficient noise exactly as phase A of the
// production kernels derives them (same draw_normals / fmaf / clamps, hence the same MUFU results
// bit for bit): flux[G][NUSED][N] in `Used` order, noise[G][5][N].  The parity tests feed these to
// the oracle, so the comparison is about the arithmetic of the path and not about a host model of
// the device's Box-Muller.
__global__ void __launch_bounds__(256)
sample_export_kernel(const __grid_constant__ ProdParams p, float* __restrict__ flux, float* __restrict__ noise) {
  const int used2line[NUSED] = {L_HA, L_HB, L_O2, L_O3, L_N2, L_S2A, L_S2B, L_S39069};
  const long long g = blockIdx.y;
  const unsigned long long gg = p.goff + (unsigned long long)g;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < p.N; s += gridDim.x * blockDim.x) {
    float z[16];
    draw_normals((uint32_t)s, gg, p, z);
#pragma unroll
    for (int u = 0; u < NUSED; ++u) {
      const float lm = p.meas[g * NLINES + used2line[u]];
      const float le = p.deterministic ? 0.0f : p.err[g * NLINES + used2line[u]];
      const float raw = fmaf(le, p.correlated ? z[5] : z[5 + u], lm);
      float v = fmaxf(raw, 0.0f);
      if (!(v <= FLT_MAX)) v = 0.0f;
      if ((p.tok & T_DROPNEG) && raw < 0.0f && raw >= -FLT_MAX) v = Math<float>::nan();
      flux[(g * NUSED + u) * p.N + s] = v;
    }
    const float sg[5] = {0.05f, 0.1f, 0.027f, 0.024f, 0.012f};
#pragma unroll
    for (int j = 0; j < 5; ++j) noise[(g * 5 + j) * p.N + s] = sg[j] * z[j];
  }
}

int launch_sample_export(const float* meas, const float* err, long long G, long long n_draw, uint32_t tok,
                         unsigned long long seed, unsigned long long galaxy_offset, int sample_mode,
                         int deterministic, float* flux_d, float* noise_d, cudaStream_t stream) {
  if (G <= 0 || n_draw <= 0) return 0;
  if (G > 65535) return set_error("mcz_debug_production_samples: at most 65535 galaxies per call");
  ProdParams p;
  memset(&p, 0, sizeof(p));
  p.meas = meas; p.err = err; p.G = G; p.N = (int)n_draw; p.tok = tok;
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.goff = galaxy_offset;
  p.correlated = sample_mode == 1; p.deterministic = deterministic != 0;
  p.calls = philox_calls(tok, p.correlated, p.deterministic);
  const int bx = (int)((n_draw + 255) / 256 < 64 ? (n_draw + 255) / 256 : 64);
  sample_export_kernel<<<dim3(bx, (unsigned)G), 256, 0, stream>>>(p, flux_d, noise_d);
  count_launch();
  return check_cuda(cudaGetLastError(), "sample_export_kernel launch");
}

}  // namespace mcz
