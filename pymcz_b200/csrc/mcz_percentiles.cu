// float64 percentile extraction of stored distributions: the statistics part of savehist()
// (reference mcz.py:273-301): finite filter, n, sum(data) <= 0 test, and
// np.percentile(data, [50, 16, 84]) with numpy's linear interpolation, bit-compatible lerp.
// One thread block per (galaxy, key) row; exact order statistics by an 8-bit MSB radix select
// over order-preserving 64-bit keys.  This is the replay/pickle path, not the throughput path
// (the fused production kernel has its own float32 selection).
#include "mcz_common.cuh"
#include "mcz_device_utils.cuh"
#include "mcz_internal.h"

namespace mcz {

static constexpr int PB = 256;

__device__ __forceinline__ unsigned long long f64_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unkey(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) finite value of row[0..N)
__device__ double radix_select(const double* __restrict__ row, long long N, long long k,
                               unsigned int* hist, unsigned long long* sh_state) {
  const int tid = threadIdx.x;
  unsigned long long prefix = 0, mask = 0;
  for (int pass = 7; pass >= 0; --pass) {
    hist[tid] = 0;
    __syncthreads();
    const int shift = 8 * pass;
    for (long long s = tid; s < N; s += PB) {
      const double v = row[s];
      if (fabs(v) <= 1.7976931348623157e308) {
        const unsigned long long key = f64_key(v);
        if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    // inclusive scan of the 256 bins (one per thread)
    unsigned int c = hist[tid], incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    __shared__ unsigned int wsum[PB / 32];
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    unsigned int base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    incl += base;
    const unsigned int excl = incl - c;
    if (c > 0 && (unsigned long long)k >= excl && (unsigned long long)k < incl) {
      sh_state[0] = prefix | ((unsigned long long)tid << shift);
      sh_state[1] = (unsigned long long)k - excl;
    }
    __syncthreads();
    prefix = sh_state[0];
    k = (long long)sh_state[1];
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return f64_unkey(prefix);
}

// numpy.lib._function_base_impl._lerp
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
  const double d = b - a;
  return (t >= 0.5) ? (b - d * (1.0 - t)) : (a + d * t);
}

__global__ void __launch_bounds__(PB)
percentiles_f64_kernel(const double* __restrict__ Z, const unsigned char* __restrict__ present,
                       long long rows, long long N, double* __restrict__ pct,
                       int* __restrict__ nvalid, signed char* __restrict__ status) {
  __shared__ unsigned int hist[PB];
  __shared__ unsigned long long sh_state[2];
  __shared__ double sh_red[4 * (PB / 32)];
  const int tid = threadIdx.x;
  int phase = 0;
  const double nan = Math<double>::nan();
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    if (present && !present[r]) {
      if (tid == 0) { status[r] = ST_ABSENT; nvalid[r] = 0; pct[3 * r] = pct[3 * r + 1] = pct[3 * r + 2] = nan; }
      continue;
    }
    const double* row = Z + r * N;
    double cnt = 0.0, sum = 0.0;
    for (long long s = tid; s < N; s += PB) {
      const double v = row[s];
      if (fabs(v) <= 1.7976931348623157e308) { cnt += 1.0; sum += v; }      // mcz.py:273
    }
    block_sum2(cnt, sum, sh_red, PB / 32, phase);
    const long long n = (long long)cnt;
    if (n == 0 || !(sum > 0.0)) {                                           // mcz.py:277-285
      if (tid == 0) {
        status[r] = n == 0 ? ST_EMPTY : ST_NONPOS;
        nvalid[r] = (int)n;
        pct[3 * r] = pct[3 * r + 1] = pct[3 * r + 2] = nan;
      }
      continue;
    }
    const double qs[3] = {50.0 / 100.0, 16.0 / 100.0, 84.0 / 100.0};        // mcz.py:288
    double out[3];
    for (int qi = 0; qi < 3; ++qi) {
      const double vi = (double)(n - 1) * qs[qi];
      const double lo = floor(vi);
      const double gamma = vi - lo;
      const long long klo = (long long)lo;
      const long long khi = klo + 1 < n ? klo + 1 : n - 1;
      const double a = radix_select(row, N, klo, hist, sh_state);
      const double b = (khi == klo) ? a : radix_select(row, N, khi, hist, sh_state);
      out[qi] = np_lerp(a, b, gamma);
    }
    if (tid == 0) {
      status[r] = ST_OK;
      nvalid[r] = (int)n;
      pct[3 * r] = out[0];
      pct[3 * r + 1] = out[1];
      pct[3 * r + 2] = out[2];
    }
  }
}

int launch_percentiles_f64(const double* Z, const unsigned char* present, long long rows, long long N,
                           double* pct, int* nvalid, signed char* status, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (N <= 0) return set_error("mcz_segmented_percentiles: N must be positive");
  const int grid = (int)(rows < 4736 ? rows : 4736);
  percentiles_f64_kernel<<<grid, PB, 0, stream>>>(Z, present, rows, N, pct, nvalid, status);
  count_launch();
  return check_cuda(cudaGetLastError(), "percentiles_f64_kernel launch");
}

}  // namespace mcz
