// Device assists for the host steps that FOLLOW the Monte-Carlo path: the reference
// spends most of its wall time after the path proper binning each (measurement, scale)
// distribution for its histograms -- Knuth's rule evaluates np.histogram(data, linspace(min, max,
// m + 1)) for dozens of bin numbers m per distribution (mcz.py:90-116, 227-239) -- and its
// sample-size diagnostic compares nested sub-samples with a two-sample Kolmogorov-Smirnov test
// (testcompleteness.py:87-113).  The optimiser, the gammaln posterior, the p-value and the plots
// stay on the host; what moves here is the data-parallel part:
//   histogram_assist_kernel  per distribution: n, min, max, mean, std of the finite values
//                            (mcz.py:273, 289) and the equal-width counts for a caller-given list of
//                            bin numbers, bit-identical to np.histogram on np.linspace edges;
//   ks_counts_kernel         the two empirical CDFs evaluated at every data point -> min / max of
//                            their difference, exactly as scipy.stats.ks_2samp forms its statistic.
#include "mcz_common.cuh"
#include "mcz_device_utils.cuh"
#include "mcz_internal.h"

namespace mcz {

static constexpr int HB = 256;            // threads per block
static constexpr int HA_MAX_BINS = 8192;  // largest bin number of one request (32 KB of counters)

__device__ __forceinline__ bool fin64(double v) { return fabs(v) <= 1.7976931348623157e308; }

// np.linspace(start, stop, m + 1)[k]: arange(k) * step + start, last element = stop exactly
// (numpy/_core/function_base.py); separate multiply and add, as numpy evaluates it.
__device__ __forceinline__ double linspace_edge(int k, int m, double start, double stop, double step) {
  return k >= m ? stop : __dadd_rn(__dmul_rn((double)k, step), start);
}

__global__ void __launch_bounds__(HB)
histogram_assist_kernel(const double* __restrict__ Z, long long rows, long long N, const int* __restrict__ mlist,
                        int nm, long long counts_per_row, double* __restrict__ stats, int* __restrict__ counts) {
  extern __shared__ int hist[];
  __shared__ double red[2 * 2 * (HB / 32)];
  __shared__ double sh_mn, sh_mx;
  const int tid = threadIdx.x, nw = HB / 32;
  int phase = 0;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* row = Z + r * N;
    // ---- n, min, max, sum of the finite values (savehist: data[np.isfinite(data)], mcz.py:273)
    double cnt = 0.0, sum = 0.0;
    double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
    for (long long s = tid; s < N; s += HB) {
      const double v = row[s];
      if (fin64(v)) { cnt += 1.0; sum += v; mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    block_sum2(cnt, sum, red, nw, phase);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = mn; red[nw + (tid >> 5)] = mx; }
    __syncthreads();
    if (tid == 0) {
      double a = red[0], b = red[nw];
      for (int w = 1; w < nw; ++w) { a = fmin(a, red[w]); b = fmax(b, red[nw + w]); }
      sh_mn = a; sh_mx = b;
    }
    __syncthreads();
    mn = sh_mn; mx = sh_mx;
    const double n = cnt, mean = n > 0.0 ? sum / n : Math<double>::nan();
    // ---- np.std: sqrt(mean(|x - mean|^2)) (mcz.py:289)
    double ssd = 0.0, unused = 0.0;
    for (long long s = tid; s < N; s += HB) {
      const double v = row[s];
      if (fin64(v)) { const double d = v - mean; ssd += d * d; }
    }
    block_sum2(ssd, unused, red, nw, phase);
    if (tid == 0) {
      double* st = stats + r * 6;
      st[0] = n; st[1] = n > 0.0 ? mn : Math<double>::nan(); st[2] = n > 0.0 ? mx : Math<double>::nan();
      st[3] = mean; st[4] = n > 0.0 ? sqrt(ssd / n) : Math<double>::nan(); st[5] = sum;
    }
    // ---- equal-width counts for every requested bin number (getknuth, mcz.py:90-99)
    int* out = counts + r * counts_per_row;
    for (int q = 0; q < nm; ++q) {
      const int m = mlist[q];
      for (int k = tid; k < m; k += HB) hist[k] = 0;
      __syncthreads();
      if (n > 0.0) {
        const double step = (mx - mn) / (double)m;          // linspace: delta / div
        for (long long s = tid; s < N; s += HB) {
          const double v = row[s];
          if (!fin64(v)) continue;
          int k;
          if (!(step > 0.0)) {
            k = m - 1;      // all edges equal: np.histogram puts everything into the last (closed) bin
          } else {
            k = (int)floor((v - mn) / step);
            k = k < 0 ? 0 : (k > m - 1 ? m - 1 : k);
            // exact searchsorted semantics on the linspace edges: bin k = [e_k, e_k+1), last bin closed
            while (k > 0 && v < linspace_edge(k, m, mn, mx, step)) --k;
            while (k < m - 1 && v >= linspace_edge(k + 1, m, mn, mx, step)) ++k;
          }
          atomicAdd(&hist[k], 1);
        }
      }
      __syncthreads();
      for (int k = tid; k < m; k += HB) out[k] = hist[k];
      out += m;
      __syncthreads();
    }
  }
}

int launch_histogram_assist(const double* Z_d, long long rows, long long N, const int* mlist_h, int nm,
                            double* stats_d, int* counts_d, int* mlist_scratch_d, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (N <= 0 || nm < 0) return set_error("mcz_histogram_assist: bad N / nm");
  long long per_row = 0;
  int maxm = 1;
  for (int q = 0; q < nm; ++q) {
    if (mlist_h[q] < 1 || mlist_h[q] > HA_MAX_BINS) return set_error("mcz_histogram_assist: bin numbers must be in [1, 8192]");
    per_row += mlist_h[q];
    if (mlist_h[q] > maxm) maxm = mlist_h[q];
  }
  if (nm > 0) {
    if (!mlist_scratch_d || !counts_d) return set_error("mcz_histogram_assist: null pointer");
    int rc = check_cuda(cudaMemcpyAsync(mlist_scratch_d, mlist_h, (size_t)nm * sizeof(int), cudaMemcpyHostToDevice, stream),
                        "bin list upload");
    if (rc) return rc;
  }
  const int grid = (int)(rows < 148 * 8 ? rows : 148 * 8);
  histogram_assist_kernel<<<grid, HB, (size_t)maxm * sizeof(int), stream>>>(Z_d, rows, N, mlist_scratch_d, nm, per_row,
                                                                           stats_d, counts_d);
  count_launch();
  return check_cuda(cudaGetLastError(), "histogram_assist_kernel launch");
}

// ---- two-sample Kolmogorov-Smirnov statistic --------------------------------------------------
// scipy.stats.ks_2samp: data_all = concat(sort(a), sort(b)); cdf1 = ECDF of a at every pooled point
// (#{a <= v} steps of 1 / n1), cdf2 likewise; cddiffs = cdf1 - cdf2; minS = clip(-min(cddiffs), 0, 1), maxS = max(cddiffs),
// D = max(minS, maxS).  The counts #{a <= v}, #{b <= v} need no sort: every thread takes one data
// point v and scans both samples (all threads of a warp read the same element: one broadcast load
// per 32 comparisons).  out[0] = min(cddiffs), out[1] = max(cddiffs) as ordered 64-bit keys.
__device__ __forceinline__ unsigned long long ord64(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(HB)
ks_counts_kernel(const double* __restrict__ a, long long n1, const double* __restrict__ b, long long n2,
                 unsigned long long* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * HB + threadIdx.x;
  double dmin = __longlong_as_double(0x7ff0000000000000LL), dmax = -dmin;
  if (i < n1 + n2) {
    const double v = i < n1 ? a[i] : b[i - n1];
    long long c1 = 0, c2 = 0;
    for (long long j = 0; j < n1; ++j) c1 += (a[j] <= v) ? 1 : 0;
    for (long long j = 0; j < n2; ++j) c2 += (b[j] <= v) ? 1 : 0;
    // ECDF values as scipy >= 1.12 forms them: np.linspace(0, 1, n + 1)[count] = count * (1 / n), last = 1
    // (older scipy divided count / n: at most one ulp away)
    const double f1 = c1 >= n1 ? 1.0 : __dmul_rn((double)c1, 1.0 / (double)n1);
    const double f2 = c2 >= n2 ? 1.0 : __dmul_rn((double)c2, 1.0 / (double)n2);
    const double d = f1 - f2;
    dmin = d; dmax = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&keys[0], ord64(dmin));
    atomicMax(&keys[1], ord64(dmax));
  }
}

__global__ void ks_finish_kernel(const unsigned long long* __restrict__ keys, double* __restrict__ out) {
  const double mn = unord64(keys[0]), mx = unord64(keys[1]);
  double minS = -mn;
  minS = minS < 0.0 ? 0.0 : (minS > 1.0 ? 1.0 : minS);          // np.clip(-cddiffs[argminS], 0, 1)
  out[0] = minS > mx ? minS : mx;                                // two-sided D
  out[1] = mx;                                                   // D+  (maxS)
  out[2] = minS;                                                 // D-  (minS)
}

int launch_ks_2samp(const double* a_d, long long n1, const double* b_d, long long n2, double* out_d,
                    unsigned long long* scratch_d, cudaStream_t stream) {
  if (n1 <= 0 || n2 <= 0) return set_error("mcz_ks_2samp: both samples must be non-empty");
  if (!a_d || !b_d || !out_d || !scratch_d) return set_error("mcz_ks_2samp: null pointer");
  const unsigned long long init[2] = {~0ull, 0ull};
  int rc = check_cuda(cudaMemcpyAsync(scratch_d, init, sizeof(init), cudaMemcpyHostToDevice, stream), "ks init");
  if (rc) return rc;
  const long long n = n1 + n2;
  ks_counts_kernel<<<(unsigned)((n + HB - 1) / HB), HB, 0, stream>>>(a_d, n1, b_d, n2, scratch_d);
  ks_finish_kernel<<<1, 1, 0, stream>>>(scratch_d, out_d);
  count_launch(2);
  return check_cuda(cudaGetLastError(), "ks kernels launch");
}

}  // namespace mcz
