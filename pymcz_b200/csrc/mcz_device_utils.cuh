// Small block/warp collectives shared by the kernels (device only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcz {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// Deterministic block-wide sum; every thread returns the same value (NaN propagates).
// `buf` must hold >= 2*nwarps doubles; `phase` alternates 0/1 between successive calls so one
// __syncthreads per call is enough.
__device__ __forceinline__ double block_sum(double v, double* buf, int nwarps, int& phase) {
  v = warp_sum(v);
  double* b = buf + phase * nwarps;
  if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nwarps; ++w) t += b[w];
  phase ^= 1;
  return t;
}

// two sums at once (one barrier)
__device__ __forceinline__ void block_sum2(double& a, double& c, double* buf, int nwarps, int& phase) {
  a = warp_sum(a);
  c = warp_sum(c);
  double* b = buf + phase * 2 * nwarps;
  if ((threadIdx.x & 31) == 0) {
    b[2 * (threadIdx.x >> 5)] = a;
    b[2 * (threadIdx.x >> 5) + 1] = c;
  }
  __syncthreads();
  double ta = 0.0, tc = 0.0;
  for (int w = 0; w < nwarps; ++w) { ta += b[2 * w]; tc += b[2 * w + 1]; }
  phase ^= 1;
  a = ta;
  c = tc;
}

}  // namespace mcz
