// Microbenchmark 2: cost of one histogram pass over columns resident in shared memory on sm_100a,
// laid out like phase D of the production kernel (18 warps, one column of 2304 floats per warp,
// 128-bit loads, 1 CTA/SM).
//  mode 0: no histogram (load + index arithmetic only)
//  mode 1: ATOMS on packed u16 counters, 1024 bins (what pass H does)
//  mode 2: ATOMS, every lane the same word
//  mode 3: ATOMS without bank conflicts (word = (bin%16)*32 + lane)
//  mode 4: per-lane private u8 counters, one bank per lane, 64 bins, LDS/IADD/STS on words
//  mode 5: per-lane private u8 counters, 128 bins (4 KB per warp)
//  mode 6: ATOMS 1024 bins, only 1 in 8 samples counted (others predicated off)
//  mode 7: __match_any_sync dedup + ATOMS by leaders
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int mode> __global__ void __launch_bounds__(576, 1) k(int iters, int nwarps, const float* __restrict__ vals, long long* out, unsigned* sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* cols = reinterpret_cast<float*>(smem);                      // 18 x 2304 floats
  unsigned* hw = reinterpret_cast<unsigned*>(smem + 17 * 2304 * 4) + warp * 1024;   // 4 KB per warp
  for (int i = threadIdx.x; i < 17 * 2304; i += 576) cols[i] = vals[blockIdx.x * 18 * 2304 + i];
  for (int i = lane; i < 1024; i += 32) hw[i] = 0;
  __syncthreads();
  const float* v = cols + (warp % 17) * 2304;
  long long t0 = clock64();
  unsigned acc = 0;
  if (warp < nwarps)
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int s0 = 4 * lane; s0 < 2304; s0 += 128) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + s0);
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = xs[j];
        const int b = __float2int_rz(fminf(fmaxf(fmaf(x, 128.0f, 512.0f), 0.0f), 1023.0f));
        if (mode == 0) acc += b;
        else if (mode == 1) atomicAdd(hw + (b >> 1), 1u << ((b & 1) << 4));
        else if (mode == 2) atomicAdd(hw + 7, 1u);
        else if (mode == 3) atomicAdd(hw + (b & 15) * 32 + lane, 1u);
        else if (mode == 4) { unsigned* h = hw + (b >> 6) * 32 + lane; *h += 1u << (((b >> 4) & 3) * 8); }
        else if (mode == 5) { unsigned* h = hw + (b >> 5) * 32 + lane; *h += 1u << (((b >> 3) & 3) * 8); }
        else if (mode == 6) { if ((b & 7) == 0) atomicAdd(hw + (b >> 1), 1u << ((b & 1) << 4)); }
        else {
          unsigned peers = __match_any_sync(0xffffffffu, b);
          if ((__ffs(peers) - 1) == lane) atomicAdd(hw + (b >> 1), (unsigned)__popc(peers) << ((b & 1) << 4));
        }
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  for (int i = lane; i < 1024; i += 32) acc += hw[i];
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
}
int main() {
  const int nb = 148; float* d; long long* o; unsigned* sink;
  size_t n = (size_t)nb * 18 * 2304;
  float* h = (float*)malloc(n * 4);
  unsigned s = 12345;
  for (size_t i = 0; i < n; ++i) { float a = 0; for (int j = 0; j < 12; ++j) { s = s * 1664525u + 1013904223u; a += (s >> 8) * (1.0f / 16777216.0f); } h[i] = (a - 6.0f) * 0.8f; }
  cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&o, nb * 16); cudaMalloc(&sink, 4);
  const int smem = 17 * 2304 * 4 + 18 * 4096;
  for (int nw : {17, 8, 4, 1})
  for (int mode = 0; mode < 8; ++mode) {
    const int iters = 20;
#define RUN(M) case M: cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<M><<<nb, 576, smem>>>(2, nw, d, o, sink); k<M><<<nb, 576, smem>>>(iters, nw, d, o, sink); break;
    switch (mode) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) }
    cudaDeviceSynchronize();
    long long ho[296]; cudaMemcpy(ho, o, nb * 16, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; ++i) avg += ho[2 * i + 1]; avg /= nb;
    printf("warps %2d mode %d: %8.1f cycles per pass (block wall); %.2f cycles per warp-instr of 32 samples per SM; %.3f samples/clk/SM  err=%s\n",
           nw, mode, avg / iters, avg / iters / (nw * 72.0), nw * 2304.0 / (avg / iters), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
