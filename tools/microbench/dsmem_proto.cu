// Checks the DSMEM / mbarrier semantics the cluster kernel relies on, and times them:
//  (1) all-to-all exchange of 32 bytes per block pair with st.async + complete_tx, the receiver's
//      arrive.expect_tx possibly AFTER the data has landed (transaction count transiently negative);
//  (2) red.async.add of a 512-word table into an owner block + 16 bytes of statistics;
//  (3) candidates appended to the owner's list: remote atomic for the position, st.async for the value.
// Every wait is bounded (prints "TIMEOUT").  nvcc -gencode arch=compute_100a,code=sm_100a -o dsmem_proto dsmem_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned cl_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_map(unsigned saddr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r; }
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned a, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_arrive(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ int g_timeouts;
__device__ __forceinline__ bool mbar_wait(unsigned a, unsigned parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(a, parity)) {
    if (clock64() - t0 > 400000000ll) { atomicAdd(&g_timeouts, 1); return false; }
  }
  return true;
}
__device__ __forceinline__ void st_async_v4(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar) : "memory"); }
__device__ __forceinline__ void st_async_u32(unsigned addr, unsigned a, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" :: "r"(addr), "r"(a), "r"(mbar) : "memory"); }
__device__ __forceinline__ void red_async_add(unsigned addr, unsigned a, unsigned mbar) {
  asm volatile("red.async.relaxed.cluster.mbarrier::complete_tx::bytes.shared::cluster.add.u32 [%0], %1, [%2];" :: "r"(addr), "r"(a), "r"(mbar) : "memory"); }
__device__ __forceinline__ unsigned cl_atom_add(unsigned addr, unsigned v) {
  unsigned o; asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); return o; }

struct Sm {
  unsigned long long xbar[2];
  int xslot[2][8][8];
  unsigned long long mb_merge, mb_cmd, mb_gather, mb_pad;
  unsigned cmd[16];
  unsigned stat[8][4];
  unsigned gcount, pad[3];
  unsigned glist[256];
  unsigned table[512];
};

// mode 0: exchange; delay_expect: the expect_tx thread waits ~delay cycles first (data lands before it)
__global__ void __launch_bounds__(576, 1) k_exchange(int iters, int delay, unsigned long long* out) {
  extern __shared__ __align__(16) unsigned char raw[];
  Sm* sm = reinterpret_cast<Sm*>(raw);
  const int t = threadIdx.x;
  const unsigned csize = cl_size(), crank = cl_rank();
  if (t == 0) {
    mbar_init((unsigned)__cvta_generic_to_shared(&sm->xbar[0]), 1);
    mbar_init((unsigned)__cvta_generic_to_shared(&sm->xbar[1]), 1);
  }
  mbar_fence_init();
  cl_sync();
  long long errs = 0;
  const long long t0 = clock64();
  bool dead = false;
  for (int q = 0; q < iters && !dead; ++q) {
    const unsigned pb = q & 1, par = (q >> 1) & 1;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&sm->xbar[pb]);
    __syncthreads();
    if (t < (int)csize) {
      const unsigned dst = cl_map((unsigned)__cvta_generic_to_shared(&sm->xslot[pb][crank][0]), t), rbar = cl_map(bar, t);
      st_async_v4(dst, q * 7 + crank, q + 1, crank, 3, rbar);
      st_async_v4(dst + 16, q, crank * 2, 0, 0, rbar);
    } else if (t == 32) {
      if (delay) { const long long d0 = clock64(); while (clock64() - d0 < delay) { } }
      mbar_expect_tx_arrive(bar, csize * 32);
    }
    if (!mbar_wait(bar, par)) { dead = true; break; }
    int s0 = 0, s1 = 0, s2 = 0, s5 = 0;
    for (unsigned r = 0; r < csize; ++r) { s0 += sm->xslot[pb][r][0]; s1 += sm->xslot[pb][r][1]; s2 += sm->xslot[pb][r][2]; s5 += sm->xslot[pb][r][5]; }
    const int cs = (int)csize, tri = cs * (cs - 1) / 2;
    if (s0 != q * 7 * cs + tri || s1 != (q + 1) * cs || s2 != tri || s5 != 2 * tri) ++errs;
  }
  const long long t1 = clock64();
  if (t == 0) { out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0); }
  if (errs) atomicAdd(&out[blockIdx.x * 4 + 1], (unsigned long long)errs);
  if (dead && t == 0) out[blockIdx.x * 4 + 2] = 1;
}

// owner/peer protocol of one column, warp 0 of every block; owner = iteration % csize
__global__ void __launch_bounds__(576, 1) k_column(int iters, unsigned long long* out) {
  extern __shared__ __align__(16) unsigned char raw[];
  Sm* sm = reinterpret_cast<Sm*>(raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned csize = cl_size(), crank = cl_rank();
  const unsigned mb_merge = (unsigned)__cvta_generic_to_shared(&sm->mb_merge), mb_cmd = (unsigned)__cvta_generic_to_shared(&sm->mb_cmd),
                 mb_gather = (unsigned)__cvta_generic_to_shared(&sm->mb_gather);
  if (t == 0) { mbar_init(mb_merge, 1); mbar_init(mb_cmd, 1); mbar_init(mb_gather, 1); }
  mbar_fence_init();
  cl_sync();
  unsigned par = 0;
  long long errs = 0;
  long long tsum = 0;
  bool dead = false;
  for (int it = 0; it < iters; ++it) {
    const unsigned owner = (unsigned)it % csize;
    if (warp == 0 && !dead) {
      for (int i = 0; i < 16; ++i) sm->table[i * 32 + lane] = (crank == owner) ? (unsigned)(i + lane) : 0u;   // owner: own contribution
      if (lane == 0) sm->gcount = 0;
    }
    cl_sync();
    const long long t0 = clock64();
    if (warp == 0 && !dead) {
      if (crank != owner) {
        const unsigned o_tab = cl_map((unsigned)__cvta_generic_to_shared(&sm->table[0]), owner), o_merge = cl_map(mb_merge, owner);
        for (int i = 0; i < 16; ++i) red_async_add(o_tab + 4 * (i * 32 + lane), (unsigned)(it + crank * (i + 1) + lane), o_merge);
        if (lane == 0) {
          st_async_v4(cl_map((unsigned)__cvta_generic_to_shared(&sm->stat[crank][0]), owner), crank, it, 5, 6, o_merge);
          mbar_expect_tx_arrive(mb_cmd, 64);
        }
        if (!mbar_wait(mb_cmd, (par >> 1) & 1)) dead = true;
        par ^= 2;
        if (!dead) {
          if (sm->cmd[0] != 1u || sm->cmd[15] != (unsigned)it) ++errs;
          // gather: every lane appends (lane % 3) values
          const unsigned mine = lane % 3;
          if (mine) {
            unsigned pos = cl_atom_add(cl_map((unsigned)__cvta_generic_to_shared(&sm->gcount), owner), mine);
            for (unsigned k = 0; k < mine; ++k)
              st_async_u32(cl_map((unsigned)__cvta_generic_to_shared(&sm->glist[0]), owner) + 4 * ((pos + k) & 255), 1000u * crank + lane, cl_map(mb_gather, owner));
          }
        }
      } else {
        if (lane == 0) mbar_expect_tx_arrive(mb_merge, (csize - 1) * (2048 + 16));
        if (!mbar_wait(mb_merge, par & 1)) dead = true;
        par ^= 1;
        if (!dead) {
          __syncwarp();
          // check the merged table
          for (int i = 0; i < 16; ++i) {
            unsigned want = (unsigned)(i + lane);
            for (unsigned r = 0; r < csize; ++r) if (r != owner) want += (unsigned)(it + r * (i + 1) + lane);
            if (sm->table[i * 32 + lane] != want) ++errs;
          }
          for (unsigned r = 0; r < csize; ++r) if (r != owner && (sm->stat[r][0] != r || sm->stat[r][1] != (unsigned)it)) ++errs;
          if ((unsigned)lane < csize && (unsigned)lane != crank) {
            const unsigned pc = cl_map((unsigned)__cvta_generic_to_shared(&sm->cmd[0]), lane), pm = cl_map(mb_cmd, lane);
            st_async_v4(pc, 1, 2, 3, 4, pm); st_async_v4(pc + 16, 5, 6, 7, 8, pm); st_async_v4(pc + 32, 9, 10, 11, 12, pm); st_async_v4(pc + 48, 13, 14, 15, it, pm);
          }
          const unsigned mine = lane % 3;
          if (mine) { unsigned pos = atomicAdd(&sm->gcount, mine); for (unsigned k = 0; k < mine; ++k) sm->glist[(pos + k) & 255] = 1000u * crank + lane; }
          const unsigned per = 31;             // sum over lanes of lane % 3 = 10 * (0+1+2) + 0 + 1 = 31
          if (lane == 0) mbar_expect_tx_arrive(mb_gather, 4 * per * (csize - 1));
          if (!mbar_wait(mb_gather, (par >> 2) & 1)) dead = true;
          par ^= 4;
          if (!dead) {
            __syncwarp();
            if (sm->gcount != per * csize) ++errs;
            // every (rank, lane) value must appear lane % 3 times
            unsigned cnt = 0;
            for (unsigned i = 0; i < per * csize; ++i) { const unsigned v = sm->glist[i]; if (v % 1000u == (unsigned)lane) ++cnt; }
            if (cnt != (lane % 3) * csize) ++errs;
          }
        }
      }
    }
    const long long t1 = clock64();
    if (crank == owner && warp == 0) tsum += t1 - t0;
    cl_sync();
  }
  if (t == 0) out[blockIdx.x * 4 + 0] = (unsigned long long)tsum;
  if (errs) atomicAdd(&out[blockIdx.x * 4 + 1], (unsigned long long)errs);
  if (dead && lane == 0 && warp == 0) out[blockIdx.x * 4 + 2] = 1;
}

template <class K, class... A> static int launch(K k, int csz, int nclus, size_t smem, A... args) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csz * nclus); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  cudaOccupancyMaxActiveClusters(&n, k, &cfg);
  printf("  [cluster %d: max active clusters %d] ", csz, n);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, args...);
  if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("sync error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 4096 * sizeof(unsigned long long));
  const size_t smem = 200 * 1024;      // one block per SM, like the production kernel
  for (int csz : {2, 5, 6, 8}) {
    for (int delay : {0, 3000}) {
      const int nclus = 8, iters = 2000;
      memset(out, 0, 4096 * sizeof(unsigned long long));
      printf("exchange csize %d delay %d:", csz, delay);
      if (launch(k_exchange, csz, nclus, smem, iters, delay, out)) return 1;
      unsigned long long cyc = 0, errs = 0, dead = 0;
      for (int b = 0; b < csz * nclus; ++b) { cyc += out[b * 4]; errs += out[b * 4 + 1]; dead += out[b * 4 + 2]; }
      printf("%.0f cycles per exchange, errors %llu, timeouts %llu\n", (double)cyc / (csz * nclus) / iters, errs, dead);
    }
    {
      const int nclus = 8, iters = 600;
      memset(out, 0, 4096 * sizeof(unsigned long long));
      printf("column   csize %d:", csz);
      if (launch(k_column, csz, nclus, smem, iters, out)) return 1;
      unsigned long long cyc = 0, errs = 0, dead = 0;
      for (int b = 0; b < csz * nclus; ++b) { cyc += out[b * 4]; errs += out[b * 4 + 1]; dead += out[b * 4 + 2]; }
      printf("%.0f cycles per column round trip (owner side), errors %llu, timeouts %llu\n", (double)cyc / nclus / iters, errs, dead);
    }
  }
  return 0;
}
