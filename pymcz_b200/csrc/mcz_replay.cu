// Replay kernel: metallicity.calculation (reference metallicity.py:86-257) on caller-supplied
// flux samples for G galaxies, one thread block per galaxy, double or float arithmetic.
// This is the parity mode: it consumes the very samples the reference drew with numpy and writes
// every per-sample output, so masks and values can be compared one to one with the oracle.
#include "mcz_physics.cuh"
#include "mcz_device_utils.cuh"
#include "mcz_internal.h"

namespace mcz {

static constexpr int RB = 256;   // threads per block

__device__ __forceinline__ double sanitise(double v, bool dropneg) {      // metallicity.py:96-102
  if (!(fabs(v) <= 1.7976931348623157e308)) v = 0.0;
  if (v < 0.0) v = dropneg ? Math<double>::nan() : 0.0;       // DROPNEGATIVES / FIXNEGATIVES
  return v;
}

template <typename R> struct GlobalPut {
  static constexpr bool WANT_AUX = true;
  double* base;       // &Z[g][0][s]
  long long N;
  double* aux;        // &aux[g][0][s] or null: the intermediate quantities (AuxKey)
  __device__ __forceinline__ void operator()(int key, R v) {
    if (key < NKEYS) base[(long long)key * N] = (double)v;
    else if (aux) aux[(long long)(key - NKEYS) * N] = (double)v;
  }
};

template <typename R>
__global__ void __launch_bounds__(RB)
replay_kernel(const double* __restrict__ flux, const double* __restrict__ noise, long long G,
              long long N, uint32_t tok, int dust, double* __restrict__ Z,
              unsigned long long* __restrict__ present_out, int* __restrict__ trips_out,
              int* __restrict__ ret_out, IterState<R>* __restrict__ scratch, double* __restrict__ aux) {
  __shared__ uint32_t sh_flags;
  __shared__ double sh_red[4 * (RB / 32)];
  const int tid = threadIdx.x;
  const int nw = RB / 32;
  const int used2line[NUSED] = {L_HA, L_HB, L_O2, L_O3, L_N2, L_S2A, L_S2B, L_S39069};
  IterState<R>* st = scratch + (long long)blockIdx.x * N;
  int phase = 0;

  for (long long g = blockIdx.x; g < G; g += gridDim.x) {
    __syncthreads();
    if (tid == 0) sh_flags = 0;
    __syncthreads();
    // ---- has* flags: any(sample > 0) over the sanitised samples (metscales.py:287-378)
    uint32_t local = 0;
    for (long long s = tid; s < N; s += RB) {
#pragma unroll
      for (int u = 0; u < NUSED; ++u) {
        const double v = sanitise(flux[((long long)used2line[u] * G + g) * N + s], (tok & T_DROPNEG) != 0u);
        const bool pos = (u >= U_S2B) ? (v > 1e-9) : (v > 0.0);
        if (pos) local |= 1u << u;
      }
    }
    local = __reduce_or_sync(0xffffffffu, local);
    if ((tid & 31) == 0 && local) atomicOr(&sh_flags, local);
    __syncthreads();
    const uint32_t flags = sh_flags;
    const int de = effective_dust(dust, flags);
    const uint64_t pm = present_mask(flags, tok);
    const bool ok = minimum_ok(flags);

    // ---- phase A
    for (long long s = tid; s < N; s += RB) {
      R f[NUSED], e[5];
#pragma unroll
      for (int u = 0; u < NUSED; ++u)
        f[u] = (R)sanitise(flux[((long long)used2line[u] * G + g) * N + s], (tok & T_DROPNEG) != 0u);
#pragma unroll
      for (int j = 0; j < 5; ++j) e[j] = noise ? (R)noise[((long long)j * G + g) * N + s] : R(0);
      double* auxs = aux ? aux + ((long long)g * NAUX) * N + s : nullptr;
      if (auxs)
        for (int k = 0; k < NAUX; ++k) auxs[(long long)k * N] = Math<double>::nan();
      GlobalPut<R> put{Z + ((long long)g * NKEYS) * N + s, N, auxs};
      Carry<R> c;
      phase_a<R>(f, e, flags, tok, de, put, c);
      if ((tok & T_KD02) && ok) {
        IterState<R> is;
        const R kd = ((pm >> K_KD02_N2O2) & 1ull) ? (R)put.base[(long long)K_KD02_N2O2 * N] : Math<R>::nan();
        make_iter<R>(c, kd, is);
        st[s] = is;
      }
    }
    // keys that are None for this galaxy read back as NaN
    for (int k = 0; k < NKEYS; ++k)
      if (!((pm >> k) & 1ull))
        for (long long s = tid; s < N; s += RB) Z[((long long)g * NKEYS + k) * N + s] = Math<double>::nan();

    int ii1 = 0, ii2 = 0;
    if ((tok & T_KD02) && ok) {
      const bool has_kn = (pm >> K_KK04_N2HA) & 1ull, has_kr = (pm >> K_KK04_R23) & 1ull;
      const bool o3o2 = (flags & F_O2) && (flags & F_O3);
      // start of the N2Ha loop: KD02_N2O2, or 8.6 if it is None / all-NaN (metscales.py:1097-1103)
      double nn = 0.0;
      if ((pm >> K_KD02_N2O2) & 1ull)
        for (long long s = tid; s < N; s += RB) nn += Math<R>::isnan(st[s].Z1) ? 0.0 : 1.0;
      nn = block_sum(nn, sh_red, nw, phase);
      if (nn == 0.0)
        for (long long s = tid; s < N; s += RB) st[s].Z1 = R(8.6);
      if (has_kn) {
        if (o3o2) {
          double conv = 100.0;
          while (conv > 1.0e-3 && ii1 < 100) {                         // metscales.py:1111
            ++ii1;
            double sum = 0.0;
            for (long long s = tid; s < N; s += RB) {
              IterState<R> is = st[s];
              sum += (double)kk04_n2ha_trip<R>(is);
              st[s].Z1 = is.Z1;
              st[s].lq1 = is.lq1;
            }
            conv = block_sum(sum, sh_red, nw, phase) / (double)N;      // metscales.py:1117
          }
        } else {
          for (long long s = tid; s < N; s += RB) st[s].Z1 = kk04_n2ha_noq<R>(st[s]);
        }
      }
      if (has_kr) {
        double conv = 100.0;
        while (conv > 1e-4 && ii2 < 100) {                             // metscales.py:1166
          ++ii2;
          double sum = 0.0;
          for (long long s = tid; s < N; s += RB) {
            IterState<R> is = st[s];
            sum += (double)kk04_r23_trip<R>(is);
            st[s].Z2 = is.Z2;
            st[s].lq2 = is.lq2;
          }
          conv = fabs(block_sum(sum, sh_red, nw, phase) / (double)N);  // metscales.py:1177
        }
      }
      // ---- phase C
      for (long long s = tid; s < N; s += RB) {
        double* col = Z + ((long long)g * NKEYS) * N + s;
        GlobalPut<R> put{col, N, nullptr};
        const R kd = ((pm >> K_KD02_N2O2) & 1ull) ? (R)col[(long long)K_KD02_N2O2 * N] : Math<R>::nan();
        const R m91 = (R)col[(long long)K_M91 * N];
        phase_c<R>(st[s], pm, ii1 >= 100, ii2 >= 100, kd, m91, put);
        // self.logq as calcKK04_N2Ha leaves it: the last logq of its loop, or 7.37177 without [OIII]/[OII]
        if (aux && has_kn) aux[((long long)g * NAUX + (A_LOGQ - NKEYS)) * N + s] = o3o2 ? (double)st[s].lq1 : 7.37177;
      }
    }
    if (tid == 0) {
      present_out[g] = pm;
      trips_out[2 * g] = ii1;
      trips_out[2 * g + 1] = ii2;
      ret_out[g] = ok ? 0 : -1;
    }
  }
}

size_t replay_scratch_bytes(long long G, long long N, int precision, int* grid_out) {
  int grid = (int)(G < 1184 ? (G > 0 ? G : 1) : 1184);   // 148 SMs x 8 resident blocks
  if (grid_out) *grid_out = grid;
  const size_t per = precision == 0 ? sizeof(IterState<double>) : sizeof(IterState<float>);
  return (size_t)grid * (size_t)N * per;
}

int launch_replay(const double* flux, const double* noise, long long G, long long N, uint32_t tok,
                  int dust, int precision, double* Z, unsigned long long* present, int* trips,
                  int* ret, void* scratch, size_t scratch_bytes, cudaStream_t stream, double* aux) {
  int grid = 0;
  const size_t need = replay_scratch_bytes(G, N, precision, &grid);
  if (scratch_bytes < need) return set_error("mcz_forward_replay: scratch too small");
  if (G <= 0 || N <= 0) return 0;
  if (precision == 0)
    replay_kernel<double><<<grid, RB, 0, stream>>>(flux, noise, G, N, tok, dust, Z, present, trips,
                                                   ret, (IterState<double>*)scratch, aux);
  else
    replay_kernel<float><<<grid, RB, 0, stream>>>(flux, noise, G, N, tok, dust, Z, present, trips,
                                                  ret, (IterState<float>*)scratch, aux);
  count_launch();
  return check_cuda(cudaGetLastError(), "replay_kernel launch");
}

}  // namespace mcz
