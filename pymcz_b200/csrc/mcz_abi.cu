// extern "C" surface of libmcz_b200.so (declared in include/mcz_b200.h).
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include "../../include/mcz_b200.h"
#include "../../include/mcz_b200_debug.h"
#include "mcz_common.cuh"
#include "mcz_internal.h"

namespace mcz {

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};

int set_error(const char* msg) {
  g_error = msg;
  return 1;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
  return 2;
}
void count_launch(int n) { g_launches += n; }

static const char* KEY_NAMES[NKEYS] = {
    "E(B-V)", "logR23", "D02", "Z94", "M91", "C01_N2S2", "C01_R23", "P05", "P01", "PP04_N2Ha",
    "PP04_O3N2", "DP00", "P10_ONS", "P10_ON", "M08_R23", "M08_N2Ha", "M08_O3Hb", "M08_O2Hb",
    "M08_O3O2", "M08_O3N2", "M13_O3N2", "M13_N2", "D13_N2S2_O3S2", "D13_N2S2_O3Hb",
    "D13_N2S2_O3O2", "D13_N2O2_O3S2", "D13_N2O2_O3Hb", "D13_N2O2_O3O2", "D13_N2Ha_O3Hb",
    "D13_N2Ha_O3O2", "KD02_N2O2", "KD02_N2S2", "KK04_N2Ha", "KK04_R23", "KD02comb", "PM14", "D16"};
static const char* LINE_NAMES[NLINES] = {"[OII]3727", "Hb", "[OIII]4959", "[OIII]5007", "[OI]6300", "Ha",
                                         "[NII]6584", "[SII]6717", "[SII]6731", "[SIII]9069", "[SIII]9532"};

}  // namespace mcz

using namespace mcz;

struct mcz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;     // result tables go back chunk by chunk while the next chunk computes
  cudaEvent_t chunk_done[16] = {};
  // device buffers (grown on demand)
  float *meas = nullptr, *err = nullptr, *pct = nullptr;
  int *nvalid = nullptr, *trips = nullptr, *ret = nullptr;
  signed char* status = nullptr;
  void* scratch = nullptr;
  size_t cap_g = 0, cap_scratch = 0;
};

extern "C" {

const char* mcz_version(void) { return "pymcz_b200 0.1 (sm_100a)"; }
const char* mcz_last_error(void) { return g_error.c_str(); }
const char* mcz_key_name(int i) { return (i >= 0 && i < NKEYS) ? KEY_NAMES[i] : ""; }
const char* mcz_line_name(int i) { return (i >= 0 && i < NLINES) ? LINE_NAMES[i] : ""; }
long long mcz_launch_count(void) { return g_launches.load(); }

int mcz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

size_t mcz_replay_scratch_bytes(long long G, long long N, int precision) {
  return replay_scratch_bytes(G, N, precision, nullptr);
}

int mcz_forward_replay(const double* flux_d, const double* noise_d, long long G, long long N,
                       unsigned tokens, int dust, int precision, double* Z_d,
                       unsigned long long* present_d, int* trips_d, int* ret_d, void* scratch_d,
                       size_t scratch_bytes, void* stream) {
  if (!flux_d || !Z_d || !present_d || !trips_d || !ret_d) return set_error("mcz_forward_replay: null pointer");
  if (precision != 0 && precision != 1) return set_error("mcz_forward_replay: precision must be 0 or 1");
  if ((tokens & (T_D02 | T_M13)) && !noise_d) return set_error("mcz_forward_replay: D02/M13 need noise_d");
  return launch_replay(flux_d, noise_d, G, N, tokens, dust, precision, Z_d, present_d, trips_d, ret_d,
                       scratch_d, scratch_bytes, (cudaStream_t)stream);
}

int mcz_forward_replay_aux(const double* flux_d, const double* noise_d, long long G, long long N,
                           unsigned tokens, int dust, int precision, double* Z_d,
                           unsigned long long* present_d, int* trips_d, int* ret_d, double* aux_d,
                           void* scratch_d, size_t scratch_bytes, void* stream) {
  if (!flux_d || !Z_d || !present_d || !trips_d || !ret_d) return set_error("mcz_forward_replay_aux: null pointer");
  if (precision != 0 && precision != 1) return set_error("mcz_forward_replay_aux: precision must be 0 or 1");
  if ((tokens & (T_D02 | T_M13)) && !noise_d) return set_error("mcz_forward_replay_aux: D02/M13 need noise_d");
  return launch_replay(flux_d, noise_d, G, N, tokens, dust, precision, Z_d, present_d, trips_d, ret_d,
                       scratch_d, scratch_bytes, (cudaStream_t)stream, aux_d);
}

int mcz_segmented_percentiles(const double* Z_d, const unsigned char* present_d, long long rows,
                              long long N, double* pct_d, int* nvalid_d, signed char* status_d,
                              void* stream) {
  if (!Z_d || !pct_d || !nvalid_d || !status_d) return set_error("mcz_segmented_percentiles: null pointer");
  return launch_percentiles_f64(Z_d, present_d, rows, N, pct_d, nvalid_d, status_d, (cudaStream_t)stream);
}

int mcz_histogram_assist(const double* Z_d, long long rows, long long N, const int* bin_numbers_h,
                         int n_bin_numbers, double* stats_d, int* counts_d, int* scratch_d, void* stream) {
  if (!Z_d || !stats_d || (n_bin_numbers > 0 && !bin_numbers_h)) return set_error("mcz_histogram_assist: null pointer");
  return launch_histogram_assist(Z_d, rows, N, bin_numbers_h, n_bin_numbers, stats_d, counts_d, scratch_d,
                                 (cudaStream_t)stream);
}

int mcz_ks_2samp(const double* a_d, long long n1, const double* b_d, long long n2, double* out_d,
                 void* scratch_d, void* stream) {
  return launch_ks_2samp(a_d, n1, b_d, n2, out_d, static_cast<unsigned long long*>(scratch_d), (cudaStream_t)stream);
}

size_t mcz_production_scratch_bytes(long long G, long long n_draw, unsigned tokens) {
  return production_scratch_bytes(G, n_draw, tokens);
}

const char* mcz_production_kernel_name(long long G, long long n_draw, unsigned tokens) {
  return production_kernel_name(G, n_draw, tokens);
}

int mcz_forward_production(const float* meas_d, const float* err_d, long long G, long long n_draw,
                           unsigned tokens, int dust, unsigned long long seed,
                           unsigned long long galaxy_offset, int sample_mode, int deterministic,
                           float* pct_d, int* nvalid_d, signed char* status_d, int* trips_d,
                           int* ret_d, float* Z_d, void* scratch_d, size_t scratch_bytes,
                           void* stream) {
  if (!meas_d || !err_d || !pct_d || !nvalid_d || !status_d || !trips_d || !scratch_d)
    return set_error("mcz_forward_production: null pointer");
  return launch_production(meas_d, err_d, G, n_draw, tokens, dust, seed, galaxy_offset, sample_mode,
                           deterministic, pct_d, nvalid_d, status_d, trips_d, ret_d, Z_d, scratch_d,
                           scratch_bytes, (cudaStream_t)stream);
}

int mcz_forward_production_gather(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                  unsigned tokens, int dust, unsigned long long seed,
                                  unsigned long long galaxy_offset, int sample_mode, int deterministic,
                                  int nranks, long long row0, void* const* pct_tables,
                                  void* const* nvalid_tables, void* const* status_tables, int* trips_d,
                                  int* ret_d, void* scratch_d, size_t scratch_bytes, void* stream) {
  if (!meas_d || !err_d || !pct_tables || !nvalid_tables || !status_tables || !trips_d || !scratch_d)
    return set_error("mcz_forward_production_gather: null pointer");
  if (nranks < 1 || nranks > 8) return set_error("mcz_forward_production_gather: 1 to 8 ranks");
  ProdDst dst;
  dst.n = nranks;
  dst.row0 = row0;
  for (int q = 0; q < nranks; ++q) {
    dst.pct[q] = static_cast<float*>(pct_tables[q]);
    dst.nvalid[q] = static_cast<int*>(nvalid_tables[q]);
    dst.status[q] = static_cast<signed char*>(status_tables[q]);
  }
  return launch_production(meas_d, err_d, G, n_draw, tokens, dust, seed, galaxy_offset, sample_mode,
                           deterministic, nullptr, nullptr, nullptr, trips_d, ret_d, nullptr, scratch_d,
                           scratch_bytes, (cudaStream_t)stream, &dst);
}

size_t mcz_samples_shared_bytes(unsigned tokens) { return wide_shared_bytes(tokens); }
size_t mcz_samples_scratch_bytes(long long n_draw, unsigned tokens) { return wide_private_bytes(n_draw, tokens); }

int mcz_forward_production_samples(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                   unsigned tokens, int dust, unsigned long long seed,
                                   unsigned long long galaxy_offset, int sample_mode, int deterministic,
                                   int nranks, int rank, void* const* shared_regions, float* pct_d,
                                   int* nvalid_d, signed char* status_d, int* trips_d, int* ret_d,
                                   void* scratch_d, size_t scratch_bytes, void* stream) {
  if (!meas_d || !err_d || !shared_regions || !pct_d || !nvalid_d || !status_d || !trips_d || !scratch_d)
    return set_error("mcz_forward_production_samples: null pointer");
  if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks)
    return set_error("mcz_forward_production_samples: 1 to 8 ranks, 0 <= rank < nranks");
  WideShare ws;
  ws.rank = rank;
  ws.world = nranks;
  for (int q = 0; q < nranks; ++q) ws.shared[q] = shared_regions[q];
  return launch_production(meas_d, err_d, G, n_draw, tokens, dust, seed, galaxy_offset, sample_mode,
                           deterministic, pct_d, nvalid_d, status_d, trips_d, ret_d, nullptr, scratch_d,
                           scratch_bytes, (cudaStream_t)stream, nullptr, nranks > 1 ? &ws : nullptr);
}

// diagnostics (not part of the public header): selection-path counters since the last reset:
// [0] columns, [1] fast path, [2] fallback: degenerate range, [3] fallback: target in an edge bin
// or > 4 target bins, [4] fallback: > 64 candidates (ties), [5] empty / non-positive columns
int mcz_debug_counters(unsigned long long* out8, int reset) { return read_sel_counters(out8, reset); }

// test entry point (not part of the public header): the production kernel's warp selection on
// caller-supplied float32 columns cols_d[ncols][N]; padded_d (ncols * roundup(N,128) floats) is
// needed for N > 2304 only
int mcz_debug_select_columns(const float* cols_d, long long ncols, long long N, float* pct_d, int* nvalid_d,
                             signed char* status_d, float* padded_d, void* stream) {
  if (!cols_d || !pct_d || !nvalid_d || !status_d) return set_error("mcz_debug_select_columns: null pointer");
  return launch_select_columns(cols_d, ncols, N, pct_d, nvalid_d, status_d, padded_d, (cudaStream_t)stream);
}

int mcz_debug_production_samples(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                 unsigned tokens, unsigned long long seed, unsigned long long galaxy_offset,
                                 int sample_mode, int deterministic, float* flux_d, float* noise_d, void* stream) {
  if (!meas_d || !err_d || !flux_d || !noise_d) return set_error("mcz_debug_production_samples: null pointer");
  return launch_sample_export(meas_d, err_d, G, n_draw, tokens, seed, galaxy_offset, sample_mode, deterministic,
                              flux_d, noise_d, (cudaStream_t)stream);
}

int mcz_production_last_stats(const void* scratch_d, unsigned long long* out8, void* stream) {
  if (!scratch_d || !out8) return set_error("mcz_production_last_stats: null pointer");
  return read_launch_stats(scratch_d, out8, (cudaStream_t)stream);
}
int mcz_debug_col_cycles(unsigned long long* out74, int reset) { return read_col_cycles(out74, reset); }
int mcz_debug_phase_cycles(unsigned long long* out8, int reset) { return read_phase_cycles(out8, reset); }

mcz_ctx* mcz_ctx_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    set_error("mcz_ctx_create: no such CUDA device (there is no CPU fallback)");
    return nullptr;
  }
  mcz_ctx* c = new mcz_ctx();
  c->device = device;
  cudaSetDevice(device);
  if (check_cuda(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream create") ||
      check_cuda(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "stream create")) {
    mcz_ctx_destroy(c);
    return nullptr;
  }
  for (auto& e : c->chunk_done) {
    if (check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create")) {
      mcz_ctx_destroy(c);
      return nullptr;
    }
  }
  return c;
}

void mcz_ctx_destroy(mcz_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->meas); cudaFree(c->err); cudaFree(c->pct); cudaFree(c->nvalid); cudaFree(c->trips);
  cudaFree(c->ret); cudaFree(c->status); cudaFree(c->scratch);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto& e : c->chunk_done) if (e) cudaEventDestroy(e);
  delete c;
}

int mcz_ctx_run_host(mcz_ctx* c, const float* meas_h, const float* err_h, long long G,
                     long long n_draw, unsigned tokens, int dust, unsigned long long seed,
                     unsigned long long galaxy_offset, int sample_mode, int deterministic,
                     float* pct_h, int* nvalid_h, signed char* status_h, int* trips_h, int* ret_h) {
  if (!c) return set_error("mcz_ctx_run_host: null context");
  if (G <= 0) return 0;
  int rc = check_cuda(cudaSetDevice(c->device), "set device");
  if (rc) return rc;
  if ((size_t)G > c->cap_g) {
    cudaFree(c->meas); cudaFree(c->err); cudaFree(c->pct); cudaFree(c->nvalid); cudaFree(c->trips);
    cudaFree(c->ret); cudaFree(c->status);
    c->meas = c->err = c->pct = nullptr; c->nvalid = c->trips = c->ret = nullptr; c->status = nullptr;
    c->cap_g = 0;
    const size_t g = (size_t)G;
    if ((rc = check_cuda(cudaMalloc(&c->meas, g * NLINES * 4), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->err, g * NLINES * 4), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->pct, g * NKEYS * 3 * 4), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->nvalid, g * NKEYS * 4), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->status, g * NKEYS), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->trips, g * 2 * 4), "malloc"))) return rc;
    if ((rc = check_cuda(cudaMalloc(&c->ret, g * 4), "malloc"))) return rc;
    c->cap_g = g;
  }
  const size_t need = production_scratch_bytes(G, n_draw, tokens);
  if (need > c->cap_scratch) {
    cudaFree(c->scratch);
    c->scratch = nullptr;
    c->cap_scratch = 0;
    if ((rc = check_cuda(cudaMalloc(&c->scratch, need), "malloc scratch"))) return rc;
    c->cap_scratch = need;
  }
  const size_t g = (size_t)G;
  cudaStream_t st = c->stream, cp = c->copy_stream;
  if ((rc = check_cuda(cudaMemcpyAsync(c->meas, meas_h, g * NLINES * 4, cudaMemcpyHostToDevice, st), "H2D meas"))) return rc;
  if ((rc = check_cuda(cudaMemcpyAsync(c->err, err_h, g * NLINES * 4, cudaMemcpyHostToDevice, st), "H2D err"))) return rc;
  // Galaxies are independent and Philox is keyed by the global galaxy index, so the table is run
  // as up to 8 launches over contiguous chunks (>= 16384 galaxies each): the result tables of a
  // chunk (641 B per galaxy) go back over PCIe on the copy stream while the next chunk computes.
  long long nchunk = G / 16384;
  nchunk = nchunk < 1 ? 1 : (nchunk > 8 ? 8 : nchunk);
  // (an error below leaves copies into the caller's buffers in flight: drain both streams before returning)
  auto fail = [&](int code) { cudaStreamSynchronize(st); cudaStreamSynchronize(cp); return code; };
  for (long long k = 0; k < nchunk; ++k) {
    const long long lo = G * k / nchunk, hi = G * (k + 1) / nchunk, n = hi - lo;
    rc = launch_production(c->meas + lo * NLINES, c->err + lo * NLINES, n, n_draw, tokens, dust, seed,
                           galaxy_offset + (unsigned long long)lo, sample_mode, deterministic,
                           c->pct + lo * NKEYS * 3, c->nvalid + lo * NKEYS, c->status + lo * NKEYS,
                           c->trips + lo * 2, c->ret + lo, nullptr, c->scratch, c->cap_scratch, st);
    if (rc) return fail(rc);
    if ((rc = check_cuda(cudaEventRecord(c->chunk_done[k], st), "event record"))) return fail(rc);
    if ((rc = check_cuda(cudaStreamWaitEvent(cp, c->chunk_done[k], 0), "stream wait"))) return fail(rc);
    const size_t m = (size_t)n;
    if ((rc = check_cuda(cudaMemcpyAsync(pct_h + lo * NKEYS * 3, c->pct + lo * NKEYS * 3, m * NKEYS * 3 * 4, cudaMemcpyDeviceToHost, cp), "D2H pct"))) return fail(rc);
    if ((rc = check_cuda(cudaMemcpyAsync(nvalid_h + lo * NKEYS, c->nvalid + lo * NKEYS, m * NKEYS * 4, cudaMemcpyDeviceToHost, cp), "D2H nvalid"))) return fail(rc);
    if ((rc = check_cuda(cudaMemcpyAsync(status_h + lo * NKEYS, c->status + lo * NKEYS, m * NKEYS, cudaMemcpyDeviceToHost, cp), "D2H status"))) return fail(rc);
    if (trips_h && (rc = check_cuda(cudaMemcpyAsync(trips_h + lo * 2, c->trips + lo * 2, m * 2 * 4, cudaMemcpyDeviceToHost, cp), "D2H trips"))) return fail(rc);
    if (ret_h && (rc = check_cuda(cudaMemcpyAsync(ret_h + lo, c->ret + lo, m * 4, cudaMemcpyDeviceToHost, cp), "D2H ret"))) return fail(rc);
  }
  if ((rc = check_cuda(cudaStreamSynchronize(st), "stream sync"))) return rc;
  return check_cuda(cudaStreamSynchronize(cp), "copy stream sync");
}

}  // extern "C"
