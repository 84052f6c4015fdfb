// Internal glue between the translation units of libmcz_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mcz {

int set_error(const char* msg);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);

size_t replay_scratch_bytes(long long G, long long N, int precision, int* grid_out);
int launch_replay(const double* flux, const double* noise, long long G, long long N, uint32_t tok,
                  int dust, int precision, double* Z, unsigned long long* present, int* trips,
                  int* ret, void* scratch, size_t scratch_bytes, cudaStream_t stream, double* aux = nullptr);

int launch_percentiles_f64(const double* Z, const unsigned char* present, long long rows, long long N,
                           double* pct, int* nvalid, signed char* status, cudaStream_t stream);

int launch_histogram_assist(const double* Z_d, long long rows, long long N, const int* mlist_h, int nm,
                            double* stats_d, int* counts_d, int* mlist_scratch_d, cudaStream_t stream);
int launch_ks_2samp(const double* a_d, long long n1, const double* b_d, long long n2, double* out_d,
                    unsigned long long* scratch_d, cudaStream_t stream);
int read_launch_stats(const void* scratch_d, unsigned long long* out8, cudaStream_t stream);
int launch_sample_export(const float* meas, const float* err, long long G, long long n_draw, uint32_t tok,
                         unsigned long long seed, unsigned long long galaxy_offset, int sample_mode,
                         int deterministic, float* flux_d, float* noise_d, cudaStream_t stream);
int read_col_cycles(unsigned long long* out, int reset);
int read_phase_cycles(unsigned long long* out8, int reset);
int read_sel_counters(unsigned long long* out8, int reset);
int launch_select_columns(const float* cols_d, long long ncols, long long N, float* pct_d, int* nvalid_d,
                          signed char* status_d, float* padded_d, cudaStream_t stream);
size_t production_scratch_bytes(long long G, long long n_draw, uint32_t tok);
const char* production_kernel_name(long long G, long long n_draw, uint32_t tok);
// Destination tables of a launch when they are not the caller's own: the gathered tables of up
// to 8 ranks (peer-mapped device pointers), rows [row0, row0 + G) of each.
struct ProdDst {
  int n;
  long long row0;
  float* pct[8];
  int* nvalid[8];
  signed char* status[8];
};
// One object's samples sharded over the GPUs of a node (wide kernel): rank, world and every rank's
// shared region (control block, merged histograms, barrier counter) as mapped into this process.
struct WideShare {
  int rank, world;
  void* shared[8];
};
int launch_production(const float* meas, const float* err, long long G, long long n_draw, uint32_t tok,
                      int dust, unsigned long long seed, unsigned long long galaxy_offset,
                      int sample_mode, int deterministic, float* pct, int* nvalid,
                      signed char* status, int* trips, int* ret, float* Z, void* scratch,
                      size_t scratch_bytes, cudaStream_t stream, const ProdDst* dst = nullptr,
                      const struct WideShare* ws = nullptr);
size_t wide_shared_bytes(uint32_t tok);
size_t wide_private_bytes(long long N, uint32_t tok);

}  // namespace mcz
