/* mcz_b200_debug.h -- test and profiling entry points of libmcz_b200.so.
 *
 * Not part of the drop-in boundary (include/mcz_b200.h): nothing here replaces a reference
 * interface.  These symbols exist so that the parity tests can feed the kernels' internals with
 * inputs the physics never produces, or read the kernels' own random samples back, and so that
 * instrumented builds (-DMCZ_PHASE_TIMING, -DMCZ_SEL_STATS) can report their counters.
 * Conventions as in mcz_b200.h (device pointers `_d`, 0 = success, mcz_last_error()).
 */
#ifndef MCZ_B200_DEBUG_H
#define MCZ_B200_DEBUG_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The production kernels' warp selection (phase D; reference mcz.py:273-301) on caller-supplied
 * float32 columns cols_d[ncols][N]; padded_d (ncols * roundup(N,128) floats) is needed for
 * N > 2304 only. */
int mcz_debug_select_columns(const float* cols_d, long long ncols, long long N, float* pct_d, int* nvalid_d,
                             signed char* status_d, float* padded_d, void* stream);

/* The production kernels' sampler on its own: the perturbed and sanitised fluxes
 * (mcz.py:442 / :589, metallicity.py:96-99) and the scaled coefficient noise
 * (metscales.py:679-680, 948-952) exactly as phase A derives them for galaxies
 * galaxy_offset .. galaxy_offset + G, bit for bit.
 *   flux_d  [G][8][n_draw] float32, lines in the order Ha, Hb, [OII]3727, [OIII]5007, [NII]6584,
 *           [SII]6717, [SII]6731, [SIII]9069 (the lines the arithmetic consumes)
 *   noise_d [G][5][n_draw] float32: D02 e1, e2, M13_N2 e1, e2, M13_O3N2 e1
 * `tokens` decides which Philox calls are made, as in mcz_forward_production. */
int mcz_debug_production_samples(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                 unsigned tokens, unsigned long long seed, unsigned long long galaxy_offset,
                                 int sample_mode, int deterministic, float* flux_d, float* noise_d, void* stream);

/* Counters of instrumented builds (all zero otherwise), summed since the last reset:
 *   mcz_debug_counters      -DMCZ_SEL_STATS: [0] columns, [1] fast path, [2] no usable 64-sample range,
 *                           [3] crowded ranges, [4] empty / non-positive columns, [5] follow-up rounds
 *   mcz_debug_phase_cycles  -DMCZ_PHASE_TIMING: [0] galaxies, [1] phase A, [2] B+C, [3] D own work,
 *                           [4] D wall, [5] B prologue, [6] KK04 loops, [7] phase C   (clock64 cycles)
 *   mcz_debug_col_cycles    -DMCZ_PHASE_TIMING: per key, selection cycles (sum) then (max): 2 x 37 */
int mcz_debug_counters(unsigned long long* out8, int reset);
int mcz_debug_phase_cycles(unsigned long long* out8, int reset);
int mcz_debug_col_cycles(unsigned long long* out74, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MCZ_B200_DEBUG_H */
