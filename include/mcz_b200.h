/* mcz_b200.h -- C ABI of the B200-native pyMCZ Monte-Carlo metallicity hot path.
 *
 * The reference (metal-sn/pyMCZ) is pure Python; the binding a maintainer adds is a ctypes stub
 * (see INTEGRATION.md).  Every entry point below replaces one piece of the reference's per-galaxy
 * Python path; citations are file:line under /root/reference/pyMCZ.
 *
 * Conventions
 *   - plain pointers and sizes only; `_d` pointers are device memory, `_h` pointers host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); device entry points only
 *     enqueue work on it and never synchronise;
 *   - every function returns 0 on success, non-zero on error; mcz_last_error() gives the message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - key order   = metallicity.py:29-56 (37 keys, mcz_key_name(i));
 *     line order  = mcz.py:45 (11 lines, mcz_line_name(i)).
 */
#ifndef MCZ_B200_H
#define MCZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCZ_NKEYS 37
#define MCZ_NLINES 11
#define MCZ_NNOISE 5

/* --md tokens (metallicity.py:161-257) as a bit mask */
#define MCZ_T_D02 (1u << 0)
#define MCZ_T_Z94 (1u << 1)
#define MCZ_T_M91 (1u << 2)
#define MCZ_T_PP04 (1u << 3)
#define MCZ_T_P10 (1u << 4)
#define MCZ_T_M08 (1u << 5)
#define MCZ_T_M08ALL (1u << 6)
#define MCZ_T_M13 (1u << 7)
#define MCZ_T_KD02 (1u << 8)
#define MCZ_T_P05 (1u << 9)
#define MCZ_T_C01 (1u << 10)
#define MCZ_T_P01 (1u << 11)
#define MCZ_T_DP00 (1u << 12)
#define MCZ_T_D16 (1u << 13)
#define MCZ_T_ALL (MCZ_T_D02 | MCZ_T_Z94 | MCZ_T_M91 | MCZ_T_PP04 | MCZ_T_P10 | MCZ_T_M08 | MCZ_T_M13 | MCZ_T_KD02)

/* dust request (metallicity.py:113-132) */
#define MCZ_DUST_NONE 0         /* --nodust */
#define MCZ_DUST_ON 1           /* correct; galaxy without Ha/Hb: E(B-V)=1e-5 applied (:116-129) */
#define MCZ_DUST_ON_IGNORING 2  /* correct; sticky IGNOREDUST already set (:127, 130-132) */

/* status codes of the percentile table (mcz.py:273-301) */
#define MCZ_ST_OK 0       /* median/p16/p84 valid */
#define MCZ_ST_ABSENT 1   /* mscales.mds[key] is None for this galaxy */
#define MCZ_ST_EMPTY 2    /* no finite sample: "-1 -1 -1" (mcz.py:277-280) */
#define MCZ_ST_NONPOS 3   /* sum(data) <= 0: "-1 -1 -1" (mcz.py:282-285) */

/* sampling semantics */
#define MCZ_SAMPLE_INDEPENDENT 0  /* one deviate per line and sample (serial run(), mcz.py:580-589) */
#define MCZ_SAMPLE_CORRELATED 1   /* one deviate per sample shared by all lines (calc(), mcz.py:442) */

/* ---- introspection ---------------------------------------------------------------------- */
const char* mcz_version(void);
const char* mcz_last_error(void);
const char* mcz_key_name(int i);   /* metallicity.get_keys()[i], metallicity.py:63 */
const char* mcz_line_name(int i);  /* mcz.alllines[i], mcz.py:45 */
int mcz_device_count(void);
/* number of kernels this library has launched since load (for the bench's gpu_launches) */
long long mcz_launch_count(void);

/* ---- replay: metallicity.calculation on caller-supplied samples ----------------------------
 * Replaces metallicity.calculation(mscales, measured, ...) (metallicity.py:86-257) and the
 * diagnostics.calc* methods it dispatches to (metscales.py:387-1264) for G galaxies at once.
 *   flux_d   [11][G][N] float64  perturbed fluxes per line (mcz.py:442 / :589), NaN row = absent
 *   noise_d  [5][G][N]  float64  coefficient noise: D02 e1,e2 (metscales.py:679-680), M13_N2 e1,e2
 *                                (:948-949), M13_O3N2 e1 (:952); may be NULL when neither is asked
 *   precision 0 = float64 arithmetic (parity mode), 1 = the production float32 arithmetic
 *   Z_d      [G][37][N] float64  mscales.mds[key] per sample; NaN where the key is None
 *   present_d[G]        uint64   bit k set <=> mds[key k] is not None
 *   trips_d  [G][2]     int32    trips of the KK04_N2Ha / KK04_R23 loops (metscales.py:1111, 1166)
 *   ret_d    [G]        int32    -1 where calculation() returns -1 (metallicity.py:145-146)
 *   scratch_d: mcz_replay_scratch_bytes(G, N, precision) bytes of device memory
 */
size_t mcz_replay_scratch_bytes(long long G, long long N, int precision);
int mcz_forward_replay(const double* flux_d, const double* noise_d, long long G, long long N,
                       unsigned tokens, int dust, int precision, double* Z_d,
                       unsigned long long* present_d, int* trips_d, int* ret_d, void* scratch_d,
                       size_t scratch_bytes, void* stream);

/* The same, also returning the intermediate per-sample quantities the reference keeps as attributes
 * of the diagnostics object (calcNIIOII, calcNIISII, calcR23, initialguess, calcKK04_N2Ha;
 * metscales.py:393-440, 468-489, 1112):
 *   aux_d [G][MCZ_NAUX][N] float64 (may be NULL): logN2O2, logN2S2, N2S2, R2, R3, R23, Z_init_guess,
 *         logq (as calcKK04_N2Ha leaves it), and the real root pair of the [NII]/[OII] quartic
 *         (larger, smaller; the real entries of N2O2_roots inside the KD02_N2O2 window); NaN where the
 *         reference would not have set the attribute.
 */
#define MCZ_NAUX 10
int mcz_forward_replay_aux(const double* flux_d, const double* noise_d, long long G, long long N,
                           unsigned tokens, int dust, int precision, double* Z_d,
                           unsigned long long* present_d, int* trips_d, int* ret_d, double* aux_d,
                           void* scratch_d, size_t scratch_bytes, void* stream);

/* ---- percentiles of stored distributions ------------------------------------------------------
 * Replaces the statistics part of savehist() (mcz.py:273-301): finite filter, n, sum<=0 test,
 * np.percentile(data, [50, 16, 84]) with linear interpolation, in float64.
 *   Z_d      [rows][N] float64;  present_d may be NULL (all rows present) or [rows] uint8
 *   pct_d    [rows][3] float64 (median, p16, p84; NaN unless status OK)
 *   nvalid_d [rows] int32, status_d [rows] int8
 */
int mcz_segmented_percentiles(const double* Z_d, const unsigned char* present_d, long long rows,
                              long long N, double* pct_d, int* nvalid_d, signed char* status_d,
                              void* stream);

/* ---- after the path: binning and sample-size assists --------------------------
 * mcz_histogram_assist replaces the data-parallel inside of the reference's histogram binning:
 * savehist's np.std (mcz.py:289) and getknuth's np.histogram(data, np.linspace(min(data),
 * max(data), m + 1)) (mcz.py:90-99), which Knuth's rule (knuthn, mcz.py:102-116) evaluates for every
 * bin number m the optimiser visits.  The optimiser and the gammaln posterior stay on the host.
 *   Z_d      [rows][N] float64 distributions (res[key][:, i] of mcz.py:600-611); non-finite = dropped
 *   bin_numbers_h  HOST array of n_bin_numbers bin numbers m, each in [1, 8192]
 *   stats_d  [rows][6] float64: n, min, max, mean, std (population, like np.std), sum of the finite values
 *   counts_d [rows][sum of all m] int32: for each row, the m counts of each requested m, one after
 *            the other in request order; bit-identical to np.histogram on the np.linspace edges
 *   scratch_d n_bin_numbers int32 of device memory (the bin list is uploaded there)
 */
int mcz_histogram_assist(const double* Z_d, long long rows, long long N, const int* bin_numbers_h,
                         int n_bin_numbers, double* stats_d, int* counts_d, int* scratch_d, void* stream);

/* mcz_ks_2samp replaces the statistic of scipy.stats.ks_2samp(x, xold) in the reference's sample-size
 * diagnostic (testcompleteness.py:111): both empirical CDFs at every data point, exactly as scipy
 * forms them (searchsorted(.., 'right') / n), then the extremes of their difference.  Neither sample
 * needs to be sorted; both must be finite and non-empty.  The p-value stays on the host.
 *   a_d [n1], b_d [n2] float64;  out_d [3] float64: D (two-sided), D+ = max(F_a - F_b), D- = max(F_b - F_a)
 *   scratch_d 16 bytes of device memory
 */
int mcz_ks_2samp(const double* a_d, long long n1, const double* b_d, long long n2, double* out_d,
                 void* scratch_d, void* stream);

/* ---- production: the fused Monte-Carlo path ---------------------------------------------------
 * Replaces, for G galaxies in one launch: the sampling and perturbation of run()/calc()
 * (mcz.py:509-519, 436-442, 580-589), metallicity.calculation (metallicity.py:86-257) and the
 * percentile extraction of savehist (mcz.py:273-301).
 *   meas_d, err_d [G][11] float32 (mcz.py:150: the reader parses fluxes as float32; NaN = absent)
 *   n_draw        samples per galaxy N' (= int(1.1 nsample), mcz.py:477-479)
 *   seed          Philox4x32-10 key; galaxy_offset + g is the Philox sub-sequence, so results do
 *                 not depend on how galaxies are sharded over GPUs
 *   sample_mode   MCZ_SAMPLE_*; deterministic != 0 reproduces nsample==1 (mcz.py:514-515)
 *   pct_d   [G][37][3] float32, nvalid_d [G][37] int32, status_d [G][37] int8,
 *   trips_d [G][2] int32, ret_d [G] int32 (may be NULL)
 *   Z_d     optional [G][37][n_draw] float32 per-sample values (NaN where absent), else NULL
 *   scratch_d: mcz_production_scratch_bytes(...) bytes
 */
size_t mcz_production_scratch_bytes(long long G, long long n_draw, unsigned tokens);
/* Statistics of the last mcz_forward_production* launch that used scratch_d (they live in the
 * header of the caller's scratch buffer: the library keeps no per-process state).  Copies them
 * to the host on `stream` and synchronises it.  out8: [0] galaxies processed, [1] KK04_N2Ha loop
 * trips executed (metscales.py:1111), [2] KK04_R23 loop trips executed (:1166), [3] galaxies whose
 * N2Ha loop was ended early because it provably runs to the reference's 100-trip cap, [4] the same
 * for the R23 loop (such galaxies are REPORTED with 100 trips and blanked, as the reference does,
 * metscales.py:1119-1121, 1179-1181; [1], [2] count what was executed), [5..7] reserved.  All zero
 * after a single-object (sample-tiled) launch. */
int mcz_production_last_stats(const void* scratch_d, unsigned long long* out8, void* stream);
/* Which kernel of the family a launch of this shape runs on the current device (the roofline
 * accounting and the profiles name it): "production_kernel<4,0>" (columns in shared memory,
 * N' <= 2304), "production_kernel<0,0>" / "<0,1>" (longer columns) or "wide_kernel" (single-object
 * runs: few galaxies, N' >= 32768, samples tiled over all thread blocks). */
const char* mcz_production_kernel_name(long long G, long long n_draw, unsigned tokens);
int mcz_forward_production(const float* meas_d, const float* err_d, long long G, long long n_draw,
                           unsigned tokens, int dust, unsigned long long seed,
                           unsigned long long galaxy_offset, int sample_mode, int deterministic,
                           float* pct_d, int* nvalid_d, signed char* status_d, int* trips_d,
                           int* ret_d, float* Z_d, void* scratch_d, size_t scratch_bytes,
                           void* stream);

/* ---- multi-GPU: production launch fused with the gather of the result tables --------------------
 * Replaces "Pool.map returns the per-galaxy results to the parent" (reference mcz.py:553-569) when
 * the galaxies are sharded over the GPUs of one node.  Rank r runs its G galaxies (global indices
 * galaxy_offset .. galaxy_offset + G) and the kernel writes each result row (641 B per galaxy)
 * straight into rows [row0, row0 + G) of EVERY rank's full-size table: pct_tables[q], nvalid_tables[q],
 * status_tables[q] are device pointers to rank q's tables as mapped into this process (NVLink peer
 * mappings, e.g. torch symmetric memory; q = own rank is the local table).  No separate collective
 * moves data; the tables are complete on every rank once every rank's launch has finished (one
 * barrier).  trips_d / ret_d stay local ([G,2], [G]).
 */
int mcz_forward_production_gather(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                  unsigned tokens, int dust, unsigned long long seed,
                                  unsigned long long galaxy_offset, int sample_mode, int deterministic,
                                  int nranks, long long row0, void* const* pct_tables,
                                  void* const* nvalid_tables, void* const* status_tables, int* trips_d,
                                  int* ret_d, void* scratch_d, size_t scratch_bytes, void* stream);

/* ---- multi-GPU, single-object runs: ONE galaxy's samples sharded over the GPUs --------------------
 * For runs with few measurements and very many samples each (the reference's per-galaxy vectors,
 * mcz.py:509-519, in the 10^6 .. 10^8 range).  Every rank passes the SAME meas/err table and gets the
 * SAME result tables; rank r draws and evaluates samples [N' r / nranks, N' (r+1) / nranks) (Philox is
 * indexed by the global sample number, so results do not depend on nranks) and everything the
 * reference takes over the whole sample vector -- the has* flags (metscales.py:287-378), the KK04 loop
 * means per trip (:1111, :1166), the order statistics (mcz.py:288) -- is reduced across ranks inside the
 * kernel through the shared regions: shared_regions[q] is rank q's region of
 * mcz_samples_shared_bytes(tokens) bytes as mapped into this process (NVLink peer mapping; q = rank is
 * the local one).  The caller zeroes its region, synchronises, and makes sure every rank has done so
 * (one host-side barrier) before any rank calls this; all ranks must call it together.
 * nranks = 1 is the single-GPU form (mcz_forward_production picks it by itself for few galaxies with
 * N' >= 32768 samples each).
 */
size_t mcz_samples_shared_bytes(unsigned tokens);
size_t mcz_samples_scratch_bytes(long long n_draw, unsigned tokens);   /* private scratch_d of each rank */
int mcz_forward_production_samples(const float* meas_d, const float* err_d, long long G, long long n_draw,
                                   unsigned tokens, int dust, unsigned long long seed,
                                   unsigned long long galaxy_offset, int sample_mode, int deterministic,
                                   int nranks, int rank, void* const* shared_regions, float* pct_d,
                                   int* nvalid_d, signed char* status_d, int* trips_d, int* ret_d,
                                   void* scratch_d, size_t scratch_bytes, void* stream);

/* ---- host-buffer convenience (what run() would call once for the whole table) -----------------
 * Same as mcz_forward_production but with HOST pointers: uploads meas/err, runs, downloads the
 * tables, synchronises.  Device buffers live in the context.  Tables of 32768 galaxies or more run
 * as up to 8 launches over contiguous chunks, and each chunk's result tables are copied back on
 * a second stream while the next chunk computes (results are identical: galaxies are independent
 * and the random stream is keyed by the global galaxy index).  Pass pinned host memory for the
 * copies to overlap.
 */
typedef struct mcz_ctx mcz_ctx;
mcz_ctx* mcz_ctx_create(int device);
void mcz_ctx_destroy(mcz_ctx* ctx);
int mcz_ctx_run_host(mcz_ctx* ctx, const float* meas_h, const float* err_h, long long G,
                     long long n_draw, unsigned tokens, int dust, unsigned long long seed,
                     unsigned long long galaxy_offset, int sample_mode, int deterministic,
                     float* pct_h, int* nvalid_h, signed char* status_h, int* trips_h, int* ret_h);

#ifdef __cplusplus
}
#endif
#endif /* MCZ_B200_H */
