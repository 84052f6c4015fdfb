"""GPU parity, production mode: the fused Philox -> physics -> percentiles kernel, through the
C ABI.  Three kinds of checks:
  1. per-sample: the kernel's own samples are re-derived with a numpy model of its Philox /
     Box-Muller stream (tests/_helpers.py), fed to the oracle in float64 and compared value by
     value (<= 1e-3 dex; masks equal except samples within float rounding of a threshold);
  2. selection: the percentile table equals np.percentile of the kernel's own columns;
  3. statistics: medians / 68 % regions agree with oracle runs within Monte-Carlo error, and
     with the paper's published annotations.
"""
import numpy as np
import pytest
import torch

from oracle import mcz_oracle as orc
from pymcz_b200 import _scales as sc
from pymcz_b200 import ops
from pymcz_b200.synth import synthetic_table
from tests._helpers import oracle_galaxy, production_flux_model, RecordingNormal

pytestmark = pytest.mark.gpu

EX_M = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
                 [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
EX_E = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
                 [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)


def run_prod(meas, err, n_draw, mds="all", dust=True, seed=0xB200, goff=0, correlated=False,
             deterministic=False, samples=True):
    tok, _, _ = sc.parse_md(mds)
    out = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n_draw, tok,
                                 sc.DUST_ON if dust else sc.DUST_NONE, seed, goff, correlated,
                                 deterministic, samples)
    torch.cuda.synchronize()
    return [o.cpu().numpy() if o is not None else None for o in out]


class FixedNoise:
    """Feeds pre-computed coefficient noise to the oracle in its np.random.normal call order."""

    def __init__(self, noise):
        self.noise = noise
        self.toggle = 0

    def __call__(self, loc, scale, size):
        row = {0.05: 0, 0.1: 1, 0.027: 2, 0.024: 3}.get(scale)
        if row is None:
            self.toggle ^= 1
            return self.noise[4] if self.toggle == 1 else np.zeros(size)
        return self.noise[row]


def oracle_on_kernel_samples(meas, err, n, seed, goff, mds, dust, correlated):
    G = meas.shape[0]
    Zo = np.full((G, sc.NKEYS, n), np.nan)
    po = np.zeros((G, sc.NKEYS), dtype=bool)
    trips = np.zeros((G, 2), dtype=np.int32)
    for g in range(G):
        flux, noise = production_flux_model(meas[g], err[g], seed, goff + g, n, correlated)
        d, ret, _ = orc.calculation(flux, mds, dust, False, "batched", FixedNoise(noise))
        trips[g] = (d.trips_n2ha, d.trips_r23)
        for j, k in enumerate(orc.KEYS):
            if d.mds[k] is not None:
                po[g, j] = True
                Zo[g, j] = d.mds[k]
    return Zo, po, trips


@pytest.mark.parametrize("mds,dust,correlated,n", [("all", True, False, 2200), ("all", False, True, 700),
                                                   ("KD02", True, False, 2304),
                                                   ("P05,C01,P01,DP00,D16,M08all,all", True, False, 1000)])
def test_per_sample_values_vs_oracle(mds, dust, correlated, n):
    sm, se = synthetic_table(40, config=4)
    meas, err = np.concatenate([EX_M, sm]), np.concatenate([EX_E, se])
    seed, goff = 0x1234ABCD5678, 1000
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, n, mds, dust, seed, goff, correlated)
    Zo, po, trips_o = oracle_on_kernel_samples(meas, err, n, seed, goff, mds, dust, correlated)
    pres = status != sc.ST_ABSENT
    assert np.array_equal(pres, po)
    same = np.all(trips == trips_o, axis=1)
    assert same.mean() >= 0.95, (trips[~same], trips_o[~same])
    Zs, Zos = Z[same].astype(np.float64), Zo[same]
    nan_mism = np.sum(np.isnan(Zs) != np.isnan(Zos))
    assert nan_mism <= 5e-5 * Zs.size + 2, nan_mism
    both = np.isfinite(Zs) & np.isfinite(Zos)
    d = np.abs(Zs[both] - Zos[both])
    # the device normals differ from the float64 model by MUFU rounding (~1e-6 sigma): values agree
    # far inside the 1e-3 dex bar except where a branch choice flips on such a rounding
    assert np.mean(d > 1e-3) < 1e-4, (np.mean(d > 1e-3), d.max())
    assert np.median(d) < 2e-5


@pytest.mark.parametrize("ngal", [30, 700])
def test_percentile_table_is_exact_selection_of_own_columns(ngal):
    # 700 galaxies x 17 columns reach every selection path many times: repeated values merged by
    # rank, split candidate lists, crowded ranges with follow-up rounds, empty columns
    sm, se = synthetic_table(ngal, config=4)
    meas, err = np.concatenate([EX_M, sm]), np.concatenate([EX_E, se])
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, 2200)
    for g in range(meas.shape[0]):
        for j in range(sc.NKEYS):
            col = Z[g, j]
            fin = col[np.isfinite(col)]
            if status[g, j] == sc.ST_ABSENT:
                assert np.all(np.isnan(col))
                continue
            assert nvalid[g, j] == fin.size, (g, j)
            if fin.size == 0:
                assert status[g, j] == sc.ST_EMPTY
            elif not (np.sum(fin.astype(np.float64)) > 0):
                assert status[g, j] == sc.ST_NONPOS
            else:
                assert status[g, j] == sc.ST_OK
                want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
                assert np.array_equal(pct[g, j], want), (g, sc.KEYS[j], pct[g, j], want)


def test_selection_adversarial_columns():
    """Ties, outliers, constant columns: exercised through deterministic/no-error inputs."""
    meas = EX_M.copy()
    err = np.zeros_like(EX_E)                       # every line sample identical -> constant columns
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, 2200)
    for g in range(2):
        for j in range(sc.NKEYS):
            if status[g, j] == sc.ST_OK:
                fin = Z[g, j][np.isfinite(Z[g, j])]
                want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
                assert np.array_equal(pct[g, j], want)
    # huge outliers: one line with SNR ~ 1 makes heavy-tailed / clipped distributions
    err2 = EX_E.copy()
    err2[:, 3] = meas[:, 3] * 0.9
    err2[:, 6] = meas[:, 6] * 0.7
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err2, 2200)
    for g in range(2):
        for j in range(sc.NKEYS):
            if status[g, j] == sc.ST_OK:
                fin = Z[g, j][np.isfinite(Z[g, j])]
                want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
                assert np.array_equal(pct[g, j], want), (g, sc.KEYS[j])


@pytest.mark.parametrize("nsample", [2000, 20000])
def test_exampledata_medians_within_mc_error(nsample):
    """Production mode vs oracle on exampledata (configs 1 and 2): same sampling mode, compare
    medians and 68 % half-widths within Monte-Carlo error (bootstrap-free bound: 5 sigma of the
    median's sampling error, sigma_med ~ 1.2533 * sd / sqrt(n))."""
    n = orc.n_draw(nsample)
    for correlated in (True, False):
        pct, nvalid, status, trips, ret, _ = run_prod(EX_M, EX_E, n, correlated=correlated, samples=False)
        rs = np.random.RandomState(99)
        out = orc.run_table(EX_M, EX_E, nsample, mode="correlated" if correlated else "shuffled",
                            rs=rs, normal=rs.normal)
        for g in range(2):
            for j, k in enumerate(sc.KEYS):
                so = out["status"][g, j]
                if so == orc.ST_ABSENT:
                    assert status[g, j] == sc.ST_ABSENT, k
                    continue
                if so not in (orc.ST_OK, orc.ST_NODIST) or out["nvalid"][g, j] < 0.5 * n:
                    continue
                assert status[g, j] == sc.ST_OK, (g, k, status[g, j])
                med, p16, p84 = out["pct"][g, j]
                half = max(0.5 * (p84 - p16), 1e-4)
                tol = 6.0 * 1.2533 * half / np.sqrt(out["nvalid"][g, j]) + 2e-4
                assert abs(pct[g, j, 0] - med) < tol, (g, k, pct[g, j], out["pct"][g, j])
                assert abs((pct[g, j, 2] - pct[g, j, 1]) - (p84 - p16)) < 8 * tol, (g, k)
                assert abs(nvalid[g, j] - out["nvalid"][g, j]) < 6 * np.sqrt(n) * 0.5 + 5


def test_published_triples():
    """Paper figure annotations, exampledata measurement 1, N=2000, correlated sampling."""
    pct, nvalid, status, trips, ret, _ = run_prod(EX_M[:1], EX_E[:1], 2200, correlated=True, samples=False)
    want = {"E(B-V)": (0.120, 0.102, 0.137), "KD02comb": (8.811, 8.799, 8.822),
            "KK04_R23": (9.059, 9.058, 9.060), "M91": (8.859, 8.849, 8.867),
            "PP04_O3N2": (8.920, 8.881, 8.976)}
    for k, (m, lo, hi) in want.items():
        got = pct[0, sc.KEY_INDEX[k]]
        assert abs(got[0] - m) < 0.01 and abs(got[1] - lo) < 0.01 and abs(got[2] - hi) < 0.012, (k, got)


def test_sharding_invariance_and_determinism():
    meas, err = synthetic_table(64, config=4)
    a = run_prod(meas, err, 2200, samples=False)
    b = run_prod(meas, err, 2200, samples=False)
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y, equal_nan=True)
    lo = run_prod(meas[:20], err[:20], 2200, goff=0, samples=False)
    hi = run_prod(meas[20:], err[20:], 2200, goff=20, samples=False)
    for i in range(4):
        assert np.array_equal(np.concatenate([lo[i], hi[i]]), a[i], equal_nan=True)


def test_flag_fallback_and_missing_lines():
    """Galaxies whose assumed has* flags are wrong (negative measurement with tiny error) or
    that miss lines entirely: present table must equal the oracle's on the kernel's own samples."""
    base_m, base_e = EX_M[0].copy(), EX_E[0].copy()
    rows_m, rows_e = [], []
    for knock in ([5], [1], [6], [0], [3], [7], [8], [7, 8], [6, 0]):
        m, e = base_m.copy(), base_e.copy()
        for j in knock:
            m[j] = np.nan
            e[j] = 0.0
        rows_m.append(m); rows_e.append(e)
    m, e = base_m.copy(), base_e.copy(); m[6] = -0.5; e[6] = 0.01     # [NII] never positive
    rows_m.append(m); rows_e.append(e)
    m, e = base_m.copy(), base_e.copy(); m[6] = -0.02; e[6] = 0.02    # [NII] positive for ~16 %
    rows_m.append(m); rows_e.append(e)
    m, e = base_m.copy(), base_e.copy(); m[0] = 0.0; e[0] = 0.0       # [OII] exactly zero
    rows_m.append(m); rows_e.append(e)
    meas, err = np.array(rows_m, dtype=np.float32), np.array(rows_e, dtype=np.float32)
    n = 600
    mds = "D02,Z94,M91,PP04,M08,M13,KD02"      # `all` minus P10 (see oracle/gen_golden.py)
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, n, mds, seed=77, goff=0)
    Zo, po, trips_o = oracle_on_kernel_samples(meas, err, n, 77, 0, mds, True, False)
    assert np.array_equal(status != sc.ST_ABSENT, po)
    assert np.array_equal(trips, trips_o)
    nan_mism = np.sum(np.isnan(Z) != np.isnan(Zo))
    assert nan_mism <= 3
    both = np.isfinite(Z) & np.isfinite(Zo)
    assert np.mean(np.abs(Z[both] - Zo[both]) > 1e-3) < 1e-3


def test_large_nsample_global_variant():
    """N' = 11000 (configs 3/5 shape) runs the global-scratch instantiation."""
    meas, err = synthetic_table(6, config=3)
    n = 11000
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, n, seed=5, goff=0)
    Zo, po, trips_o = oracle_on_kernel_samples(meas, err, n, 5, 0, "all", True, False)
    assert np.array_equal(status != sc.ST_ABSENT, po)
    same = np.all(trips == trips_o, axis=1)
    assert same.all()
    both = np.isfinite(Z) & np.isfinite(Zo)
    assert np.mean(np.abs(Z[both] - Zo[both]) > 1e-3) < 1e-4
    for g in range(meas.shape[0]):
        for j in range(sc.NKEYS):
            if status[g, j] == sc.ST_OK:
                fin = Z[g, j][np.isfinite(Z[g, j])]
                want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
                assert np.array_equal(pct[g, j], want), (g, sc.KEYS[j])


def test_deterministic_single_evaluation():
    """nsample == 1 (mcz.py:514-515): one evaluation at the measured fluxes."""
    pct, nvalid, status, trips, ret, Z = run_prod(EX_M, EX_E, 1, "KD02,Z94,PP04", deterministic=True)
    for g in range(2):
        flux = np.repeat(EX_M[g].astype(np.float64)[:, None], 1, axis=1)
        po, Zo, _, _, _ = oracle_galaxy(flux, "KD02,Z94,PP04", True)
        for j in range(sc.NKEYS):
            if po[j] and np.isfinite(Zo[j, 0]):
                assert abs(Z[g, j, 0] - Zo[j, 0]) < 1e-4, (g, sc.KEYS[j])


@pytest.mark.parametrize("ngal,n", [(50, 2200), (40000, 64)])     # one launch / three pipelined chunks
def test_host_buffer_entry_point(ngal, n):
    """mcz_ctx_run_host (the call run() makes): same tables as the device-pointer entry point,
    also when the table is split into chunks whose results are copied back while the next runs."""
    import ctypes
    from pymcz_b200 import _lib
    L = _lib.lib()
    meas, err = synthetic_table(ngal, config=4)
    tok = sc.T_ALL
    ref = run_prod(meas, err, n, samples=False)
    ctx = L.mcz_ctx_create(0)
    assert ctx
    G = meas.shape[0]
    pct = np.empty((G, sc.NKEYS, 3), np.float32); nvalid = np.empty((G, sc.NKEYS), np.int32)
    status = np.empty((G, sc.NKEYS), np.int8); trips = np.empty((G, 2), np.int32); ret = np.empty(G, np.int32)
    rc = L.mcz_ctx_run_host(ctx, meas.ctypes.data, err.ctypes.data, G, n, tok, sc.DUST_ON, 0xB200, 0, 0, 0,
                            pct.ctypes.data, nvalid.ctypes.data, status.ctypes.data, trips.ctypes.data,
                            ret.ctypes.data)
    assert rc == 0, L.mcz_last_error()
    L.mcz_ctx_destroy(ctx)
    assert np.array_equal(pct, ref[0], equal_nan=True)
    assert np.array_equal(nvalid, ref[1]) and np.array_equal(status, ref[2]) and np.array_equal(trips, ref[3])


@pytest.mark.parametrize("ngal,n,min_capped", [(360, 440, 10), (150, 2400, 4)])   # shared / global-scratch variant
def test_divergent_galaxies_match_reference_cap(ngal, n, min_capped):
    """The kernel ends an N2Ha loop early when it can PROVE non-convergence (elliptic Moebius
    maps: mean |dlogq| bounded below by more than the tolerance) and reports the reference's
    outcome: 100 trips, galaxy blanked (metscales.py:1119-1121).  Every such galaxy must be a
    100-trip galaxy in the oracle, and every 100-trip galaxy of the oracle must be capped here."""
    meas, err = synthetic_table(ngal, config=4)
    seed = 0xD1CE
    pct, nvalid, status, trips, ret, Z = run_prod(meas, err, n, "KD02", True, seed, 0, False)
    Zo, po, trips_o = oracle_on_kernel_samples(meas, err, n, seed, 0, "KD02", True, False)
    capped, capped_o = trips[:, 0] >= 100, trips_o[:, 0] >= 100
    assert capped_o.sum() >= min_capped              # the table does contain divergent galaxies
    assert np.array_equal(capped, capped_o), (np.nonzero(capped != capped_o), trips[capped != capped_o], trips_o[capped != capped_o])
    j = sc.KEY_INDEX["KK04_N2Ha"]
    assert np.all(np.isnan(Z[capped, j])) and np.all(status[capped, j] == sc.ST_EMPTY)
    # galaxies that are not capped keep the oracle's trip counts (float32 vs float64 decisions)
    same = trips[~capped] == trips_o[~capped]
    assert same.mean() > 0.97
