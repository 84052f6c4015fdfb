"""GPU parity, replay mode: the CUDA kernels consume the reference's own numpy-drawn flux samples
(through the C ABI) and are compared with the oracle sample by sample.

Bar (BASELINE.json): present table + NaN/validity masks bit-exact, values within 1e-3 dex
(float64 arithmetic actually agrees to ~1e-10); percentiles exact vs numpy.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mcz_oracle as orc
from pymcz_b200 import _scales as sc
from pymcz_b200 import ops
from pymcz_b200.synth import synthetic_table
from tests._helpers import oracle_galaxy, compare_samples

pytestmark = pytest.mark.gpu

EX_M = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
                 [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
EX_E = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
                 [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)


def oracle_table(meas, err, nsample, seed, mode, mds, dust_corr):
    """Oracle over a table with the legacy numpy stream; returns flux[11,G,N], noise[5,G,N], ..."""
    np.random.seed(seed)
    G = meas.shape[0]
    samples = orc.draw_samples(G, nsample)
    n = len(samples[0])
    flux = np.zeros((11, G, n))
    noise = np.zeros((5, G, n))
    Zo = np.zeros((G, sc.NKEYS, n))
    po = np.zeros((G, sc.NKEYS), dtype=bool)
    trips = np.zeros((G, 2), dtype=np.int32)
    rets = np.zeros(G, dtype=np.int32)
    for g in range(G):
        fl = orc.perturb(meas[g], err[g], samples[g], mode)
        flux[:, g] = fl
        po[g], Zo[g], noise[:, g], trips[g], rets[g] = oracle_galaxy(fl, mds, dust_corr)
    return flux, noise, Zo, po, trips, rets


def gpu_replay(flux, noise, mds, dust_corr, precision):
    tok, _, _ = sc.parse_md(mds)
    Z, present, trips, ret = ops.forward_replay(
        torch.from_numpy(flux).cuda(), torch.from_numpy(noise).cuda(), tok,
        sc.DUST_ON if dust_corr else sc.DUST_NONE, precision)
    torch.cuda.synchronize()
    return (Z.cpu().numpy(), ops.present_to_bool(present).cpu().numpy(), trips.cpu().numpy(),
            ret.cpu().numpy())


CASES = [("example_corr", EX_M, EX_E, 200, 200, "correlated", "all", True),
         ("example_shuf", EX_M, EX_E, 200, 201, "shuffled", "all", True),
         ("example_nodust", EX_M, EX_E, 200, 202, "correlated", "all", False),
         ("tokens", *synthetic_table(12, config=3), 100, 203, "shuffled",
          "P05,C01,P01,DP00,D16,M08all,KD02,D02,M13", True),
         ("kd02only", *synthetic_table(12, config=3), 100, 204, "correlated", "KD02", True),
         ("synth_all", *synthetic_table(96, config=4), 300, 205, "shuffled", "all", True)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_replay_f64_matches_oracle(case):
    _, meas, err, nsample, seed, mode, mds, dust = case
    flux, noise, Zo, po, trips_o, rets_o = oracle_table(meas, err, nsample, seed, mode, mds, dust)
    Z, pm, trips, ret = gpu_replay(flux, noise, mds, dust, "f64")
    assert np.array_equal(ret, rets_o)
    assert np.array_equal(trips, trips_o)
    mism, inf_mism, maxdiff, same_inf = compare_samples(Z, Zo, pm, po)
    assert mism == 0 and inf_mism == 0 and same_inf          # validity masks bit-exact
    assert maxdiff < 1e-8, maxdiff                           # (bar: 1e-3 dex)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_replay_f32_production_arithmetic(case):
    """The float32 (production) instantiation against the oracle on the SAME float32-representable
    samples (the reference's numpy draws rounded to float32, as the production sampler's are): trip
    counts identical, every galaxy included in the 1e-3 dex bar; a sample whose ratio sits within
    float rounding of a validity threshold may flip its mask (counted, bounded)."""
    from tests._helpers import oracle_on_samples
    _, meas, err, nsample, seed, mode, mds, dust = case
    flux, noise, _, _, _, rets_o = oracle_table(meas, err, nsample, seed, mode, mds, dust)
    flux = flux.astype(np.float32).astype(np.float64)
    noise = noise.astype(np.float32).astype(np.float64)
    Zo, po, trips_o, _ = oracle_on_samples(np.ascontiguousarray(flux.transpose(1, 0, 2)),
                                           np.ascontiguousarray(noise.transpose(1, 0, 2)), mds, dust, procs=1)
    Z, pm, trips, ret = gpu_replay(flux, noise, mds, dust, "f32")
    assert np.array_equal(ret, rets_o)
    assert np.array_equal(pm, po)
    assert np.array_equal(trips, trips_o)
    nan_mism = int(np.sum(np.isnan(Z) != np.isnan(Zo)))
    assert nan_mism <= 1e-5 * Z.size + 1, nan_mism
    both = np.isfinite(Z) & np.isfinite(Zo)
    assert np.max(np.abs(Z[both] - Zo[both])) < 1e-3
    print(case[0], "mask flips:", nan_mism, "of", Z.size, "max |dZ|:", float(np.max(np.abs(Z[both] - Zo[both]))))


def test_replay_golden_fixtures_from_reference(golden_dir):
    """Straight against the committed reference outputs (not only the oracle)."""
    g = np.load(os.path.join(golden_dir, "example_corr_n200.npz"))
    flux, noise, Zo, po, trips_o, _ = oracle_table(g["meas"], g["err"], int(g["nsample"]), int(g["seed"]),
                                                   str(g["mode"]), str(g["mds"]), bool(g["dust_corr"]))
    assert np.array_equal(Zo, g["Z"], equal_nan=True)        # oracle == reference on this fixture
    Z, pm, trips, ret = gpu_replay(flux, noise, "all", True, "f64")
    assert np.array_equal(pm, g["present"])
    assert np.array_equal(trips, g["trips"])
    assert np.array_equal(np.isnan(Z), np.isnan(g["Z"]))
    fin = np.isfinite(g["Z"])
    assert np.max(np.abs(Z[fin] - g["Z"][fin])) < 1e-8


def test_replay_dropnegatives(golden_dir):
    """DROPNEGATIVES (metallicity.py:23-26) as the T_DROPNEG flag of the token mask: the replay kernel on
    the reference's own samples against the reference's outputs with the switch on; and production
    mode against the oracle on the kernel's own samples."""
    g = np.load(os.path.join(golden_dir, "dropneg_n60.npz"))
    meas, err, mds = g["meas"], g["err"], str(g["mds"])
    tok, _, _ = sc.parse_md(mds)
    assert "D02" in mds.split(",") and "M13" in mds.split(",")
    # the reference's stream: samples, then per galaxy the shuffles of perturb() and the coefficient
    # noise of calculation(), interleaved -- replayed through the oracle with a recording normal()
    np.random.seed(int(g["seed"]))
    samples = orc.draw_samples(meas.shape[0], int(g["nsample"]))
    n = len(samples[0])
    flux = np.zeros((11, meas.shape[0], n))
    noise = np.zeros((5, meas.shape[0], n))
    Zo = np.zeros((meas.shape[0], sc.NKEYS, n))
    for i in range(meas.shape[0]):
        flux[:, i] = orc.perturb(meas[i], err[i], samples[i], str(g["mode"]))
        rec = []

        def normal(loc, scale, size, rec=rec):
            v = np.random.normal(loc, scale, size)
            rec.append((scale, v))
            return v
        d, ret, _ = orc.calculation(flux[:, i].copy(), mds, True, False, "batched", normal, dropnegatives=True)
        rows = {0.05: 0, 0.1: 1, 0.027: 2, 0.024: 3}
        seen12 = 0
        for scale, v in rec:
            if scale in rows:
                noise[rows[scale], i] = v
            elif seen12 == 0:
                noise[4, i] = v
                seen12 = 1
        for j, k in enumerate(orc.KEYS):
            Zo[i, j] = d.mds[k] if d.mds[k] is not None else np.nan
    assert np.array_equal(Zo, g["Z"], equal_nan=True)            # oracle == reference on this fixture
    Z, present, trips, ret = ops.forward_replay(torch.from_numpy(flux).cuda(), torch.from_numpy(noise).cuda(),
                                                tok | sc.T_DROPNEG, sc.DUST_ON, "f64")
    torch.cuda.synchronize()
    Z = Z.cpu().numpy()
    assert np.array_equal(ops.present_to_bool(present).cpu().numpy(), g["present"])
    assert np.array_equal(trips.cpu().numpy(), g["trips"])
    assert np.array_equal(np.isnan(Z), np.isnan(g["Z"]))
    fin = np.isfinite(g["Z"])
    assert np.max(np.abs(Z[fin] - g["Z"][fin])) < 1e-8
    # without the flag the same samples give other masks (the negatives are zeroed, not dropped)
    Z0 = ops.forward_replay(torch.from_numpy(flux).cuda(), torch.from_numpy(noise).cuda(), tok, sc.DUST_ON, "f64")[0]
    assert not np.array_equal(np.isnan(Z0.cpu().numpy()), np.isnan(g["Z"]))

    # production mode: the kernel's own Philox samples, the sampler's export carries the NaNs
    from tests._helpers import kernel_samples, oracle_on_samples
    seed, goff, N = 77, 3, 600
    out = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), N, tok | sc.T_DROPNEG,
                                 sc.DUST_ON, seed, goff, False, False, True)
    torch.cuda.synchronize()
    Zp = out[5].cpu().numpy()
    fl, nz = kernel_samples(meas, err, N, seed, goff, tok | sc.T_DROPNEG)
    assert np.isnan(fl[:, [0, 1, 3, 5, 6, 7, 8]]).any()          # dropped samples are there
    fl0 = np.where(np.isnan(fl), -1.0, fl)                       # the oracle drops them itself
    Zo2 = np.full(Zp.shape, np.nan)
    for i in range(meas.shape[0]):
        from tests._helpers import FixedNoise
        d, _, _ = orc.calculation(fl0[i].copy(), mds, True, False, "batched", FixedNoise(nz[i]), dropnegatives=True)
        for j, k in enumerate(orc.KEYS):
            if d.mds[k] is not None:
                Zo2[i, j] = d.mds[k]
    it = [sc.KEY_INDEX[k] for k in ("KK04_N2Ha", "KK04_R23", "KD02comb")]
    rest = [j for j in range(sc.NKEYS) if j not in it]
    flips = np.isnan(Zp[:, rest]) != np.isnan(Zo2[:, rest])
    assert flips.sum() <= 2, flips.sum()
    both = np.isfinite(Zp[:, rest]) & np.isfinite(Zo2[:, rest])
    assert np.max(np.abs(Zp[:, rest][both] - Zo2[:, rest][both])) < 1e-3


def test_replay_missing_lines(golden_dir):
    g = np.load(os.path.join(golden_dir, "missing_lines_n50.npz"))
    for i in range(g["meas"].shape[0]):
        mds = str(g["mds"][i])
        flux, noise, Zo, po, trips_o, rets_o = oracle_table(g["meas"][i:i + 1], g["err"][i:i + 1],
                                                            int(g["nsample"]), int(g["seed"][i]),
                                                            "correlated", mds, True)
        assert np.array_equal(Zo[0], g["Z"][i], equal_nan=True)
        Z, pm, trips, ret = gpu_replay(flux, noise, mds, True, "f64")
        assert ret[0] == g["ret"][i]
        assert np.array_equal(pm[0], g["present"][i]), i
        assert np.array_equal(trips[0], g["trips"][i])
        mism, inf_mism, maxdiff, same_inf = compare_samples(Z, Zo, pm, po)
        assert mism == 0 and inf_mism == 0 and same_inf and maxdiff < 1e-8, (i, mism, maxdiff)


@pytest.mark.parametrize("mode", ["correlated", "shuffled"])
def test_config2_exampledata_n20000(mode):
    """BASELINE config 2: exampledata, nsample=20000 (N'=22000), all scales, replay check."""
    flux, noise, Zo, po, trips_o, rets_o = oracle_table(EX_M, EX_E, 20000, 20000, mode, "all", True)
    Z, pm, trips, ret = gpu_replay(flux, noise, "all", True, "f64")
    assert np.array_equal(trips, trips_o)
    mism, inf_mism, maxdiff, same_inf = compare_samples(Z, Zo, pm, po)
    assert mism == 0 and inf_mism == 0 and same_inf and maxdiff < 1e-8
    # percentiles of the replayed distributions, exact vs numpy (mcz.py:273-301)
    pct, nvalid, status = ops.segmented_percentiles(torch.from_numpy(Z).cuda(), torch.from_numpy(pm).cuda())
    pct, nvalid, status = pct.cpu().numpy(), nvalid.cpu().numpy(), status.cpu().numpy()
    for g in range(2):
        for j, k in enumerate(sc.KEYS):
            if not po[g, j]:
                assert status[g, j] == sc.ST_ABSENT
                continue
            st, med, p16, p84, n = orc.savehist_stats(Zo[g, j])
            want = st if st != orc.ST_NODIST else orc.ST_OK
            assert status[g, j] == want, (g, k)
            assert nvalid[g, j] == n
            if st in (orc.ST_OK, orc.ST_NODIST):
                assert np.allclose(pct[g, j], [med, p16, p84], rtol=0, atol=1e-8), (g, k)


def test_percentiles_edge_cases():
    rng = np.random.default_rng(5)
    rows = []
    rows.append(np.full(777, np.nan))                               # empty
    rows.append(np.where(rng.random(777) < 0.5, np.nan, -rng.random(777)))   # negative sum
    rows.append(np.full(777, 8.69))                                 # no distribution
    r = rng.normal(8.7, 0.1, 777); r[::7] = np.inf; r[1::7] = -np.inf; r[2::7] = np.nan
    rows.append(r)
    r = rng.normal(8.7, 0.1, 777); r[5:] = np.nan
    rows.append(r)                                                  # n = 5
    r = np.full(777, np.nan); r[3] = 8.1
    rows.append(r)                                                  # n = 1
    r = np.round(rng.normal(8.7, 0.1, 777), 1)
    rows.append(r)                                                  # heavy ties
    Z = np.array(rows)
    pct, nvalid, status = ops.segmented_percentiles(torch.from_numpy(Z).cuda())
    pct, nvalid, status = pct.cpu().numpy(), nvalid.cpu().numpy(), status.cpu().numpy()
    for i, row in enumerate(rows):
        st, med, p16, p84, n = orc.savehist_stats(row)
        want = st if st != orc.ST_NODIST else orc.ST_OK
        assert status[i] == want, i
        assert nvalid[i] == n
        if want == orc.ST_OK:
            assert np.array_equal(pct[i], [med, p16, p84]), (i, pct[i], med, p16, p84)
