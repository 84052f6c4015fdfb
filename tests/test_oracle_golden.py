"""Pin the oracle (oracle/mcz_oracle.py) to vectors produced by the UNMODIFIED reference.

The .npz files under tests/golden were written by oracle/gen_golden.py, which imports
/root/reference/pyMCZ/{metallicity,metscales}.py as-is and exec's mcz.py's savehist.  Here the
oracle replays the same seeded legacy-numpy RNG stream and must reproduce every output
bit-for-bit (None-ness, NaN positions and float64 values), plus the iteration trip counts.
"""
import os

import numpy as np
import pytest

from oracle import mcz_oracle as orc

PER_SAMPLE = ["example_corr_n200", "example_shuf_n200", "example_nodust_n200", "tokens_n100",
              "kd02only_n100", "synth_all_n100", "missing_lines_sticky_n50"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _replay(meas, err, nsample, seed, mode, mds, dust_corr, roots_mode, dropnegatives=False):
    np.random.seed(int(seed))
    G = meas.shape[0]
    samples = orc.draw_samples(G, int(nsample))
    n = len(samples[0])
    present = np.zeros((G, len(orc.KEYS)), dtype=bool)
    Z = np.full((G, len(orc.KEYS), n), np.nan)
    rets = np.zeros(G, dtype=np.int32)
    trips = np.zeros((G, 2), dtype=np.int32)
    ign = False
    for g in range(G):
        flux = orc.perturb(meas[g], err[g], samples[g], str(mode))
        d, ret, ign = orc.calculation(flux, str(mds), bool(dust_corr), ign, roots_mode, dropnegatives=dropnegatives)
        rets[g] = -1 if ret == -1 else 0
        trips[g] = (d.trips_n2ha, d.trips_r23)
        for j, k in enumerate(orc.KEYS):
            if d.mds[k] is not None:
                present[g, j] = True
                Z[g, j] = d.mds[k]
    return present, Z, rets, trips


@pytest.mark.parametrize("name", PER_SAMPLE)
@pytest.mark.parametrize("roots_mode", ["loop", "batched"])
def test_oracle_matches_reference_bit_exact(golden_dir, name, roots_mode):
    g = _load(golden_dir, name)
    present, Z, rets, trips = _replay(g["meas"], g["err"], g["nsample"], g["seed"], g["mode"],
                                      g["mds"], g["dust_corr"], roots_mode)
    assert np.array_equal(present, g["present"])
    assert np.array_equal(rets, g["ret"])
    assert np.array_equal(trips, g["trips"])
    assert np.array_equal(np.isnan(Z), np.isnan(g["Z"]))
    # bit-exact, not allclose: same numpy ops in the same order
    assert np.array_equal(Z, g["Z"], equal_nan=True)


@pytest.mark.parametrize("roots_mode", ["loop", "batched"])
def test_oracle_dropnegatives_matches_reference(golden_dir, roots_mode):
    """The reference run with its module switch DROPNEGATIVES = True (metallicity.py:23-26): negative
    flux samples become NaN instead of 0 (low-SNR row, negative-[NII] row, exampledata row 1)."""
    g = _load(golden_dir, "dropneg_n60")
    present, Z, rets, trips = _replay(g["meas"], g["err"], g["nsample"], g["seed"], g["mode"],
                                      g["mds"], g["dust_corr"], roots_mode, dropnegatives=True)
    assert np.array_equal(present, g["present"])
    assert np.array_equal(rets, g["ret"])
    assert np.array_equal(trips, g["trips"])
    assert np.array_equal(Z, g["Z"], equal_nan=True)
    # and the switch matters on these rows: with FIXNEGATIVES the masks differ
    _, Z0, _, _ = _replay(g["meas"], g["err"], g["nsample"], g["seed"], g["mode"], g["mds"], g["dust_corr"], roots_mode)
    assert not np.array_equal(np.isnan(Z0), np.isnan(g["Z"]))


def test_oracle_missing_lines_rows(golden_dir):
    g = _load(golden_dir, "missing_lines_n50")
    for i in range(g["meas"].shape[0]):
        present, Z, rets, trips = _replay(g["meas"][i:i + 1], g["err"][i:i + 1], g["nsample"],
                                          g["seed"][i], g["mode"], g["mds"][i], g["dust_corr"], "batched")
        assert np.array_equal(present[0], g["present"][i]), i
        assert rets[0] == g["ret"][i]
        assert np.array_equal(trips[0], g["trips"][i])
        assert np.array_equal(Z[0], g["Z"][i], equal_nan=True), i


@pytest.mark.parametrize("name", ["pct_example_n2000", "pct_example_n20000", "pct_synth_n400"])
def test_oracle_percentile_strings(golden_dir, name):
    """savehist's result strings (mcz.py:280-301, 423) from the reference's own savehist."""
    g = _load(golden_dir, name)
    np.random.seed(int(g["seed"]))
    meas, err = g["meas"], g["err"]
    G = meas.shape[0]
    samples = orc.draw_samples(G, int(g["nsample"]))
    ign = False
    for gi in range(G):
        flux = orc.perturb(meas[gi], err[gi], samples[gi], str(g["mode"]))
        d, ret, ign = orc.calculation(flux, str(g["mds"]), bool(g["dust_corr"]), ign, "batched")
        assert (d.trips_n2ha, d.trips_r23) == tuple(g["trips"][gi])
        for j, k in enumerate(orc.KEYS):
            want = str(g["strings"][gi, j])
            if d.mds[k] is None:
                assert want == "ABSENT"
                continue
            st, med, p16, p84, n = orc.savehist_stats(d.mds[k])
            assert orc.savehist_string(st, med, p16, p84) == want, (gi, k)
            if st in (orc.ST_OK, orc.ST_NODIST):
                assert n == g["nfinite"][gi, j]


def test_published_triples_correlated_mode():
    """Coarse sanity vs the paper's figure annotations: exampledata
    measurement 1, N=2000, correlated sampling; one unseeded MC realisation each -> +-0.01."""
    meas = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan]], dtype=np.float32)
    err = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00]], dtype=np.float32)
    out = orc.run_table(meas, err, 2000, rs=np.random.RandomState(7), normal=np.random.RandomState(8).normal)
    want = {"E(B-V)": (0.120, 0.102, 0.137), "KD02comb": (8.811, 8.799, 8.822),
            "KK04_R23": (9.059, 9.058, 9.060), "M91": (8.859, 8.849, 8.867),
            "PP04_O3N2": (8.920, 8.881, 8.976)}
    for k, (m, lo, hi) in want.items():
        got = out["pct"][0, orc.KEY_INDEX[k]]
        assert abs(got[0] - m) < 0.01 and abs(got[1] - lo) < 0.01 and abs(got[2] - hi) < 0.012, (k, got)


def test_batched_roots_equal_loop():
    rng = np.random.default_rng(0)
    for name, c in (("n2o2", orc.N2O2_COEF), ("n2ha", orc.M08_COEFS['N2Ha']), ("o3o2", orc.M08_COEFS['O3O2'])):
        coef = np.array([c] * 400, dtype=float)
        y = rng.uniform(-3, 2, 400)
        y[::50] = np.nan
        y[1::50] = np.inf
        coef[:, 0] -= y
        a = orc.fz_roots(coef.copy(), "loop")
        b = orc.fz_roots(coef.copy(), "batched")
        assert np.array_equal(a, b), name
