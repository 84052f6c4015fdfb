"""The reference's Python surface on top of the kernels: these tests read like a session with the
reference (build the `measured` dict as mcz.calc() does, call metallicity.calculation on a
metscales.diagnostics object, look at `.mds`) and compare with the oracle run on the same seed."""
import io
import os
import pickle
import shutil

import numpy as np
import pytest

from oracle import mcz_oracle as orc
from pymcz_b200 import _scales as sc
from pymcz_b200 import mcz, metallicity
from pymcz_b200 import metscales as ms
from pymcz_b200.synth import synthetic_table

pytestmark = pytest.mark.gpu

EX_M = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
                 [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
EX_E = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
                 [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)


def measured_dict(meas_row, err_row, sample):
    """fluxi of mcz.calc() (mcz.py:436-442)."""
    return {k: meas_row[j] * np.ones(len(sample)) + err_row[j] * sample for j, k in enumerate(sc.LINES)}


@pytest.mark.parametrize("mds,dust", [("all", True), ("all", False), ("KD02", True), ("D02,M13,PP04,P05,C01,D16", True), ("M08all,Z94", True)])
def test_calculation_matches_oracle_same_seed(mds, dust):
    sm, se = synthetic_table(6, config=4)
    meas, err = np.concatenate([EX_M, sm]), np.concatenate([EX_E, se])
    n = 330
    for g in range(meas.shape[0]):
        np.random.seed(100 + g)
        sample = np.random.normal(0, 1, n)
        fluxi = measured_dict(meas[g], err[g], sample)
        scales = ms.diagnostics(n, io.StringIO(), 2)
        metallicity.IGNOREDUST = False
        state = np.random.get_state()
        ret = metallicity.calculation(scales, fluxi, meas.shape[0], mds, 2, io.StringIO(), dust_corr=dust)
        # the oracle, restarted from the same RNG state, consumes the same coefficient-noise draws
        np.random.set_state(state)
        flux = np.array([meas[g, j] * np.ones(n) + err[g, j] * sample for j in range(11)])
        d, ret_o, _ = orc.calculation(flux, mds, dust, False, "batched")
        assert (ret == -1) == (ret_o == -1)
        assert scales.trips == (d.trips_n2ha, d.trips_r23)
        for k in sc.KEYS:
            a, b = scales.mds[k], d.mds[k]
            assert (a is None) == (b is None), (g, k)
            if a is None:
                continue
            assert np.array_equal(np.isnan(a), np.isnan(b)), (g, k)          # validity masks bit-exact
            fin = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), fin)
            if fin.any():
                assert np.max(np.abs(a[fin] - b[fin])) < 1e-8, (g, k)       # bar: 1e-3 dex
        # calculation() sanitised the caller's arrays in place, like the reference (metallicity.py:96-99)
        assert all(np.all(v >= 0) for v in fluxi.values())
        # the intermediate attributes the reference's always-run precompute leaves on the object
        # (calcNIIOII, calcNIISII, calcR23, initialguess: metscales.py:393-440, 468-489) and self.logq
        if ret != -1:
            for name in ("logN2O2", "logN2S2", "N2S2", "R2", "R3", "R23", "Z_init_guess"):
                a, b = getattr(scales, name), getattr(d, name, None)
                assert (a is None) == (b is None), (g, name)
                if a is not None:
                    assert np.allclose(a, b, rtol=1e-9, atol=1e-9, equal_nan=True), (g, name)
            if "KD02" in mds or "all" in mds:
                if scales.hasN2 and scales.hasHa and d.logq is not None and np.ndim(d.logq):
                    assert np.allclose(scales.logq, d.logq, rtol=1e-8, atol=1e-8, equal_nan=True), g
            if scales.hasN2O2:
                # the real roots of the [NII]/[OII] quartic, as np.roots lists them (descending)
                coef = np.array(orc.N2O2_COEF, dtype=np.float64)
                for i in (0, n // 2, n - 1):
                    c = coef.copy(); c[0] -= d.logN2O2[i]
                    r = np.roots(c[::-1]) if np.isfinite(c[0]) else np.array([])
                    real = np.sort(r[np.abs(r.imag) == 0].real)[::-1]
                    got = scales.N2O2_roots[i]
                    if len(real) == 2 and np.all(np.isfinite(got)):
                        # (the closed form is fitted for roots inside the KD02_N2O2 window [7.5, 9.4]; a
                        # partner root outside it is good to ~1e-5)
                        tol = np.where((real >= 7.5) & (real <= 9.4), 1e-7, 5e-5)
                        assert np.all(np.abs(got - real) <= tol), (g, i, got, real)


def test_calculation_dropnegatives_switch():
    """metallicity.DROPNEGATIVES = True (reference metallicity.py:23-26, 101-102): calculation() turns the
    negative samples of the caller's arrays into NaN and the scales are NaN for those samples."""
    n = 400
    meas, err = EX_M[0].copy(), EX_E[0].copy()
    err[:] = np.where(np.isnan(meas), 0.0, np.abs(meas) * 0.6)            # SNR 1.7: many negative samples
    np.random.seed(7)
    sample = np.random.normal(0, 1, n)
    mds = "D02,Z94,M91,PP04,M08,M13,KD02"
    old = metallicity.DROPNEGATIVES
    try:
        metallicity.DROPNEGATIVES = True
        metallicity.IGNOREDUST = False
        fluxi = measured_dict(meas, err, sample)
        neg = {k: v < 0 for k, v in fluxi.items()}
        scales = ms.diagnostics(n, io.StringIO(), 2)
        state = np.random.get_state()
        metallicity.calculation(scales, fluxi, 1, mds, 2, io.StringIO(), dust_corr=True)
        for k in ("Ha", "Hb", "[NII]6584", "[OII]3727"):
            assert neg[k].any() and np.all(np.isnan(fluxi[k][neg[k]]))    # dropped in place, like the reference
        np.random.set_state(state)
        flux = np.array([meas[j] * np.ones(n) + err[j] * sample for j in range(11)])
        d, _, _ = orc.calculation(flux, mds, True, False, "batched", dropnegatives=True)
        assert scales.trips == (d.trips_n2ha, d.trips_r23)
        for k in sc.KEYS:
            a, b = scales.mds[k], d.mds[k]
            assert (a is None) == (b is None), k
            if a is None:
                continue
            assert np.array_equal(np.isnan(a), np.isnan(b)), k
            fin = np.isfinite(b)
            if fin.any():
                assert np.max(np.abs(a[fin] - b[fin])) < 1e-8, k
        assert np.isnan(scales.mds["E(B-V)"][neg["Ha"] | neg["Hb"]]).all()
    finally:
        metallicity.DROPNEGATIVES = old
        metallicity.IGNOREDUST = False


def test_calculation_log_lines_and_warnings():
    """The log the reference's calculation() writes when nps == 1 (printsafemulti, metallicity.py:71-78):
    "calculating X" per scale in dispatch order, and the KD02 / KK04 applicability warning of calcNIIOII
    (metscales.py:405-414) for a measurement whose mean log([NII]/[OII]) is below -1.2."""
    n = 200
    np.random.seed(3)
    sample = np.random.normal(0, 1, n)
    low_m, low_e = synthetic_table(3, config=4)                 # row 2: the low-metallicity template
    for row, expect_warning in ((0, False), (2, True)):
        fluxi = measured_dict(low_m[row], low_e[row], sample)
        log = io.StringIO()
        scales = ms.diagnostics(n, log, 1)
        metallicity.IGNOREDUST = False
        metallicity.calculation(scales, fluxi, 3, "all", 1, log)
        text = log.getvalue()
        want = ["calculating E(B-V)", "calculating R23", "calculating D02", "calculating Z94", "calculating M91",
                "calculating PP04", "calculating P10", "calculating M08", "calculating M13", "calculating KD02_N2O2",
                "calculating KK04_N2Ha", "calculating KK04_R23", "calculating KD_combined"]
        pos = [text.find(w) for w in want]
        assert all(p >= 0 for p in pos) and pos == sorted(pos), text
        assert ("should only be used for  log([NII]6564/[OII]3727) > -1.2" in text) == expect_warning
        assert (np.mean(scales.logN2O2) < -1.2) == expect_warning


def test_individual_calc_methods():
    n = 200
    np.random.seed(5)
    sample = np.random.normal(0, 1, n)
    f = measured_dict(EX_M[1], EX_E[1], sample)
    for v in f.values():
        v[~np.isfinite(v)] = 0.0
        v[v < 0] = 0.0
    s = ms.diagnostics(n, io.StringIO(), 2)
    s.setHab(f['Ha'], f['Hb'])
    s.setOlines(f['[OII]3727'], f['[OIII]5007'], f['[OI]6300'], f['[OIII]5007'] / 3.)
    s.setSII(f['[SII]6717'], f['[SII]6731'], f['[SIII]9069'], f['[SIII]9532'])
    s.setNII(f['[NII]6584'])
    assert s.hasHa and s.hasN2 and s.hasO3O2 and not s.hasS39069
    s.calcEB_V()
    s.calcR23()
    s.calcZ94()
    s.calcPP04()
    s.calcKDcombined()
    flux = np.array([f[k] for k in sc.LINES])
    d, _, _ = orc.calculation(flux.copy(), "Z94,PP04,KD02", True, False, "batched")
    for k in ("E(B-V)", "logR23", "Z94", "PP04_N2Ha", "PP04_O3N2", "KD02_N2O2", "KK04_N2Ha", "KK04_R23", "M91", "KD02comb"):
        assert np.allclose(s.mds[k], d.mds[k], rtol=0, atol=1e-8, equal_nan=True), k
    assert s.mds["D02"] is None
    s.calcM08(allM08=True)                                    # metscales.py:969-1059
    d2, _, _ = orc.calculation(flux.copy(), "M08all", True, False, "batched")
    for k in ("M08_R23", "M08_N2Ha", "M08_O3O2", "M08_O3Hb", "M08_O2Hb", "M08_O3N2"):
        assert np.allclose(s.mds[k], d2.mds[k], rtol=0, atol=1e-8, equal_nan=True), k
    assert np.all(np.isnan(s.mds["M08_O3Hb"])) and np.isfinite(s.mds["M08_O3N2"]).any()
    assert s.calcpyqz() == -1 and s.calcPM14() == -1          # third-party scales: warn, stay None


def test_missing_balmer_lines_and_sticky_ignoredust(golden_dir):
    """metallicity.py:116-132: first galaxy without Ha gets E(B-V)=1e-5 *applied*, sets the
    sticky IGNOREDUST; the next one without Hb takes the 'no correction at all' branch.  Compared
    with the reference outputs committed in tests/golden (rows run in one session)."""
    g = np.load(os.path.join(golden_dir, "missing_lines_sticky_n50.npz"))
    np.random.seed(int(g["seed"]))
    G = g["meas"].shape[0]
    samples = orc.draw_samples(G, int(g["nsample"]))
    metallicity.IGNOREDUST = False
    for i in range(G):
        flux = orc.perturb(g["meas"][i], g["err"][i], samples[i], "correlated")
        fluxi = {k: flux[j].copy() for j, k in enumerate(sc.LINES)}
        scales = ms.diagnostics(flux.shape[1], io.StringIO(), 2)
        ret = metallicity.calculation(scales, fluxi, G, "all", 2, io.StringIO(), dust_corr=True)
        assert (ret == -1) == (g["ret"][i] == -1)
        for j, k in enumerate(sc.KEYS):
            assert (scales.mds[k] is not None) == bool(g["present"][i, j]), (i, k)
            if scales.mds[k] is not None:
                ref = g["Z"][i, j]
                assert np.array_equal(np.isnan(scales.mds[k]), np.isnan(ref)), (i, k)
                fin = np.isfinite(ref)
                if k in ("D02", "M13_N2", "M13_O3N2"):
                    continue          # these carry np.random coefficient noise drawn after `samples`
                assert np.max(np.abs(scales.mds[k][fin] - ref[fin]), initial=0.0) < 1e-8, (i, k)
    assert metallicity.IGNOREDUST is True
    metallicity.IGNOREDUST = False


def test_run_exampledata_products(golden_dir, tmp_path, capsys):
    """mcz.run on the reference's example input: table, ascii and pickle products (mcz.py:618-666),
    medians within Monte-Carlo error of an oracle run in the same sampling mode, --unpickle path."""
    root = tmp_path / "data"
    shutil.copytree(os.path.join(golden_dir, "input"), root / "input")
    mcz.ASCIIOUTPUT = True
    mcz.main(["exampledata", "2000", "--path", str(root), "--multiproc", "--asciiout", "--seed", "11"])
    out = capsys.readouterr().out
    assert "measurement  1/2" in out and "KD02comb" in out and "E(B-V)" in out
    binp = root / "output" / "exampledata"
    res = pickle.load(open(binp / "exampledata_n2000.pkl", "rb"))
    assert set(res) == set(sc.KEYS) and res["KD02comb"].shape == (2200, 2)
    assert np.all(np.isnan(res["D13_N2S2_O3S2"]))                  # absent scales are NaN columns (mcz.py:604-605)
    lines = open(binp / "exampledata_n2000_1.txt").read().splitlines()
    assert lines[0].startswith("exampledata\t Median Oxygen abundance")
    got = {l.split("\t")[0]: [float(x) for x in l.split("\t")[1:]] for l in lines[1:]}
    rs = np.random.RandomState(3)
    ora = orc.run_table(EX_M, EX_E, 2000, mode="correlated", rs=rs, normal=rs.normal)
    for j, k in enumerate(sc.KEYS):
        if ora["status"][0, j] in (orc.ST_OK, orc.ST_NODIST) and ora["nvalid"][0, j] > 1100:
            med, p16, p84 = ora["pct"][0, j]
            assert k in got, k
            tol = 6 * 1.2533 * max(0.5 * (p84 - p16), 1e-4) / np.sqrt(ora["nvalid"][0, j]) + 2e-3
            assert abs(got[k][0] - med) < tol, (k, got[k], med)
    # published annotations, correlated mode
    assert abs(got["KD02comb"][0] - 8.811) < 0.01 and abs(got["E(B-V)"][0] - 0.120) < 0.01
    first = mcz.last_result
    mcz.main(["exampledata", "2000", "--path", str(root), "--unpickle"])
    again = mcz.last_result
    ok = first["status"] == sc.ST_OK
    assert np.array_equal(again["status"] == sc.ST_OK, ok)
    assert np.allclose(again["pct"][ok], first["pct"][ok], rtol=0, atol=1e-6)   # f64 op on the pickled f32 samples
    mcz.ASCIIOUTPUT = False
