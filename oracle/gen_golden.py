"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

Test infrastructure, not product.  ``/root/reference`` exists only in the build container, so
this script is run there once and its outputs are committed; nothing at test/bench time reads
``/root/reference``.  Usage:  python oracle/gen_golden.py [outdir]

What is executed from the reference, unmodified:
  * ``pyMCZ/metallicity.py::calculation`` and ``pyMCZ/metscales.py::diagnostics`` -- imported
    as-is; ``measured`` is a dict subclass providing ``iterkeys()`` (the only py2-ism on the
    path, metallicity.py:94).
  * ``pyMCZ/mcz.py::savehist`` (+ getknuth/knuthn/getbinsize) -- mcz.py itself is py2-only syntax
    (mcz.py:432, 470), so the py3-valid line ranges 90-116, 227-239, 261-429 are exec'd verbatim
    from the file with NOPLOT=True and BINMODE='t' (binning is not on the path; 't' avoids the
    slow Knuth optimiser without touching the percentile lines 273-301).
The sampling lines of ``run()``/``calc()`` (mcz.py:436-442, 509-519, 580-589) cannot be imported
(py2 tuple parameters); they are five lines and are restated in oracle/mcz_oracle.py::perturb.
"""
import io
import os
import sys

import numpy as np

REF = "/root/reference/pyMCZ"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

import metallicity as ref_met          # noqa: E402  (the reference, unmodified)
import metscales as ref_ms             # noqa: E402

from oracle import mcz_oracle as orc   # noqa: E402
from pymcz_b200.synth import synthetic_table  # noqa: E402


class PyTwoDict(dict):
    def iterkeys(self):
        return iter(list(self.keys()))


def _exec_savehist():
    src = open(os.path.join(REF, "mcz.py")).read().split("\n")
    chunk = "\n".join(src[89:116] + src[226:239] + src[260:429])
    g = {"np": np, "os": os, "NOPLOT": True, "BINMODE": "t", "VERBOSE": False,
         "__name__": "mcz_excerpt"}
    import scipy.stats as stats
    from scipy.special import gammaln
    from scipy import optimize
    g.update(stats=stats, gammaln=gammaln, optimize=optimize)
    exec(compile(chunk, "mcz_excerpt", "exec"), g)
    return g["savehist"]


ref_savehist = _exec_savehist()


def ref_calculation(flux, mds, dust_corr, nps=2):
    """Run the reference on one galaxy's [11,N] flux samples. Returns (mds dict, ret, trips)."""
    n = flux.shape[1]
    measured = PyTwoDict()
    for j, name in enumerate(orc.LINES):
        measured[name] = flux[j].copy()
    sc = ref_ms.diagnostics(n, io.StringIO(), nps)
    calls = []
    orig = sc.calclogq

    def counting(Z):
        calls.append(1)
        return orig(Z)
    sc.calclogq = counting
    marks = {}
    o1, o2 = sc.calcKK04_N2Ha, sc.calcKK04_R23

    def w1():
        a = len(calls); r = o1(); marks['n2ha'] = len(calls) - a; return r

    def w2():
        a = len(calls); r = o2(); marks['r23'] = len(calls) - a; return r
    sc.calcKK04_N2Ha, sc.calcKK04_R23 = w1, w2
    so = sys.stdout
    sys.stdout = io.StringIO()
    try:
        with np.errstate(all='ignore'):
            ret = ref_met.calculation(sc, measured, 2, mds, nps, io.StringIO(), dust_corr=dust_corr)
    finally:
        sys.stdout = so
    return sc.mds, ret, (marks.get('n2ha', 0), marks.get('r23', 0))


def pack(mdsdict, n):
    present = np.array([mdsdict[k] is not None for k in orc.KEYS])
    Z = np.full((len(orc.KEYS), n), np.nan)
    for j, k in enumerate(orc.KEYS):
        if mdsdict[k] is not None:
            Z[j] = mdsdict[k]
    return present, Z


EXAMPLE_MEAS = np.array([
    [2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
    [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
EXAMPLE_ERR = np.array([
    [0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
    [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)


def check_example_file():
    """exampledata as parsed float32 (mcz.py:150) must equal the constants above."""
    for name, arr in (("meas", EXAMPLE_MEAS), ("err", EXAMPLE_ERR)):
        raw = np.loadtxt("/root/reference/input/exampledata_%s.txt" % name, comments=';', dtype=np.float32)
        assert np.array_equal(raw[:, 1:], arr, equal_nan=True), name


def run_case(meas, err, nsample, seed, mode, mds, dust_corr, reset_ignoredust=True, dropnegatives=False):
    """Per-sample golden: reference run over a table with seeded legacy numpy RNG.
    dropnegatives: the reference's module switch DROPNEGATIVES (metallicity.py:23-26) for this case."""
    if reset_ignoredust:
        ref_met.IGNOREDUST = False
    ref_met.DROPNEGATIVES = bool(dropnegatives)
    ref_met.FIXNEGATIVES = not dropnegatives                               # metallicity.py:25-26
    np.random.seed(seed)
    G = meas.shape[0]
    samples = orc.draw_samples(G, nsample)
    n = len(samples[0])
    present = np.zeros((G, len(orc.KEYS)), dtype=bool)
    Z = np.full((G, len(orc.KEYS), n), np.nan)
    rets = np.zeros(G, dtype=np.int32)
    trips = np.zeros((G, 2), dtype=np.int32)
    for g in range(G):
        flux = orc.perturb(meas[g], err[g], samples[g], mode)
        m, ret, tr = ref_calculation(flux, mds, dust_corr)
        present[g], Z[g] = pack(m, n)
        rets[g] = -1 if ret == -1 else 0
        trips[g] = tr
    ref_met.DROPNEGATIVES, ref_met.FIXNEGATIVES = False, True
    return dict(meas=meas, err=err, nsample=nsample, seed=seed, mode=mode, mds=mds,
                dust_corr=dust_corr, present=present, Z=Z, ret=rets, trips=trips)


def run_pct_case(meas, err, nsample, seed, mode, mds, dust_corr):
    """Percentile golden: reference calculation + reference savehist strings."""
    ref_met.IGNOREDUST = False
    np.random.seed(seed)
    G = meas.shape[0]
    samples = orc.draw_samples(G, nsample)
    strings = np.empty((G, len(orc.KEYS)), dtype=object)
    nfin = np.zeros((G, len(orc.KEYS)), dtype=np.int32)
    trips = np.zeros((G, 2), dtype=np.int32)
    so = sys.stdout
    for g in range(G):
        flux = orc.perturb(meas[g], err[g], samples[g], mode)
        m, ret, tr = ref_calculation(flux, mds, dust_corr)
        trips[g] = tr
        for j, k in enumerate(orc.KEYS):
            if m[k] is None:
                strings[g, j] = "ABSENT"
                continue
            sys.stdout = io.StringIO()
            try:
                s, data, _ = ref_savehist(np.asarray(m[k], dtype=float), "x", k, nsample, g, "/tmp", G, ["a"] * G)
            finally:
                sys.stdout = so
            strings[g, j] = s
            nfin[g, j] = len(data)
    return dict(meas=meas, err=err, nsample=nsample, seed=seed, mode=mode, mds=mds,
                dust_corr=dust_corr, strings=strings.astype(str), nfinite=nfin, trips=trips)


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    check_example_file()
    save = lambda name, d: np.savez_compressed(os.path.join(outdir, name + ".npz"), **d)

    # A-C: exampledata, N'=220, both sampling modes, dust on/off
    save("example_corr_n200", run_case(EXAMPLE_MEAS, EXAMPLE_ERR, 200, 200, "correlated", "all", True))
    save("example_shuf_n200", run_case(EXAMPLE_MEAS, EXAMPLE_ERR, 200, 201, "shuffled", "all", True))
    save("example_nodust_n200", run_case(EXAMPLE_MEAS, EXAMPLE_ERR, 200, 202, "correlated", "all", False))

    # D: explicit tokens (the scales outside `all`), on galaxies with every line present
    m, e = synthetic_table(12, config=3)
    save("tokens_n100", run_case(m, e, 100, 203, "shuffled",
                                 "P05,C01,P01,DP00,D16,M08all,KD02,D02,M13", True))
    save("kd02only_n100", run_case(m, e, 100, 204, "correlated", "KD02", True))

    # E: synthetic table slice incl. divergent (100-trip) galaxies
    m, e = synthetic_table(36, config=4)
    save("synth_all_n100", run_case(m, e, 100, 205, "shuffled", "all", True))

    # F: missing-line variants of exampledata row 1 (+ one low-SNR row that clips samples to 0)
    base_m, base_e = EXAMPLE_MEAS[0].copy(), EXAMPLE_ERR[0].copy()
    rows_m, rows_e = [], []
    for knock in ([orc.L_HA], [orc.L_HB], [orc.L_N2], [orc.L_O2], [orc.L_O3], [orc.L_S2A],
                  [orc.L_S2B], [orc.L_S2A, orc.L_S2B], [orc.L_N2, orc.L_O2], [orc.L_N2, orc.L_S2A]):
        mm, ee = base_m.copy(), base_e.copy()
        for j in knock:
            mm[j] = np.nan
            ee[j] = 0.0
        rows_m.append(mm)
        rows_e.append(ee)
    mm, ee = base_m.copy(), base_e.copy()
    ee[:] = mm * 0.6                      # SNR 1.7: many samples clip to zero
    ee[np.isnan(ee)] = 0.0
    rows_m.append(mm)
    rows_e.append(ee)
    mm, ee = base_m.copy(), base_e.copy()
    mm[orc.L_N2] = -0.05                  # negative measurement, mostly clipped
    ee[orc.L_N2] = 0.03
    rows_m.append(mm)
    rows_e.append(ee)
    rows_m, rows_e = np.array(rows_m, dtype=np.float32), np.array(rows_e, dtype=np.float32)
    # each row with a fresh IGNOREDUST (the non-sticky semantics the product implements) ...
    # rows without [OII] or [OIII] (but with Hb and [NII]) make the reference's calcP10 build a
    # ragged object array (metscales.py:815), which numpy >= 1.24 refuses: run those rows with
    # every `all` scale except P10 as explicit tokens (same call order => same RNG consumption).
    NOP10 = "D02,Z94,M91,PP04,M08,M13,KD02"
    mds_rows = [NOP10 if (np.isnan(rows_m[i, orc.L_O2]) or np.isnan(rows_m[i, orc.L_O3])) else "all"
                for i in range(len(rows_m))]
    outs = [run_case(rows_m[i:i + 1], rows_e[i:i + 1], 50, 300 + i, "correlated", mds_rows[i], True)
            for i in range(len(rows_m))]
    merged = dict(outs[0])
    for key in ("present", "Z", "ret", "trips"):
        merged[key] = np.concatenate([o[key] for o in outs])
    merged["meas"], merged["err"] = rows_m, rows_e
    merged["seed"] = np.array([300 + i for i in range(len(rows_m))])
    merged["mds"] = np.array(mds_rows)
    save("missing_lines_n50", merged)
    # ... and the first two rows in one run, so the sticky global (metallicity.py:127) is pinned too
    save("missing_lines_sticky_n50", run_case(rows_m[:3], rows_e[:3], 50, 320, "correlated", "all", True))

    # F2: DROPNEGATIVES = True (metallicity.py:23-26): negative samples become NaN instead of 0 -- the
    # low-SNR row, the negative-[NII] row and exampledata row 1, P10 left out (ragged arrays, see above)
    save("dropneg_n60", run_case(np.concatenate([rows_m[10:12], EXAMPLE_MEAS[:1]]),
                                 np.concatenate([rows_e[10:12], EXAMPLE_ERR[:1]]), 60, 330, "shuffled",
                                 NOP10, True, dropnegatives=True))

    # G: percentile strings from the reference's own savehist
    save("pct_example_n2000", run_pct_case(EXAMPLE_MEAS, EXAMPLE_ERR, 2000, 2000, "correlated", "all", True))
    save("pct_example_n20000", run_pct_case(EXAMPLE_MEAS, EXAMPLE_ERR, 20000, 20000, "correlated", "all", True))
    m, e = synthetic_table(9, config=4)
    save("pct_synth_n400", run_pct_case(m, e, 400, 206, "shuffled", "all", True))
    print("golden written to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "tests", "golden"))
