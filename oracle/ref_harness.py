"""Runs the UNMODIFIED reference modules staged in oracle/_ref/pyMCZ (oracle/build_ref.py) on one
measurement, the way the reference's --multiproc worker does (TEST / BASELINE INFRASTRUCTURE ONLY).

What is executed from the reference, unmodified: metallicity.calculation and metscales.diagnostics
(imported as they are) and savehist's statistics (mcz.py is py2-only syntax at mcz.py:432, 470, so
its py3-valid line ranges 90-116, 227-239, 261-429 are exec'd verbatim with NOPLOT=True and
BINMODE='t': binning is not on the path and 't' avoids the Knuth optimiser without touching the
percentile lines 273-301).  What is restated here (five lines the py2 file cannot give us): the
Pool worker's perturbation, mcz.py:436-442, and its sample vector, mcz.py:509-519.
Only bench.py's CPU legs and tests/ may import this.
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref", "pyMCZ")

LINES = ['[OII]3727', 'Hb', '[OIII]4959', '[OIII]5007', '[OI]6300', 'Ha', '[NII]6584',
         '[SII]6717', '[SII]6731', '[SIII]9069', '[SIII]9532']                     # mcz.py:45


def available():
    return os.path.isfile(os.path.join(REFDIR, "metallicity.py")) and os.path.isfile(os.path.join(REFDIR, "metscales.py"))


_mods = None


class _PyTwoDict(dict):
    def iterkeys(self):                      # metallicity.py:94
        return iter(list(self.keys()))


def _load():
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref/pyMCZ is not staged: run `python oracle/build_ref.py` in the build container")
    sys.path.insert(0, REFDIR)
    try:
        import metallicity as ref_met
        import metscales as ref_ms
    finally:
        sys.path.remove(REFDIR)
    savehist = None
    mczpy = os.path.join(REFDIR, "mcz.py")
    if os.path.isfile(mczpy):
        src = open(mczpy).read().split("\n")
        chunk = "\n".join(src[89:116] + src[226:239] + src[260:429])
        import scipy.stats as stats
        from scipy.special import gammaln
        from scipy import optimize
        g = {"np": np, "os": os, "NOPLOT": True, "BINMODE": "t", "VERBOSE": False, "__name__": "mcz_excerpt",
             "stats": stats, "gammaln": gammaln, "optimize": optimize}
        exec(compile(chunk, "mcz_excerpt", "exec"), g)
        savehist = g["savehist"]
    _mods = (ref_met, ref_ms, savehist)
    return _mods


def run_measurement(meas_row, err_row, nsample, mds="all", dust_corr=True, rs=None, stats=True):
    """One measurement through the reference: N' = int(1.1 nsample) correlated perturbations
    (calc(), mcz.py:436-442), metallicity.calculation (per-sample np.roots loops and all), and the
    savehist statistics of every non-empty scale.  Returns (mds dict of arrays-or-None, ret, table)
    where table maps key -> savehist's "median  -err  +err" string."""
    ref_met, ref_ms, savehist = _load()
    rs = rs if rs is not None else np.random
    n = int(nsample + 0.1 * nsample) if nsample > 1 else nsample                # mcz.py:477-479
    sample = rs.normal(0, 1, n) if nsample > 1 else np.array([0.])               # mcz.py:514-517
    measured = _PyTwoDict()
    for j, name in enumerate(LINES):                                           # mcz.py:436-442 (no shuffle)
        measured[name] = float(meas_row[j]) + float(err_row[j]) * sample
    sc = ref_ms.diagnostics(n, io.StringIO(), 2)
    so = sys.stdout
    sys.stdout = io.StringIO()
    table = {}
    try:
        with np.errstate(all='ignore'):
            ret = ref_met.calculation(sc, measured, 2, mds, 2, io.StringIO(), dust_corr=dust_corr)
            if stats:
                for key in ref_met.get_keys():
                    v = sc.mds[key]
                    if v is None or np.sum(~np.isnan(v)) == 0:                   # mcz.py:604-607, 651-652
                        continue
                    if savehist is not None:
                        r = savehist(np.asarray(v, dtype=np.float64), "bench", key, nsample, 0, os.path.join(HERE, "_ref"),
                                     1, ["m"])                                   # mcz.py:261, called as at :653
                        table[key] = r[0]
                    else:
                        d = v[np.isfinite(v)]
                        table[key] = np.percentile(d, [50, 16, 84]) if d.size else None   # mcz.py:273, 288
    finally:
        sys.stdout = so
    return sc.mds, ret, table
