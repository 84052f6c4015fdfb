"""GPU parity of the post-path assists, through the C ABI: equal-width counts and
moments against np.histogram / np.std on np.linspace edges -- exactly what getknuth computes
(/root/reference/pyMCZ/mcz.py:90-99) -- Knuth's rule against a numpy restatement of knuthn
(mcz.py:102-116), and the KS statistic against scipy.stats.ks_2samp (testcompleteness.py:111)."""
import numpy as np
import pytest
import torch
from scipy import stats, optimize
from scipy.special import gammaln

from pymcz_b200 import ops, binning, _scales as sc
from pymcz_b200.synth import synthetic_table

pytestmark = pytest.mark.gpu


def columns():
    """Real output distributions of the production kernel plus adversarial rows."""
    meas, err = synthetic_table(6, config=4)
    out = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), 2200, sc.T_ALL,
                                 return_samples=True)
    Z = out[5].double().reshape(-1, 2200).cpu().numpy()
    rs = np.random.RandomState(3)
    extra = np.stack([
        np.full(2200, 8.5),                                   # constant: all linspace edges equal
        np.where(rs.rand(2200) < 0.5, 8.0, 9.0),              # two-point
        np.round(rs.normal(8.7, 0.1, 2200), 2),               # heavy ties on bin edges
        np.concatenate([rs.normal(8.7, 0.1, 2190), [1e30] * 10]),
        np.full(2200, np.nan),                                # empty
        np.concatenate([[1.0], np.full(2199, np.nan)]),       # a single value
        np.arange(2200) / 7.0,                                # values exactly on computed edges
    ])
    return np.concatenate([Z, extra])


def test_counts_equal_np_histogram_on_linspace_edges():
    Z = columns()
    ms = [1, 2, 3, 7, 26, 50, 64, 100, 233, 1000]
    st, counts = ops.histogram_assist(torch.from_numpy(Z).cuda(), ms)
    st = st.cpu().numpy()
    counts = [c.cpu().numpy() for c in counts]
    checked = 0
    for r in range(Z.shape[0]):
        data = Z[r][np.isfinite(Z[r])]
        assert st[r, 0] == data.size
        if data.size == 0:
            assert all(c[r].sum() == 0 for c in counts)
            continue
        assert st[r, 1] == data.min() and st[r, 2] == data.max()
        assert abs(st[r, 3] - data.mean()) <= 1e-12 * max(1.0, abs(data.mean()))
        assert abs(st[r, 4] - data.std()) <= 1e-10 * max(1.0, data.std())
        for m, c in zip(ms, counts):
            want, _ = np.histogram(data, np.linspace(min(data), max(data), m + 1))     # mcz.py:94-96
            assert np.array_equal(c[r], want), (r, m)
            checked += 1
    assert checked > 500


def _ref_knuthn(data):
    """knuthn / getknuth of mcz.py:90-116 restated with numpy (test-side checker)."""
    N = data.size

    def gk(m, data, N):
        m = int(m)
        if m > N:
            return [-1]
        bins = np.linspace(min(data), max(data), int(m) + 1)
        try:
            nk, bins = np.histogram(data, bins)
            return -(N * np.log(m) + gammaln(0.5 * m) - m * gammaln(0.5) - gammaln(N + 0.5 * m) + np.sum(gammaln(nk + 0.5)))
        except Exception:
            return [-1]
    maxM = 5 * np.sqrt(N)
    m0 = 2.0 * N ** (1. / 3.)
    mk = optimize.fmin(gk, m0, args=(data, N), disp=False, maxiter=30)[0]
    if mk > maxM or mk < 0.3 * np.sqrt(N):
        return m0, 't'
    return mk, 0


def test_knuth_rule_matches_reference_rule():
    Z = columns()[:74]
    tab = binning.KnuthTable(Z)
    n_checked = 0
    for r in range(Z.shape[0]):
        data = Z[r][np.isfinite(Z[r])]
        if data.size < 100 or data.min() == data.max():
            continue
        for m in (1, 5, 26, 100, tab.maxM):
            bins = np.linspace(min(data), max(data), m + 1)
            nk, _ = np.histogram(data, bins)
            want = -(data.size * np.log(m) + gammaln(0.5 * m) - m * gammaln(0.5) - gammaln(data.size + 0.5 * m) + np.sum(gammaln(nk + 0.5)))
            assert tab.getknuth(r, m) == want
        got, want = tab.knuthn(r), _ref_knuthn(data)
        assert got[1] == want[1] and got[0] == want[0], (r, got, want)
        n_checked += 1
    assert n_checked >= 30


def test_ks_statistic_equals_scipy():
    rs = np.random.RandomState(5)
    cases = [(rs.normal(0, 1, 2200), np.zeros(10)),                       # the diagnostic's first comparison
             (rs.normal(0, 1, 5500), rs.normal(0.05, 1, 2200)),
             (np.round(rs.normal(8.7, 0.1, 3000), 2), np.round(rs.normal(8.7, 0.1, 1500), 2)),   # ties
             (rs.normal(0, 1, 22000), rs.normal(0, 1.02, 16500)),
             (np.array([1.0]), np.array([2.0, 3.0]))]
    for a, b in cases:
        got = ops.ks_2samp(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
        want = stats.ks_2samp(a, b, method="asymp")            # (the raw statistic: 'exact' re-rounds it)
        assert got[0] == want.statistic, (got, want)
        D, p = binning.ks_2samp(a, b)                          # the call the reference makes (method 'auto')
        w = stats.ks_2samp(a, b)
        assert D == w.statistic and abs(p - w.pvalue) <= 1e-9 * w.pvalue + 1e-300, (D, p, w)
    # the nested sub-sample table of testcompleteness.py:87-113
    x = rs.normal(8.7, 0.1, 22000)
    subs = [np.sort(rs.choice(x, int(f * x.size), replace=False)) for f in (0.1, 0.25, 0.5, 0.75, 1)]
    tab = binning.ks_table(subs)
    for (D, p), (xa, xb) in zip(tab, zip(subs[1:], subs[:-1])):
        w = stats.ks_2samp(xa, xb)
        assert D == w.statistic and abs(p - w.pvalue) <= 1e-9 * w.pvalue + 1e-300
