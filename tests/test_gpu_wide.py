"""The cooperative "wide" kernel (one galaxy at a time, samples tiled across all thread blocks, for
single-object runs with very many samples) against the one-block-per-galaxy
kernels on the same inputs: same Philox indexing and device functions, so status, trip counts,
n and the percentile table must be identical whenever the trip counts agree (the KK04 means are
summed in a different order, which can move a borderline trip decision)."""
import os

import numpy as np
import pytest
import torch

from pymcz_b200 import _scales as sc
from pymcz_b200 import ops
from pymcz_b200.synth import synthetic_table

pytestmark = pytest.mark.gpu

EX_M = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
                 [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
EX_E = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
                 [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)


def run(meas, err, n, tok, wide, dust=sc.DUST_ON, correlated=False, deterministic=False, samples=False):
    old = os.environ.get("MCZ_WIDE")
    os.environ["MCZ_WIDE"] = "1" if wide else "0"
    try:
        out = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n, tok, dust,
                                     0xB200, 3, correlated, deterministic, samples)
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["MCZ_WIDE"]
        else:
            os.environ["MCZ_WIDE"] = old
    return [o.cpu().numpy() if o is not None else None for o in out]


def tables():
    sm, se = synthetic_table(30, config=4)
    meas, err = np.concatenate([EX_M, sm]), np.concatenate([EX_E, se])
    meas, err = meas.copy(), err.copy()
    meas[5, 1] = np.nan; err[5, 1] = np.nan          # no Hb: E(B-V) fallback, most scales None
    meas[6, [0, 3]] = np.nan                         # no [OII], [OIII]
    err[7] = 0.0                                     # every sample identical: constant columns
    err[8, 5] = 3.0 * meas[8, 5]                     # Ha with SNR 1/3: clipped samples, heavy ties
    return meas, err


@pytest.mark.parametrize("n,mds", [(2200, "all"), (129, "all"), (40000, "all"), (70001, "KD02"),
                                   (5000, "P05,C01,P01,DP00,D16,M08all,all"), (300000, "all")])
def test_wide_kernel_equals_per_galaxy_kernel(n, mds):
    meas, err = tables()
    if n >= 70000:
        meas, err = meas[:10], err[:10]
    tok = sc.parse_md(mds)[0]
    ref = run(meas, err, n, tok, wide=False)
    wid = run(meas, err, n, tok, wide=True)
    assert np.array_equal(ref[2], wid[2])                        # status
    assert np.array_equal(ref[4], wid[4])                        # ret
    same = np.all(ref[3] == wid[3], axis=1)
    assert same.mean() >= 0.9, (ref[3][~same], wid[3][~same])
    assert np.array_equal(ref[1][same], wid[1][same])            # n per scale
    a, b = ref[0][same], wid[0][same]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    # same per-sample device functions, but the per-galaxy kernel runs a copy of phase A specialised on
    # the launch's flags / tokens and the compiler fuses a few multiply-adds differently in the two
    # kernels: the tables agree to float32 rounding (both are checked against the oracle elsewhere)
    assert np.allclose(a, b, rtol=0, atol=6e-6, equal_nan=True), np.nanmax(np.abs(a - b))


@pytest.mark.parametrize("n,ngal,mds", [(40000, 24, "all"), (150000, 6, "all"), (33000, 40, "KD02")])
def test_wide_kernel_vs_oracle(n, ngal, mds):
    """The wide kernel against the ORACLE on the kernel's own samples (read back from the device):
    present table, trip counts, NaN masks and per-sample values to the same bar as the per-galaxy
    kernel (tests/_helpers.py::compare_with_oracle), and its percentile table against numpy on its
    own float32 columns."""
    from tests._helpers import kernel_samples, oracle_on_samples, compare_with_oracle
    meas, err = tables()
    sm, se = synthetic_table(max(ngal, 1), config=4)
    meas, err = np.concatenate([meas[:10], sm])[:ngal], np.concatenate([err[:10], se])[:ngal]
    tok = sc.parse_md(mds)[0]
    wid = run(meas, err, n, tok, wide=True, samples=True)
    pct, nvalid, status, trips, ret, Z = wid
    flux, noise = kernel_samples(meas, err, n, 0xB200, 3, tok)
    Zo, po, trips_o, _ = oracle_on_samples(flux, noise, mds, True)
    compare_with_oracle(Z, status, trips, Zo, po, trips_o, what="wide n=%d %s" % (n, mds))
    for g in range(meas.shape[0]):
        for j in range(sc.NKEYS):
            if status[g, j] == sc.ST_OK:
                fin = Z[g, j][np.isfinite(Z[g, j])]
                want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
                assert np.array_equal(pct[g, j], want), (g, sc.KEYS[j])


def test_wide_kernel_per_sample_output():
    """Z[G,37,N'] (pickle / --asciidistrib product) from the wide kernel equals the per-galaxy kernel's,
    bit for bit outside the KK04 iteration."""
    meas, err = tables()
    meas, err = meas[:12], err[:12]
    ref = run(meas, err, 5000, sc.T_ALL, wide=False, samples=True)
    wid = run(meas, err, 5000, sc.T_ALL, wide=True, samples=True)
    same = np.all(ref[3] == wid[3], axis=1)
    it = [sc.KEY_INDEX[k] for k in ("KK04_N2Ha", "KK04_R23", "KD02comb")]
    rest = [j for j in range(sc.NKEYS) if j not in it]
    assert np.array_equal(np.isnan(ref[5][:, rest]), np.isnan(wid[5][:, rest]))
    assert np.allclose(ref[5][:, rest], wid[5][:, rest], rtol=0, atol=4e-6, equal_nan=True)
    # (the iteration amplifies a last-bit difference of its inputs on nearly parabolic samples)
    d = np.abs(ref[5][same][:, it] - wid[5][same][:, it])
    assert np.nanmedian(d) < 1e-6 and np.nanmean(d > 1e-4) < 1e-4


def test_wide_kernel_is_chosen_for_single_object_runs():
    """One object, 2 x 10^5 samples: the dispatcher takes the wide kernel by itself; the percentiles
    agree with numpy on the per-sample output of the per-galaxy kernel."""
    meas, err = EX_M[:1].copy(), EX_E[:1].copy()
    n = 200000
    tok = sc.T_ALL
    auto = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n, tok, sc.DUST_ON, 7)
    torch.cuda.synchronize()
    pct = auto[0].cpu().numpy()
    Z = ops.forward_production(torch.from_numpy(meas).cuda(), torch.from_numpy(err).cuda(), n, tok, sc.DUST_ON, 7,
                               return_samples=True)[5].cpu().numpy()
    st = auto[2].cpu().numpy()
    for j in range(sc.NKEYS):
        if st[0, j] == sc.ST_OK:
            fin = Z[0, j][np.isfinite(Z[0, j])]
            want = np.percentile(fin.astype(np.float64), [50, 16, 84]).astype(np.float32)
            assert np.array_equal(pct[0, j], want), sc.KEYS[j]
