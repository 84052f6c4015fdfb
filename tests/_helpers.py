"""Shared test helpers (oracle-side; tests only)."""
import numpy as np

from oracle import mcz_oracle as orc


class RecordingNormal:
    """Stand-in for np.random.normal that records the coefficient-noise draws of calcD02 / calcM13
    (metscales.py:679-680, 948-949, 952-953) so they can be replayed into the CUDA kernels.
    noise rows: D02 e1, D02 e2, M13_N2 e1, M13_N2 e2, M13_O3N2 e1 (its e2 is never used)."""

    def __init__(self, n, rs=None):
        self.rs = rs if rs is not None else np.random
        self.noise = np.zeros((5, n))
        self._o3n2_toggle = 0

    def __call__(self, loc, scale, size):
        v = self.rs.normal(loc, scale, size)
        if scale == 0.05:
            self.noise[0] = v
        elif scale == 0.1:
            self.noise[1] = v
        elif scale == 0.027:
            self.noise[2] = v
        elif scale == 0.024:
            self.noise[3] = v
        elif scale == 0.012:
            if self._o3n2_toggle == 0:
                self.noise[4] = v
            self._o3n2_toggle ^= 1
        else:
            raise AssertionError(scale)
        return v


def oracle_galaxy(flux, mds="all", dust_corr=True, ignoredust=False, rs=None):
    """Run the oracle on one galaxy's [11,N] flux samples. Returns (present[37], Z[37,N], noise[5,N],
    trips(2), ret). `flux` is NOT modified."""
    flux = np.array(flux, dtype=np.float64)
    n = flux.shape[1]
    rec = RecordingNormal(n, rs)
    d, ret, ign = orc.calculation(flux.copy(), mds, dust_corr, ignoredust, "batched", rec)
    present = np.array([d.mds[k] is not None for k in orc.KEYS])
    Z = np.full((len(orc.KEYS), n), np.nan)
    for j, k in enumerate(orc.KEYS):
        if d.mds[k] is not None:
            Z[j] = d.mds[k]
    return present, Z, rec.noise, (d.trips_n2ha, d.trips_r23), (-1 if ret == -1 else 0)


def compare_samples(Z, Zo, present, presento, tol=1e-3, what=""):
    """Parity bar of BASELINE.json: present table and NaN masks bit-exact, values within tol dex."""
    assert np.array_equal(present, presento), what + " present table differs"
    nan_a, nan_b = np.isnan(Z), np.isnan(Zo)
    mism = int(np.sum(nan_a != nan_b))
    both = ~nan_a & ~nan_b
    inf_mism = int(np.sum(np.isinf(Z[both]) != np.isinf(Zo[both])))
    fin = both & np.isfinite(Z) & np.isfinite(Zo)
    maxdiff = float(np.max(np.abs(Z[fin] - Zo[fin]))) if fin.any() else 0.0
    same_inf = np.all(Z[both & ~fin] == Zo[both & ~fin]) if (both & ~fin).any() else True
    return mism, inf_mism, maxdiff, bool(same_inf)
