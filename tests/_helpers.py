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


# ---- numpy model of the production sampler (csrc/mcz_philox.cuh + draw_normals) ---------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011); uint32 arrays in, 4 uint32 arrays out."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def box_muller(a, b):
    u1 = (((a >> np.uint32(9)) | np.uint32(0x3f800000)).view(np.float32) - np.float32(0.99999994)).astype(np.float64)
    u2 = (((b >> np.uint32(9)) | np.uint32(0x3f800000)).view(np.float32) - np.float32(1.5)).astype(np.float64)
    r = np.sqrt(np.maximum(-2.0 * np.log(u1), 0.0))
    th = 2.0 * np.pi * u2
    return r * np.cos(th), r * np.sin(th)


def production_normals(seed, galaxy, n):
    """The 16 standard normals per sample the production kernel derives for global galaxy index
    `galaxy` (float64 model of the float32/MUFU device arithmetic): returns z[16, n]."""
    s = np.arange(n, dtype=np.uint32)
    z = np.zeros((16, n))
    for c in range(4):
        o = philox4x32_10(s, np.full(n, c, np.uint32), np.full(n, galaxy & 0xFFFFFFFF, np.uint32),
                          np.full(n, galaxy >> 32, np.uint32), seed & 0xFFFFFFFF, seed >> 32)
        z[4 * c + 0], z[4 * c + 1] = box_muller(o[0], o[1])
        z[4 * c + 2], z[4 * c + 3] = box_muller(o[2], o[3])
    return z


USED_LINES = [5, 1, 0, 3, 6, 7, 8, 9]      # Ha, Hb, [OII], [OIII]5007, [NII], [SII]6717, [SII]6731, [SIII]9069
NOISE_SIGMA = [0.05, 0.1, 0.027, 0.024, 0.012]


def production_flux_model(meas_row, err_row, seed, galaxy, n, correlated=False):
    """flux [11, n] float64 and noise [5, n] as the production kernel samples them (float32
    fma of float32 inputs); lines the kernel never reads are left NaN-free copies of meas."""
    z = production_normals(seed, galaxy, n)
    flux = np.zeros((11, n))
    for j in range(11):
        flux[j] = meas_row[j]
    for u, line in enumerate(USED_LINES):
        zz = z[5] if correlated else z[5 + u]
        flux[line] = (np.float32(meas_row[line]) + np.float32(err_row[line]) * zz.astype(np.float32)).astype(np.float64)
    noise = np.array([NOISE_SIGMA[j] * z[j] for j in range(5)])
    return flux, noise


# ---- the production kernels' own samples, read back from the device --------------------------
def kernel_samples(meas, err, n, seed, goff, tokens, correlated=False, deterministic=False):
    """flux [G, 11, n] float64 and noise [G, 5, n] float64 exactly as the production kernels sample
    them (mcz_debug_production_samples: same device code as phase A, bit for bit).  Lines the
    arithmetic never reads keep their measured value ([OIII]4959, [OI]6300, [SIII]9532)."""
    import torch
    from pymcz_b200 import _lib
    L = _lib.lib()
    meas = np.ascontiguousarray(meas, dtype=np.float32)
    err = np.ascontiguousarray(err, dtype=np.float32)
    G = meas.shape[0]
    flux = np.repeat(meas.astype(np.float64)[:, :, None], n, axis=2)
    noise = np.zeros((G, 5, n))
    for lo in range(0, G, 4096):
        hi = min(G, lo + 4096)
        m, e = torch.from_numpy(meas[lo:hi]).cuda(), torch.from_numpy(err[lo:hi]).cuda()
        f = torch.empty((hi - lo, 8, n), dtype=torch.float32, device="cuda")
        z = torch.empty((hi - lo, 5, n), dtype=torch.float32, device="cuda")
        rc = L.mcz_debug_production_samples(m.data_ptr(), e.data_ptr(), hi - lo, n, int(tokens), int(seed),
                                            int(goff) + lo, 1 if correlated else 0, 1 if deterministic else 0,
                                            f.data_ptr(), z.data_ptr(), None)
        _lib.check(rc, "mcz_debug_production_samples")
        torch.cuda.synchronize()
        f = f.cpu().numpy().astype(np.float64)
        for u, line in enumerate(USED_LINES):
            flux[lo:hi, line] = f[:, u]
        noise[lo:hi] = z.cpu().numpy().astype(np.float64)
    return flux, noise


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


def _oracle_one(args):
    flux, noise, mds, dust = args
    d, ret, _ = orc.calculation(flux, mds, dust, False, "batched", FixedNoise(noise))
    n = flux.shape[1]
    present = np.zeros(len(orc.KEYS), dtype=bool)
    Z = np.full((len(orc.KEYS), n), np.nan)
    for j, k in enumerate(orc.KEYS):
        if d.mds[k] is not None:
            present[j] = True
            Z[j] = d.mds[k]
    return present, Z, (d.trips_n2ha, d.trips_r23), (d.conv_n2ha, d.conv_r23)


def oracle_on_samples(flux, noise, mds, dust, procs=None):
    """The oracle on [G, 11, n] flux samples (float64) + [G, 5, n] coefficient noise, galaxies spread
    over a process pool.  Returns Z [G,37,n], present [G,37], trips [G,2] and the per-trip
    convergence measures of the two KK04 loops (list of (conv_n2ha, conv_r23) per galaxy)."""
    import multiprocessing as mp
    import os
    G, _, n = flux.shape
    jobs = [(flux[g], noise[g], mds, dust) for g in range(G)]
    procs = procs or min(os.cpu_count() or 1, 16)
    if procs > 1 and G >= 8:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_oracle_one, jobs, chunksize=max(1, G // (8 * procs)))
    else:
        res = [_oracle_one(j) for j in jobs]
    Zo = np.stack([r[1] for r in res])
    po = np.stack([r[0] for r in res])
    trips = np.array([r[2] for r in res], dtype=np.int32)
    conv = [r[3] for r in res]
    return Zo, po, trips, conv


ITER_KEYS = ("KK04_N2Ha", "KK04_R23", "KD02comb")      # outputs of the two KK04 fixed-point loops


def compare_with_oracle(Z, status, trips, Zo, po, trips_o, what=""):
    """The parity bar for the float32 production arithmetic, over EVERY galaxy of the run:
      * present table identical;
      * loop trip counts identical for >= 99.9 % of the galaxies, and a galaxy may only differ when the
        loop in question ran for 9 trips or more on both sides (the N2Ha map of some samples is then
        elliptic -- a rotation that never converges -- and whether the MEAN dips below the tolerance at
        some late trip depends on rotation phases that float32 cannot track for 50+ trips);
      * |dZ| < 1e-3 dex for every sample of every scale of every galaxy, except a counted handful
        (< 2e-6 of the loop outputs -- measured 5e-7 --, < 3e-8 of the other scales) that sit within float32 rounding of a
        branch threshold (Z_init_guess class, P10 regime, root choice, lower / upper R23 branch) or on
        a nearly parabolic fixed-point map (error amplification 1 / (1 - multiplier)); the median
        |dZ| is below 2e-6;
      * NaN-mask flips counted and bounded (samples within float32 rounding of a validity threshold).
    Returns the counts for the report."""
    assert np.array_equal(status != 1, po), what + ": present table differs"
    G = trips.shape[0]
    same = np.all(trips == trips_o, axis=1)
    nbad = int((~same).sum())
    assert nbad <= max(1, int(1e-3 * G)), (what, nbad, trips[~same][:10], trips_o[~same][:10])
    for g in np.nonzero(~same)[0]:
        for k in range(2):
            if trips[g, k] != trips_o[g, k]:
                assert min(trips[g, k], trips_o[g, k]) >= 9, (what, g, trips[g], trips_o[g])
    Zd = Z.astype(np.float64)
    flips = np.isnan(Zd) != np.isnan(Zo)
    flips_same = int(flips[same].sum())
    assert flips_same <= 2e-6 * flips[same].size + 3, (what, flips_same)
    both = np.isfinite(Zd) & np.isfinite(Zo)
    d = np.abs(np.where(both, Zd - Zo, 0.0))
    it = [orc.KEY_INDEX[k] for k in ITER_KEYS]
    rest = [j for j in range(len(orc.KEYS)) if j not in it]
    over_rest = int((d[:, rest] > 1e-3).sum())
    assert over_rest <= 3e-8 * d[:, rest].size + 1, (what, over_rest, d[:, rest].max())
    over = int((d[:, it] > 1e-3).sum())
    assert over <= 2e-6 * d[:, it].size + 2, (what, over, d[:, it].max())
    assert np.median(d[both]) < 2e-6
    worst = {orc.KEYS[j]: float(d[:, j].max()) for j in range(len(orc.KEYS)) if d[:, j].max() > 1e-3}
    rep = dict(galaxies=G, trip_mismatches=nbad, mask_flips_in_trip_equal_galaxies=flips_same,
               mask_flips_total=int(flips.sum()), iteration_values_over_1e3=over,
               other_values_over_1e3=over_rest, scales_with_exceedances=worst,
               p999_dz=float(np.quantile(d[both], 0.999)))
    print(what, rep)
    return rep
