"""CPU oracle for the pyMCZ Monte-Carlo metallicity hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy/float64 restatement of the reference's algorithm for the path
``BASELINE.json:north_star`` names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``pymcz_b200/`` imports it and the product path has no CPU fallback.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified reference modules
(``/root/reference/pyMCZ/metallicity.py`` + ``metscales.py``) in the build container and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against those
vectors bit-for-bit (same numpy operations in the same order => identical float64 results).

Every function cites the reference ``file:line`` (paths relative to ``/root/reference/pyMCZ``) it
follows.  Reference quirks are reproduced on purpose; none is "fixed" here.
"""
from __future__ import annotations

import numpy as np
import numpy.polynomial.polynomial as nppoly

# --------------------------------------------------------------------------------------------
# Vocabulary (mcz.py:45, metallicity.py:29-56)
# --------------------------------------------------------------------------------------------
LINES = ['[OII]3727', 'Hb', '[OIII]4959', '[OIII]5007', '[OI]6300', 'Ha', '[NII]6584',
         '[SII]6717', '[SII]6731', '[SIII]9069', '[SIII]9532']                     # mcz.py:45
L_O2, L_HB, L_O34959, L_O3, L_O1, L_HA, L_N2, L_S2A, L_S2B, L_S39069, L_S39532 = range(11)

KEYS = ["E(B-V)", "logR23", "D02", "Z94", "M91", "C01_N2S2", "C01_R23", "P05", "P01",
        "PP04_N2Ha", "PP04_O3N2", "DP00", "P10_ONS", "P10_ON",
        "M08_R23", "M08_N2Ha", "M08_O3Hb", "M08_O2Hb", "M08_O3O2", "M08_O3N2",
        "M13_O3N2", "M13_N2",
        "D13_N2S2_O3S2", "D13_N2S2_O3Hb", "D13_N2S2_O3O2", "D13_N2O2_O3S2",
        "D13_N2O2_O3Hb", "D13_N2O2_O3O2", "D13_N2Ha_O3Hb", "D13_N2Ha_O3O2",
        "KD02_N2O2", "KD02_N2S2", "KK04_N2Ha", "KK04_R23", "KD02comb", "PM14", "D16"]  # metallicity.py:29-56
KEY_INDEX = {k: i for i, k in enumerate(KEYS)}

# CCM Rv=3.1 extinction constants (metscales.py:10-22)
k_Ha = 2.535
k_Hb = 3.609
k_O2 = 4.771
k_O35007 = 3.341
k_O34959 = 3.384
k_O3 = (k_O35007 + k_O34959) / 2.
k_N2 = 2.443
k_S2 = 2.381
k_S3 = 1

# Maiolino+08 coefficient table (metscales.py:61-66)
M08_COEFS = {'R23': [0.7462, -0.7149, -0.9401, -0.6154, -0.2524],
             'N2Ha': [-0.7732, 1.2357, -0.2811, -0.7201, -0.3330],
             'O3Hb': [0.1549, -1.5031, -0.9790, -0.0297],
             'O3O2': [-0.2839, -1.3881, -0.3172],
             'O2Hb': [0.5603, 0.0450, -1.8017, -1.8434, -0.6549],
             'O3N2': [0.4520, -2.6096, -0.7170, 0.1347]}

N2O2_COEF = [1106.8660, -532.15451, 96.373260, -7.8106123, 0.23928247]   # metscales.py:146, 419

# dust handling modes (metallicity.py:113-132)
DUST_MEASURED = 0     # E(B-V) from the Balmer decrement (metallicity.py:113-115)
DUST_FALLBACK = 1     # dust wanted, Ha/Hb missing, first time: E=1e-5 *applied* (metallicity.py:116-129)
DUST_OFF = 2          # --nodust or sticky IGNOREDUST: corrections are exactly 0/1 (metallicity.py:130-132)


# --------------------------------------------------------------------------------------------
# Root finding (metscales.py:248-263)
# --------------------------------------------------------------------------------------------
def fz_roots(coef, mode="loop"):
    """All complex roots of each row of ascending-order coefficients (metscales.py:257-263).

    Non-finite coefficients are zeroed first (metscales.py:259).  ``mode="loop"`` is the
    reference's per-sample ``np.roots`` loop (the reference's 98 % hot spot); ``mode="batched"``
    feeds the same companion matrices to one stacked ``eigvals`` call (the same LAPACK ``geev``
    per matrix; tests check it is identical to the loop) and falls back to ``np.roots`` for rows
    where ``np.roots`` would strip zero coefficients.
    """
    coef = np.array(coef, dtype=float)
    coef[~np.isfinite(coef)] = 0.0
    n, m = coef.shape
    rts = np.zeros((n, m - 1), dtype=complex)
    if mode == "loop":
        for i in range(n):
            rts[i] = np.roots(coef[i][::-1])
        return rts
    desc = coef[:, ::-1]
    generic = (desc[:, 0] != 0) & (desc[:, -1] != 0)
    if generic.any():
        p = desc[generic]
        # np.roots: A = diag(ones(N-2), -1); A[0, :] = -p[1:] / p[0]; eigvals(A)
        A = np.zeros((p.shape[0], m - 1, m - 1))
        idx = np.arange(m - 2)
        A[:, idx + 1, idx] = 1.0
        A[:, 0, :] = -p[:, 1:] / p[:, :1]
        rts[generic] = np.linalg.eigvals(A)
    for i in np.nonzero(~generic)[0]:
        rts[i] = np.roots(desc[i])
    return rts


def _first_valid(sols, extra=None):
    """The double-cumsum "first valid root" selection (metscales.py:980-991)."""
    ok = (sols.real >= 7.1) * (sols.real <= 9.4) * (sols.imag == 0)
    if extra is not None:
        ok = ok * extra
    indx = ok.cumsum(1).cumsum(1) == 1
    return indx


# --------------------------------------------------------------------------------------------
# One galaxy: metallicity.calculation + metscales.diagnostics restated
# --------------------------------------------------------------------------------------------
class Diag:
    """Per-galaxy state of ``metscales.diagnostics`` (metscales.py:86-162), functional flavour."""

    def __init__(self, n, roots_mode="loop", normal=None):
        self.n = n
        self.roots_mode = roots_mode
        self.normal = normal if normal is not None else np.random.normal
        self.mds = {k: None for k in KEYS}
        self.dust_on = True
        self.has = dict(Ha=False, Hb=False, O2=False, O3=False, S2=False, N2=False, O3Hb=False,
                        O3O2=False, N2O2=False, N2S2=False, S26731=False, S39532=False,
                        S39069=False, S2Hb=False)
        for name in ("Ha Hb O2 O3 N2 S2a S2b S39069 S39532 O34959p5007 O35007O2 O2O35007 "
                     "logO35007O2 logO2O35007 logO2Hb O3Hb logO3Hb OIII_OII logO3O2 OIII_Hb "
                     "logN2Ha NII_SII NII_OII logS2Ha OIII_SII N2S2 logN2S2 logN2O2 N2O2_roots "
                     "R2 R3 R23 logR23 logS23 P S2Hb logO3O2sq logq Z_init_guess").split():
            setattr(self, name, None)
        self.trips_n2ha = 0
        self.trips_r23 = 0
        self.conv_n2ha = []          # the loops' convergence measure per trip (diagnostics for the tests)
        self.conv_r23 = []

    # ---- dust (metscales.py:273-282)
    def dc(self, l1, l2, flux=False):
        if self.dust_on:
            if not flux:
                return 0.4 * self.mds['E(B-V)'] * (l1 - l2)
            return 10 ** (0.4 * self.mds['E(B-V)'] * (l1 - l2))
        return 1.0 if flux else 0

    # ---- setters (metscales.py:284-384)
    def set_hab(self, Ha, Hb):                                             # metscales.py:284-290
        self.Ha, self.Hb = Ha, Hb
        self.has['Ha'] = bool(np.sum(Ha > 0))
        self.has['Hb'] = bool(np.sum(Hb > 0))

    def set_olines(self, O2, O3, O34959):                                  # metscales.py:292-343
        h = self.has
        self.O2, self.O3 = O2, O3
        h['O3'] = bool(np.sum(O3 > 0))
        h['O2'] = bool(np.sum(O2 > 0))
        if h['O2'] and h['O3']:
            self.O35007O2 = (O3 / O2) * self.dc(k_O3, k_O2, flux=True)
            self.O2O35007 = (O2 / O3) * self.dc(k_O2, k_O3, flux=True)
            self.logO35007O2 = np.log10(self.O35007O2)
            self.logO2O35007 = np.log10(self.O2O35007)
        if h['Hb']:
            if h['O2']:
                self.logO2Hb = np.log10(O2 / self.Hb) + self.dc(k_O2, k_Hb, flux=False)
            if h['O3']:
                self.O3Hb = (O3 / self.Hb) * self.dc(k_O35007, k_Hb, flux=True)
                self.logO3Hb = np.log10(self.O3Hb)
                h['O3Hb'] = True
        if h['O2'] and h['O3']:
            self.OIII_OII = np.log10(O3 / O2 * self.dc(k_O35007, k_O2, flux=True))
            if O34959 is not None and np.sum(O34959 > 0) > 0:
                self.O34959p5007 = (O34959 + O3)
                self.logO3O2 = np.log10(self.O34959p5007 / O2) + self.dc(k_O3, k_O2, flux=False)
                h['O3O2'] = True
        if h['Hb']:
            self.OIII_Hb = np.log10(O3 / self.Hb * self.dc(k_O35007, k_Hb, flux=True))

    def set_sii(self, S2a, S2b, S39069, S39532):                           # metscales.py:361-384
        h = self.has
        if S2a is not None and np.sum(S2a > 0) > 0:
            self.S2a = S2a
            h['S2'] = True
            if h['Ha']:
                self.logS2Ha = np.log10(S2a / self.Ha) + self.dc(k_S2, k_Ha, flux=False)
        if S2b is not None and np.sum(S2b > 1e-9) > 0:
            self.S2b = S2b
            h['S26731'] = True
        if S39069 is not None and np.sum(S39069 > 1e-9) > 0:
            self.S39069 = S39069
            h['S39069'] = True
        if S39532 is not None and np.sum(S39532 > 1e-9) > 0:
            self.S39532 = S39532
            h['S39532'] = True
        if h['S2']:
            # hasN2 is still False here (setSII runs before setNII, metallicity.py:141-143)
            if h['N2'] and self.NII_SII is None and h['S26731']:
                self.NII_SII = np.log10(self.N2 / (self.S2a + self.S2b))
            if h['O3'] and self.OIII_SII is None and h['S26731']:
                self.OIII_SII = np.log10(self.O3 / (self.S2a + self.S2b) * self.dc(k_O3, k_S2, flux=True))

    def set_nii(self, N2):                                                 # metscales.py:345-359
        h = self.has
        if N2 is not None and np.sum(N2 > 0):
            self.N2 = N2
            h['N2'] = True
            if h['Ha']:
                self.logN2Ha = np.log10(N2 / self.Ha)
            if h['S2'] and h['S26731'] and h['N2']:
                self.NII_SII = np.log10(N2 / (self.S2a + self.S2b))
            if h['O2'] and h['N2']:
                self.NII_OII = np.log10(N2 / self.O2 * self.dc(k_N2, k_O2, flux=True))

    # ---- always-run precompute (metallicity.py:148-154)
    def calc_ebv(self):                                                    # metscales.py:387-390
        e = np.log10(2.86 * self.Hb / self.Ha) / (0.4 * (k_Ha - k_Hb))
        e[e <= 0] = 1e-5
        self.mds['E(B-V)'] = e

    def check_minimum(self, red_corr, ignoredust):                         # metscales.py:233-246
        h = self.has
        if red_corr and not ignoredust:
            if not h['Ha'] or not h['Hb']:
                return -1
        if not h['N2'] and not (h['O2'] and h['O3']):
            return -1
        return None

    def calc_niioii(self):                                                 # metscales.py:402-423
        h = self.has
        if h['N2'] and h['O2']:
            self.logN2O2 = np.log10(self.N2 / self.O2) + self.dc(k_N2, k_O2, flux=False)
            h['N2O2'] = True
        if not h['N2O2']:
            self.N2O2_roots = np.zeros(self.n) + float('NaN')
        else:
            coef = np.array([N2O2_COEF] * self.n).T
            coef[0] -= self.logN2O2
            self.N2O2_roots = fz_roots(coef.T, self.roots_mode)

    def calc_niisii(self):                                                 # metscales.py:393-399
        h = self.has
        if h['S2'] and h['N2']:
            self.N2S2 = self.N2 / self.S2a * self.dc(k_N2, k_S2, flux=True)
            self.logN2S2 = np.log10(self.N2 / self.S2a) + self.dc(k_N2, k_S2, flux=False)
            h['N2S2'] = True

    def calc_r23(self):                                                    # metscales.py:426-440
        h = self.has
        if h['O3'] and h['O2'] and h['Hb']:
            # NOTE: the reference dereferences O34959p5007 here; it is None only if O3>0 but
            # O3/3 underflows to 0 for every sample, which cannot happen for float32 inputs.
            self.R2 = (self.O2 / self.Hb) * self.dc(k_O2, k_Hb, flux=True)
            self.R3 = (self.O34959p5007 / self.Hb) * self.dc(k_O3, k_Hb, flux=True)
            self.R23 = self.R2 + self.R3
            self.logR23 = np.log10(self.R23)
            self.mds['logR23'] = self.logR23

    def calc_s23(self):                                                    # metscales.py:443-453
        h = self.has
        if h['S2'] and h['S39069'] and h['Hb']:
            self.logS23 = np.log10((self.S2a / self.Hb) * self.dc(k_S2, k_Hb, flux=True) +
                                   (self.S39069 / self.Hb) * self.dc(k_S3, k_Hb, flux=True))

    def calclogq(self, Z):                                                 # metscales.py:458-465
        if not self.has['O3O2']:
            return -1
        if self.logO3O2sq is None:
            self.logO3O2sq = self.logO3O2 ** 2
        y, y2 = self.logO3O2, self.logO3O2sq
        return (32.81 - 1.153 * y2 + Z * (-3.396 - 0.025 * y + 0.1444 * y2)) / \
               (4.603 - 0.3119 * y - 0.163 * y2 + Z * (-0.48 + 0.0271 * y + 0.02037 * y2))

    def initialguess(self):                                                # metscales.py:468-489
        h = self.has
        self.Z_init_guess = np.zeros(self.n) + 8.7
        if h['Ha'] and h['N2']:
            self.Z_init_guess[(self.logN2Ha < -1.3) & (self.N2 != 0.0)] = 8.2
            self.Z_init_guess[(self.logN2Ha < -1.1) & (self.logN2Ha >= -1.3) & (self.N2 != 0.0)] = 8.4
            self.Z_init_guess[(self.logN2Ha >= -1.1) & (self.N2 != 0.0)] = 8.7
        if h['N2'] and h['O2']:
            N2O2 = np.zeros(self.n) + float('nan')
            if h['Hb']:
                N2O2 = self.N2 * self.Ha * self.Hb * self.O2
            self.Z_init_guess[(self.logN2O2 < -1.2) & (N2O2 != 0.0)] = 8.2
            self.Z_init_guess[(self.logN2O2 >= -1.2) & (N2O2 != 0.0)] = 8.7

    # ---- scales
    def calc_d02(self):                                                    # metscales.py:674-685
        e1 = self.normal(0, 0.05, self.n)
        e2 = self.normal(0, 0.1, self.n)
        if self.has['N2'] and self.has['Ha']:
            self.mds['D02'] = 9.12 + e1 + (0.73 + e2) * self.logN2Ha

    def calc_pp04(self):                                                   # metscales.py:688-707
        if self.has['N2'] and self.has['Ha']:
            v = nppoly.polyval(self.logN2Ha, [9.37, 2.03, 1.26, 0.32])
            index = (self.logN2Ha > -2.5) * (self.logN2Ha < -0.3)
            v[~index] = float('NaN')
            self.mds['PP04_N2Ha'] = v
            if self.has['O3Hb']:
                w = 8.73 - 0.32 * (self.logO3Hb - self.logN2Ha)
                w[self.logO3Hb > 2] = float('NaN')
                self.mds['PP04_O3N2'] = w

    def _need_r23(self):
        if self.logR23 is None:
            self.calc_r23()
        return self.logR23 is not None

    def calc_z94(self):                                                    # metscales.py:710-725
        if not self._need_r23():
            return -1
        v = nppoly.polyval(self.logR23, [9.265, -0.33, -0.202, -0.207, -0.333])
        v[self.logR23 > 0.9] = None
        self.mds['Z94'] = v

    def calc_p(self):                                                      # metscales.py:729-740
        if self.P is None:
            if not self._need_r23():
                return -1
            self.P = self.R3 / self.R23

    def calc_p05(self):                                                    # metscales.py:743-761
        if self.calc_p() == -1:
            return -1
        if self.Z_init_guess is None:
            self.initialguess()
        P, R23 = self.P, self.R23
        Psq = P * P
        up = (R23 + 726.1 + 842.2 * P + 337.5 * Psq) / (85.96 + 82.76 * P + 43.98 * Psq + 1.793 * R23)
        low = (R23 + 106.4 + 106.8 * P - 3.40 * Psq) / (17.72 + 6.60 * P + 6.95 * Psq - 0.302 * R23)
        self.mds['P05'] = up
        sel = self.Z_init_guess < 8.4
        self.mds['P05'][sel] = low[sel]

    def calc_p10(self):                                                    # metscales.py:764-841
        h = self.has
        if not h['Hb']:
            return -1
        n = self.n
        ons = np.zeros(n) + float('NaN')
        on = np.zeros(n) + float('NaN')
        lR3 = np.zeros(n) + float('NaN')
        lR2 = np.zeros(n) + float('NaN')
        lN2 = np.zeros(n) + float('NaN')
        lS2 = np.zeros(n) + float('NaN')
        self.calc_p()
        if self.R2 is not None:
            lR2 = np.log(self.R2)                    # natural log (sic, metscales.py:788)
        if self.R3 is not None:
            lR3 = np.log(self.R3)                    # natural log (sic, metscales.py:791)
        if h['N2']:
            lN2 = (np.log((self.N2 * 1.33) / self.Hb) + self.dc(k_N2, k_Hb, flux=False))  # :798
        if h['S2'] and h['S26731']:
            self.S2Hb = ((self.S2a + self.S2b) / self.Hb * self.dc(k_S2, k_Hb, flux=True))
            h['S2Hb'] = True
            lS2 = np.log10(self.S2Hb)                # base 10 (sic, metscales.py:805)
        lN2S2 = lN2 - lS2
        lN2R2 = lN2 - lR2
        lS2R2 = lS2 - lR2
        cONS = [np.array([8.277, 0.657, -0.399, -0.061, 0.005]),
                np.array([8.816, -0.733, 0.454, 0.710, -0.337]),
                np.array([8.774, -1.855, 1.517, 0.304, 0.328])]
        cON = [np.array([8.606, -0.105, -0.410, -0.150]),
               np.array([8.642, 0.077, 0.411, 0.601]),
               np.array([8.013, 0.905, 0.602, 0.751])]
        # When P is None (no R23) the reference builds an object array here, which numpy >= 1.24
        # rejects; under the numpy of the reference's era the outcome was all-NaN P10 arrays.
        # The oracle follows that outcome (ONS untouched = NaN; ON = dot with NaN rows = NaN).
        if self.P is not None:
            vsONS = np.array([np.ones(n), self.P, lR3, lN2R2, lS2R2]).T
        vsON = np.array([np.ones(n), lR3, lR2, lN2R2]).T
        regimes = [lN2 > -0.1,
                   (lN2 < -0.1) * (lN2S2 > -0.25),
                   (lN2 < -0.1) * (lN2S2 < -0.25)]
        for indx, a, b in zip(regimes, cONS, cON):
            if self.P is not None:
                ons[indx] = np.dot(vsONS[indx], a)
            on[indx] = np.dot(vsON[indx], b)
        bad = ~((ons > 7.1) * (on > 7.1) * (ons < 9.4) * (on < 9.4))
        if self.P is not None:
            ons[bad] = float('NaN')
        on[bad] = float('NaN')
        self.mds['P10_ONS'], self.mds['P10_ON'] = ons, on

    def calc_p01(self):                                                    # metscales.py:844-865
        h = self.has
        if self.Z_init_guess is None:
            self.initialguess()
        if h['O3O2'] and h['O3'] and h['O2']:
            P = 10 ** self.logO3O2 / (1 + 10 ** self.logO3O2)
            if not self._need_r23():
                return -1
            Psq = P ** 2
            old = (self.R23 + 54.2 + 59.45 * P + 7.31 * Psq) / (6.07 + 6.71 * P + 0.371 * Psq + 0.243 * self.R23)
            v = np.zeros(self.n) + float('NaN')
            sel = self.Z_init_guess >= 8.4
            v[sel] = old[sel]
            self.mds['P01'] = v

    def calc_c01(self):                                                    # metscales.py:868-894
        h = self.has
        x2 = x3 = None
        if h['O3'] and h['O2'] and h['O3Hb']:
            x2 = self.O2O35007 / 1.5
            x3 = (10 ** self.logO3Hb) * 0.5
            v = np.zeros(self.n) + float('NaN')
            lo = self.O2O35007 < 0.8
            v[lo] = np.log10(3.78e-4 * (x2[lo]) ** 0.17 * x3[lo] ** (-0.44)) + 12.0
            hi = self.O2O35007 >= 0.8
            v[hi] = np.log10(3.96e-4 * x3[hi] ** (-0.46)) + 12.0
            self.mds['C01_R23'] = v
        if not h['N2S2']:
            self.calc_niisii()
        if h['N2S2'] and h['O3'] and h['O2'] and h['O3Hb']:
            self.mds['C01_N2S2'] = np.log10(5.09e-4 * (x2 ** 0.17) * ((self.N2S2 / 0.85) ** 1.17)) + 12

    def calc_m91(self):                                                    # metscales.py:897-936
        self.mds['M91'] = np.zeros(self.n) + float('NaN')
        if not self._need_r23():
            return -1
        if self.Z_init_guess is None:
            self.initialguess()
        x, y = self.logR23, self.logO3O2
        low = nppoly.polyval(x, [12.0 - 4.944, 0.767, 0.602]) - \
            y * nppoly.polyval(x, [0.29, 0.332, -0.331])
        up = nppoly.polyval(x, [12.0 - 2.939, -0.2, -0.237, -0.305, -0.0283]) - \
            y * nppoly.polyval(x, [0.0047, -0.0221, -0.102, -0.0817, -0.00717])
        indx = (np.abs(y) > 0) * (np.abs(x) > 0) * (self.Z_init_guess < 8.4)
        self.mds['M91'][indx] = low[indx]
        indx = (np.abs(y) > 0) * (np.abs(x) > 0) * (self.Z_init_guess >= 8.4)
        self.mds['M91'][indx] = up[indx]
        self.mds['M91'][(up < low)] = float('NaN')

    def calc_m13(self):                                                    # metscales.py:939-959
        h = self.has
        if not h['Ha'] or not h['N2']:
            return -1
        e1 = self.normal(0, 0.027, self.n)
        e2 = self.normal(0, 0.024, self.n)
        self.mds["M13_N2"] = 8.743 + e1 + (0.462 + e2) * self.logN2Ha
        if h['Hb'] and h['O3']:
            e1 = self.normal(0, 0.012, self.n)
            e2 = self.normal(0, 0.012, self.n)          # drawn, never used (metscales.py:953)
            O3N2 = self.logO3Hb - self.logN2Ha
            v = 8.533 + e1 - (0.214 + e1) * O3N2        # e1 twice (sic, metscales.py:955)
            v[O3N2 > 1.7] = float('NaN')
            v[O3N2 < -1.1] = float('NaN')
            self.mds["M13_O3N2"] = v

    def _m08_one(self, name, yvals, extra_fn=None):
        coefs = np.array([M08_COEFS[name]] * self.n).T
        coefs[0] = coefs[0] - yvals
        sols = fz_roots(coefs.T, self.roots_mode) + 8.69
        return sols

    def calc_m08(self, allM08=False):                                      # metscales.py:969-1059
        n = self.n
        highZ = None
        if self.logO35007O2 is not None:
            v = np.zeros(n) + float('NaN')
            sols = self._m08_one('O3O2', self.logO35007O2)
            indx = _first_valid(sols)
            v[(indx.sum(1)) > 0] = sols[indx].real
            self.mds['M08_O3O2'] = v
            highZ = np.median(self.logO35007O2) < 0
        if self.logN2Ha is not None:
            v = np.zeros(n) + float('NaN')
            sols = self._m08_one('N2Ha', self.logN2Ha)
            indx = _first_valid(sols)
            v[(indx.sum(1)) > 0] = sols[indx].real
            self.mds['M08_N2Ha'] = v
            if highZ is None:
                highZ = np.median(self.logN2Ha) > -1.3
        if self.logR23 is None:
            self.calc_r23()
        if self.logR23 is not None:
            v = np.zeros(n) + float('NaN')
            sols = self._m08_one('R23', self.logR23)
            # `highZ is True` is never true for a numpy.bool_ (metscales.py:1013-1018): dead branch
            if highZ is True:
                indx = _first_valid(sols, sols.real >= 8.0)
                v[(indx.sum(1)) > 0] = sols[indx].real
            elif highZ is False:
                indx = _first_valid(sols, sols.real <= 8.0)
                v[(indx.sum(1)) > 0] = sols[indx].real
            self.mds['M08_R23'] = v
        if not allM08:
            return
        for name, yv, cut in (('O3Hb', self.logO3Hb, 7.9), ('O2Hb', self.logO2Hb, 8.7)):  # :1024-1047
            if yv is not None:
                v = np.zeros(n) + float('NaN')
                sols = self._m08_one(name, yv)
                if highZ is True:
                    indx = _first_valid(sols, sols.real >= cut)
                    v[(indx.sum(1)) > 0] = sols[indx].real
                elif highZ is False:
                    indx = _first_valid(sols, sols.real <= cut)
                    v[(indx.sum(1)) > 0] = sols[indx].real
                self.mds['M08_' + name] = v
        if self.has['O3'] and self.has['N2']:                              # metscales.py:1051-1059
            v = np.zeros(n) + float('NaN')
            coefs = np.array([M08_COEFS['O3N2']] * n).T
            coefs[0] = coefs[0] - np.log10(self.O3 / self.N2) + self.dc(k_O35007, k_N2, flux=False)
            sols = fz_roots(coefs.T, self.roots_mode) + 8.69
            indx = _first_valid(sols)
            v[(indx.sum(1)) > 0] = sols[indx].real
            self.mds['M08_O3N2'] = v

    def calc_kd02_n2o2(self):                                              # metscales.py:1062-1090
        h = self.has
        if h['N2'] and h['O2'] and h['Ha'] and h['Hb']:
            self.mds['KD02_N2O2'] = np.zeros(self.n) + float('NaN')
            if not h['N2O2']:
                self.calc_niioii()
            r = self.N2O2_roots
            if (not h['N2O2'] or r is None or
                    np.sum(np.isnan(r.flatten())) == len(r.flatten())):
                return -1
            roots = r.T
            for k in range(4):
                indx = (abs(roots[k]) >= 7.5) * (abs(roots[k]) <= 9.4) * (roots[k][:].imag == 0.0)
                self.mds['KD02_N2O2'][indx] = abs(roots[k][indx])
        return 1

    def _kk04_n2ha_step(self, logq):                                       # metscales.py:1114-1116
        n = self.logN2Ha
        return nppoly.polyval(n, [7.04, 5.28, 6.28, 2.37]) - \
            logq * nppoly.polyval(n, [-2.44, -2.01, -0.325, +0.128]) + \
            10 ** (n - 0.2) * logq * (-3.16 + 4.65 * n)

    def calc_kk04_n2ha(self):                                              # metscales.py:1093-1131
        h = self.has
        if self.mds['KD02_N2O2'] is None:
            self.calc_kd02_n2o2()
        if self.mds['KD02_N2O2'] is None or np.sum(np.isnan(self.mds['KD02_N2O2'])) == self.n:
            Z = np.zeros(self.n) + 8.6
        else:
            Z = self.mds['KD02_N2O2'].copy()
        if h['N2'] and h['Ha']:
            logq_save = np.zeros(self.n)
            convergence, tol, ii = 100, 1.0e-3, 0
            if h['O3O2']:
                while convergence > tol and ii < 100:
                    ii += 1
                    self.logq = self.calclogq(Z)
                    Z = self._kk04_n2ha_step(self.logq)
                    convergence = np.abs(self.logq - logq_save).mean()
                    self.conv_n2ha.append(float(convergence))
                    logq_save = self.logq.copy()
                if ii >= 100:
                    Z = np.zeros(self.n) + float('NaN')
            else:
                self.logq = 7.37177 * np.ones(self.n)
                Z = self._kk04_n2ha_step(self.logq)
            self.trips_n2ha = ii
            self.mds['KK04_N2Ha'] = Z
            self.mds['KK04_N2Ha'][self.logN2Ha > 0.8] = float('NaN')

    def calc_kk04_r23(self):                                               # metscales.py:1134-1187
        Zmax = np.zeros(self.n)
        if not self.has['O3O2']:
            return
        if self.Z_init_guess is None:
            self.initialguess()
        Z = self.Z_init_guess.copy()
        if not self._need_r23():
            return
        x = self.logR23
        logqold, convergence, ii = np.zeros(self.n) + 100, 100, 0
        tol = 1e-4
        while convergence > tol and ii < 100:
            Zmax = Zmax * 0.0
            ii += 1
            logq = self.calclogq(Z)
            Zmax[(logq >= 6.7) * (logq < 8.3)] = 8.4
            Z = nppoly.polyval(x, [9.72, -0.777, -0.951, -0.072, -0.811]) - \
                logq * nppoly.polyval(x, [0.0737, -0.0713, -0.141, 0.0373, -0.058])
            indx = self.Z_init_guess <= Zmax
            Z[indx] = nppoly.polyval(x[indx], [9.40, 4.65, -3.17]) - \
                logq[indx] * nppoly.polyval(x[indx], [0.272, 0.547, -0.513])
            convergence = np.abs((logqold - logq).mean())
            self.conv_r23.append(float(convergence))
            logqold = logq.copy()
        if ii >= 100:
            Z = np.zeros(self.n) + float('NaN')
        lims = [nppoly.polyval(x, [9.40, 4.65, -3.17]) -
                logq * nppoly.polyval(x, [0.272, 0.547, -0.513]),
                nppoly.polyval(x, [9.72, -0.777, -0.951, -0.072, -0.811]) -
                logq * nppoly.polyval(x, [0.0737, -0.0713, -0.141, 0.0373, -0.058])]
        Z[(lims[0] > lims[1])] = None
        self.trips_r23 = ii
        self.mds['KK04_R23'] = Z

    def calc_kdcombined(self):                                             # metscales.py:1190-1264
        m = self.mds
        if m['KD02_N2O2'] is None:
            self.calc_kd02_n2o2()
        if m['KK04_N2Ha'] is None:
            self.calc_kk04_n2ha()
        if self.logR23 is None:
            self.calc_r23()
        if m['M91'] is None:
            self.calc_m91()
        if m['KK04_R23'] is None:
            self.calc_kk04_r23()
        # metscales.py:1232-1239 computes a local `logq` that is never used (dead code); when the
        # reference evaluates `calclogq(None)` there it raises -- reproduced by the guard below.
        h = self.has
        if h['N2'] and h['O2'] and h['Hb'] and h['Ha'] and h['O3O2']:
            pass            # needs mds['KD02_N2O2'] (always set under this guard) and self.logq
        else:
            if self.Z_init_guess is None:
                self.initialguess()
        v = np.zeros(self.n) + float('NaN')
        ig = self.Z_init_guess > 8.4
        if m['KD02_N2O2'] is not None:
            v[ig] = m['KD02_N2O2'][ig].copy()
        if m['KK04_N2Ha'] is not None:
            v[~ig] = m['KK04_N2Ha'][~ig].copy()
        if m['KK04_R23'] is not None and m['M91'] is not None:
            indx = (~np.isnan(m['KK04_R23'])) * (~np.isnan(m['M91'])) * (~ig)
            v[indx] = 0.5 * (m['KK04_R23'][indx].copy() + m['M91'][indx].copy())
        m['KD02comb'] = v

    def calc_dp00(self):                                                   # metscales.py:658-671
        if self.logS23 is None:
            self.calc_s23()
            if self.logS23 is None:
                return -1
        self.mds['DP00'] = 1.53 * self.logS23 + 8.27 + 1.0 / (2.0 - 9.0 * self.logS23 ** 3)

    def calc_d16(self):                                                    # metscales.py:1343-1356
        h = self.has
        if not h['Ha'] or not h['N2'] or not h['S2']:
            return -1
        y = self.logN2S2 + 0.264 * self.logN2Ha
        v = 8.77 + y - 0.45 * pow(y + 0.3, 5)
        v[y < -1.] = float('NaN')
        v[y > 0.5] = float('NaN')
        self.mds["D16"] = v


def calculation(flux, mds='all', dust_corr=True, ignoredust=False, roots_mode="loop", normal=None,
                dropnegatives=False):
    """``metallicity.calculation`` (metallicity.py:86-257) for ONE galaxy.

    flux : float64 array [11, N] in ``LINES`` order (NaN rows = line absent) -- it is sanitised
           in place like the reference does (metallicity.py:94-99).
    Returns (diag, ret, ignoredust') where ``ret`` is -1/None as in the reference and
    ``ignoredust'`` is the updated sticky ``IGNOREDUST`` global (metallicity.py:127).
    pyqz (D13) and HII-CHI-mistry (PM14) are third-party and absent: their keys stay None, as in
    the reference when ``PYQZ_DIR`` / ``HIICHI_DIR`` are unset (metallicity.py:163-171, 227-232).
    """
    flux = np.asarray(flux)
    assert flux.dtype == np.float64 and flux.shape[0] == 11
    n = flux.shape[1]
    d = Diag(n, roots_mode=roots_mode, normal=normal)
    d.dust_on = True                                                       # metallicity.py:89
    with np.errstate(all='ignore'):
        for k in range(11):                                                # metallicity.py:94-99
            row = flux[k]
            row[~np.isfinite(row)] = 0.0
            row[row < 0] = np.nan if dropnegatives else 0.0                # DROPNEGATIVES / FIXNEGATIVES, metallicity.py:23-26, 98-102
        o34959 = flux[L_O3] / 3.                                           # metallicity.py:108
        d.set_hab(flux[L_HA], flux[L_HB])
        if dust_corr and d.has['Ha'] and d.has['Hb']:                      # metallicity.py:113-115
            d.calc_ebv()
            d.dust_mode = DUST_MEASURED
        elif dust_corr and not ignoredust:                                 # metallicity.py:116-129
            ignoredust = True
            dust_corr = False
            d.mds['E(B-V)'] = np.ones(n) * 1e-5
            d.dust_mode = DUST_FALLBACK
        else:                                                              # metallicity.py:130-132
            d.dust_on = False
            d.mds['E(B-V)'] = np.ones(n) * 1e-5
            d.dust_mode = DUST_OFF
        d.set_olines(flux[L_O2], flux[L_O3], o34959)                       # metallicity.py:141-143
        d.set_sii(flux[L_S2A], flux[L_S2B], flux[L_S39069], flux[L_S39532])
        d.set_nii(flux[L_N2])
        if d.check_minimum(dust_corr, ignoredust) == -1:                   # metallicity.py:145-146
            return d, -1, ignoredust
        d.calc_niioii()                                                    # metallicity.py:148-154
        d.calc_niisii()
        d.calc_r23()
        d.initialguess()
        toks = mds.split(',')                                              # metallicity.py:155
        if 'all' in toks:                                                  # metallicity.py:161-188
            d.calc_d02()
            d.calc_z94()
            d.calc_m91()
            d.calc_pp04()
            d.calc_p10()
            d.calc_m08()
            d.calc_m13()
            d.calc_kd02_n2o2()
            d.calc_kk04_n2ha()
            d.calc_kk04_r23()
            d.calc_kdcombined()
        if 'DP00' in toks:                                                 # metallicity.py:190-257
            d.calc_dp00()
        if 'P01' in toks:
            d.calc_p01()
        if 'D02' in toks:
            d.calc_d02()
        if 'PP04' in toks:
            d.calc_pp04()
        if 'Z94' in toks:
            d.calc_z94()
        if 'M91' in toks:
            d.calc_m91()
        if 'P10' in toks:
            d.calc_p10()
        if 'M13' in toks:
            d.calc_m13()
        if 'M08all' in toks:
            d.calc_m08(allM08=True)
        elif 'M08' in toks:
            d.calc_m08()
        if 'P05' in toks:
            d.calc_p05()
        if 'C01' in toks:
            d.calc_c01()
        if 'KD02' in toks:
            d.calc_kd02_n2o2()
            d.calc_kk04_n2ha()
            d.calc_kk04_r23()
            d.calc_kdcombined()
        if 'D16' in toks:
            d.calc_d16()
    return d, None, ignoredust


# --------------------------------------------------------------------------------------------
# Sampling (mcz.py:476-479, 509-519, 436-442, 580-589)
# --------------------------------------------------------------------------------------------
def n_draw(nsample):
    """N' = int(1.1 nsample) for nsample > 1 (mcz.py:477-479)."""
    return int(nsample + 0.1 * nsample) if nsample > 1 else nsample


def draw_samples(nm, nsample, rs=np.random):
    """One N(0,1) vector per galaxy, all drawn up front (mcz.py:514-517)."""
    if nsample == 1:
        return [np.array([0]) for _ in range(nm)]
    nn = n_draw(nsample)
    return [rs.normal(0, 1, nn) for _ in range(nm)]


def perturb(meas_row, err_row, sample, mode, keys=None, rs=np.random):
    """flux_k = meas_k + err_k * sample for one galaxy -> float64 [11, N].

    mode "correlated": the ``--multiproc`` worker, no shuffle (mcz.py:442);
    mode "shuffled":   the serial path, ``shuffle(sample)`` before every line (mcz.py:582-589).
    ``keys`` is the iteration order of the file header dict (default: ``LINES`` order).
    meas/err are float32 scalars promoted on multiply exactly like the structured-array fields.
    """
    keys = LINES if keys is None else keys
    out = np.full((11, len(sample)), np.nan)
    for k in keys:
        j = LINES.index(k)
        if mode == "shuffled":
            rs.shuffle(sample)
        out[j] = meas_row[j] * np.ones(len(sample)) + err_row[j] * sample
    return out


# --------------------------------------------------------------------------------------------
# Percentile extraction (mcz.py:273-301, 423)
# --------------------------------------------------------------------------------------------
ST_OK, ST_ABSENT, ST_EMPTY, ST_NONPOS, ST_NODIST = 0, 1, 2, 3, 4


def savehist_stats(data):
    """(status, median, p16, p84, n) following ``savehist`` (mcz.py:273-301).

    status: ST_EMPTY  -> n == 0, string "-1 -1 -1" (mcz.py:277-280);
            ST_NONPOS -> sum(data) <= 0, string "-1 -1 -1" (mcz.py:282-285);
            ST_NODIST -> the three percentiles agree to 6 decimals (mcz.py:297-301);
            ST_OK otherwise.
    """
    data = np.asarray(data, dtype=float)
    data = data[np.isfinite(data)]
    n = data.shape[0]
    if not n > 0:
        return ST_EMPTY, np.nan, np.nan, np.nan, 0
    if np.sum(data) <= 0:
        return ST_NONPOS, np.nan, np.nan, np.nan, n
    median, pc16, pc84 = np.percentile(data, [50, 16, 84])
    if round(pc84, 6) == round(pc16, 6) and round(pc16, 6) == round(median, 6):
        return ST_NODIST, median, pc16, pc84, n
    return ST_OK, median, pc16, pc84, n


def savehist_string(status, median, pc16, pc84):
    """The tab-separated string ``savehist`` returns (mcz.py:280, 285, 301, 423)."""
    if status in (ST_EMPTY, ST_NONPOS):
        return "-1\t -1\t -1"
    return "%f\t %f\t %f" % (round(median, 3), round(median - pc16, 3), round(pc84 - median, 3))


def run_table(meas, err, nsample, mds='all', dust_corr=True, mode="correlated",
              roots_mode="batched", rs=np.random, normal=None, keep_samples=False):
    """Whole-table driver: ``run()`` minus I/O (mcz.py:470-611) + per-key ``savehist`` stats.

    meas, err : float32 [G, 11].  Returns dict with pct [G,37,3], nvalid [G,37], status [G,37],
    trips [G,2], ret [G] and, if keep_samples, Z [G,37,N'] (NaN where None, mcz.py:604-605).
    """
    meas = np.asarray(meas, dtype=np.float32)
    err = np.asarray(err, dtype=np.float32)
    G = meas.shape[0]
    samples = draw_samples(G, nsample, rs)
    nn = len(samples[0])
    pct = np.full((G, len(KEYS), 3), np.nan)
    nvalid = np.zeros((G, len(KEYS)), dtype=np.int32)
    status = np.full((G, len(KEYS)), ST_ABSENT, dtype=np.int8)
    trips = np.zeros((G, 2), dtype=np.int32)
    rets = np.zeros(G, dtype=np.int32)
    Z = np.full((G, len(KEYS), nn), np.nan) if keep_samples else None
    ignoredust = False
    for g in range(G):
        flux = perturb(meas[g], err[g], samples[g], mode, rs=rs)
        d, ret, ignoredust = calculation(flux, mds, dust_corr, ignoredust, roots_mode, normal)
        rets[g] = -1 if ret == -1 else 0
        trips[g] = (d.trips_n2ha, d.trips_r23)
        for j, key in enumerate(KEYS):
            v = d.mds[key]
            if v is None:
                continue
            if keep_samples:
                Z[g, j] = v
            st, med, p16, p84, n = savehist_stats(v)
            status[g, j] = st
            nvalid[g, j] = n
            pct[g, j] = (med, p16, p84)
    return dict(pct=pct, nvalid=nvalid, status=status, trips=trips, ret=rets, Z=Z)
