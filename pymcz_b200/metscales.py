"""Drop-in for the reference's ``metscales`` module (/root/reference/pyMCZ/metscales.py).

``diagnostics`` keeps the reference's constructor, ``mds`` dict (key -> float64 array or None),
``has*`` flags, line setters and ``calc*`` methods -- but every ``calc*`` is a thin request to the
CUDA replay op (``pymcz_b200.ops.forward_replay``): the samples the caller set are uploaded once,
the scale(s) are evaluated on the GPU in float64 and the affected ``mds`` entries are filled.
There is no numpy implementation of the physics here and no CPU fallback.

Differences from the reference, all deliberate:
  * no module-global ``DUSTCORRECT`` (metscales.py:24-25, 265-271): the switch is per instance;
  * ``calcpyqz`` / ``calcPM14`` only warn: their arithmetic lives in third-party code the
    reference does not vendor (pyqz, HII-CHI-mistry; metscales.py:497-655, 1268-1340).
"""
from __future__ import annotations

import numpy as np

from . import _scales as sc

k_Ha, k_Hb = 2.535, 3.609                      # metscales.py:10-11 (CCM Rv=3.1)
k_O2, k_O35007, k_O34959 = 4.771, 3.341, 3.384  # metscales.py:14-16
k_O3 = (k_O35007 + k_O34959) / 2.               # metscales.py:17
k_N2, k_S2, k_S3 = 2.443, 2.381, 1              # metscales.py:19-22

#: which mds keys each calc* method may fill (subset that the GPU reports as present)
_GROUP_KEYS = {
    sc.T_D02: ["D02"], sc.T_Z94: ["Z94"], sc.T_M91: ["M91"], sc.T_PP04: ["PP04_N2Ha", "PP04_O3N2"],
    sc.T_P10: ["P10_ONS", "P10_ON"], sc.T_M08: ["M08_R23", "M08_N2Ha", "M08_O3O2"],
    sc.T_M13: ["M13_N2", "M13_O3N2"], sc.T_P05: ["P05"], sc.T_C01: ["C01_R23", "C01_N2S2"],
    sc.T_P01: ["P01"], sc.T_DP00: ["DP00"], sc.T_D16: ["D16"],
}


def _warn(string, logf, nps):
    from .metallicity import printsafemulti
    printsafemulti(string, logf, nps)


class diagnostics:
    """Same construction and public attributes as metscales.py:86-162."""

    def __init__(self, num, logf, nps):
        self.nm = num
        self.logf = logf
        self.nps = nps
        self.Ha = self.Hb = None
        for flag in ("hasHa hasHb hasO2 hasO3 hasS2 hasN2 hasO3Hb hasO3O2 hasN2O2 hasN2S2 hasS26731 "
                     "hasS39532 hasS39069 hasS2Hb").split():
            setattr(self, flag, False)
        self.O23727 = self.O35007 = self.O16300 = self.O34959 = None
        self.N26584 = self.S26717 = self.S26731 = self.S39069 = self.S39532 = None
        self.mds = {Z: None for Z in sc.KEYS}
        self._dust = sc.DUST_ON          # what calculation() decided (metallicity.py:113-132)
        self._dropneg = False            # calculation() ran with DROPNEGATIVES (metallicity.py:24, 101-102)
        self._noise = np.zeros((5, 1, 1))
        self.trips = (0, 0)              # trips of the KK04_N2Ha / KK04_R23 loops of the last run
        self.ret = None
        # intermediate quantities (metscales.py:100-147): filled by the calc* that set them in the reference
        self.logN2O2 = self.N2O2_roots = self.logN2S2 = self.N2S2 = None
        self.R2 = self.R3 = self.R23 = self.logR23 = None
        self.Z_init_guess = self.logq = None
        self._aux = None

    # ---- dust switch (metscales.py:265-271), per instance
    def setdustcorrect(self):
        self._dust = sc.DUST_ON

    def unsetdustcorrect(self):
        self._dust = sc.DUST_NONE

    # ---- setters (metscales.py:284-384): keep the arrays, set the has* flags like the reference
    def setHab(self, Ha, Hb):
        self.Ha, self.Hb = Ha, Hb
        self.hasHa = bool(np.sum(np.asarray(Ha) > 0))
        self.hasHb = bool(np.sum(np.asarray(Hb) > 0))

    def setOlines(self, O23727, O35007, O16300, O34959):
        self.O23727, self.O35007, self.O16300, self.O34959 = O23727, O35007, O16300, O34959
        self.hasO3 = bool(np.sum(np.asarray(O35007) > 0))
        self.hasO2 = bool(np.sum(np.asarray(O23727) > 0))
        self.hasO3Hb = self.hasHb and self.hasO3
        self.hasO3O2 = self.hasO2 and self.hasO3 and O34959 is not None and bool(np.sum(np.asarray(O34959) > 0) > 0)

    def setSII(self, S26717, S26731, S39069, S39532):
        if S26717 is not None and np.sum(np.asarray(S26717) > 0) > 0:
            self.S26717, self.hasS2 = S26717, True
        if S26731 is not None and np.sum(np.asarray(S26731) > 1e-9) > 0:
            self.S26731, self.hasS26731 = S26731, True
        if S39069 is not None and np.sum(np.asarray(S39069) > 1e-9) > 0:
            self.S39069, self.hasS39069 = S39069, True
        if S39532 is not None and np.sum(np.asarray(S39532) > 1e-9) > 0:
            self.S39532, self.hasS39532 = S39532, True

    def setNII(self, N26584):
        if N26584 is not None and np.sum(np.asarray(N26584) > 0):
            self.N26584, self.hasN2 = N26584, True
        self.hasN2O2 = self.hasN2 and self.hasO2
        self.hasN2S2 = self.hasN2 and self.hasS2

    def checkminimumreq(self, red_corr, ignoredust):                       # metscales.py:233-246
        if red_corr and not ignoredust:
            if not self.hasHa or not self.hasHb:
                return -1
        if not self.hasN2 and not (self.hasO2 and self.hasO3):
            return -1

    # ---- the GPU request
    def _n(self):
        for a in (self.Ha, self.Hb, self.O23727, self.O35007, self.N26584):
            if a is not None and np.ndim(a) == 1 and len(a) > 1:
                return len(a)
        return int(self.nm)

    def _flux(self):
        n = self._n()
        flux = np.full((sc.NLINES, 1, n), np.nan)
        lines = {'Ha': self.Ha, 'Hb': self.Hb, '[OII]3727': self.O23727, '[OIII]5007': self.O35007,
                 '[OI]6300': self.O16300, '[NII]6584': self.N26584, '[SII]6717': self.S26717,
                 '[SII]6731': self.S26731, '[SIII]9069': self.S39069, '[SIII]9532': self.S39532}
        for name, arr in lines.items():
            if arr is None:
                continue
            arr = np.asarray(arr, dtype=np.float64)
            if arr.shape == (n,):
                # DROPNEGATIVES: calculation() has already turned the negative samples into NaN in the
                # caller's arrays (after zeroing the non-finite ones); the kernel sanitises what it is
                # given, so they travel as negatives and become NaN again there (T_DROPNEG)
                flux[sc.LINES.index(name), 0] = np.where(np.isnan(arr), -1.0, arr) if self._dropneg else arr
        return flux

    def _run(self, tokens, keys=None):
        """Evaluate the token mask on the GPU; fill ``mds`` (all present keys, or only ``keys``)."""
        import torch
        from . import ops
        flux = self._flux()
        n = flux.shape[2]
        noise = self._noise if self._noise.shape == (5, 1, n) else np.zeros((5, 1, n))
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._dropneg:
            tokens |= sc.T_DROPNEG
        Z, present, trips, ret, aux = ops.forward_replay(torch.from_numpy(flux).to(dev), torch.from_numpy(noise).to(dev),
                                                         tokens, self._dust, "f64", return_aux=True)
        pres = ops.present_to_bool(present)[0].cpu().numpy()
        Zh = Z[0].cpu().numpy()
        self._aux = dict(zip(ops.AUX_NAMES, aux[0].cpu().numpy()))
        self.trips = tuple(int(t) for t in trips[0].cpu().numpy())
        self.ret = int(ret[0].item())
        for j, key in enumerate(sc.KEYS):
            if keys is not None and key not in keys:
                continue
            if pres[j]:
                self.mds[key] = Zh[j].copy()
        return self.ret

    def _draw(self, sigmas):
        """np.random.normal draws in the reference's order (metscales.py:679-680, 948-953)."""
        return [np.random.normal(0, s, self._n()) for s in sigmas]

    def _set_noise(self, rows, values):
        n = self._n()
        if self._noise.shape != (5, 1, n):
            self._noise = np.zeros((5, 1, n))
        for r, v in zip(rows, values):
            self._noise[r, 0] = v

    # ---- the reference's intermediate attributes, from the replay op's aux output
    def _say(self, string):
        _warn(string, self.logf, self.nps)

    def _take_niioii(self):                                                # metscales.py:402-423
        if self.hasN2 and self.hasO2:
            self.logN2O2 = self._aux["logN2O2"].copy()
            self.hasN2O2 = True
        if not self.hasN2O2 or np.mean(self.logN2O2) < -1.2:
            try:
                self._say('''WARNING: the KD02 and KK04 (+M08) methods should only be used for  log([NII]6564/[OII]3727) > -1.2, 
                the mean log([NII]6564/[OII]3727)= %f''' % np.mean(self.logN2O2))
            except TypeError:
                self._say('''WARNING: the KD02 and KK04 (+M08) methods 
                should only be used for  log([NII]6564/[OII]3727) >-1.2, 
                the mean log([NII]6564/[OII]3727)= %s''' % self.logN2O2)
        if not self.hasN2O2:
            self.N2O2_roots = np.zeros(self._n()) + float('NaN')
        else:
            # the REAL root pair of the quartic (descending, as np.roots lists it), NaN where the pair
            # is complex or outside the KD02_N2O2 window; the reference also keeps the complex pair(s),
            # which nothing reads (documented deviation)
            self.N2O2_roots = np.stack([self._aux["N2O2_root_hi"], self._aux["N2O2_root_lo"]], axis=1)

    def _take_niisii(self):                                                # metscales.py:393-399
        if self.hasS2 and self.hasN2:
            self.N2S2 = self._aux["N2S2"].copy()
            self.logN2S2 = self._aux["logN2S2"].copy()
            self.hasN2S2 = True
        else:
            self._say("WARNING: needs SII6717 and NII6584 to calculate NIISII: did you run setN2() and setS?")

    def _take_r23(self):                                                   # metscales.py:426-440
        self._say("calculating R23")
        if self.hasO3 and self.hasO2 and self.hasHb:
            self.R2, self.R3, self.R23 = (self._aux[k].copy() for k in ("R2", "R3", "R23"))
            self.logR23 = self.mds['logR23']
        else:
            self._say("WARNING: need O3, O2, Hb")

    def _take_initialguess(self):                                          # metscales.py:468-489
        self.Z_init_guess = self._aux["Z_init_guess"].copy()

    # ---- calc* (metscales.py:387-1356), each a request to the kernels
    def calcEB_V(self):
        self._say("calculating E(B-V)")
        return self._run(0, ["E(B-V)"])

    def calcNIIOII(self):
        self._run(0, [])
        self._take_niioii()

    def calcNIISII(self):
        self._run(0, [])
        self._take_niisii()

    def calcR23(self):
        self._run(0, ["logR23"])
        self._take_r23()

    def calcS23(self):
        self._say("calculating S23")
        return None          # folded into DP00 on the GPU (metscales.py:443-453)

    def initialguess(self):
        self._run(0, [])
        if self.hasN2 and self.hasO2 and not self.hasN2O2:
            self._say("WARNING: must calculate logN2O2 first")
            self._take_niioii()
        self._take_initialguess()

    def calcD02(self):
        self._say("calculating D02")
        self._set_noise([0, 1], self._draw([0.05, 0.1]))                   # metscales.py:679-680
        return self._run(sc.T_D02, _GROUP_KEYS[sc.T_D02])

    def calcPP04(self):
        self._say("calculating PP04")
        return self._run(sc.T_PP04, _GROUP_KEYS[sc.T_PP04])

    def calcZ94(self):
        self._say("calculating Z94")
        return self._run(sc.T_Z94, _GROUP_KEYS[sc.T_Z94] + ["logR23"])

    def calcP05(self):
        self._say("calculating P05")
        return self._run(sc.T_P05, _GROUP_KEYS[sc.T_P05] + ["logR23"])

    def calcP10(self):
        self._say("calculating P10")
        return self._run(sc.T_P10, _GROUP_KEYS[sc.T_P10])

    def calcP01(self):
        self._say("calculating old P05")
        return self._run(sc.T_P01, _GROUP_KEYS[sc.T_P01] + ["logR23"])

    def calcC01_ZR23(self):
        self._say("calculating C01")
        return self._run(sc.T_C01, _GROUP_KEYS[sc.T_C01])

    def calcM91(self):
        self._say("calculating M91")
        return self._run(sc.T_M91, _GROUP_KEYS[sc.T_M91] + ["logR23"])

    def calcM13(self):
        self._say("calculating M13")
        if not self.hasHa or not self.hasN2:                               # metscales.py:943-946
            _warn("WARNING: need O3, N2, Ha and Hb, or at least N2 and Ha", self.logf, self.nps)
            return -1
        self._set_noise([2, 3], self._draw([0.027, 0.024]))                # metscales.py:948-949
        if self.hasHb and self.hasO3:
            e1, _unused = self._draw([0.012, 0.012])                       # metscales.py:952-953
            self._set_noise([4], [e1])
        return self._run(sc.T_M13, _GROUP_KEYS[sc.T_M13])

    def calcM08(self, allM08=False):
        self._say("calculating M08")
        if allM08:                                                         # metscales.py:1019-1059
            return self._run(sc.T_M08ALL, _GROUP_KEYS[sc.T_M08] + ["M08_O3Hb", "M08_O2Hb", "M08_O3N2", "logR23"])
        return self._run(sc.T_M08, _GROUP_KEYS[sc.T_M08] + ["logR23"])

    def calcKD02_N2O2(self):
        self._say("calculating KD02_N2O2")
        self._run(sc.T_KD02, ["KD02_N2O2"])
        return 1

    def _loop_warnings(self, which):
        if self.trips[which] >= 100:                                       # metscales.py:1119-1120, 1179-1180
            self._say("WARNING: loop did not converge")

    def calcKK04_N2Ha(self):
        self._say("calculating KK04_N2Ha")
        r = self._run(sc.T_KD02, ["KD02_N2O2", "KK04_N2Ha"])
        if self.hasN2 and self.hasHa:
            self.logq = self._aux["logq"].copy()                           # metscales.py:1112, 1124
            self._loop_warnings(0)
        return r

    def calcKK04_R23(self):
        self._say("calculating KK04_R23")
        r = self._run(sc.T_KD02, ["KK04_R23", "logR23"])
        self._loop_warnings(1)
        return r

    def calcKDcombined(self):
        self._say("calculating KD_combined")
        # fills whatever of KD02_N2O2 / KK04_N2Ha / M91 / KK04_R23 is still None (metscales.py:1202-1216)
        return self._run(sc.T_KD02, ["KD02_N2O2", "KK04_N2Ha", "KK04_R23", "M91", "KD02comb", "logR23"])

    def calcDP00(self):
        self._say("calculating DP00")
        return self._run(sc.T_DP00, _GROUP_KEYS[sc.T_DP00])

    def calcD16(self):
        return self._run(sc.T_D16, _GROUP_KEYS[sc.T_D16])

    def calcpyqz(self, plot=False, allD13=False):
        _warn("WARNING: D13 needs the third-party pyqz package, which this build does not vendor", self.logf, self.nps)
        return -1

    def calcPM14(self):
        _warn("WARNING: PM14 needs the third-party HII-CHI-mistry script, which this build does not vendor",
              self.logf, self.nps)
        return -1

    def printme(self, verbose=False):                                      # metscales.py:164-231
        for name in ("Ha", "Hb", "O23727", "O35007"):
            arr = getattr(self, name)
            if arr is not None:
                print("\n" + name, np.mean(arr))
        for k, v in self.mds.items():
            if v is not None:
                print("\n", k, np.nanmean(v), np.nanstd(v))
