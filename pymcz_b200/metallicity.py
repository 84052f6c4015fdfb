"""Drop-in for the reference's ``metallicity`` module (/root/reference/pyMCZ/metallicity.py).

``calculation`` keeps the reference's signature, return value, in-place sanitising of
``measured``, sticky ``IGNOREDUST`` global and np.random consumption order -- and hands the
arithmetic to the CUDA replay op in ONE launch for the whole ``--md`` list.
"""
from __future__ import print_function

import numpy as np

from . import _scales as sc

IGNOREDUST = False        # metallicity.py:21 (sticky, set at :127)
MP = True
FIXNEGATIVES = True       # metallicity.py:23
DROPNEGATIVES = False     # metallicity.py:24: negative flux samples become NaN instead of 0 (wins over FIXNEGATIVES, :25-26)

Zs = list(sc.KEYS)        # metallicity.py:29-56
Zserr = list(sc.ERRKEYS)  # metallicity.py:59


def get_keys():           # metallicity.py:63
    return Zs


def get_errkeys():        # metallicity.py:67
    return Zserr


def printsafemulti(string, logf, nps):     # metallicity.py:71-78
    if nps == 1:
        logf.write(string + "\n")
    else:
        print(string)


def dispatch_order(toks):
    """The scales the reference's dispatch evaluates for an --md token list, in its order
    (metallicity.py:161-257), by the name each calc* logs ("calculating <name>")."""
    KD = ["KD02_N2O2", "KK04_N2Ha", "KK04_R23", "KD_combined"]
    order = []
    if 'all' in toks:
        order += ["D02", "Z94", "M91", "PP04", "P10", "M08", "M13"] + KD
    for t, names in (("DP00", ["DP00"]), ("P01", ["old P05"]), ("D02", ["D02"]), ("PP04", ["PP04"]), ("Z94", ["Z94"]),
                     ("M91", ["M91"]), ("P10", ["P10"]), ("M13", ["M13"])):
        if t in toks:
            order += names
    if 'M08all' in toks:
        order += ["M08", "other M08s"]
    elif 'M08' in toks:
        order += ["M08"]
    for t, names in (("P05", ["P05"]), ("C01", ["C01"]), ("KD02", KD), ("KD02comb", KD)):
        if t in toks:
            order += names
    return order


def calculation(mscales, measured, num, mds, nps, logf, dust_corr=True, disp=False, verbose=False):
    """metallicity.py:86-257.  ``measured``: dict line name -> float64 array of samples."""
    global IGNOREDUST
    mscales.setdustcorrect()                                               # metallicity.py:89
    raw_lines = {'[OIII]5007': np.array([float('NaN')]), 'Hb': np.array([float('NaN')])}
    for k in list(measured.keys()):                                        # metallicity.py:94-104
        measured[k][~(np.isfinite(measured[k][:]))] = 0.0
        if FIXNEGATIVES and not DROPNEGATIVES:
            measured[k][measured[k] < 0] = 0.0
        if DROPNEGATIVES:                                                  # metallicity.py:101-102
            measured[k][measured[k] < 0] = np.nan
        raw_lines[k] = measured[k]
    mscales._dropneg = bool(DROPNEGATIVES)
    raw_lines['[OIII]4959'] = raw_lines['[OIII]5007'] / 3.                 # metallicity.py:108
    mscales.setHab(raw_lines['Ha'], raw_lines['Hb'])                       # KeyError without 'Ha', as the reference

    if dust_corr and mscales.hasHa and mscales.hasHb:                      # metallicity.py:113-115
        dust = sc.DUST_ON
        printsafemulti("calculating E(B-V)", logf, nps)                    # metscales.py:388
    elif dust_corr and not IGNOREDUST:                                     # metallicity.py:116-129
        # the reference prompts on stdin when nps == 1; a library must not block: always warn
        print('''WARNING: reddening correction cannot be done
            without both H_alpha and H_beta measurement!!''')
        IGNOREDUST = True
        dust_corr = False
        dust = sc.DUST_ON                 # E(B-V) = 1e-5 *applied*: the kernel's fallback for DUST_ON
    else:                                                                  # metallicity.py:130-132
        mscales.unsetdustcorrect()
        dust = sc.DUST_ON_IGNORING if dust_corr else sc.DUST_NONE
    mscales._dust = dust

    n = len(raw_lines['Ha'])
    for k in ['[OII]3727', '[OIII]5007', '[OI]6300', '[OIII]4959', '[SII]6717', '[SII]6731',
              '[SIII]9069', '[SIII]9532', '[NII]6584']:                    # metallicity.py:134-139
        if k not in raw_lines or (len(raw_lines[k]) == 1 and np.isnan(raw_lines[k][0]) and n != 1):
            raw_lines[k] = None
    mscales.setOlines(raw_lines['[OII]3727'], raw_lines['[OIII]5007'], raw_lines['[OI]6300'], raw_lines['[OIII]4959'])
    mscales.setSII(raw_lines['[SII]6717'], raw_lines['[SII]6731'], raw_lines['[SIII]9069'], raw_lines['[SIII]9532'])
    mscales.setNII(raw_lines['[NII]6584'])

    if mscales.checkminimumreq(dust_corr, IGNOREDUST) == -1:               # metallicity.py:145-146
        mscales._run(0, ["E(B-V)"])
        return -1

    tokens, third, _unknown = sc.parse_md(mds)
    toks = mds.split(',')
    # coefficient-noise draws in the order the reference's dispatch makes them (metallicity.py:161-257)
    draws = []
    if 'all' in toks:
        draws += ['D02', 'M13']
        printsafemulti('''WARNING: CANNOT CALCULATE pyqz:
            this build does not vendor pyqz (D13 scales stay empty)\n''', logf, nps)
    if 'D02' in toks:
        draws.append('D02')
    if 'M13' in toks:
        draws.append('M13')
    for t in third:
        printsafemulti("WARNING: --md %s needs third-party code this build does not vendor" % t, logf, nps)
    for d in draws:
        if d == 'D02':
            mscales._set_noise([0, 1], mscales._draw([0.05, 0.1]))
        elif mscales.hasHa and mscales.hasN2:
            mscales._set_noise([2, 3], mscales._draw([0.027, 0.024]))
            if mscales.hasHb and mscales.hasO3:
                e1, _unused = mscales._draw([0.012, 0.012])
                mscales._set_noise([4], [e1])
    mscales._run(tokens)
    # the always-run precompute of metallicity.py:148-154: attributes, warnings and log lines as the
    # reference's calcNIIOII / calcNIISII / calcR23 / initialguess leave them
    mscales._take_niioii()
    mscales._take_niisii()
    mscales._take_r23()
    mscales._take_initialguess()
    if verbose:
        print("calculating metallicity diagnostic scales: ", toks)
    for name in dispatch_order(toks):
        printsafemulti("calculating " + name, logf, nps)
        if name == "KK04_N2Ha" and mscales.hasN2 and mscales.hasHa:
            mscales.logq = mscales._aux["logq"].copy()                     # metscales.py:1112, 1124
            mscales._loop_warnings(0)
        elif name == "KK04_R23":
            mscales._loop_warnings(1)
    return None
