"""Names shared by the host side: output keys, line order, ``--md`` tokens -> kernel bit mask.

Mirrors /root/reference/pyMCZ/metallicity.py:29-56 (keys), mcz.py:45 (lines) and the token
dispatch of metallicity.py:155-257.  The numeric values must match csrc/mcz_common.cuh.
"""
from __future__ import annotations

LINES = ['[OII]3727', 'Hb', '[OIII]4959', '[OIII]5007', '[OI]6300', 'Ha', '[NII]6584',
         '[SII]6717', '[SII]6731', '[SIII]9069', '[SIII]9532']

KEYS = ["E(B-V)", "logR23", "D02", "Z94", "M91", "C01_N2S2", "C01_R23", "P05", "P01",
        "PP04_N2Ha", "PP04_O3N2", "DP00", "P10_ONS", "P10_ON",
        "M08_R23", "M08_N2Ha", "M08_O3Hb", "M08_O2Hb", "M08_O3O2", "M08_O3N2",
        "M13_O3N2", "M13_N2",
        "D13_N2S2_O3S2", "D13_N2S2_O3Hb", "D13_N2S2_O3O2", "D13_N2O2_O3S2",
        "D13_N2O2_O3Hb", "D13_N2O2_O3O2", "D13_N2Ha_O3Hb", "D13_N2Ha_O3O2",
        "KD02_N2O2", "KD02_N2S2", "KK04_N2Ha", "KK04_R23", "KD02comb", "PM14", "D16"]
KEY_INDEX = {k: i for i, k in enumerate(KEYS)}
ERRKEYS = ['PM14err']
NKEYS = len(KEYS)
NLINES = len(LINES)

# token bits (csrc/mcz_common.cuh :: enum Token)
T_D02, T_Z94, T_M91, T_PP04, T_P10, T_M08, T_M08ALL, T_M13, T_KD02, T_P05, T_C01, T_P01, T_DP00, T_D16 = (
    1 << i for i in range(14))
T_ALL = T_D02 | T_Z94 | T_M91 | T_PP04 | T_P10 | T_M08 | T_M13 | T_KD02
#: not a scale: the reference's DROPNEGATIVES sanitise mode (metallicity.py:23-26) rides in the token mask
T_DROPNEG = 1 << 30

_TOKEN_BITS = {
    'all': T_ALL, 'D02': T_D02, 'Z94': T_Z94, 'M91': T_M91, 'PP04': T_PP04, 'P10': T_P10,
    'M08': T_M08, 'M08all': T_M08ALL, 'M13': T_M13, 'KD02': T_KD02, 'P05': T_P05, 'C01': T_C01, 'P01': T_P01,
    'DP00': T_DP00, 'D16': T_D16,
    # README.md:55-56 advertises "KD02comb"; the reference's dispatch matches nothing for it
    # (metallicity.py:251).  Documented deviation: treated as an alias of KD02.
    'KD02comb': T_KD02,
}
#: tokens whose arithmetic lives in third-party code the reference does not vendor
THIRD_PARTY_TOKENS = ('D13', 'D13all', 'PM14')

DUST_NONE, DUST_ON, DUST_ON_IGNORING = 0, 1, 2
ST_OK, ST_ABSENT, ST_EMPTY, ST_NONPOS, ST_NODIST = 0, 1, 2, 3, 4


def parse_md(mds: str):
    """``--md`` string -> (token bit mask, third-party tokens seen, unknown tokens).

    Matching is by exact token membership after ``split(',')`` like metallicity.py:155-257.
    """
    mask, third, unknown = 0, [], []
    for tok in mds.split(','):
        if tok in _TOKEN_BITS:
            mask |= _TOKEN_BITS[tok]
        elif tok in THIRD_PARTY_TOKENS:
            third.append(tok)
        else:
            unknown.append(tok)
    return mask, third, unknown
