"""Generate pymcz_b200/csrc/mcz_root_tables.h: closed-form replacements for the reference's
per-sample ``np.roots`` calls (/root/reference/pyMCZ/metscales.py:248-263, callers :423, :979, :997).

Both quartics have FIXED coefficients except the constant term, i.e. the roots solve f(x) = L for
a fixed quartic f with exactly one real extremum x0 (checked below).  With t = x - x0 and
f(x0+t) - f(x0) = sgn * t^2 * h(t),  h(t) = |d2 + d3 t + d4 t^2| > 0, the two real roots are
t(+s) and t(-s) with s = sqrt(sgn*(L - f(x0))), where t(s) is the analytic inverse of
s = t*sqrt(h(t)).  We tabulate t(s) as a polynomial in u = (s-mid)/half on the s-range the
reference's validity windows allow; float32 uses it as is (max error printed below), float64
polishes with two Newton steps on s = t*sqrt(h(t)).

Run:  python tools/gen_root_tables.py   (needs mpmath; the output header is committed)
"""
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "pymcz_b200", "csrc", "mcz_root_tables.h")

N2O2 = ["1106.8660", "-532.15451", "96.373260", "-7.8106123", "0.23928247"]   # metscales.py:146,419
M08_N2HA = ["-0.7732", "1.2357", "-0.2811", "-0.7201", "-0.3330"]             # metscales.py:62
M08_O3N2 = ["0.4520", "-2.6096", "-0.7170", "0.1347", "0"]                    # metscales.py:66 (cubic)


def as_double_exact(strs):
    """The reference parses these literals as float64: use exactly those doubles."""
    return [mp.mpf(float(s)) for s in strs]


def real_extremum(c, window=None):
    d = [c[i] * i for i in range(1, 5)]
    while d and d[-1] == 0:
        d.pop()                                   # cubic padded to five coefficients
    roots = mp.polyroots(d[::-1], maxsteps=200, extraprec=200)
    real = [mp.re(r) for r in roots if abs(mp.im(r)) < mp.mpf(10) ** -40]
    if window is not None:                        # a cubic has two extrema: take the one in the window
        real = [r for r in real if window[0] <= r <= window[1]]
    assert len(real) == 1, roots
    return real[0]


def recentre(c, x0):
    """Taylor coefficients of f about x0 (f(x0), f'(x0)~0, d2, d3, d4)."""
    out = []
    for k in range(5):
        s = mp.mpf(0)
        for i in range(k, 5):
            s += c[i] * mp.binomial(i, k) * x0 ** (i - k)
        out.append(s)
    return out


def fval(c, x):
    return sum(c[i] * mp.mpf(x) ** i for i in range(5))


def fit(d, sgn, s_lo, s_hi, deg):
    """Chebyshev-node least-squares polynomial for t(s) on [s_lo, s_hi] in u=(s-mid)/half."""
    d2, d3, d4 = [float(sgn * v) for v in d[2:5]]

    def s_of_t(t):
        return t * np.sqrt(d2 + d3 * t + d4 * t * t)

    def t_of_s(s):
        # solve by bisection/Newton in float64 with mp fallback
        t = mp.findroot(lambda tt: tt * mp.sqrt(sgn * (d[2] + d[3] * tt + d[4] * tt * tt)) - s,
                        mp.mpf(s) / mp.sqrt(sgn * d[2]))
        return float(t)
    mid, half = 0.5 * (s_lo + s_hi), 0.5 * (s_hi - s_lo)
    nodes = np.cos(np.pi * (np.arange(400) + 0.5) / 400)
    tv = np.array([t_of_s(mid + half * u) for u in nodes])
    cheb = np.polynomial.chebyshev.chebfit(nodes, tv, deg)
    mono = np.polynomial.chebyshev.cheb2poly(cheb)
    # error on a dense grid, float64 and float32 Horner
    ug = np.linspace(-1, 1, 4001)
    tg = np.array([t_of_s(mid + half * u) for u in ug[::8]])
    p64 = np.polynomial.polynomial.polyval(ug[::8], mono)
    u32 = ug[::8].astype(np.float32)
    acc = np.zeros_like(u32) + np.float32(mono[-1])
    for cf in mono[-2::-1]:
        acc = acc * u32 + np.float32(cf)
    return mid, half, mono, np.max(np.abs(p64 - tg)), np.max(np.abs(acc.astype(np.float64) - tg))


def emit(name, c_strs, lo_x, hi_x, breaks, deg, window=None):
    c = as_double_exact(c_strs)
    x0 = real_extremum(c, window)
    d = recentre(c, x0)
    sgn = 1 if d[2] > 0 else -1
    f0 = d[0]
    f_lo, f_hi = fval(c, lo_x), fval(c, hi_x)
    s_lo = -float(mp.sqrt(sgn * (f_lo - f0)))
    s_hi = float(mp.sqrt(sgn * (f_hi - f0)))
    # a margin so that values a few ulp past a window edge still evaluate sanely
    edges = [s_lo - 0.02] + list(breaks) + [s_hi + 0.02]
    P = name.upper()
    lines = ["// %s: f(x)=L, extremum x0, f(x0+t)-f0 = %+d * t^2 * h(t); s in [%.6f, %.6f]" % (name, sgn, s_lo, s_hi)]
    lines.append("#define %s_X0   %s" % (P, mp.nstr(x0, 20)))
    lines.append("#define %s_F0   %s" % (P, mp.nstr(f0, 20)))
    lines.append("#define %s_SGN  %d.0" % (P, sgn))
    lines.append("#define %s_H0   %s" % (P, mp.nstr(sgn * d[2], 20)))
    lines.append("#define %s_H1   %s" % (P, mp.nstr(sgn * d[3], 20)))
    lines.append("#define %s_H2   %s" % (P, mp.nstr(sgn * d[4], 20)))
    lines.append("#define %s_F_LO %s   // f(%s)" % (P, mp.nstr(f_lo, 20), lo_x))
    lines.append("#define %s_F_HI %s   // f(%s)" % (P, mp.nstr(f_hi, 20), hi_x))
    lines.append("#define %s_NSEG %d" % (P, len(edges) - 1))
    lines.append("#define %s_DEG %d" % (P, deg))
    worst64 = worst32 = 0.0
    for k, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
        mid, half, mono, e64, e32 = fit(d, sgn, a, b, deg)
        worst64, worst32 = max(worst64, e64), max(worst32, e32)
        lines.append("// segment %d: s in [%.6f, %.6f): max |err| float64 %.3g, float32 Horner %.3g" % (k, a, b, e64, e32))
        lines.append("#define %s_SEG%d_HI %.17g" % (P, k, b))
        lines.append("#define %s_SEG%d_MID %.17g" % (P, k, mid))
        lines.append("#define %s_SEG%d_RHALF %.17g" % (P, k, 1.0 / half))
        lines.append("#define %s_SEG%d_POLY(X) \\" % (P, k))
        lines.append("    " + " ".join("X(%.17g)" % v for v in mono))
    print(name, "x0", mp.nstr(x0, 17), "f0", mp.nstr(f0, 17), "sgn", sgn, "s range", s_lo, s_hi,
          "worst err64 %.3g err32 %.3g" % (worst64, worst32))
    return "\n".join(lines)


def main():
    parts = ["// GENERATED by tools/gen_root_tables.py -- do not edit.",
             "// Closed-form root tables replacing np.roots for the KD02 [NII]/[OII] quartic",
             "// (/root/reference/pyMCZ/metscales.py:419-423), the M08 [NII]/Ha quartic (:995-999) and the",
             "// M08 [OIII]/[NII] cubic of --md M08all (:1051-1059).",
             "#pragma once", ""]
    # KD02_N2O2: valid |root| in [7.5, 9.4] (metscales.py:1086)
    parts.append(emit("rt_n2o2", N2O2, "7.5", "9.4", [0.55], 10))
    parts.append("")
    # M08_N2Ha: valid root+8.69 in [7.1, 9.4] (metscales.py:998): x in [-1.59, 0.71]
    lo = mp.mpf(float("7.1")) - mp.mpf(float("8.69"))
    hi = mp.mpf(float("9.4")) - mp.mpf(float("8.69"))
    parts.append(emit("rt_m08n2", M08_N2HA, mp.nstr(lo, 25), mp.nstr(hi, 25), [-1.3, -1.0, -0.6], 10))
    parts.append("")
    # M08_O3N2 (M08all, metscales.py:1051-1059): cubic, local maximum inside the window; the other
    # extremum (x = 4.87) and the third root (x ~ 7.4 .. 8) are far outside [7.1, 9.4] - 8.69
    parts.append(emit("rt_m08o3n2", M08_O3N2, mp.nstr(lo, 25), mp.nstr(hi, 25), [0.6, 1.3], 10,
                      window=(lo, hi)))
    parts.append("")
    open(OUT, "w").write("\n".join(parts))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
