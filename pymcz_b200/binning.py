"""Histogram binning of the output distributions: the reference's rules (``getknuth``, ``knuthn``,
``getbinsize``; /root/reference/pyMCZ/mcz.py:90-116, 227-239) with the data-parallel inside -- the
equal-width counts for every candidate bin number and the moments -- taken from the GPU assist
(``ops.histogram_assist``), for all (measurement, scale) distributions of a run in ONE launch.
The optimiser (scipy ``fmin``), the ``gammaln`` posterior and the plots stay on the host, as
BASELINE.json's north_star says; nothing here is on the Monte-Carlo path itself.

Also the statistic of the sample-size diagnostic (testcompleteness.py:87-113): ``ks_table``.
"""
from __future__ import annotations

import numpy as np


def knuth_objective(N, m, nk):
    """-(log posterior) of Knuth's rule for m bins with counts nk (mcz.py:97)."""
    from scipy.special import gammaln
    return -(N * np.log(m) + gammaln(0.5 * m) - m * gammaln(0.5) - gammaln(N + 0.5 * m) + np.sum(gammaln(nk + 0.5)))


class KnuthTable:
    """Counts of every bin number 1 .. maxM for a batch of distributions, computed on the GPU once;
    ``getknuth(row, m)`` then is a table lookup with the reference's return values."""

    def __init__(self, Z, maxM=None):
        """Z: [rows, N] float64 array (host or cuda); non-finite entries are dropped per row, as
        savehist does (mcz.py:273)."""
        import torch
        from . import ops
        if not torch.is_tensor(Z):
            Z = torch.from_numpy(np.ascontiguousarray(Z, dtype=np.float64))
        if not Z.is_cuda:
            Z = Z.cuda()
        self.N_total = int(Z.shape[-1])
        if maxM is None:
            maxM = int(5 * np.sqrt(self.N_total)) + 1                          # knuthn's cap, mcz.py:106-107
        self.maxM = int(min(max(maxM, 1), 8192))
        stats, counts = ops.histogram_assist(Z.reshape(-1, self.N_total), range(1, self.maxM + 1))
        self.stats = stats.cpu().numpy()
        self.counts = [c.cpu().numpy() for c in counts]

    def n(self, row):
        return int(self.stats[row, 0])

    def getknuth(self, row, m):
        """getknuth(m, data, N) of mcz.py:90-99 for distribution ``row``."""
        m = int(m)
        N = self.n(row)
        if m > N or m < 1:          # (m < 1: np.linspace / np.log fail inside the reference's try -> [-1])
            return [-1]
        if m > self.maxM:
            raise ValueError("bin number %d beyond the table (maxM=%d)" % (m, self.maxM))
        return knuth_objective(N, m, self.counts[m - 1][row])

    def knuthn(self, row, maxM=None, disp=False):
        """knuthn(data, maxM) of mcz.py:102-116: (number of bins, 't' or 0)."""
        from scipy import optimize
        N = self.n(row)
        if not maxM:
            maxM = 5 * np.sqrt(N)
        m0 = 2.0 * N ** (1. / 3.)
        mkall = optimize.fmin(lambda m: self.getknuth(row, m[0] if np.ndim(m) else m), m0, disp=disp, maxiter=30)
        mk = mkall[0]
        if mk > maxM or mk < 0.3 * np.sqrt(N):
            return m0, 't'
        return mk, 0


def ks_2samp(x, y):
    """scipy.stats.ks_2samp(x, y) as the reference calls it (default two-sided, method 'auto';
    testcompleteness.py:111) with the statistic from the GPU: returns (D, p).

    The device gives max |F_x - F_y| over all data points, formed as scipy forms it.  scipy's 'exact'
    branch (taken when max(n1, n2) <= 10000) counts lattice paths for p; for larger samples it uses the asymptotic Kolmogorov distribution.
    Both are scalar host work and are taken from scipy itself."""
    import math
    import torch
    from scipy import stats
    from . import ops
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d = float(ops.ks_2samp(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())[0].item())
    n1, n2 = len(x), len(y)
    if max(n1, n2) <= 10000:                       # scipy's MAX_AUTO_N
        try:
            from scipy.stats._stats_py import _attempt_exact_2kssamp
            g = math.gcd(n1, n2)
            success, _d_lattice, p = _attempt_exact_2kssamp(n1, n2, g, d, 'two-sided')
            if success:
                return d, float(np.clip(p, 0, 1))
        except ImportError:
            pass
    m, n = sorted([float(n1), float(n2)], reverse=True)
    en = m * n / (m + n)
    return d, float(np.clip(stats.distributions.kstwo.sf(d, np.round(en)), 0, 1))


def ks_table(x_list):
    """(D, p) between consecutive samples of ``x_list`` (1-D arrays), as the reference's diagnostic
    compares each sub-sample with the previous one (testcompleteness.py:87-113)."""
    out = []
    prev = None
    for x in x_list:
        if prev is not None:
            out.append(ks_2samp(x, prev))
        prev = x
    return out
