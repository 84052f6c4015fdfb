"""The production kernel's percentile selection (csrc/mcz_select.cuh) on columns the physics never
produces: run through the test entry point mcz_debug_select_columns, which stages each column the
way phase D sees it and calls the same warp_percentiles().  The bar is exactness:
np.percentile(finite values in float64, [50, 16, 84]) cast to float32 (reference mcz.py:273-301)."""
import ctypes

import numpy as np
import pytest
import torch

from pymcz_b200 import _lib
from pymcz_b200 import _scales as sc

pytestmark = pytest.mark.gpu


def select(cols):
    """cols [ncols, N] float32 -> pct [ncols, 3], nvalid, status through the C ABI."""
    L = _lib.lib()
    fn = L.mcz_debug_select_columns
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    c = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.float32)).cuda()
    n, N = c.shape
    pct = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    nv = torch.empty((n,), dtype=torch.int32, device="cuda")
    st = torch.empty((n,), dtype=torch.int8, device="cuda")
    pad = torch.empty((n, (N + 127) // 128 * 128), dtype=torch.float32, device="cuda") if N > 2304 else None
    rc = fn(c.data_ptr(), n, N, pct.data_ptr(), nv.data_ptr(), st.data_ptr(),
            pad.data_ptr() if pad is not None else None, None)
    _lib.check(rc, "mcz_debug_select_columns")
    torch.cuda.synchronize()
    return pct.cpu().numpy(), nv.cpu().numpy(), st.cpu().numpy()


def expect(col):
    fin = col[np.isfinite(col)].astype(np.float64)
    if fin.size == 0:
        return None, 0, sc.ST_EMPTY
    if not (np.sum(fin) > 0):
        return None, fin.size, sc.ST_NONPOS
    return np.percentile(fin, [50, 16, 84]).astype(np.float32), fin.size, sc.ST_OK


def adversarial_columns(N, rs):
    cols, names = [], []

    def add(name, x):
        x = np.asarray(x, dtype=np.float32)
        assert x.shape == (N,)
        cols.append(x)
        names.append(name)

    g = rs.normal(8.5, 0.1, N)
    add("gaussian", g)
    add("sorted", np.sort(g))
    add("reverse sorted", np.sort(g)[::-1])
    add("constant", np.full(N, 8.69))
    add("constant tiny", np.full(N, 1e-5))
    add("two values", np.where(rs.rand(N) < 0.5, 8.1, 8.9))
    add("two values 90/10", np.where(rs.rand(N) < 0.9, 8.1, 8.9))
    add("three values", rs.choice([7.5, 8.0, 9.25], N))
    add("mostly clamp", np.where(rs.rand(N) < 0.97, 1e-5, np.abs(rs.normal(0.2, 0.1, N))))
    add("half clamp", np.where(rs.rand(N) < 0.5, 1e-5, np.abs(rs.normal(0.2, 0.1, N)) + 1e-4))
    add("clamp absent from the first 64", np.concatenate([rs.normal(0.3, 0.01, min(64, N)),
                                                           np.full(max(N - 64, 0), 1e-5)])[:N])
    add("first 64 constant, rest spread", np.concatenate([np.full(min(64, N), 8.0), rs.normal(8.5, 0.3, max(N - 64, 0))])[:N])
    add("bimodal narrow peaks", np.where(rs.rand(N) < 0.5, rs.normal(7.6, 1e-4, N), rs.normal(8.8, 1e-4, N)))
    add("bimodal 1e-6 peaks", np.where(rs.rand(N) < 0.3, rs.normal(7.6, 1e-6, N), rs.normal(8.8, 1e-6, N)))
    add("spike plus wide", np.where(rs.rand(N) < 0.6, rs.normal(8.5, 1e-5, N), rs.uniform(7, 9.5, N)))
    add("heavy tails", 8.5 + 0.05 * rs.standard_cauchy(N))
    add("huge outliers", np.where(rs.rand(N) < 0.02, 1e30, g))
    add("huge negative outliers", np.where(rs.rand(N) < 0.02, -1e30, g))
    add("wild first 64", np.concatenate([rs.uniform(-1e20, 1e20, min(64, N)), rs.normal(8.5, 0.1, max(N - 64, 0))])[:N])
    add("uniform", rs.uniform(7, 9.5, N))
    add("exponential", rs.exponential(1.0, N))
    add("lognormal wide", rs.lognormal(0, 5, N))
    add("denormals", rs.uniform(1e-45, 1e-39, N))
    add("near float max", rs.uniform(1e38, 3e38, N))
    add("integers with ties", rs.randint(0, 40, N).astype(np.float32) + 1)
    add("all 64-sample prefix NaN", np.concatenate([np.full(min(64, N), np.nan), rs.normal(8.5, 0.1, max(N - 64, 0))])[:N])
    add("90 % NaN", np.where(rs.rand(N) < 0.9, np.nan, g))
    add("one finite", np.concatenate([[8.25], np.full(N - 1, np.nan)]) if N > 1 else np.array([8.25]))
    add("two finite", np.concatenate([np.full(N - 2, np.nan), [8.25, 9.0]]) if N > 2 else np.full(N, 8.25))
    add("all NaN", np.full(N, np.nan))
    add("infs are not data", np.where(rs.rand(N) < 0.1, np.inf, np.where(rs.rand(N) < 0.1, -np.inf, g)))
    add("negative sum", rs.normal(-0.3, 0.1, N))
    add("zero sum exactly", np.zeros(N))
    add("mixed sign positive sum", rs.normal(0.05, 1.0, N) + 0.2)
    add("steps of one ulp", np.float32(8.5) + np.arange(N) % 7 * np.spacing(np.float32(8.5)))
    add("geometric", 1.0001 ** np.arange(N))
    return np.stack(cols), names


@pytest.mark.parametrize("N", [1, 2, 5, 63, 64, 65, 127, 128, 129, 700, 2199, 2200, 2304, 2305, 11000, 70000])
def test_selection_is_exact_on_adversarial_columns(N):
    rs = np.random.RandomState(1000 + N)
    cols, names = adversarial_columns(N, rs)
    pct, nv, st = select(cols)
    for i, name in enumerate(names):
        want, n, status = expect(cols[i])
        assert st[i] == status, (N, name, st[i], status)
        if status != sc.ST_EMPTY or True:
            assert nv[i] == n or status == sc.ST_EMPTY, (N, name, nv[i], n)
        if want is None:
            assert np.all(np.isnan(pct[i])), (N, name, pct[i])
        else:
            assert np.array_equal(pct[i], want), (N, name, pct[i], want)


def test_selection_random_mixtures():
    """2000 random mixtures of point masses, narrow and wide components, NaN fractions."""
    rs = np.random.RandomState(77)
    N = 2200
    cols = np.empty((2000, N), dtype=np.float32)
    for i in range(cols.shape[0]):
        k = rs.randint(1, 5)
        w = rs.dirichlet(np.ones(k))
        comp = rs.choice(k, N, p=w)
        mu = rs.uniform(7, 9.5, k)
        sig = 10.0 ** rs.uniform(-7, 0, k) * (rs.rand(k) < 0.8)      # some components are point masses
        x = mu[comp] + sig[comp] * rs.normal(0, 1, N)
        x[rs.rand(N) < rs.choice([0, 0, 0.05, 0.5, 0.95])] = np.nan
        cols[i] = x
    pct, nv, st = select(cols)
    for i in range(cols.shape[0]):
        want, n, status = expect(cols[i])
        assert st[i] == status and (status == sc.ST_EMPTY or nv[i] == n), i
        if want is not None:
            assert np.array_equal(pct[i], want), (i, pct[i], want)
