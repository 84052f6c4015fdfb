"""CPU check of the per-sample CUDA physics (csrc/mcz_physics.cuh compiled for the host by g++)
against the oracle, in double (replay arithmetic) and float (production arithmetic).

This harness is test-only: it lets the formulas be debugged without a GPU.  The GPU parity tests
(test_gpu_*.py) are the ones that exercise the real kernels through the C ABI.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import mcz_oracle as orc
from pymcz_b200 import _scales as sc
from pymcz_b200.synth import synthetic_table
from tests._helpers import oracle_galaxy, compare_samples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostemu") / "hostemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")])
    lib = ctypes.CDLL(out)
    lib.hostemu_galaxy.restype = ctypes.c_int
    return lib


def run_emu(lib, use_float, flux, noise, tok, dust):
    flux = np.ascontiguousarray(flux, dtype=np.float64)
    noise = np.ascontiguousarray(noise, dtype=np.float64)
    n = flux.shape[1]
    Z = np.full((sc.NKEYS, n), np.nan)
    present = ctypes.c_ulonglong(0)
    trips = (ctypes.c_int * 2)()
    ret = ctypes.c_int(0)
    P = ctypes.POINTER(ctypes.c_double)
    lib.hostemu_galaxy(int(use_float), flux.ctypes.data_as(P), noise.ctypes.data_as(P), n,
                       ctypes.c_uint(tok), int(dust), Z.ctypes.data_as(P), ctypes.byref(present),
                       trips, ctypes.byref(ret))
    pm = np.array([(present.value >> k) & 1 for k in range(sc.NKEYS)], dtype=bool)
    Z[~pm] = np.nan
    return pm, Z, (trips[0], trips[1]), ret.value


def tables():
    ex_m = np.array([[2.875, 1.251, 0.168, 0.064, np.nan, 4.026, 0.781, 0.821, 0.573, np.nan, np.nan],
                     [1.842, 0.958, np.nan, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, np.nan, np.nan]], dtype=np.float32)
    ex_e = np.array([[0.1005, 0.0435, 0.028, 0.0245, 0.00, 0.069, 0.0335, 0.0345, 0.0315, 0.00, 0.00],
                     [0.053, 0.032, 0.00, 0.029, 0.0205, 0.0255, 0.022, 0.0205, 0.0185, 0.00, 0.00]], dtype=np.float32)
    sm, se = synthetic_table(60, config=4)
    return np.concatenate([ex_m, sm]), np.concatenate([ex_e, se])


@pytest.mark.parametrize("use_float", [0, 1])
@pytest.mark.parametrize("mds,dust", [("all", True), ("all", False), ("KD02", True),
                                      ("P05,C01,P01,DP00,D16,M08all,D02,M13,Z94,PP04,P10,M91", True)])
def test_hostemu_vs_oracle(emu, use_float, mds, dust):
    meas, err = tables()
    rs = np.random.RandomState(11)
    tok, _, _ = sc.parse_md(mds)
    tot_mism = tot = 0
    worst = 0.0
    for g in range(meas.shape[0]):
        z = rs.normal(0, 1, 300)
        flux = orc.perturb(meas[g], err[g], z, "shuffled", rs=rs)
        po, Zo, noise, trips_o, ret_o = oracle_galaxy(flux, mds, dust, rs=rs)
        pm, Z, trips, ret = run_emu(emu, use_float, flux, noise, tok, sc.DUST_ON if dust else sc.DUST_NONE)
        assert ret == ret_o
        mism, inf_mism, maxdiff, same_inf = compare_samples(Z, Zo, pm, po, what="galaxy %d" % g)
        assert same_inf and inf_mism == 0
        if not use_float:
            # double arithmetic: NaN masks bit-exact, trip counts identical, values ~1e-12
            assert mism == 0, (g, mism)
            assert trips == trips_o, (g, trips, trips_o)
            assert maxdiff < 1e-9, (g, maxdiff)
        else:
            # float arithmetic: a sample within float rounding of a validity threshold may flip
            tot_mism += mism
            tot += Z.size
            if trips == trips_o:
                assert maxdiff < 1e-3, (g, maxdiff)
        worst = max(worst, maxdiff)
    if use_float:
        assert tot_mism <= 2e-5 * tot, (tot_mism, tot)
    print("worst |dZ| = %.3g" % worst)


def test_hostemu_missing_lines(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "missing_lines_n50.npz"))
    for i in range(g["meas"].shape[0]):
        np.random.seed(int(g["seed"][i]))
        z = orc.draw_samples(1, int(g["nsample"]))[0]
        flux = orc.perturb(g["meas"][i], g["err"][i], z, "correlated")
        mds = str(g["mds"][i])
        po, Zo, noise, trips_o, ret_o = oracle_galaxy(flux, mds, True)
        assert np.array_equal(Zo, g["Z"][i], equal_nan=True)       # oracle == reference golden
        tok, _, _ = sc.parse_md(mds)
        pm, Z, trips, ret = run_emu(emu, 0, flux, noise, tok, sc.DUST_ON)
        assert ret == ret_o == g["ret"][i]
        mism, inf_mism, maxdiff, same_inf = compare_samples(Z, Zo, pm, po, what="row %d" % i)
        assert mism == 0 and inf_mism == 0 and same_inf and maxdiff < 1e-9, (i, mism, maxdiff)
        assert trips == trips_o


def test_philox_known_answers(emu):
    """Random123 kat_vectors for philox4x32-10, against the header the kernel compiles."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    from tests._helpers import philox4x32_10
    for c, k, want in kats:
        out = (ctypes.c_uint * 4)()
        emu.hostemu_philox(*[ctypes.c_uint(x) for x in c], ctypes.c_uint(k[0]), ctypes.c_uint(k[1]), out)
        assert tuple(out) == want
        model = philox4x32_10(*[np.array([x], np.uint32) for x in c], k[0], k[1])
        assert tuple(int(v[0]) for v in model) == want


REF_COEFS = {0: [1106.8660, -532.15451, 96.373260, -7.8106123, 0.23928247],      # metscales.py:146, 419
             1: [-0.7732, 1.2357, -0.2811, -0.7201, -0.3330],                    # metscales.py:62
             2: [-0.2839, -1.3881, -0.3172],                                     # metscales.py:64
             3: [0.4520, -2.6096, -0.7170, 0.1347]}                              # metscales.py:66


def _reference_selection(which, y):
    """np.roots + the reference's selection rule for one sample (metscales.py:979-990, 1085-1087)."""
    c = np.array(REF_COEFS[which], dtype=np.float64)
    c[0] -= y
    r = np.roots(c[::-1])
    if which == 0:                                   # last root with |r| in [7.5, 9.4], imag == 0
        out = np.nan
        for k in range(4):
            if abs(r[k]) >= 7.5 and abs(r[k]) <= 9.4 and r[k].imag == 0:
                out = abs(r[k])
        return out
    s = r + 8.69                                     # first root with real in [7.1, 9.4], imag == 0
    for k in range(len(s)):
        if 7.1 <= s[k].real <= 9.4 and s[k].imag == 0:
            return s[k].real
    return np.nan


@pytest.mark.parametrize("which,lo,hi", [(0, -2.5, 2.5), (1, -3.5, 0.5), (2, -3.0, 2.0), (3, -2.5, 3.0)])
def test_closed_form_roots_vs_np_roots(emu, which, lo, hi):
    """Dense scan of every closed-form root selection against np.roots with the reference's rule,
    including both sides of every window edge; float64 to 1e-9, float32 to 2e-6."""
    y = np.linspace(lo, hi, 6001)
    want = np.array([_reference_selection(which, v) for v in y])
    P = ctypes.POINTER(ctypes.c_double)
    for use_float, tol in ((0, 1e-9), (1, 3e-6)):
        got = np.empty_like(y)
        lib_y = np.ascontiguousarray(y)
        emu.hostemu_roots(which, use_float, lib_y.ctypes.data_as(P), len(y), got.ctypes.data_as(P))
        nan_w, nan_g = np.isnan(want), np.isnan(got)
        flips = np.nonzero(nan_w != nan_g)[0]
        # a grid point within rounding of a window edge / the double root may flip; nothing else
        assert len(flips) <= (2 if not use_float else 6), (which, use_float, y[flips][:8])
        both = ~nan_w & ~nan_g
        # near the double root np.roots itself is only good to ~1e-8: compare away from it
        err = np.abs(got[both] - want[both])
        assert np.percentile(err, 99.5) < tol * 50 and np.median(err) < tol, (which, use_float, err.max())
