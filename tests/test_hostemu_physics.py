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
                                      ("P05,C01,P01,DP00,D16,M08,D02,M13,Z94,PP04,P10,M91", True)])
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
