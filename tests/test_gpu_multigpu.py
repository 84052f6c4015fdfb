"""Two-GPU check of the kernel-fused gather of the result tables (pymcz_b200.dist.GatheredTables):
identical, bit for bit, to an NCCL all_gather of the per-rank tables.  Needs two GPUs in the box;
skipped otherwise (the gloo world-size-2 test in test_host_logic.py covers the host-side sharding
logic on CPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_gather_matches_all_gather():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tests", "multigpu", "check_fused_gather.py"), "2000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "fused gather == all_gather: True" in out.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sample_sharded_single_object_run():
    """One object's samples over two GPUs (reductions inside the kernel over NVLink) == one GPU."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29542",
           os.path.join(ROOT, "tests", "multigpu", "check_sample_sharded.py"), "300001"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "True (N" in out.stdout and "bitwise True" in out.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """mcz_ctx_create(0) and mcz_ctx_create(1) in ONE process: the per-device kernel state (SM count,
    the 227 KB shared-memory opt-in of every kernel, the cooperative grid limit) is keyed by device,
    so the second device launches like the first and gives the same tables.  Shapes: the shared-memory
    kernel, the global-scratch variant and the cooperative wide kernel."""
    import numpy as np
    from pymcz_b200 import _lib, _scales as sc
    from pymcz_b200.synth import synthetic_table
    L = _lib.lib()
    for ngal, n in ((60, 2200), (20, 4700), (1, 70001)):
        meas, err = synthetic_table(max(ngal, 8), config=4)
        meas, err = np.ascontiguousarray(meas[:ngal]), np.ascontiguousarray(err[:ngal])
        out = []
        for dev in (0, 1):
            ctx = L.mcz_ctx_create(dev)
            assert ctx, L.mcz_last_error()
            pct = np.empty((ngal, sc.NKEYS, 3), np.float32); nvalid = np.empty((ngal, sc.NKEYS), np.int32)
            status = np.empty((ngal, sc.NKEYS), np.int8); trips = np.empty((ngal, 2), np.int32)
            ret = np.empty(ngal, np.int32)
            rc = L.mcz_ctx_run_host(ctx, meas.ctypes.data, err.ctypes.data, ngal, n, sc.T_ALL, sc.DUST_ON, 0xB200, 0, 0, 0,
                                    pct.ctypes.data, nvalid.ctypes.data, status.ctypes.data, trips.ctypes.data,
                                    ret.ctypes.data)
            assert rc == 0, (dev, L.mcz_last_error())
            L.mcz_ctx_destroy(ctx)
            out.append((pct, nvalid, status, trips))
        assert np.array_equal(out[0][0], out[1][0], equal_nan=True), (ngal, n)
        for k in (1, 2, 3):
            assert np.array_equal(out[0][k], out[1][k]), (ngal, n, k)
