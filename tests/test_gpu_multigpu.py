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
