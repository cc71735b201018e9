"""GPU, >= 2 devices: the NCCL row-sharded path (one process per GPU, launched like the driver does)
against the single-GPU result.  Skipped on single-GPU boxes; the gloo CPU test covers the sharding
arithmetic there."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_nccl_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "scripts", "dist_check.py"), "1000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dE ") == 2 and out.stdout.count("batched replicas ok") == 2 and out.stdout.count("sharded core gradient ok") == 2
    assert out.stdout.count("sharded host path ok") == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_nccl_only_matches_single_gpu():
    """The same check with the peer mapping switched off: NCCL all-reduce / all-gather everywhere."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29545", os.path.join(ROOT, "scripts", "dist_check.py"), "600"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, UPOL_NO_P2P="1"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("sharded host path ok") == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_periodic_box_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "scripts", "dist_periodic_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count(" OK") == 6 and "FAIL" not in out.stdout
