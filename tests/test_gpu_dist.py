"""NCCL frame-sharding parity, driver-run: launches ``tests/dist_check.py`` under torchrun on every
visible GPU (2..8) and requires its verdict.  Skips on a single-GPU box (the gloo world-size-2 test
in ``test_distributed_cpu.py`` covers the host-side combine logic there)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_nccl_sharded_fit_equals_single_gpu_fit():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = Path(__file__).resolve().parent / "dist_check.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 8)}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    assert "-> OK" in run.stdout
