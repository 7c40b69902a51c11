"""SURVEY T10 on real GPUs: the striped multi-GPU canvas (NCCL neighbour strip exchange, overlapped with the interior
tile solves) equals the single-GPU canvas BIT FOR BIT.  Needs >= 2 GPUs on the box (skipped otherwise; the gloo
world-size 2/3 version of the same protocol runs in tests/test_sharded_cpu.py)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on one box")
def test_sharded_canvas_equals_single_gpu_bit_exact_over_nccl():
    n = 2 if torch.cuda.device_count() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tools" / "check_sharded_gpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT),
                         env={**os.environ, "MASTER_ADDR": "127.0.0.1"})
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("sharded == single-GPU: True") == n, res.stdout[-2000:]
