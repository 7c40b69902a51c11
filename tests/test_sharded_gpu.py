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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on one box")
def test_launches_follow_the_tensors_device_not_the_current_device():
    """libtdx launches on the CURRENT device; `_lib.call` makes the tensor's device current for the call.  A
    model and a canvas on cuda:1 driven while cuda:0 is current must give what they give with cuda:1 current."""
    from oracle import unet as ounet
    from terrain_diffusion_b200.inference.canvas import BlendCanvas
    from terrain_diffusion_b200.inference.tiling import linear_weight_window
    from terrain_diffusion_b200.models import EDMUnet2D
    cfg = ounet.DECODER_CFG
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
    m = m.to("cuda:1")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 5, 64, 64, generator=g).to("cuda:1")
    t = torch.full((1,), 1.1, device="cuda:1")
    torch.cuda.set_device(0)
    y0 = m(x, t, []).cpu()
    cv = BlendCanvas(1, 96, 96, "cuda:1")
    cv.accumulate(torch.ones(1, 64, 64, device="cuda:1"), 16, 16, linear_weight_window(64, "cuda:1"))
    n0 = cv.normalized().cpu()
    with torch.cuda.device(1):
        y1 = m(x, t, []).cpu()
        cv1 = BlendCanvas(1, 96, 96, "cuda:1")
        cv1.accumulate(torch.ones(1, 64, 64, device="cuda:1"), 16, 16, linear_weight_window(64, "cuda:1"))
        n1 = cv1.normalized().cpu()
    assert torch.equal(y0, y1) and torch.isfinite(y0).all()
    assert torch.equal(n0[:, 16:80, 16:80], n1[:, 16:80, 16:80])
