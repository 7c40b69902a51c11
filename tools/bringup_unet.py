"""GPU bring-up of the whole U-Net program: per-stage comparison against the fp32 oracle (oracle/unet.py).

Run on the GPU box:  python tools/bringup_unet.py [size] [batch]
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle import unet as O
from terrain_diffusion_b200.layout import from_nc8hw8
from terrain_diffusion_b200.models import EDMUnet2D


def rel(a, b):
    return float((a - b).square().mean().sqrt() / (b.square().mean().sqrt() + 1e-30))


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cfg = O.DECODER_CFG
    sd = O.procedural_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 5, size, size, generator=g)
    t = torch.atan(torch.exp(torch.randn(n, generator=g) * 1.5) / 0.5)
    trace = {}
    t0 = time.time()
    ref = O.unet_forward(sd, cfg, x, t, [], trace=trace)
    print(f"oracle forward {time.time()-t0:.2f}s")
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    m.use_cuda_graph = False
    out = m(x.cuda(), t.cuda(), [])
    torch.cuda.synchronize()
    prog, bufs = m._plans[("fwd", n, size, size, False)]
    arena = prog.arena
    worst = 0.0
    for key, want in trace.items():
        got = from_nc8hw8(arena[key + ".raw"]).cpu()
        r = rel(got, want)
        worst = max(worst, r)
        flag = "" if r < 2e-2 else "   <-- BAD"
        print(f"{key:28s} rel_rms={r:.3e} ref_std={float(want.std()):.3f}{flag}")
    r = rel(out.cpu(), ref)
    print(f"OUTPUT rel_rms={r:.3e}  max_abs={float((out.cpu()-ref).abs().max()):.3e} ref_std={float(ref.std()):.3f}")
    print(f"launches per forward: {prog.n_launch} (igemm {prog.n_igemm})")
    # graph replay + timing
    m.use_cuda_graph = True
    xs, ts = x.cuda(), t.cuda()
    for _ in range(3):
        out2 = m(xs, ts, [])
    torch.cuda.synchronize()
    print("graph vs eager max diff:", float((out2 - out).abs().max()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prog.run(True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"forward {size}x{size} N={n}: {ms*1e3:.1f} us  -> {n/ms*1e3:.1f} tile-steps/s, "
          f"{343.94*(size/256)**2*n/ms:.1f} TFLOP/s")
    return 0 if r < 1e-2 else 1


if __name__ == "__main__":
    sys.exit(main())
