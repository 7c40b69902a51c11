"""Host-side profile (cProfile) of the e2e leg of bench.py: sample_decoder_diffusion_tiled with pinned-host inputs."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import unet as ounet
from terrain_diffusion_b200.inference import sample_decoder_diffusion_tiled
from terrain_diffusion_b200.models import EDMUnet2D
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler

dev = torch.device("cuda:0")
cfg = ounet.DECODER_CFG
m = EDMUnet2D(**cfg).eval()
m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
m = m.to(dev)
sched = EDMDPMSolverMultistepScheduler()
g = torch.Generator().manual_seed(1)
noise_h = (torch.randn(1, 1, 256, 256, generator=g) * 80).pin_memory()
cond_h = torch.randn(1, 4, 256, 256, generator=g).pin_memory()
out_h = torch.empty(1, 1, 256, 256).pin_memory()


def one():
    nz = noise_h.to(dev, non_blocking=True)
    cd = cond_h.to(dev, non_blocking=True)
    y = sample_decoder_diffusion_tiled(m, sched, cd, nz, 256, 256, num_steps=20)
    out_h.copy_(y, non_blocking=True)
    torch.cuda.synchronize()


for _ in range(3):
    one()
t0 = time.perf_counter()
for _ in range(10):
    one()
print(f"e2e solve: {(time.perf_counter() - t0) * 100:.3f} ms each (20 steps)")
# host time until the graph is launched = time the GPU idles per solve
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    one()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:7000])
