"""Host-side profile (cProfile) of WorldPipeline.get() at cold locations: where the TTFT goes that is not kernels."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
from oracle import unet as ounet
from terrain_diffusion_b200.inference import WorldPipeline
from terrain_diffusion_b200.models import EDMUnet2D


def build(cfg):
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
    return m


def cond_fn(i1, i2, j1, j2):
    gg = torch.Generator().manual_seed((i1 * 7919 + j1 + 12345) & 0x7FFFFFFF)
    return torch.randn(5, i2 - i1, j2 - j1, generator=gg)


def main():
    pipe = WorldPipeline.from_local_models(build(B.COARSE_CFG), build(B.BASE_CFG), build(ounet.DECODER_CFG), seed=42,
                                           latents_batch_size=[1, 2, 4, 8, 16], torch_compile=True, dtype="bf16",
                                           cache_limit=None, conditioning_fn=cond_fn)
    pipe.to("cuda").bind("TEMP")
    for k in range(4):
        pipe.get(k * 100000, 0, k * 100000 + 512, 512, with_climate=False)
        pipe.empty_cache()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for k in range(5):
        b = (k + 10) * 100000 + 137 * k
        pipe.get(b, 777 * k, b + 512, 777 * k + 512, with_climate=False)
        torch.cuda.synchronize()
        pipe.empty_cache()
    pr.disable()
    print(f"5 cold gets: {(time.perf_counter() - t0) * 1e3 / 5:.1f} ms each")
    for cv in ("_coarse", "_latents_init", "_latents", "_residual"):
        c = getattr(pipe, cv, None)
        if c is not None:
            print(cv, "windows computed so far:", c.windows_computed)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
    print(s.getvalue()[:5000])


if __name__ == "__main__":
    main()
