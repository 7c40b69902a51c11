"""Measure the best (n_per_item, k_split) of every implicit-GEMM launch shape of the shipped models on THIS GPU and
write terrain_diffusion_b200/tuned_shapes.json (the planner's lookup table; the library's cost model is the fallback).

    python tools/tune_igemm.py            # on a B200 (gpurun); merges into the existing table
Shapes: decoder at 256^2 for 1/2/4/8/16 tiles, 512^2 for 1/4 tiles, 64^2 and 128^2; coarse model 64^2; base (latent)
model 64^2 for the padded batch sizes 1/2/4/8/16.
"""
import json
import os
import sys

os.environ["TDX_AUTOTUNE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import unet as O
from terrain_diffusion_b200.models import EDMUnet2D, plan
from tests.test_oracle_golden import BASE_CFG, COARSE_CFG


def run(cfg, cases, cond=None):
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(O.procedural_state_dict(cfg, seed=0))
    m = m.cuda()
    for n, hw in cases:
        x = torch.randn(n, cfg["in_channels"], hw, hw, device="cuda")
        t = torch.full((n,), 1.1, device="cuda")
        ci = cond(n) if cond else []
        m(x, t, ci)
        torch.cuda.synchronize()
        print(cfg["model_channels"], n, hw, len(plan._TUNED_NEW), flush=True)
    del m
    torch.cuda.empty_cache()


def main():
    which = sys.argv[1:] or ["decoder", "coarse", "base"]
    if "decoder" in which:
        run(O.DECODER_CFG, [(1, 256), (16, 256), (2, 256), (4, 256), (8, 256), (1, 512), (4, 512), (1, 64), (1, 128),
                            (2, 128), (4, 64), (16, 64)])
    if "coarse" in which:
        run(COARSE_CFG, [(1, 64)], cond=lambda n: [torch.zeros(n, device="cuda") for _ in range(5)])
    if "base" in which:
        run(BASE_CFG, [(16, 64), (1, 64), (2, 64), (4, 64), (8, 64)], cond=lambda n: [torch.randn(n, 58, device="cuda")])
    table = {}
    try:
        table = json.load(open(plan.TUNED_PATH)).get("shapes", {})
    except Exception:
        pass
    table.update(plan._TUNED_NEW)
    json.dump({"device": torch.cuda.get_device_name(0), "how": "tools/tune_igemm.py: median of 5 x 12 back-to-back "
               "launches per valid (n_per_item, k_split)", "shapes": dict(sorted(table.items()))},
              open(plan.TUNED_PATH, "w"), indent=0)
    print("shapes in table:", len(table), "new:", len(plan._TUNED_NEW))


if __name__ == "__main__":
    main()
