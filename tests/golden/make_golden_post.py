"""Golden vectors for the elevation read-out, produced by the reference's own code:
`terrain_diffusion/data/laplacian_encoder.py` is imported unchanged (it needs only torch + torchvision, both present in
this container) and `WorldPipeline._compute_elev` is extracted from the reference source with `ast` and called with a
stand-in `self` (world_pipeline.py itself cannot be imported: infinite_tensor, h5py, rasterio ... are absent).
Only outputs are stored (tests/golden/post_golden.npz); inputs are regenerated from seeds by tests/_post_inputs.py.

    python tests/golden/make_golden_post.py
"""
from __future__ import annotations

import ast
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path[:0] = [str(ROOT / "oracle" / "_stub"), str(REF), str(ROOT)]

from terrain_diffusion.data.laplacian_encoder import laplacian_decode, laplacian_denoise  # noqa: E402

from tests._post_inputs import (ELEV_WINDOWS, RESIDUAL_MEAN, RESIDUAL_STD, coarse_canvas, elev_canvases,  # noqa: E402
                                laplacian_case)

torch.set_grad_enabled(False)


class TorchCanvas:
    def __init__(self, fake):
        self.fake = fake

    def __getitem__(self, key):
        _, ys, xs = key
        return torch.from_numpy(self.fake.planes(ys.start, ys.stop, xs.start, xs.stop).copy())


def main():
    out = {}
    for name in ("rect", "square", "wide"):
        r, l = laplacian_case(name)
        r_t, l_t = torch.from_numpy(r), torch.from_numpy(l)
        r2, l2 = laplacian_denoise(r_t, l_t, sigma=5)
        out[f"lap_{name}_lowres"] = l2.numpy()
        out[f"lap_{name}_elev"] = laplacian_decode(r2, l2).numpy()
        out[f"lap_{name}_decode_extrap"] = laplacian_decode(r_t, l_t, extrapolate=True).numpy()
    src = (REF / "terrain_diffusion/inference/world_pipeline.py").read_text()
    tree = ast.parse(src)
    body = []
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name == "WorldPipeline":
            body += [m for m in n.body if isinstance(m, ast.FunctionDef) and m.name in ("_compute_elev", "_compute_climate")]
    assert len(body) == 2, "WorldPipeline._compute_elev / _compute_climate not found"
    # local_baseline_temperature_torch lives in inference/postprocessing.py, which imports matplotlib (absent here):
    # the function definition alone is extracted the same way
    pp = ast.parse((REF / "terrain_diffusion/inference/postprocessing.py").read_text())
    body += [n for n in pp.body if isinstance(n, ast.FunctionDef) and n.name == "local_baseline_temperature_torch"]
    assert len(body) == 3
    import torch.nn.functional as F
    ns = {"torch": torch, "np": np, "F": F, "laplacian_denoise": laplacian_denoise, "laplacian_decode": laplacian_decode}
    exec(compile(ast.Module(body=body, type_ignores=[]), "world_pipeline_extract", "exec"), ns)
    resid, lat = elev_canvases()
    fake = SimpleNamespace(kwargs={"residual_mean": RESIDUAL_MEAN, "residual_std": RESIDUAL_STD}, latents=TorchCanvas(lat),
                           coarse=TorchCanvas(coarse_canvas()))
    for name, (i1, j1, i2, j2) in ELEV_WINDOWS.items():
        elev = ns["_compute_elev"](fake, i1, j1, i2, j2, TorchCanvas(resid), 8)
        out[f"elev_{name}"] = elev.numpy()
        clim = ns["_compute_climate"](fake, i1, j1, i2, j2, elev, 8)
        out[f"climate_{name}"] = clim.numpy()[:, ::2, ::2]        # every other pixel: keeps the fixture small
        print(name, tuple(elev.shape), float(elev.abs().max()), tuple(clim.shape), [round(float(c.mean()), 4) for c in clim])
    np.savez_compressed(HERE / "post_golden.npz", **out)
    print("wrote", HERE / "post_golden.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
