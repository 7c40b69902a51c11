"""In-graph timeline of the igemm launches of one forward: per-launch duration and the gap to the previous launch
(globaltimer stamps written by the kernels themselves; debug hook tdx_debug_set_igemm_timeline)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the hooks this tool reads are compiled out of the production kernel: (re)build libtdx.so with them
# (run `python -m terrain_diffusion_b200.build` afterwards to get the production library back)
os.environ["TDX_DEBUG_HOOKS"] = "1"
from terrain_diffusion_b200.build import build as _build  # noqa: E402
import importlib, terrain_diffusion_b200.build as _b  # noqa: E402
importlib.reload(_b).build()
import torch

from oracle import unet as O
from terrain_diffusion_b200 import _lib as L
from terrain_diffusion_b200.models import EDMUnet2D


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = L.lib()
    lib.tdx_debug_set_igemm_timeline.argtypes = [C.c_void_p, C.c_int]
    cfg = O.DECODER_CFG
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(O.procedural_state_dict(cfg, seed=0))
    m = m.cuda()
    cap = 128
    tl = torch.zeros(cap, 2, dtype=torch.int64, device="cuda")
    tl[:, 0] = torch.iinfo(torch.int64).max
    lib.tdx_debug_set_igemm_timeline(tl.data_ptr(), cap)      # the plan built next records its launches
    x = torch.randn(n, 5, size, size, device="cuda")
    t = torch.full((n,), 1.2, device="cuda")
    m(x, t, [])                                                # builds + captures + runs once
    lib.tdx_debug_set_igemm_timeline(None, 0)
    for _ in range(3):
        tl[:, 0] = torch.iinfo(torch.int64).max
        tl[:, 1] = 0
        torch.cuda.synchronize()
        m(x, t, [])
        torch.cuda.synchronize()
    a = tl.cpu()
    rows = [(int(a[i, 0]), int(a[i, 1])) for i in range(cap) if int(a[i, 1]) > 0]
    t0 = rows[0][0]
    tot_d = tot_g = 0.0
    out = []
    for i, (s, e) in enumerate(rows):
        gap = (s - rows[i - 1][1]) / 1e3 if i else 0.0
        d = (e - s) / 1e3
        tot_d += d
        tot_g += gap
        out.append(f"{d:.1f}/{gap:+.1f}")
    print(f"{len(rows)} igemm launches; span {(rows[-1][1]-t0)/1e3:.1f} us; sum durations {tot_d:.1f} us; sum gaps {tot_g:.1f} us")
    print("duration/gap-before (us):", " ".join(out))


if __name__ == "__main__":
    main()
