"""Seeded inputs shared by tests/golden/make_golden_post.py (reference side) and the post-step tests (oracle / CUDA)."""
import numpy as np


def field(seed, h, w, lo=0.0, amp=1.0):
    """Smooth-ish random field: low-frequency sinusoids + white noise (deterministic in numpy, float32)."""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.zeros((h, w))
    for _ in range(4):
        fy, fx, ph = rng.uniform(0.01, 0.12), rng.uniform(0.01, 0.12), rng.uniform(0, 6.28)
        out += rng.uniform(0.3, 1.0) * np.sin(fy * y + fx * x + ph)
    out += 0.25 * rng.randn(h, w)
    return (lo + amp * out).astype(np.float32)


def laplacian_case(name):
    """(residual [H, W], lowres [h, w]) of the direct laplacian_denoise / laplacian_decode cases."""
    shapes = {"rect": (96, 80, 12, 10), "square": (128, 128, 16, 16), "wide": (64, 136, 8, 17)}
    H, W, h, w = shapes[name]
    seed = sorted(shapes).index(name) + 1
    return field(10 * seed, H, W, 0.0, 3.0), field(10 * seed + 1, h, w, -31.4, 38.6)


class FakeCanvas:
    """canvas[:, a:b, c:d] over world coordinates (negative allowed) -> planes [C+1, b-a, d-c]: values * w and w, like the
    un-normalised (sum x*w, sum w) tensors the pipeline's lazy canvases return."""

    def __init__(self, seed, channels, origin, size, lo, amp):
        self.origin, self.size = origin, size
        h, w = size
        self.wt = (0.5 + np.abs(field(seed + 99, h, w))).astype(np.float32)
        self.val = np.stack([field(seed + c, h, w, lo, amp) for c in range(channels)])

    def planes(self, a, b, c, d):
        oy, ox = self.origin
        assert a >= oy and c >= ox and b <= oy + self.size[0] and d <= ox + self.size[1], "window outside the fake canvas"
        v = self.val[:, a - oy:b - oy, c - ox:d - ox]
        w = self.wt[a - oy:b - oy, c - ox:d - ox]
        return np.concatenate([v * w[None], w[None]], axis=0)


ELEV_WINDOWS = {"neg_square": (-37, 20, 91, 148), "wide": (8, -60, 136, 140), "aligned": (0, 0, 64, 64)}
RESIDUAL_MEAN, RESIDUAL_STD = 0.12, 1.35


def elev_canvases():
    """HR residual canvas (1 value plane + weight) and LR latents canvas (5 value planes + weight), scale 8."""
    resid = FakeCanvas(1000, 1, (-160, -200), (480, 560), 0.0, 0.8)
    lat = FakeCanvas(2000, 5, (-20, -25), (60, 70), 0.2, 0.6)
    return resid, lat


def coarse_canvas():
    """Coarse canvas at 1/256 resolution: 6 value planes (0 = signed-sqrt elevation, 2 = temperature, ...) + weight."""
    c = FakeCanvas(3000, 6, (-12, -12), (30, 32), 0.0, 1.0)
    c.val[0] = field(3100, 30, 32, 5.0, 18.0)        # signed sqrt of metres: mostly land, some ocean (< 0)
    c.val[2] = field(3200, 30, 32, 12.0, 6.0)        # temperature, deg C
    return c
