"""Unbounded, lazy, device-resident canvas -- the arithmetic and window protocol of the reference's external canvas
engine (`infinite_tensor>=0.3.0`, requirements.txt:32; absent from /root/reference), as observable from its call sites
(world_pipeline.py:982-992,1146-1201,1259-1270; annotated_infinite_panorama.py:141-226):

  * `TensorWindow(size, stride, offset)`: window index k along an axis covers [k*stride+offset, k*stride+offset+size);
  * slicing `canvas[:, a:b, c:d]` (integers are real, possibly negative, world coordinates) returns the SUM over all
    windows intersecting the slice of f's outputs, computing missing windows on demand; each window's f receives the
    slices of its dependencies taken at the SAME window index through their own TensorWindow;
  * with `batch_size`, f receives lists of up to that many window indices (and lists of dependency slices).

PARITY UNPINNED against the library itself (it is not on disk and the reference has no tests for it); the semantics
above are pinned by tests/test_lazy_canvas_gpu.py against a brute-force evaluation, and the arithmetic (fp32 sums in
row-major window order) by the bounded-canvas tests.  Everything stays in HBM: blocks of the canvas are fp32 CUDA
tensors, windows are added with tdx_canvas_add.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import _lib as L
from .tiling import window_range


@dataclass(frozen=True)
class TensorWindow:
    size: tuple      # (C, h, w)
    stride: tuple    # (C, sh, sw)
    offset: tuple = (0, 0, 0)


class LazyCanvas:
    def __init__(self, channels: int, f, output_window: TensorWindow, device, args=(), args_windows=(),
                 batch_size: int | None = None, block: int = 512):
        self.channels = channels
        self.f = f
        self.win = output_window
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.TdxError("LazyCanvas lives in GPU memory; there is no CPU path")
        self.args, self.args_windows = tuple(args), tuple(args_windows)
        assert len(self.args) == len(self.args_windows)
        self.batch_size = batch_size
        self.block = block
        self.blocks: dict = {}
        self.done: set = set()
        self.windows_computed = 0

    # ------------------------------------------------------------------ storage
    def _block(self, by, bx):
        key = (by, bx)
        if key not in self.blocks:
            self.blocks[key] = torch.zeros((self.channels, self.block, self.block), dtype=torch.float32,
                                           device=self.device)
        return self.blocks[key]

    def _add(self, tile: torch.Tensor, y0: int, x0: int):
        c, th, tw = tile.shape
        tile = tile.contiguous()
        b = self.block
        for by in range(y0 // b, (y0 + th - 1) // b + 1):
            for bx in range(x0 // b, (x0 + tw - 1) // b + 1):
                blk = self._block(by, bx)
                L.call(L.lib().tdx_canvas_add, blk.device, blk.data_ptr(), c, b, b, tile.data_ptr(), th, tw, y0 - by * b,
                       x0 - bx * b)

    def clear_cache(self):
        self.blocks.clear()
        self.done.clear()

    # ------------------------------------------------------------------ windows
    def window_origin(self, i: int, j: int) -> tuple[int, int]:
        return (i * self.win.stride[1] + self.win.offset[1], j * self.win.stride[2] + self.win.offset[2])

    def windows_for(self, a: int, b: int, c: int, d: int) -> list[tuple[int, int]]:
        """Window indices (row-major) whose extent intersects rows [a,b) x columns [c,d)."""
        rows = window_range(a, b, self.win.size[1], self.win.stride[1], self.win.offset[1])
        cols = window_range(c, d, self.win.size[2], self.win.stride[2], self.win.offset[2])
        return [(i, j) for i in rows for j in cols]

    def _dep_slices(self, i, j):
        out = []
        for dep, w in zip(self.args, self.args_windows):
            y0 = i * w.stride[1] + w.offset[1]
            x0 = j * w.stride[2] + w.offset[2]
            out.append(dep[:, y0:y0 + w.size[1], x0:x0 + w.size[2]])
        return out

    def _ensure(self, idxs):
        missing = [ij for ij in idxs if ij not in self.done]
        if not missing:
            return
        if self.batch_size is None:
            for (i, j) in missing:
                tile = self.f((0, i, j), *self._dep_slices(i, j))
                self._add(tile.to(self.device, torch.float32), *self.window_origin(i, j))
                self.done.add((i, j))
                self.windows_computed += 1
        else:
            for g0 in range(0, len(missing), self.batch_size):
                grp = missing[g0:g0 + self.batch_size]
                deps = [self._dep_slices(i, j) for (i, j) in grp]
                lists = [list(col) for col in zip(*deps)] if deps and deps[0] else []
                tiles = self.f([(0, i, j) for (i, j) in grp], *lists)
                for (i, j), tile in zip(grp, tiles):
                    self._add(tile.to(self.device, torch.float32), *self.window_origin(i, j))
                    self.done.add((i, j))
                    self.windows_computed += 1

    # ------------------------------------------------------------------ read
    def __getitem__(self, key) -> torch.Tensor:
        if not (isinstance(key, tuple) and len(key) == 3):
            raise IndexError("LazyCanvas is indexed as canvas[:, a:b, c:d]")
        cs, ys, xs = key
        a, b, c, d = ys.start, ys.stop, xs.start, xs.stop
        if None in (a, b, c, d) or b <= a or d <= c:
            raise IndexError("row and column slices need explicit start < stop (world coordinates, may be negative)")
        self._ensure(self.windows_for(a, b, c, d))
        out = torch.zeros((self.channels, b - a, d - c), dtype=torch.float32, device=self.device)
        blk = self.block
        for by in range(a // blk, (b - 1) // blk + 1):
            for bx in range(c // blk, (d - 1) // blk + 1):
                if (by, bx) not in self.blocks:
                    continue
                y0, y1 = max(a, by * blk), min(b, (by + 1) * blk)
                x0, x1 = max(c, bx * blk), min(d, (bx + 1) * blk)
                out[:, y0 - a:y1 - a, x0 - c:x1 - c] = self.blocks[(by, bx)][:, y0 - by * blk:y1 - by * blk,
                                                                               x0 - bx * blk:x1 - bx * blk]
        return out[cs]
