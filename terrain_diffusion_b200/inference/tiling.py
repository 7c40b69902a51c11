"""Tile geometry (host, integer, bit-exact with the reference).

  tile_starts            terrain_diffusion/training/evaluation/__init__.py:16-22   (bounded canvas, last tile clamped)
  window_range           window-index rule of the unbounded infinite_tensor canvas as observable from its call sites
                         (world_pipeline.py:917-920,1091,1147,1230,1259-1260): window k covers
                         [k*stride+offset, k*stride+offset+size), negative coordinates are real coordinates
  linear_weight_window   world_pipeline.py:117-124 == training/evaluation/__init__.py:3-10
  padded_batch_size      world_pipeline.py:393-398
"""
from __future__ import annotations

import functools

import torch


def tile_starts(length: int, tile_size: int, stride: int) -> list[int]:
    if length <= tile_size:
        return [0]
    starts = list(range(0, max(1, length - tile_size + 1), max(1, stride)))
    if starts[-1] != length - tile_size:
        starts.append(length - tile_size)
    return starts


def window_range(a: int, b: int, size: int, stride: int, offset: int = 0) -> range:
    """Window indices k whose extent intersects [a, b) (pure integer floor/ceil division)."""
    k_lo = (a - offset - size) // stride + 1
    k_hi = -((-(b - offset)) // stride) - 1
    return range(k_lo, k_hi + 1)


def window_origin(k: int, stride: int, offset: int = 0) -> int:
    return k * stride + offset


@functools.lru_cache(maxsize=32)
def _linear_weight_window_cached(size: int, device: str, dtype) -> torch.Tensor:
    mid = (size - 1) / 2
    y, x = torch.meshgrid(torch.arange(size), torch.arange(size), indexing="ij")
    eps = 1e-3
    wy = 1 - (1 - eps) * torch.clamp(torch.abs(y - mid).to(dtype) / mid, 0, 1)
    wx = 1 - (1 - eps) * torch.clamp(torch.abs(x - mid).to(dtype) / mid, 0, 1)
    return (wy * wx).to(device)


def linear_weight_window(size: int, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """[size, size] blend weights; computed on the host in the reference's op order (fp32-identical), then moved.
    Cached per (size, device, dtype): every sampler call asks for the same window (the 256^2 host evaluation plus a
    pageable copy was ~0.3 ms of each end-to-end solve).  Treat the result as read-only."""
    return _linear_weight_window_cached(int(size), str(torch.device(device)), dtype)


def padded_batch_size(n: int, max_batch: int) -> int:
    """Smallest power of two >= n, capped at max_batch (world_pipeline.py:393-398)."""
    p = 1
    while p < n:
        p *= 2
    return min(p, max_batch)


def shard_rows(n_rows: int, world_size: int, rank: int) -> range:
    """Contiguous stripe of tile rows owned by `rank` (1-D partition of the tile grid)."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))
