"""N > 1 host logic on CPU: world_size-2/3 gloo process groups run the striped canvas + halo exchange and must
reproduce the single-process canvas BIT FOR BIT.  The canvas arithmetic injected here is the oracle's
(test infrastructure); on GPUs the same protocol drives BlendCanvas / NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tiling as otile
from terrain_diffusion_b200.inference.sharded import ShardedCanvas
from terrain_diffusion_b200.inference.tiling import shard_rows, tile_starts


class CpuCanvas:
    """Same surface as inference.canvas.BlendCanvas, arithmetic from the oracle (CPU, tests only)."""

    def __init__(self, channels, height, width, device, origin=(0, 0)):
        self.origin = origin
        self.val = torch.zeros(channels, height, width)
        self.wsum = torch.zeros(height, width)

    def accumulate(self, tile, y0, x0, window=None):
        t = tile.shape[-1]
        win = otile.linear_weight_window(t)
        y, x = y0 - self.origin[0], x0 - self.origin[1]
        lo = max(0, -y)                                    # rows above this canvas are clipped (strip canvases)
        h = min(t, self.val.shape[1] - y)
        self.val[:, y + lo:y + h, x:x + t] += (tile * win)[:, lo:h]
        self.wsum[y + lo:y + h, x:x + t] += win[lo:h]


def _tile_value(i0, j0, c, t):
    g = torch.Generator().manual_seed(i0 * 100003 + j0)
    return torch.randn(c, t, t, generator=g)


def _worker(rank, world, port, h, w, t, stride, c, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cv = ShardedCanvas(c, h, w, t, stride, "cpu", canvas_factory=CpuCanvas)
        tiles = cv.my_tiles()
        assert tiles[:cv.n_boundary_tiles()] == [ij for ij in sorted(tiles) if ij[0] + t > cv.own_hi and rank + 1 < world]
        for k, (i0, j0) in enumerate(tiles):
            cv.add_tile(_tile_value(i0, j0, c, t), i0, j0)
            if k + 1 == cv.n_boundary_tiles():
                cv.start_exchange()                       # strips travel while the interior tiles are "solved"
        cv.finalize()
        full = cv.gather(0)
        if rank == 0:
            torch.save(full, out_path)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,h,w,t,stride", [(2, 160, 96, 64, 48), (3, 256, 64, 64, 32), (2, 100, 70, 64, 48)])
def test_striped_canvas_matches_single_process_bit_exact(tmp_path, world, h, w, t, stride):
    c = 2
    out = tmp_path / "full.pt"
    mp.spawn(_worker, args=(world, _free_port(), h, w, t, stride, c, str(out)), nprocs=world, join=True)
    got = torch.load(out)
    val, ws = torch.zeros(1, c, h, w), torch.zeros(1, 1, h, w)
    win = otile.linear_weight_window(t)[None, None]
    for i0 in tile_starts(h, t, stride):
        for j0 in tile_starts(w, t, stride):
            otile.accumulate(val, ws, _tile_value(i0, j0, c, t)[None], win, i0, j0)
    assert torch.equal(got, (val / ws)[0])


def test_shard_rows_partition():
    for n, world in [(23, 8), (4, 2), (5, 5), (7, 3)]:
        rows = [list(shard_rows(n, world, r)) for r in range(world)]
        assert sum(rows, []) == list(range(n))
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1


def test_too_many_ranks_is_rejected():
    with pytest.raises(ValueError):
        ShardedCanvas(1, 64, 64, 64, 48, "cpu", canvas_factory=CpuCanvas, rank=0, world=2)
