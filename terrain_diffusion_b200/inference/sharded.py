"""Multi-GPU canvas: tile rows striped across ranks, neighbour exchange of the overlap strips.

The reference has no inference-time multi-GPU; this is the natural sharding of its tile loop: tiles within
a phase are independent, so each rank solves the tile rows `tiling.shard_rows` gives it and owns the canvas rows from its
first tile row to the next rank's first tile row.  Only a rank's LAST tile rows overhang into the next rank's pixels.
The (sum x*w, sum w) partial sums of that strip go to the neighbour with one point-to-point message per boundary
(torch.distributed send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests) -- no collective in the data path.

Bit-exactness: fp32 addition is order dependent, and the single-GPU order is row-major over tiles.  A
rank therefore first INSTALLS the strip it receives from its upper neighbour and only then accumulates its own tiles
in row-major order, so every pixel sees exactly the single-GPU sequence of additions.

Overlap: the strip a rank sends only holds contributions of its OWN overhanging tile rows (the constructor rejects
partitions in which a strip could reach past the next rank), so nothing chains from rank to rank.  `my_tiles()` lists
the overhanging (boundary) tile rows FIRST; as soon as they are solved, `start_exchange()` blends them into a separate
strip canvas, and posts the send to the lower neighbour together with the receive from the upper one (one batched
isend/irecv, NCCL's own stream) -- the transfer runs while the interior tiles are being solved, and `finalize()` only
waits for it before the (cheap) blend of the rank's own rows.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .canvas import BlendCanvas
from .tiling import shard_rows, tile_starts


class ShardedCanvas:
    def __init__(self, channels: int, height: int, width: int, tile_size: int, stride: int, device, group=None,
                 canvas_factory=BlendCanvas, rank: int | None = None, world: int | None = None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.channels, self.height, self.width, self.tile = channels, height, width, tile_size
        self.device = torch.device(device)
        self.row_starts = tile_starts(height, tile_size, stride)
        self.col_starts = tile_starts(width, tile_size, stride)
        if len(self.row_starts) < self.world:
            raise ValueError(f"{len(self.row_starts)} tile rows cannot be striped over {self.world} ranks")
        self.rows = shard_rows(len(self.row_starts), self.world, self.rank)
        first = [self.row_starts[shard_rows(len(self.row_starts), self.world, r)[0]] for r in range(self.world)]
        first[0] = 0
        self.bounds = first + [height]                       # rank r owns pixel rows [bounds[r], bounds[r+1])
        self.own_lo, self.own_hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.cover_hi = min(height, self.row_starts[self.rows[-1]] + tile_size)   # how far this rank's tiles reach
        if self.rank + 1 < self.world and self.cover_hi > self.bounds[self.rank + 2]:
            raise ValueError("stripes are thinner than the tile overhang; use fewer ranks for this canvas")
        # local canvas covers [own_lo, cover_hi) (the rows past own_hi are scratch: the neighbour owns them)
        self.local = canvas_factory(channels, self.cover_hi - self.own_lo, width, self.device,
                                    origin=(self.own_lo, 0))
        self._factory = canvas_factory
        self._tiles: list = []
        self._work, self._recv_buf, self._send_buf = None, None, None

    def boundary_rows(self) -> list[int]:
        """Tile rows of this rank whose tiles overhang into the next rank's pixel rows."""
        if self.rank + 1 >= self.world:
            return []
        return [r for r in self.rows if self.row_starts[r] + self.tile > self.own_hi]

    def my_tiles(self) -> list[tuple[int, int]]:
        """Tile origins this rank must solve: the boundary rows first (their strip can then travel while the interior
        is being solved), each group row-major."""
        brows = self.boundary_rows()
        order = brows + [r for r in self.rows if r not in brows]
        return [(self.row_starts[r], j0) for r in order for j0 in self.col_starts]

    def n_boundary_tiles(self) -> int:
        return len(self.boundary_rows()) * len(self.col_starts)

    def add_tile(self, tile: torch.Tensor, i0: int, j0: int) -> None:
        self._tiles.append((tile, i0, j0))

    def _strip_rows(self) -> int:
        return max(0, self.cover_hi - self.own_hi)

    def _upper_rows(self) -> int:
        if self.rank == 0:
            return 0
        upper_cover = min(self.height,
                          self.row_starts[shard_rows(len(self.row_starts), self.world, self.rank - 1)[-1]] + self.tile)
        return max(0, upper_cover - self.own_lo)

    def start_exchange(self) -> None:
        """Call once every boundary tile has been added (any time later is also correct, just less overlapped)."""
        if self._work is not None or self.world == 1:
            return
        ops = []
        n_up = self._upper_rows()
        if n_up > 0:
            self._recv_buf = torch.empty((self.channels + 1, n_up, self.width), dtype=torch.float32,
                                         device=self.device)
            ops.append(dist.P2POp(dist.irecv, self._recv_buf, self._global(self.rank - 1), self.group))
        n = self._strip_rows()
        if self.rank + 1 < self.world and n > 0:
            # the overhang only ever holds this rank's own boundary tiles: blend them (row-major) in a strip canvas
            strip = self._factory(self.channels, n, self.width, self.device, origin=(self.own_hi, 0))
            for tile, i0, j0 in sorted(self._tiles, key=lambda t: (t[1], t[2])):
                if i0 + self.tile > self.own_hi:
                    strip.accumulate(tile, i0, j0)
            self._send_buf = torch.cat([strip.val, strip.wsum[None]], dim=0).contiguous()
            ops.append(dist.P2POp(dist.isend, self._send_buf, self._global(self.rank + 1), self.group))
        self._work = dist.batch_isend_irecv(ops) if ops else []

    def finalize(self) -> None:
        """Wait for the strips, install the upper neighbour's, accumulate the own tiles (row-major)."""
        self.start_exchange()
        for w in (self._work or []):
            w.wait()
        if self._recv_buf is not None:
            n = self._recv_buf.shape[1]
            self.local.val[:, :n] = self._recv_buf[:-1]
            self.local.wsum[:n] = self._recv_buf[-1]
        for tile, i0, j0 in sorted(self._tiles, key=lambda t: (t[1], t[2])):
            self.local.accumulate(tile, i0, j0)
        self._tiles.clear()
        self._work, self._recv_buf, self._send_buf = None, None, None

    def _global(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def owned(self) -> tuple[torch.Tensor, torch.Tensor]:
        """(sum x*w [C, rows, W], sum w [rows, W]) for the pixel rows this rank owns."""
        n = self.own_hi - self.own_lo
        return self.local.val[:, :n], self.local.wsum[:n]

    def normalized_owned(self, divisor: float = 1.0) -> torch.Tensor:
        val, w = self.owned()
        out = val / w
        return out if divisor == 1.0 else out / divisor

    def gather(self, dst: int = 0):
        """Full normalised canvas on rank `dst` (None elsewhere); for writers that want one array."""
        part = self.normalized_owned().contiguous()
        if self.rank == dst:
            parts = [part]
            for r in range(self.world):
                if r == dst:
                    continue
                buf = torch.empty((self.channels, self.bounds[r + 1] - self.bounds[r], self.width),
                                  dtype=torch.float32, device=self.device)
                dist.recv(buf, src=self._global(r), group=self.group)
                parts.append(buf)
            order = [dst] + [r for r in range(self.world) if r != dst]
            parts = [p for _, p in sorted(zip(order, parts), key=lambda t: t[0])]
            return torch.cat(parts, dim=1)
        dist.send(part, dst=self._global(dst), group=self.group)
        return None
