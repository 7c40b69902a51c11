"""TerrainPipeline: the three chained lazy canvases of the product path (coarse -> latents x2 phases -> decoder), wired
exactly like WorldPipeline._build_coarse_stage / _build_latent_stage / _build_decoder_stage
(reference inference/world_pipeline.py:961-992, 1133-1203, 1244-1270), with every stage on the GPU.

What is NOT here (out of the hot-path scope): the Perlin/WorldClim conditioning synthesis (pass a
`conditioning_fn(i1, i2, j1, j2) -> [5, h, w]`), HDF5
tile stores, the CLI / HTTP front-ends.  Window geometry, seeds, phase times and batching follow the reference.
"""
from __future__ import annotations

import math

import torch

from ..scheduler import EDMDPMSolverMultistepScheduler
from .lazy_canvas import LazyCanvas, TensorWindow
from .stages import coarse_stage_tile, decoder_stage_tile, latent_stage_tiles
from .tiling import linear_weight_window

# seed offsets of the latent stage's noise fields: _build_latent_stage (world_pipeline.py:1133-1203) passes 5819 to the
# init phase and 5820 + i to the i-th T_INTER phase (pinned against the reference source by tests/golden/stages_golden.npz)
LATENT_INIT_SEED_OFFSET = 5819
LATENT_STEP_SEED_OFFSET = 5820


class TerrainPipeline:
    def __init__(self, coarse_model, base_model, decoder_model, seed: int, conditioning_fn, *, coarse_means, coarse_stds,
                 cond_snr, histogram_raw, latents_means, latents_stds, latents_batch_size: int = 16,
                 decoder_tile_size: int = 512, decoder_tile_stride: int = 384, latent_compression: int = 8,
                 t_inter: float | None = None, residual_mean: float | None = None, residual_std: float | None = None,
                 native_resolution: float = 90.0):
        self.device = decoder_model.device
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.native_resolution = native_resolution          # metres per pixel (world_pipeline.py:293,331)
        self.coarse_model, self.base_model, self.decoder_model = coarse_model, base_model, decoder_model
        self.conditioning_fn = conditioning_fn
        self.kw = dict(coarse_means=coarse_means, coarse_stds=coarse_stds)
        self.cond_snr = torch.as_tensor(cond_snr, dtype=torch.float32)
        self.histogram_raw = torch.as_tensor(histogram_raw, dtype=torch.float32).view(1, -1)
        self.lat_means = torch.as_tensor(latents_means, dtype=torch.float32)
        self.lat_stds = torch.as_tensor(latents_stds, dtype=torch.float32)
        self.lc = latent_compression
        self.residual_mean, self.residual_std = residual_mean, residual_std   # the reference's model kwargs
        sched = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
        self.t_init = math.atan(float(sched.sigmas[0]) / 0.5)
        self.t_inter = math.atan(0.35 / 0.5) if t_inter is None else t_inter   # world_pipeline.py:1144-1145
        dev = self.device

        # ---- coarse: 64^2 tiles, stride 48, 20-step DPM-Solver++ (world_pipeline.py:961-992)
        ww64 = linear_weight_window(64, dev)
        self._set_cond(self.cond_snr)
        coarse_sched = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)

        def f_coarse(ctx):
            _, i, j = ctx
            smap = self.conditioning_fn(i * 48, i * 48 + 64, j * 48, j * 48 + 64)
            return coarse_stage_tile(self.coarse_model, coarse_sched, self.seed, ctx, smap, self._t_cond,
                                     self._cond_inputs, ww64, coarse_means, coarse_stds)

        self.coarse = LazyCanvas(7, f_coarse, TensorWindow((7, 64, 64), (7, 48, 48)), dev)

        # ---- latents: 64^2 tiles, stride 32, two consistency phases (world_pipeline.py:1133-1203)
        out_w = TensorWindow((6, 64, 64), (6, 32, 32))
        coarse_w = TensorWindow((7, 4, 4), (7, 1, 1), (0, -1, -1))

        def f_lat1(ctxs, coarse_windows):
            return latent_stage_tiles(self.base_model, self.seed, ctxs, None, coarse_windows, self.t_init, ww64,
                                      self.histogram_raw, self.lat_means, self.lat_stds, seed_offset=LATENT_INIT_SEED_OFFSET)

        self.latents_init = LazyCanvas(6, f_lat1, out_w, dev, args=(self.coarse,), args_windows=(coarse_w,),
                                       batch_size=latents_batch_size)

        def f_lat2(ctxs, prev_windows, coarse_windows):
            return latent_stage_tiles(self.base_model, self.seed, ctxs, prev_windows, coarse_windows, self.t_inter,
                                      ww64, self.histogram_raw, self.lat_means, self.lat_stds,
                                      seed_offset=LATENT_STEP_SEED_OFFSET)

        self.latents = LazyCanvas(6, f_lat2, out_w, dev, args=(self.latents_init, self.coarse),
                                  args_windows=(out_w, coarse_w), batch_size=latents_batch_size)

        # ---- decoder: T^2 pixel tiles over (T/lc)^2 latent windows, 1-step consistency (world_pipeline.py:1244-1270)
        T, S = decoder_tile_size, decoder_tile_stride
        wwT = linear_weight_window(T, dev)

        def f_dec(ctx, latents_window):
            return decoder_stage_tile(self.decoder_model, self.seed, ctx, latents_window, wwT, [self.t_init], T, S,
                                      latent_compression=self.lc)

        self.residual = LazyCanvas(2, f_dec, TensorWindow((2, T, T), (2, S, S)), dev, args=(self.latents,),
                                   args_windows=(TensorWindow((6, T // self.lc, T // self.lc),
                                                              (6, S // self.lc, S // self.lc)),))

    # ------------------------------------------------------------------ small WorldPipeline API (host state only)
    def _set_cond(self, cond_snr) -> None:
        self.cond_snr = torch.as_tensor(cond_snr, dtype=torch.float32)
        self._t_cond = torch.atan(self.cond_snr)
        self._cond_inputs = [v.view(-1) for v in torch.log(torch.tan(self._t_cond) / 8.0)]

    def empty_cache(self) -> None:
        """Drop every cached window of every stage (WorldPipeline.empty_cache, world_pipeline.py:697-704)."""
        for canvas in (self.coarse, self.latents_init, self.latents, self.residual):
            canvas.clear_cache()

    def change_seed(self, seed: int | None = None) -> bool:
        """New world seed (masked to 64 bits; None draws one like portable_rng.next_seed(None)) and all cached tiles
        dropped; False (no-op) when the seed is unchanged (world_pipeline.py:743-763).  A seed-dependent
        `conditioning_fn` must read `pipeline.seed` itself -- the conditioning synthesis is the caller's."""
        from .noise import next_seed
        new_seed = (int(seed) & 0xFFFFFFFFFFFFFFFF) if seed is not None else next_seed(None)
        if new_seed == self.seed:
            return False
        self.seed = new_seed
        self.empty_cache()
        return True

    def set_cond_snr(self, cond_snr) -> None:
        """Per-channel conditioning SNR (exactly five values) and a rebuild (world_pipeline.py:765-779)."""
        if len(cond_snr) != 5:
            raise ValueError("cond_snr must contain exactly 5 values.")
        self._set_cond([float(x) for x in cond_snr])
        self.empty_cache()

    def close(self) -> None:
        """Release the cached tiles (the reference also closes its HDF5 tile store here, world_pipeline.py:706-712)."""
        self.empty_cache()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()
        return False

    def get_elev(self, i1: int, j1: int, i2: int, j2: int, residual_mean: float | None = None,
                 residual_std: float | None = None, as_int16: bool = False):
        """Elevation in metres over pixel rows [i1,i2) x columns [j1,j2), computed on the device: the `elev` entry of
        WorldPipeline.get (reference inference/world_pipeline.py:1277-1313, 1367-1384; residual_mean / residual_std are
        the reference's model kwargs, given here or to the constructor).  With as_int16 also returns the int16 tensor
        the HTTP API ships (api.py:73-77)."""
        from .postproc import compute_elev
        mean = self.residual_mean if residual_mean is None else residual_mean
        std = self.residual_std if residual_std is None else residual_std
        if mean is None or std is None:
            raise ValueError("get_elev needs residual_mean / residual_std (constructor or call arguments)")
        return compute_elev(self.residual, self.latents, i1, j1, i2, j2, self.lc, mean, std, as_int16=as_int16)

    def get(self, i1: int, j1: int, i2: int, j2: int, with_climate: bool = True) -> dict:
        """WorldPipeline.get (world_pipeline.py:1367-1384), computed on the device: {'elev': fp32 [H, W] in metres,
        'climate': fp32 [5, H, W] or None} -- CUDA tensors (the reference returns CPU tensors; call .cpu() to match)."""
        from .postproc import compute_climate
        elev = self.get_elev(i1, j1, i2, j2)
        climate = compute_climate(self.coarse, i1, j1, i2, j2, elev, self.lc) if with_climate else None
        return {"elev": elev, "climate": climate}

    def residual_normalized(self, i1: int, j1: int, i2: int, j2: int) -> torch.Tensor:
        """Blended decoder output over pixel rows [i1,i2) x columns [j1,j2): residual[0] / residual[1]."""
        r = self.residual[:, i1:i2, j1:j2]
        return r[0] / r[1]
