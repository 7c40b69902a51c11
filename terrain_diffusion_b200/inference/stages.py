"""The three InfiniteDiffusion stage functions of the product pipeline, device-resident.

Mirrors of WorldPipeline._coarse_inference / _latent_inference / _decoder_inference
(reference inference/world_pipeline.py:909-959, 1052-1131, 1209-1242): same arguments in spirit, same seeds, same tile
geometry (window index ctx -> origin ctx*stride), same TrigFlow consistency / DPM-Solver arithmetic and the same packed
return value cat([x*w, w]) -- but every tensor stays on the GPU (the reference crosses the host/device boundary twice
per tile: numba noise + H2D in, .cpu() out), the noise comes from the bit-compatible GPU generator, and the U-Net /
scheduler run on libtdx.  What the reference computes OUTSIDE the hot path is taken as an argument: the synthetic
conditioning map of the coarse stage (`_conditioning_model_input`, Perlin/WorldClim machinery) and the already blended
dependency windows that the lazy canvas engine (`infinite_tensor`, absent) hands to each callback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .noise import gaussian_noise_patch, standard_normal
from .samplers import get_diffusion_solve
from .tiling import padded_batch_size


def _mp_concat(args, dim=1):
    """mp_concat with equal weights (models/mp_layers.py:65-86)."""
    w = torch.full((len(args),), 1 / len(args), dtype=args[0].dtype, device=args[0].device)
    n_tot = torch.tensor(sum(a.shape[dim] for a in args), dtype=args[0].dtype, device=args[0].device)
    c = torch.sqrt(n_tot / torch.sum(torch.square(w)))
    return torch.concat([a * (c / np.sqrt(a.shape[dim]) * w[i]) for i, a in enumerate(args)], dim=dim)


def process_latent_conditioning(cond_img, histogram_raw, cond_means, cond_stds, noise_level, seed, seed_offset=0):
    """_process_latent_conditioning (world_pipeline.py:1018-1050): [n,7,4,4] coarse window -> [n,58] condition vector.
    Runs on cond_img's device; the NaN fill uses the portable stream standard_normal(seed + 9999 + seed_offset)."""
    dev = cond_img.device
    cond_img = (cond_img - cond_means.to(dev).view(1, -1, 1, 1)) / cond_stds.to(dev).view(1, -1, 1, 1)
    cond_img[0:1] = cond_img[0:1].nan_to_num(float(cond_means[0]))
    cond_img[1:2] = cond_img[1:2].nan_to_num(float(cond_means[1]))
    means_crop = cond_img[:, 0:1]
    p5_crop = cond_img[:, 1:2]
    climate_means_crop = cond_img[:, 2:6, 1:3, 1:3].mean(dim=(2, 3))
    mask_crop = cond_img[:, 6:7]
    nan_mask = torch.isnan(climate_means_crop)
    nan_count = int(nan_mask.sum().item())
    if nan_count > 0:
        fill = standard_normal(seed + 9999 + seed_offset, nan_count, device=dev) if dev.type == "cuda" else None
        if fill is None:
            raise RuntimeError("process_latent_conditioning: NaN fill needs the GPU noise generator")
        climate_means_crop[nan_mask] = fill.to(climate_means_crop.dtype)
    noise_level_norm = (noise_level - 0.5) * np.sqrt(12)
    return _mp_concat([means_crop.flatten(1), p5_crop.flatten(1), climate_means_crop.flatten(1),
                       mask_crop.flatten(1), histogram_raw.to(dev), noise_level_norm.view(-1, 1).to(dev)], dim=1).float()


@torch.no_grad()
def decoder_stage_tile(model, seed: int, ctx, latents: torch.Tensor, weight_window: torch.Tensor, t_list,
                       tile_size: int = 512, tile_stride: int = 384, sigma_data: float = 0.5,
                       latent_compression: int = 8) -> torch.Tensor:
    """_decoder_inference (world_pipeline.py:1209-1242).  latents: [6, T/lc, T/lc] packed (sum x*w, sum w) window of the
    latent canvas (any device); returns the packed [2, T, T] decoder tile on the model's device."""
    dev = model.device
    t_ = tile_size
    lc = latent_compression
    lat = latents.to(dev, torch.float32)
    lat = (lat[:-1] / lat[-1:])[:4].view(1, 4, t_ // lc, t_ // lc)
    up = torch.nn.functional.interpolate(lat, size=(t_, t_), mode="nearest")
    sample = torch.zeros((1, 1, t_, t_), device=dev, dtype=torch.float32)
    for i, t in enumerate(t_list):
        t = float(t)
        z = gaussian_noise_patch(seed + 5819 + i, ctx[1] * tile_stride, ctx[2] * tile_stride, t_, t_, 1, t_, t_,
                                 device=dev)[None] * sigma_data
        x_t = math.cos(t) * sample + math.sin(t) * z
        model_in = torch.cat([x_t / sigma_data, up], dim=1)
        pred = -model(model_in, torch.tensor([t], device=dev, dtype=torch.float32), [])
        sample = math.cos(t) * x_t - math.sin(t) * sigma_data * pred
    sample = sample.float() / sigma_data
    w = weight_window.to(dev)
    return torch.cat([sample[0] * w[None], w[None]], dim=0)


@torch.no_grad()
def latent_stage_tiles(model, seed: int, ctxs, samples, cond_imgs, t: float, weight_window: torch.Tensor,
                       histogram_raw, cond_means, cond_stds, seed_offset: int = 0, sigma_data: float = 0.5,
                       pad_batch_to=None) -> list:
    """_latent_inference (world_pipeline.py:1052-1131) for a batch of window indices `ctxs`.
    samples: None (first phase) or list of packed [6,64,64] windows of the previous phase's canvas;
    cond_imgs: list of packed [7,4,4] coarse windows.  Returns a list of packed [6,64,64] tiles on the device."""
    dev = model.device
    tile, stride = 64, 32
    if samples is None:
        samples = [None] * len(ctxs)
    model_in, cond_vecs, kept = [], [], []
    for ctx, sample, cond_img in zip(ctxs, samples, cond_imgs):
        if sample is None:
            sample = torch.zeros((1, 5, tile, tile), device=dev, dtype=torch.float32)
        else:
            sample = torch.as_tensor(sample, device=dev, dtype=torch.float32)
            sample = (sample[:-1] / sample[-1:] * sigma_data)[None]
        cond_img = torch.as_tensor(cond_img, device=dev, dtype=torch.float32)
        cond_img = cond_img[:-1] / cond_img[-1:]
        cond_img = torch.cat([cond_img, torch.ones((1, 4, 4), device=dev)], dim=0)[None]
        cond_vecs.append(process_latent_conditioning(cond_img, histogram_raw, cond_means, cond_stds, torch.tensor(0.0),
                                                     seed, seed_offset=ctx[1] * 65536 + ctx[2]))
        z = gaussian_noise_patch(seed + seed_offset, ctx[1] * stride, ctx[2] * stride, tile, tile, 5, tile, tile,
                                 device=dev)[None] * sigma_data
        x_t = math.cos(t) * sample + math.sin(t) * z
        model_in.append(x_t / sigma_data)
        kept.append(x_t)
    if not model_in:
        return []
    n = len(model_in)
    x = torch.cat(model_in, dim=0)
    c = torch.cat(cond_vecs, dim=0)
    if pad_batch_to:   # the reference pads to {1,2,4,8,16} under torch.compile (world_pipeline.py:393-398,1107-1118)
        m = padded_batch_size(n, pad_batch_to)
        if m > n:
            x = torch.cat([x, x[:1].repeat(m - n, 1, 1, 1)], dim=0)
            c = torch.cat([c, c[:1].repeat(m - n, 1)], dim=0)
    pred = -model(x, torch.full((x.shape[0],), float(t), device=dev, dtype=torch.float32), [c])
    w = weight_window.to(dev)
    outs = []
    for i in range(n):
        s = (math.cos(t) * kept[i] - math.sin(t) * sigma_data * pred[i:i + 1]).float() / sigma_data
        outs.append(torch.cat([s[0] * w[None], w[None]], dim=0))
    return outs


@torch.no_grad()
def coarse_stage_tile(model, scheduler, seed: int, ctx, synthetic_map: torch.Tensor, t_cond: torch.Tensor,
                      cond_inputs, weight_window: torch.Tensor, coarse_means, coarse_stds, num_steps: int = 20,
                      pool_size: int = 1) -> torch.Tensor:
    """_coarse_inference (world_pipeline.py:909-959).  synthetic_map: [5,64,64] raw conditioning map for this tile (the
    reference builds it with its Perlin/WorldClim machinery, out of scope).  Returns the packed [7,64,64] tile."""
    if pool_size != 1:
        raise NotImplementedError("coarse_pooling > 1 is host-side pooling of the finished tile; not on the GPU path")
    dev = model.device
    tile, stride = 64, 48
    means = torch.as_tensor(coarse_means, dtype=torch.float32, device=dev)
    stds = torch.as_tensor(coarse_stds, dtype=torch.float32, device=dev)
    _, i, j = ctx
    i1, j1 = i * stride, j * stride
    smap = (synthetic_map.to(dev, torch.float32) - means[[0, 2, 3, 4, 5], None, None]) / stds[[0, 2, 3, 4, 5], None, None]
    cond_noise = gaussian_noise_patch(seed, i1, j1, tile, tile, 5, tile, tile, device=dev)[None]
    tc = t_cond.to(dev, torch.float32).view(1, -1, 1, 1)
    cond_img = torch.cos(tc) * smap[None] + torch.sin(tc) * cond_noise
    scheduler.set_timesteps(num_steps)
    noise = gaussian_noise_patch(seed + 1, i1, j1, tile, tile, 6, tile, tile, device=dev)[None]
    sample0 = noise * float(scheduler.sigmas[0])
    solve = get_diffusion_solve(model, scheduler, 1, tile, tile, num_steps)
    sample = solve.run(sample0, cond_img, conditional_inputs=[c.to(dev, torch.float32) for c in cond_inputs])
    sample = sample.float() / scheduler.config.sigma_data
    sample = sample * stds.view(1, -1, 1, 1) + means.view(1, -1, 1, 1)
    sample[0, 1] = sample[0, 0] - sample[0, 1]
    w = weight_window.to(dev)
    return torch.cat([sample[0] * w[None], w[None]], dim=0)
