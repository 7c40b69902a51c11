"""Lazy unbounded canvas (infinite_tensor semantics) and the three-stage pipeline wiring."""
import math

import pytest
import torch

from oracle import unet as ounet
from terrain_diffusion_b200.inference import LazyCanvas, TensorWindow, TerrainPipeline, decoder_stage_tile
from terrain_diffusion_b200.inference.tiling import linear_weight_window, window_range
from terrain_diffusion_b200.models import EDMUnet2D
from tests.test_oracle_golden import BASE_CFG, COARSE_CFG

pytestmark = pytest.mark.gpu


def _tile(i, j, c, t):
    g = torch.Generator().manual_seed((i + 1000) * 100003 + (j + 1000))
    return torch.randn(c, t, t, generator=g)


def _brute(a, b, c, d, ch, size, stride, off):
    """Sum over every window intersecting the slice, row-major, fp32 -- straight from the Appendix C definition."""
    out = torch.zeros(ch, b - a, d - c)
    for i in window_range(a, b, size, stride, off):
        for j in window_range(c, d, size, stride, off):
            t = _tile(i, j, ch, size)
            y0, x0 = i * stride + off, j * stride + off
            ys, ye, xs, xe = max(a, y0), min(b, y0 + size), max(c, x0), min(d, x0 + size)
            out[:, ys - a:ye - a, xs - c:xe - c] += t[:, ys - y0:ye - y0, xs - x0:xe - x0]
    return out


@pytest.mark.parametrize("size,stride,off", [(64, 48, 0), (64, 32, 0), (16, 16, -5)])
def test_lazy_canvas_equals_bruteforce_window_sum_incl_negative_coordinates(size, stride, off):
    calls = []

    def f(ctx):
        calls.append(ctx)
        return _tile(ctx[1], ctx[2], 3, size).cuda()

    cv = LazyCanvas(3, f, TensorWindow((3, size, size), (3, stride, stride), (0, off, off)), "cuda", block=128)
    for k, (a, b, c, d) in enumerate([(-70, 95, -130, 20), (0, 64, 0, 64), (100, 101, -3, 300)]):
        got = cv[:, a:b, c:d].cpu()
        want = _brute(a, b, c, d, 3, size, stride, off)
        # every read sums its windows in row-major order, whatever order they were computed in: bit-identical
        assert torch.equal(got, want), (k, a, b, c, d)
    n = len(calls)
    cv[:, 0:64, 0:64]                      # cached: no window is computed twice
    assert len(calls) == n and len(set(calls)) == n


def test_cache_limit_evicts_least_recently_used_windows_and_recomputes_them():
    """Byte-limited window cache (the reference's MemoryTileStore(cache_size_bytes), world_pipeline.py:666-674): HBM use
    stays bounded while exploring, evicted windows are recomputed on the next request, results do not change."""
    calls = []

    def f(ctx):
        calls.append(ctx)
        return _tile(ctx[1], ctx[2], 2, 32).cuda()

    tile_bytes = 2 * 32 * 32 * 4
    cv = LazyCanvas(2, f, TensorWindow((2, 32, 32), (2, 24, 24)), "cuda", cache_limit=6 * tile_bytes)
    n_win = len(cv.windows_for(0, 60, 0, 60))             # window indices -1..2 per axis: 16 windows > the limit of 6
    assert n_win == 16
    first = cv[:, 0:60, 0:60].cpu()
    assert torch.equal(first, _brute(0, 60, 0, 60, 2, 32, 24, 0))
    assert cv.cache_bytes <= 6 * tile_bytes and cv.windows_evicted == n_win - 6 and len(cv.done) == 6
    n = len(calls)
    for k in range(8):                                    # walk away: memory stays bounded
        a = 200 * (k + 1)
        assert torch.equal(cv[:, a:a + 40, -a:-a + 40].cpu(), _brute(a, a + 40, -a, -a + 40, 2, 32, 24, 0))
        assert cv.cache_bytes <= 6 * tile_bytes
    again = cv[:, 0:60, 0:60].cpu()                       # everything was evicted meanwhile: recomputed, same values
    assert torch.equal(again, first)
    assert len(calls) >= n + n_win
    unbounded = LazyCanvas(2, f, TensorWindow((2, 32, 32), (2, 24, 24)), "cuda")
    unbounded[:, 0:60, 0:60]
    assert unbounded.windows_evicted == 0 and len(unbounded.done) == n_win


def test_dependency_windows_use_the_same_window_index_and_batches():
    """latent <- coarse wiring of the reference: size 4, stride 1, offset -1 (world_pipeline.py:1147)."""
    base = LazyCanvas(1, lambda ctx: torch.full((1, 8, 8), float(ctx[1] * 100 + ctx[2])).cuda(),
                      TensorWindow((1, 8, 8), (1, 8, 8)), "cuda", block=64)
    seen = {}

    def f(ctxs, deps):
        assert isinstance(ctxs, list) and len(ctxs) <= 3
        outs = []
        for ctx, dep in zip(ctxs, deps):
            seen[ctx] = dep.clone()
            outs.append(torch.ones(1, 4, 4, device="cuda"))
        return outs

    top = LazyCanvas(1, f, TensorWindow((1, 4, 4), (1, 2, 2)), "cuda", args=(base,),
                     args_windows=(TensorWindow((1, 4, 4), (1, 1, 1), (0, -1, -1)),), batch_size=3, block=64)
    top[:, 0:6, 0:6]
    for (_, i, j), dep in seen.items():
        assert torch.equal(dep, base[:, i - 1:i + 3, j - 1:j + 3])


def test_a_request_fills_the_batches_of_its_dependencies():
    """A canvas that needs several missing windows first makes each dependency compute the UNION of what those windows
    will read (row-major), so a batched dependency sees full batches instead of one small request per window -- and
    exactly the windows the per-window reads would have computed, with the same values."""
    sizes = []

    def f_mid(ctxs):
        sizes.append(len(ctxs))
        return [torch.full((1, 8, 8), float(c[1] * 10 + c[2]), device="cuda") for c in ctxs]

    def build(batch):
        sizes.clear()
        mid = LazyCanvas(1, f_mid, TensorWindow((1, 8, 8), (1, 4, 4)), "cuda", batch_size=batch)
        top = LazyCanvas(1, lambda ctx, dep: dep * 2.0, TensorWindow((1, 8, 8), (1, 8, 8)), "cuda", args=(mid,),
                         args_windows=(TensorWindow((1, 8, 8), (1, 8, 8)),))
        return mid, top

    mid, top = build(16)
    out = top[:, 0:32, 0:32]                                   # 4 x 4 top windows, each reads 3 x 3 mid windows
    assert mid.windows_computed == 9 * 9 == sum(sizes)         # the union: mid windows -1..7 per axis, each ONCE
    assert sizes[:5] == [16] * 5 and sizes[5] == 1 and len(sizes) == 6
    one_by_one = LazyCanvas(1, lambda ctx: f_mid([ctx])[0], TensorWindow((1, 8, 8), (1, 4, 4)), "cuda")
    ref = LazyCanvas(1, lambda ctx, dep: dep * 2.0, TensorWindow((1, 8, 8), (1, 8, 8)), "cuda", args=(one_by_one,),
                     args_windows=(TensorWindow((1, 8, 8), (1, 8, 8)),))
    assert torch.equal(out, ref[:, 0:32, 0:32])
    # a dependency whose cache limit cannot hold the union is left alone: windows are computed on demand as before
    sizes.clear()
    small = LazyCanvas(1, f_mid, TensorWindow((1, 8, 8), (1, 4, 4)), "cuda", batch_size=16, cache_limit=20 * 8 * 8 * 4)
    top2 = LazyCanvas(1, lambda ctx, dep: dep * 2.0, TensorWindow((1, 8, 8), (1, 8, 8)), "cuda", args=(small,),
                      args_windows=(TensorWindow((1, 8, 8), (1, 8, 8)),))
    assert torch.equal(top2[:, 0:32, 0:32], out) and max(sizes) <= 9


def test_three_stage_pipeline_slice_equals_direct_stage_evaluation():
    def build(cfg):
        m = EDMUnet2D(**cfg).eval()
        m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
        return m.cuda()

    coarse, base, dec = build(COARSE_CFG), build(BASE_CFG), build(ounet.DECODER_CFG)
    g = torch.Generator().manual_seed(3)

    def cond_fn(i1, i2, j1, j2):
        gg = torch.Generator().manual_seed(i1 * 7919 + j1 + 12345)
        return torch.randn(5, i2 - i1, j2 - j1, generator=gg)

    pipe = TerrainPipeline(coarse, base, dec, seed=7, conditioning_fn=cond_fn,
                           coarse_means=(torch.randn(6, generator=g) * 0.1).tolist(),
                           coarse_stds=(torch.rand(6, generator=g) + 0.5).tolist(), cond_snr=[0.3, 0.5, 1.0, 2.0, 4.0],
                           histogram_raw=torch.randn(5, generator=g), latents_means=torch.zeros(7),
                           latents_stds=torch.ones(7), decoder_tile_size=128, decoder_tile_stride=96)
    out = pipe.residual[:, 40:90, -20:70]          # rows: decoder window 0 only; columns: windows -1 and 0
    assert out.shape == (2, 50, 90) and torch.isfinite(out).all() and float(out[1].min()) > 0
    # the decoder window (0,0,0) recomputed directly from the blended latent canvas must be what the canvas summed
    lat = pipe.latents[:, 0:16, 0:16]
    t0 = decoder_stage_tile(dec, 7, (0, 0, 0), lat, linear_weight_window(128, "cuda"), [pipe.t_init], 128, 96)
    lat_m1 = pipe.latents[:, 0:16, -12:4]
    t1 = decoder_stage_tile(dec, 7, (0, 0, -1), lat_m1, linear_weight_window(128, "cuda"), [pipe.t_init], 128, 96)
    # summation order in the canvas is row-major over window indices: (0,-1) then (0,0)
    ref = torch.zeros(2, 50, 90, device="cuda")
    ref[:, :, :52] += t1[:, 40:90, 76:128]         # window -1 covers columns [-96,32)
    ref[:, :, 20:] += t0[:, 40:90, 0:70]           # window 0 covers columns [0,128)
    assert torch.equal(out, ref)
    assert pipe.coarse.windows_computed >= 1 and pipe.latents_init.windows_computed > pipe.latents.windows_computed
    n = pipe.residual_normalized(40, -20, 90, 70)
    assert torch.isfinite(n).all()


def test_world_pipeline_drop_in_surface_from_pretrained_to_bind_get(tmp_path):
    """SURVEY 8(b) Pipeline row: WorldPipeline.from_pretrained(path, seed=..., latents_batch_size=..., torch_compile=...,
    dtype=..., cache_limit=..., **kw).to(device).bind() -- the call chain of api.py:37-52 / tiff_export.py:95-139 /
    latency.py:39-94 -- then .get() -> CPU tensors and .residual / .latents / .coarse -> un-normalised CPU fp32 windows
    (negative indices legal), equal to the device-resident TerrainPipeline with the same seed."""
    from terrain_diffusion_b200.inference import WorldPipeline

    def build(cfg):
        m = EDMUnet2D(**cfg).eval()
        m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
        return m

    def cond_fn(i1, i2, j1, j2):
        gg = torch.Generator().manual_seed(i1 * 7919 + j1 + 12345)
        return torch.randn(5, i2 - i1, j2 - j1, generator=gg)

    kw = dict(decoder_tile_size=128, decoder_tile_stride=96, residual_mean=0.1, residual_std=1.2)
    src = WorldPipeline.from_local_models(build(COARSE_CFG), build(BASE_CFG), build(ounet.DECODER_CFG), seed=5, **kw)
    src.save_pretrained(tmp_path)
    assert (tmp_path / "config.json").exists() and (tmp_path / "base_model" / "config.json").exists()
    world = WorldPipeline.from_pretrained(str(tmp_path), seed=11, latents_batch_size=[1, 2, 4, 8, 16],
                                          torch_compile=True, dtype="bf16", caching_strategy="direct",
                                          cache_limit=64 << 20, conditioning_fn=cond_fn)
    assert world.seed == 11 and world.decoder_tile_size == 128 and world.residual is None
    world = world.to("cuda").bind(hdf5_file=None)
    assert world.device.type == "cuda"
    out = world.get(-30, 10, 34, 106, with_climate=True)
    assert out["elev"].device.type == "cpu" and out["elev"].shape == (64, 96) and out["elev"].dtype == torch.float32
    assert out["climate"].device.type == "cpu" and out["climate"].shape == (5, 64, 96)
    assert torch.isfinite(out["elev"]).all() and torch.isfinite(out["climate"]).all()
    r = world.residual[:, -30:34, 10:106]
    assert r.device.type == "cpu" and r.shape == (2, 64, 96) and float(r[1].min()) > 0
    assert world.latents[:, -4:4, 0:12].shape == (6, 8, 12) and world.coarse[:, 0:2, -1:1].shape == (7, 2, 2)
    # same seed, device-resident variant: identical windows
    ref = TerrainPipeline(world.coarse_model, world.base_model, world.decoder_model, seed=11, conditioning_fn=cond_fn,
                          coarse_means=world.kwargs["coarse_means"], coarse_stds=world.kwargs["coarse_stds"],
                          cond_snr=world.kwargs["cond_snr"], histogram_raw=world.kwargs["histogram_raw"],
                          latents_means=world.cond_input_mean, latents_stds=world.cond_input_std,
                          decoder_tile_size=128, decoder_tile_stride=96, residual_mean=0.1, residual_std=1.2)
    ref_out = ref.get(-30, 10, 34, 106, with_climate=True)        # the same request sequence -> the same tile batches
    assert ref_out["elev"].is_cuda and torch.equal(ref_out["elev"].cpu(), out["elev"])
    assert torch.equal(ref_out["climate"].cpu(), out["climate"])
    assert torch.equal(ref.residual[:, -30:34, 10:106].cpu(), r)
    assert world.change_seed(12) is True and len(world._residual.done) == 0
    with pytest.raises(NotImplementedError):
        WorldPipeline(caching_strategy="indirect").from_local_models(None, None, None, caching_strategy="indirect").bind("TEMP")
