"""SURVEY 8(a) row a17 on the GPU: the multi-phase InfiniteDiffusion sampler (phases split with the reference's
build_timestep_ranges; every phase = fused sub-solves per tile + blend; the next phase reads the blended canvas)
against the dense fp32 oracle restating annotated_infinite_panorama.py:176-226 / infinite_consistency.py:207-239, with
the decoder U-Net as the denoiser (SD-v1.5 is not available offline).  Bounded and lazy (unbounded) forms."""
import math

import pytest
import torch

from oracle import scheduler as osched
from oracle import tiling as otile
from oracle import unet as ounet
from terrain_diffusion_b200.inference import (infinite_diffusion_canvases, sample_decoder_diffusion_tiled,
                                              sample_infinite_diffusion)
from terrain_diffusion_b200.inference.multiphase import normalize
from terrain_diffusion_b200.inference.noise import gaussian_noise_patch
from terrain_diffusion_b200.models import EDMUnet2D
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler

pytestmark = pytest.mark.gpu
TOL = 1.0e-2
TH = (0.25 * math.log(5.0), 0.25 * math.log(0.3))        # 6-step schedule -> phases of 3 + 1 + 2 steps


def rel_rms(a, b):
    return float((a - b).square().mean().sqrt() / (b.square().mean().sqrt() + 1e-30))


@pytest.fixture(scope="module")
def decoder():
    cfg = ounet.DECODER_CFG
    sd = ounet.procedural_state_dict(cfg, seed=0)
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(sd)
    return m.cuda(), sd, cfg


@pytest.mark.parametrize("tile_batch", [1, 3])
def test_three_phase_bounded_canvas_matches_dense_oracle(decoder, tile_batch):
    m, sd, cfg = decoder
    g = torch.Generator().manual_seed(41)
    noise = torch.randn(1, 1, 128, 96, generator=g) * 80
    cond = torch.randn(1, 4, 128, 96, generator=g)
    sched = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
    got = sample_infinite_diffusion(m, sched, cond.cuda(), noise.cuda(), 64, 32, num_steps=6, thresholds=TH,
                                    tile_batch=tile_batch).cpu()
    ref = otile.sample_infinite_diffusion(lambda x, t: ounet.unet_forward(sd, cfg, x, t, []), osched.OracleScheduler,
                                          cond, noise, 64, 32, 6, TH)
    assert got.shape == ref.shape == (1, 1, 128, 96)
    assert rel_rms(got, ref) < TOL
    # blending between phases matters: the single-phase sampler gives a different canvas
    one = sample_decoder_diffusion_tiled(m, sched, cond.cuda(), noise.cuda(), 64, 32, num_steps=6).cpu()
    assert rel_rms(one, ref) > 5 * rel_rms(got, ref)


def test_no_thresholds_is_the_single_phase_sampler(decoder):
    m, sd, cfg = decoder
    g = torch.Generator().manual_seed(42)
    noise = (torch.randn(1, 1, 96, 96, generator=g) * 80).cuda()
    cond = torch.randn(1, 4, 96, 96, generator=g).cuda()
    sched = EDMDPMSolverMultistepScheduler()
    a = sample_infinite_diffusion(m, sched, cond, noise, 64, 32, num_steps=4, thresholds=())
    b = sample_decoder_diffusion_tiled(m, sched, cond, noise, 64, 32, num_steps=4)
    assert torch.equal(a, b)


def test_lazy_chain_of_phase_canvases_equals_the_dense_oracle_on_the_same_windows(decoder):
    """Unbounded form: windows (i, j) at origin (32 i, 32 j); a slice far enough inside a 3 x 3 block of windows only
    depends on them, so the oracle can be evaluated densely on that block (same tile-seeded noise field)."""
    m, sd, cfg = decoder
    sched = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
    seed, T, S = 99, 64, 32
    cond_field = torch.randn(4, 512, 512, generator=torch.Generator().manual_seed(43))

    def cond_fn(y0, x0, t):
        return cond_field[:, y0 + 128:y0 + 128 + t, x0 + 128:x0 + 128 + t]

    cvs = infinite_diffusion_canvases(m, sched, seed, channels=1, tile_size=T, tile_stride=S, num_steps=6, thresholds=TH,
                                      cond_fn=cond_fn, batch_size=4, cache_limit=64 << 20)
    assert len(cvs) == 3
    # rows/cols [64, 96): covered by windows with origin in {32, 64} x {32, 64} (the 2 x 2 that overlap it) in the last
    # phase; those read phase-2 pixels [32, 128) ... -> dense block [-64, 224) contains every dependency exactly when
    # evaluated with window origins on the same 32-grid: use the block of windows i, j in -2 .. 5
    got = normalize(cvs[-1][:, 64:96, 64:96]).cpu()
    lo, n = -2, 8
    size = (n - 1) * S + T
    y0 = lo * S
    noise = gaussian_noise_patch(seed, y0, y0, size, size, 1, 256, 256, device="cuda").cpu()[None] * float(sched.sigmas[0])
    cond = cond_field[None, :, y0 + 128:y0 + 128 + size, y0 + 128:y0 + 128 + size]
    ref = otile.sample_infinite_diffusion(lambda x, t: ounet.unet_forward(sd, cfg, x, t, []), osched.OracleScheduler,
                                          cond, noise, T, S, 6, TH)
    want = ref[0, :, 64 - y0:96 - y0, 64 - y0:96 - y0]
    assert rel_rms(got, want) < TOL
