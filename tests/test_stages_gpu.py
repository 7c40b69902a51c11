"""the three pipeline stage callbacks on the GPU path vs goldens produced by the
reference's own method bodies (tests/golden/make_golden_stages.py).  Tolerance: bf16 U-Net vs the fp32 reference,
rel-RMS <= 1e-2 on the value planes; the weight plane (last channel) and the conditioning vector are exact / fp32."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import unet as ounet
from terrain_diffusion_b200.inference import (coarse_stage_tile, decoder_stage_tile, latent_stage_tiles,
                                              process_latent_conditioning)
from terrain_diffusion_b200.inference.tiling import linear_weight_window
from terrain_diffusion_b200.models import EDMUnet2D
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler
from tests._stage_inputs import SEED, stage_inputs
from tests.test_oracle_golden import BASE_CFG, COARSE_CFG

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).resolve().parent / "golden" / "stages_golden.npz")
TOL = 1.0e-2


def rel_rms(a, b):
    return float((a - b).square().mean().sqrt() / (b.square().mean().sqrt() + 1e-30))


def build(cfg):
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
    return m.cuda()


def check_packed(got, want):
    got, want = got.cpu(), torch.from_numpy(want)
    assert got.shape == want.shape
    assert torch.equal(got[-1], want[-1])                       # blend weights: bit-exact
    assert rel_rms(got[:-1], want[:-1]) < TOL


def test_decoder_stage_matches_reference_method():
    inp = stage_inputs()
    m = build(ounet.DECODER_CFG)
    ww = linear_weight_window(128)
    sig0 = float(EDMDPMSolverMultistepScheduler().sigmas[0])
    import math
    t0 = math.atan(sig0 / 0.5)
    check_packed(decoder_stage_tile(m, SEED, (0, 2, -1), inp["dec_latents"].clone(), ww, [t0], 128, 96),
                 G["decoder_1step"])
    check_packed(decoder_stage_tile(m, SEED, (0, 2, -1), inp["dec_latents"].clone(), ww,
                                    [t0, math.atan(0.065 / 0.5)], 128, 96), G["decoder_2step"])


def test_latent_conditioning_vector_and_stage_match_reference_method():
    import math
    inp = stage_inputs()
    c0 = torch.cat([inp["lat_cond"][0][:-1] / inp["lat_cond"][0][-1:], torch.ones(1, 4, 4)])[None].cuda()
    vec = process_latent_conditioning(c0, inp["lat_hist"], inp["lat_means"], inp["lat_stds"], torch.tensor(0.0), SEED,
                                      seed_offset=65538).cpu()
    np.testing.assert_allclose(vec.numpy(), G["latent_condvec"], rtol=0, atol=2e-6)
    m = build(BASE_CFG)
    ww = linear_weight_window(64)
    sig0 = float(EDMDPMSolverMultistepScheduler().sigmas[0])
    ctxs = [(0, 1, 2), (0, -1, 0)]
    p1 = latent_stage_tiles(m, SEED, ctxs, None, [c.clone() for c in inp["lat_cond"]], math.atan(sig0 / 0.5), ww,
                            inp["lat_hist"], inp["lat_means"], inp["lat_stds"], seed_offset=int(G["latent_seed_offsets"][0]),
                            pad_batch_to=16)
    for got, want in zip(p1, G["latent_phase1"]):
        check_packed(got, want)
    # phase 2 consumes the REFERENCE's phase-1 tiles so the two phases are checked independently
    prev = [torch.from_numpy(x) for x in G["latent_phase1"]]
    p2 = latent_stage_tiles(m, SEED, ctxs, prev, [c.clone() for c in inp["lat_cond"]], math.atan(0.35 / 0.5), ww,
                            inp["lat_hist"], inp["lat_means"], inp["lat_stds"], seed_offset=int(G["latent_seed_offsets"][1]))
    for got, want in zip(p2, G["latent_phase2"]):
        check_packed(got, want)


def test_coarse_stage_20_step_solve_matches_reference_method():
    inp = stage_inputs()
    m = build(COARSE_CFG)
    t_cond = torch.atan(inp["cond_snr"])
    cond_inputs = [v.view(-1) for v in torch.log(torch.tan(t_cond) / 8.0)]
    got = coarse_stage_tile(m, EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5), SEED,
                            (0, 1, -2), inp["coarse_map"].clone(), t_cond, cond_inputs, linear_weight_window(64),
                            inp["coarse_means"].tolist(), inp["coarse_stds"].tolist())
    check_packed(got, G["coarse"])
