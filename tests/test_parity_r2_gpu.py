"""Round-2 parity holes (task 4), all through the C ABI on the GPU:

* BASELINE config 3 geometry -- 1664^2 canvas, tile 512 / stride 384, 16 tiles, `tile_batch` 1 and 4 -- against the
  reference model + reference scheduler run tile by tile (tests/golden/make_golden_r2.py, stored every 8th pixel);
* one 512 x 512 forward (the product tile: 1 375 GFLOP, several rounds of work items per SM) against the reference;
* batch-16 rows of a 256 x 256 forward == the same rows evaluated alone (the configuration the throughput numbers use);
* the decoder stage at the product tile 512 / stride 384 against the reference's own `_decoder_inference`;
* the second half of the SURVEY 8(c) tolerance: our error vs fp32 <= 1.25 x the reference's OWN bf16-vs-fp32 error on
  the same inputs (recorded by the generator from the unmodified reference in bf16).
"""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import unet as ounet
from terrain_diffusion_b200.inference import decoder_stage_tile, sample_decoder_diffusion_tiled
from terrain_diffusion_b200.inference.tiling import linear_weight_window, tile_starts
from terrain_diffusion_b200.models import EDMUnet2D
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler
from tests._stage_inputs import SEED, stage_inputs

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
G2 = np.load(HERE / "golden" / "parity_r2_golden.npz")
GREF = np.load(HERE / "golden" / "reference_golden.npz")
GST = np.load(HERE / "golden" / "stages_golden.npz")
TOL = 1.0e-2


def rel_rms(a, b):
    return float((a - b).square().mean().sqrt() / (b.square().mean().sqrt() + 1e-30))


def _gen_inputs(cfg, n, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cfg["in_channels"], hw, hw, generator=g)
    t = torch.atan(torch.exp(torch.randn(n, generator=g) * 1.5) / 0.5)
    return x, t


@pytest.fixture(scope="module")
def decoder():
    cfg = ounet.DECODER_CFG
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(ounet.procedural_state_dict(cfg, seed=0))
    return m.cuda(), cfg


def test_forward_512_product_tile_matches_reference(decoder):
    m, cfg = decoder
    x, t = _gen_inputs(cfg, 1, 512, seed=11)
    y = m(x.cuda(), t.cuda(), []).cpu()
    ref = torch.from_numpy(G2["fwd512_sub4"])
    assert float(ref.std()) > 0.3
    assert rel_rms(y[:, :, ::4, ::4], ref) < TOL


@pytest.mark.parametrize("tile_batch", [1, 4])
def test_config3_geometry_16_tiles_matches_reference(decoder, tile_batch):
    m, cfg = decoder
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(1, 1, 1664, 1664, generator=g)
    cond = torch.randn(1, 4, 1664, 1664, generator=g)
    sched = EDMDPMSolverMultistepScheduler(sigma_min=0.002, sigma_max=80, sigma_data=0.5)
    sched.set_timesteps(2)
    assert tile_starts(1664, 512, 384) == [int(v) for v in G2["cfg3_starts"]] == [0, 384, 768, 1152]
    got = sample_decoder_diffusion_tiled(m, sched, cond.cuda(), (noise * float(sched.sigmas[0])).cuda(), 512, 384,
                                         num_steps=2, tile_batch=tile_batch).cpu()
    ref = torch.from_numpy(G2["cfg3_sub8"])
    assert got.shape == (1, 1, 1664, 1664)
    assert rel_rms(got[:, :, ::8, ::8], ref) < TOL


def test_batch16_rows_equal_single_tile_rows(decoder):
    """Batching tiles only changes how work items are spread over CTAs (other N / split-K choices): every row of a
    16-tile evaluation must be the single-tile result up to the accumulation order."""
    m, cfg = decoder
    x, t = _gen_inputs(cfg, 16, 256, seed=31)
    yb = m(x.cuda(), t.cuda(), []).cpu()
    for i in (0, 7, 15):
        y1 = m(x[i:i + 1].cuda(), t[i:i + 1].cuda(), []).cpu()
        # both are bf16 pipelines: other work-item shapes only change fp32 summation order, which moves bf16 roundings of
        # intermediates (measured 3.4e-3 end to end; the error against fp32 is ~5e-3 for either)
        assert rel_rms(yb[i:i + 1], y1) < 6e-3
    ref = ounet.unet_forward(ounet.procedural_state_dict(cfg, seed=0), cfg, x[3:4], t[3:4], [])
    assert rel_rms(yb[3:4], ref) < TOL


def test_decoder_stage_product_tile_512_384_matches_reference_method(decoder):
    m, cfg = decoder
    inp = stage_inputs()
    t0 = math.atan(float(EDMDPMSolverMultistepScheduler().sigmas[0]) / 0.5)
    got = decoder_stage_tile(m, SEED, (0, -1, 3), inp["dec_latents_512"].clone(), linear_weight_window(512), [t0], 512,
                             384).cpu()
    want = torch.from_numpy(GST["decoder_512_sub4"])
    assert got.shape == (2, 512, 512)
    assert torch.equal(got[-1, ::4, ::4], want[-1])                       # blend weights: bit-exact
    assert rel_rms(got[:-1, ::4, ::4], want[:-1]) < TOL


@pytest.mark.parametrize("n,hw,seed,key", [(1, 64, 1, "decoder.y"), (2, 128, 2, "decoder128.y"), (1, 256, 7, None)])
def test_error_is_within_1p25x_of_the_references_own_bf16_error(decoder, n, hw, seed, key):
    m, cfg = decoder
    x, t = _gen_inputs(cfg, n, hw, seed)
    ref = torch.from_numpy(GREF[key]) if key else ounet.unet_forward(ounet.procedural_state_dict(cfg, seed=0), cfg, x, t, [])
    err = rel_rms(m(x.cuda(), t.cuda(), []).cpu(), ref)
    ref_bf16 = float(G2[f"ref_bf16_err_{hw}"])
    assert 2e-3 < ref_bf16 < 2e-2, ref_bf16
    assert err <= 1.25 * ref_bf16, (err, ref_bf16)
    assert err < TOL
