"""Pin the oracle (oracle/*.py, oracle/rng.c) against golden vectors recorded from the UNMODIFIED reference
(tests/golden/reference_golden.npz, written by tests/golden/make_golden.py) and against SURVEY.md Appendix E."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import rng as orng  # noqa: E402
from oracle import scheduler as osched  # noqa: E402
from oracle import tiling as otile  # noqa: E402
from oracle import unet as ounet  # noqa: E402

G = np.load(ROOT / "tests" / "golden" / "reference_golden.npz")


def _gen_inputs(cfg, n, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cfg["in_channels"], hw, hw, generator=g)
    t = torch.atan(torch.exp(torch.randn(n, generator=g) * 1.5) / 0.5)
    cond = []
    for kind, dim, _w in cfg.get("conditional_inputs") or []:
        if kind == "tensor":
            cond.append(torch.randn(n, dim, generator=g))
        elif kind == "float":
            cond.append(torch.randn(n, generator=g))
        else:
            cond.append(torch.randint(0, dim, (n,), generator=g))
    return x, t, cond


TINY_CFGS = {
    "tiny_dec": dict(image_size=32, in_channels=5, out_channels=1, model_channels=16, model_channel_mults=[1, 2],
                     layers_per_block=1, attn_resolutions=[], midblock_attention=False, concat_balance=0.5,
                     conditional_inputs=[], fourier_scale="pos"),
    "tiny_attn_cond": dict(image_size=16, in_channels=3, out_channels=3, model_channels=16,
                           model_channel_mults=[1, 2], layers_per_block=1, attn_resolutions=[8],
                           midblock_attention=True, concat_balance=0.3,
                           conditional_inputs=[["tensor", 7, 0.5], ["float", 8, 0.2], ["embedding", 5, 0.3]],
                           fourier_scale=1, block_kwargs={"channels_per_head": 8}),
}
BASE_CFG = dict(image_size=512, in_channels=5, out_channels=5, model_channels=192, model_channel_mults=[1, 2, 3, 4],
                layers_per_block=3, attn_resolutions=[8, 16], midblock_attention=True, concat_balance=0.5,
                conditional_inputs=[["tensor", 58, 1.0]], fourier_scale="pos", block_kwargs={"dropout": 0.1})
COARSE_CFG = dict(image_size=16, in_channels=11, out_channels=6, model_channels=128, model_channel_mults=[1],
                  layers_per_block=2, attn_resolutions=[], midblock_attention=False, concat_balance=0.5,
                  conditional_inputs=[["float", 64, 0.2]] * 5, fourier_scale="pos", block_kwargs={})


# ------------------------------------------------------------------------------------------------ U-Net
@pytest.mark.parametrize("name", list(TINY_CFGS))
def test_unet_tiny_matches_reference(name):
    cfg = TINY_CFGS[name]
    sd = ounet.procedural_state_dict(cfg, seed=3)
    x, t, cond = _gen_inputs(cfg, 2, cfg["image_size"], seed=11)
    y = ounet.unet_forward(sd, cfg, x, t, cond)
    np.testing.assert_allclose(y.numpy(), G[f"{name}.y"], rtol=0, atol=2e-5)


def test_unet_decoder_full_config_matches_reference():
    cfg = ounet.DECODER_CFG
    sd = ounet.procedural_state_dict(cfg, seed=0)
    assert sum(v.numel() for v in sd.values()) == 27922533 + 32 + 128 + 128  # params + freqs/phases buffers
    x, t, cond = _gen_inputs(cfg, 1, 64, seed=1)
    y = ounet.unet_forward(sd, cfg, x, t, cond)
    assert float(y.std()) > 0.5  # non-vacuous (emb_gain / out_gain are non-zero)
    np.testing.assert_allclose(y.numpy(), G["decoder.y"], rtol=0, atol=2e-5)


def test_unet_decoder_128_batch2_matches_reference():
    cfg = ounet.DECODER_CFG
    sd = ounet.procedural_state_dict(cfg, seed=0)
    x, t, cond = _gen_inputs(cfg, 2, 128, seed=2)
    y = ounet.unet_forward(sd, cfg, x, t, cond)
    np.testing.assert_allclose(y.numpy(), G["decoder128.y"], rtol=0, atol=3e-5)


def test_unet_coarse_config_matches_reference():
    sd = ounet.procedural_state_dict(COARSE_CFG, seed=0)
    x, t, cond = _gen_inputs(COARSE_CFG, 1, 64, seed=1)
    y = ounet.unet_forward(sd, COARSE_CFG, x, t, cond)
    np.testing.assert_allclose(y.numpy(), G["coarse.y"], rtol=0, atol=2e-5)


@pytest.mark.slow
def test_unet_base_config_matches_reference():
    sd = ounet.procedural_state_dict(BASE_CFG, seed=0)
    x, t, cond = _gen_inputs(BASE_CFG, 1, 64, seed=1)
    y = ounet.unet_forward(sd, BASE_CFG, x, t, cond)
    np.testing.assert_allclose(y.numpy(), G["base.y"], rtol=0, atol=5e-5)


def test_block_plan_matches_survey_names():
    enc, dec = ounet.block_plan(ounet.DECODER_CFG)
    assert [b["name"] for b in enc][:5] == ["512x512_conv", "512x512_block0", "512x512_block1", "512x512_block2",
                                            "256x256_down"]
    assert len(enc) == 16 and len(dec) == 21
    assert [b["cin"] for b in dec if b["name"].startswith("64x64_block")] == [512, 512, 512, 448]
    assert [b["cin"] for b in dec if b["name"].startswith("512x512_block")] == [192, 128, 128, 128]


# ------------------------------------------------------------------------------------------------ scheduler
@pytest.mark.parametrize("n", [4, 12, 20])
def test_scheduler_tables_and_steps(n):
    s = osched.OracleScheduler()
    s.set_timesteps(n)
    np.testing.assert_array_equal(s.sigmas.numpy(), G[f"sched{n}.sigmas"])
    np.testing.assert_array_equal(s.timesteps.numpy(), G[f"sched{n}.timesteps"])
    np.testing.assert_array_equal(s.trigflow_precondition_noise(s.sigmas[:-1]).numpy(), G[f"sched{n}.cnoise"])
    traj = torch.from_numpy(G[f"sched{n}.traj"])
    g = torch.Generator().manual_seed(100 + n)
    x = torch.randn(1, 1, 8, 8, generator=g) * 80
    for i, (t, sigma) in enumerate(zip(s.timesteps, s.sigmas)):
        f = torch.randn(1, 1, 8, 8, generator=g)
        assert torch.equal(s.precondition_inputs(x, sigma), traj[2 * i, 0])
        assert torch.equal(f, traj[2 * i, 1])
        x = s.step(f, t, x)
        assert torch.equal(x, traj[2 * i + 1, 0]), f"step {i}"


def test_scheduler_closed_form_coefficients():
    """The fp64 closed form (what the CUDA step kernel is fed) reproduces the reference trajectory to fp32 accuracy."""
    for n in (4, 12, 20):
        s = osched.OracleScheduler()
        s.set_timesteps(n)
        co = osched.step_coefficients(s.sigmas.double().tolist())
        traj = torch.from_numpy(G[f"sched{n}.traj"])
        g = torch.Generator().manual_seed(100 + n)
        x = (torch.randn(1, 1, 8, 8, generator=g) * 80).double()
        x0_prev = torch.zeros_like(x)
        for i, c in enumerate(co):
            f = torch.randn(1, 1, 8, 8, generator=g).double()
            x0 = c["c_skip"] * x + c["c_out"] * f
            x = c["r"] * x + (1 - c["r"]) * x0 + c["k"] * (x0 - x0_prev)
            x0_prev = x0
            ref = traj[2 * i + 1, 0].double()
            assert float((x - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max())), (n, i)
    # Appendix E: 4-step sigma table
    s = osched.OracleScheduler()
    s.set_timesteps(4)
    assert s.sigmas.tolist() == [79.99998474121094, 9.723200798034668, 0.46997925639152527, 0.0019999996293336153,
                                 0.0]


def test_scheduler_requires_set_timesteps():
    s = osched.OracleScheduler()
    with pytest.raises(ValueError):
        s.step(torch.zeros(1), torch.tensor(0.0), torch.zeros(1))


# ------------------------------------------------------------------------------------------------ tiling
def test_tile_starts_golden():
    cases, lens, flat = G["tile_starts.cases"], G["tile_starts.lens"], G["tile_starts.flat"]
    off = 0
    for (length, tile, stride), n in zip(cases.tolist(), lens.tolist()):
        assert otile.tile_starts(length, tile, stride) == flat[off:off + n].tolist()
        off += n
    assert otile.tile_starts(1664, 512, 384) == [0, 384, 768, 1152]
    assert otile.tile_starts(100, 64, 48) == [0, 36]


def test_weight_window_golden():
    for size in (4, 8, 64):
        np.testing.assert_array_equal(otile.linear_weight_window(size).numpy(), G[f"window{size}"])
    w = otile.linear_weight_window(512)
    probe = np.array([w[0, 0], w[0, 255], w[255, 255], w[511, 300], w[17, 401]], dtype=np.float32)
    np.testing.assert_array_equal(probe, G["window512.probe"])
    assert float(w.double().sum()) == float(G["window512.sum64"])
    assert float(w[0, 0]) == 9.999743042499176e-07


def test_window_range_rule():
    # cold get(0,0,512,512) -> decoder windows k in [-1, 1]; latents [-48,112) -> k in [-3, 3]
    assert list(otile.window_range(-48, 560, 512, 384)) == [-1, 0, 1]
    assert list(otile.window_range(-48, 112, 64, 32)) == [-3, -2, -1, 0, 1, 2, 3]
    assert list(otile.window_range(16336, 24624, 512, 384)) == list(range(42, 65))
    # brute force cross-check incl. negative coords and offsets
    for (a, b, size, stride, off) in [(-100, 37, 64, 48, 0), (5, 6, 4, 1, -1), (-7, -3, 8, 3, 2), (0, 512, 512, 384, 0)]:
        brute = [k for k in range(-400, 400) if k * stride + off < b and k * stride + off + size > a]
        assert list(otile.window_range(a, b, size, stride, off)) == brute


# ------------------------------------------------------------------------------------------------ RNG
def test_rng_appendix_e_vectors():
    assert orng.tile_seed(1, 0, 0) == 7046029251746621361
    assert orng.tile_seed(1, -1, 2) == 18446744066279796218
    assert orng.tile_seed(123456789, 5, -7) == 4397873024654090011
    assert orng.next_seed(1) == 14210067475669473140 and orng.next_seed(42) == 1039766031909981117
    assert orng.standard_normal(42, 8).tolist() == [
        -0.06246672943234444, -0.6764206290245056, 0.3714064657688141, 0.4223460555076599, 1.0910028219223022,
        -1.0840394496917725, -0.6641873717308044, 0.9288548827171326]
    assert orng.standard_normal(0x5EED0001, 4, np.float64).tolist() == [
        -0.7930057917628103, -2.0632548555210044, 1.2479833527356219, -0.992191172690882]


def test_rng_matches_reference_numba_streams():
    for key in G.files:
        if key.startswith("normal."):
            _, seed, n = key.split(".")
            np.testing.assert_array_equal(orng.standard_normal(int(seed), int(n)), G[key])
    np.testing.assert_array_equal(orng.standard_normal(0x5EED0001, 64, np.float64), G["normal64.1592590337.64"])
    assert [orng.next_seed(s) for s in (1, 42, 2 ** 63 + 5)] == G["next_seed"].tolist()


def test_rng_tile_seed_and_patches_match_reference():
    for (base, (_, ty, tx)), want in zip(zip(G["tile_seed.base_u64"].tolist(), G["tile_seed.args"].tolist()),
                                         G["tile_seed.out"].tolist()):
        assert orng.tile_seed(base, ty, tx) == want
        assert orng.py_tile_seed(base, ty, tx) == want
    for i, (seed, y0, x0, h, w, c, th, tw) in enumerate(G["patch.args"].tolist()):
        np.testing.assert_array_equal(orng.gaussian_noise_patch(seed, y0, x0, h, w, c, th, tw), G[f"patch.{i}"])


def test_rng_python_restatement_agrees_with_c():
    got = np.array(orng.py_standard_normal(987654321, 300))
    np.testing.assert_array_equal(got.astype(np.float32), orng.standard_normal(987654321, 300))


# ------------------------------------------------------------------------------------------------ tiled samplers
def test_cfg1_single_tile_4_step_solve_matches_reference():
    """BASELINE configs[0]: one 64x64 tile, 4-step scheduler, decoder U-Net, fp32 CPU."""
    cfg = ounet.DECODER_CFG
    sd = ounet.procedural_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(1, 1, 64, 64, generator=g) * 80
    cond = torch.randn(1, 4, 64, 64, generator=g)
    y = otile.sample_decoder_diffusion_tiled(lambda x, t: ounet.unet_forward(sd, cfg, x, t, []),
                                             osched.OracleScheduler, cond, noise, 64, 64, num_steps=4)
    np.testing.assert_allclose(y.numpy(), G["cfg1.y"], rtol=0, atol=1e-4)


def test_consistency_multi_tile_blend_matches_reference():
    cfg = ounet.DECODER_CFG
    sd = ounet.procedural_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(1, 1, 96, 96, generator=g)
    cond = torch.randn(1, 4, 96, 96, generator=g)
    y = otile.sample_decoder_consistency_tiled(lambda x, t: ounet.unet_forward(sd, cfg, x, t, []), 0.5,
                                               float(osched.karras_sigmas(1000)[0]), cond,
                                               noise, 64, 32, intermediate_t=[0.6])
    np.testing.assert_allclose(y.numpy(), G["consistency96.y"], rtol=0, atol=1e-4)
