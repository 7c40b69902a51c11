"""Deterministic inputs shared by tests/golden/make_golden_stages.py (reference side) and tests/test_stages_gpu.py."""
import torch

SEED = 20240607


def stage_inputs():
    """Deterministic inputs shared with tests/test_stages_gpu.py."""
    g = torch.Generator().manual_seed(77)
    d = {}
    d["dec_latents"] = torch.cat([torch.randn(5, 16, 16, generator=g) * 0.6, torch.rand(1, 16, 16, generator=g) + 0.5])
    d["dec_latents"][:5] *= d["dec_latents"][5:]
    lat_c = []
    for _ in range(2):
        c = torch.cat([torch.randn(6, 4, 4, generator=g), torch.rand(1, 4, 4, generator=g) + 0.5])
        c[3, 1, 2] = float("nan")     # a NaN climate cell inside the 2x2 centre crop -> exercises the seeded NaN fill
        c[:6] *= c[6:]
        lat_c.append(c)
    d["lat_cond"] = torch.stack(lat_c)
    d["lat_hist"] = torch.randn(1, 5, generator=g)
    d["lat_means"] = torch.randn(7, generator=g) * 0.1
    d["lat_stds"] = torch.rand(7, generator=g) + 0.5
    d["coarse_map"] = torch.randn(5, 64, 64, generator=g)
    d["coarse_means"] = torch.randn(6, generator=g) * 0.2
    d["coarse_stds"] = torch.rand(6, generator=g) + 0.5
    d["cond_snr"] = torch.tensor([0.3, 0.5, 1.0, 2.0, 4.0])
    # appended last so the tensors above keep their values: the product decoder window (512 / 8 = 64 latent pixels)
    d["dec_latents_512"] = torch.cat([torch.randn(5, 64, 64, generator=g) * 0.6, torch.rand(1, 64, 64, generator=g) + 0.5])
    d["dec_latents_512"][:5] *= d["dec_latents_512"][5:]
    return d


