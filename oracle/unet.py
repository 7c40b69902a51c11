"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference's EDM2 magnitude-preserving U-Net.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
It is a functional re-statement (state_dict + config in, tensor out) of xandergos/terrain-diffusion @ 82a0431:

  normalize / resample / mp_silu / mp_sum / mp_concat      terrain_diffusion/models/mp_layers.py:9-86
  MPPositionalEmbedding / MPFourier                        terrain_diffusion/models/mp_layers.py:88-131
  MPConv.forward (weight re-normalisation on every call)   terrain_diffusion/models/mp_layers.py:201-221
  UNetBlock.forward / UNetBlock.attn                       terrain_diffusion/models/unet_block.py:102-156
  EDMUnet2D.__init__ (module plan) / compute_embeddings /
  forward                                                  terrain_diffusion/models/edm_unet.py:67-184

Parity pinned: tests/test_oracle_golden.py checks this file against outputs of the unmodified reference recorded by
tests/golden/make_golden.py (tiny and full decoder/base/coarse configs, procedural weights).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- mp_layers.py
def normalize(x: torch.Tensor, dim=None, eps: float = 1e-4) -> torch.Tensor:
    """mp_layers.py:9-12 -- x / (eps + ||x|| * sqrt(norm.numel / x.numel))."""
    norm = torch.linalg.vector_norm(x, dim=dim, keepdim=True)
    norm = torch.add(eps, norm, alpha=np.sqrt(norm.numel() / x.numel()))
    return x / norm


def resample(x: torch.Tensor, mode: str) -> torch.Tensor:
    """mp_layers.py:15-30 with the 'pooling' filters written out: a depthwise 1x1 ones conv with stride 2 is a
    subsample; a depthwise 2x2 ones transposed conv with stride 2 is nearest x2 (also onnx/export.py:42-45)."""
    if mode == "keep":
        return x
    if mode == "down":
        return x[:, :, ::2, ::2]
    if mode == "up":
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    raise ValueError(mode)


def mp_silu(x):
    return F.silu(x) / 0.596  # mp_layers.py:33-34


def mp_sum(args, w):
    """mp_layers.py:47-62; weights are materialised in the activation dtype (fp32 here)."""
    if w is None:
        wt = torch.full((len(args),), 1 / len(args), dtype=args[0].dtype, device=args[0].device)
    elif isinstance(w, float):
        wt = torch.tensor([1 - w, w], dtype=args[0].dtype, device=args[0].device)
    else:
        wt = torch.tensor(w, dtype=args[0].dtype, device=args[0].device)
    acc = sum(a * wi for a, wi in zip(args, wt))
    return acc / torch.linalg.vector_norm(wt)


def mp_concat(args, w, dim=1):
    """mp_layers.py:65-86."""
    if isinstance(w, float):
        wt = torch.tensor([1 - w, w], dtype=args[0].dtype, device=args[0].device)
    else:
        wt = torch.tensor(w, dtype=args[0].dtype, device=args[0].device)
    n_tot = torch.tensor(sum(a.shape[dim] for a in args), dtype=args[0].dtype, device=args[0].device)
    c = torch.sqrt(n_tot / torch.sum(torch.square(wt)))
    return torch.cat([a * (c / np.sqrt(a.shape[dim]) * wt[i]) for i, a in enumerate(args)], dim=dim)


def mp_concat_scales(n_a: int, n_b: int, t: float) -> tuple[float, float]:
    """The two per-tensor scalars mp_concat applies (closed form of mp_layers.py:83-86), in fp64."""
    c = math.sqrt((n_a + n_b) / ((1 - t) ** 2 + t ** 2))
    return c / math.sqrt(n_a) * (1 - t), c / math.sqrt(n_b) * t


def positional_embedding(t: torch.Tensor, num_channels: int) -> torch.Tensor:
    """MPPositionalEmbedding, mp_layers.py:88-107."""
    half = num_channels // 2
    freqs = torch.exp(torch.arange(half, device=t.device) * -(math.log(10) / (half - 1)))
    y = t.to(torch.float32).outer(freqs.to(torch.float32))
    return (torch.cat([torch.sin(y), torch.cos(y)], dim=1) * np.sqrt(2)).to(t.dtype)


def fourier_embedding(x: torch.Tensor, freqs: torch.Tensor, phases: torch.Tensor) -> torch.Tensor:
    """MPFourier.forward, mp_layers.py:115-131."""
    y = x.to(torch.float32).outer(freqs.to(torch.float32)) + phases.to(torch.float32)
    return (y.cos() * np.sqrt(2)).to(x.dtype)


def effective_weight(w: torch.Tensor, gain=1.0) -> torch.Tensor:
    """The weight MPConv actually convolves with, mp_layers.py:203-213 (global RMS normalisation, then
    gain / sqrt(fan_in)); fp32."""
    w = normalize(w.to(torch.float32))
    return w * (gain / np.sqrt(w[0].numel()))


def mp_conv(x: torch.Tensor, w: torch.Tensor, gain=1.0) -> torch.Tensor:
    """MPConv.forward, mp_layers.py:201-221 (groups=1, padding k//2)."""
    we = effective_weight(w, gain).to(x.dtype)
    if we.ndim == 2:
        return F.linear(x, we)
    return F.conv2d(x, we, padding=we.shape[-1] // 2)


# ---------------------------------------------------------------------------------------------- model plan
def block_plan(cfg: dict) -> tuple[list, list]:
    """Module order/shape plan of EDMUnet2D.__init__, edm_unet.py:105-137.  Returns (enc, dec) lists of dicts."""
    mults = cfg.get("model_channel_mults") or [1, 2, 3, 4]
    mc = cfg.get("model_channels", 128)
    lpb = cfg.get("layers_per_block", 2)
    if isinstance(lpb, int):
        lpb = [lpb] * len(mults)
    attn_res = cfg.get("attn_resolutions") or []
    image_size = cfg["image_size"]
    block_channels = [mc * m for m in mults]
    enc, dec = [], []
    cout = cfg["in_channels"] + 1
    for level, (channels, nb) in enumerate(zip(block_channels, lpb)):
        res = image_size // 2 ** level
        if level == 0:
            cin, cout = cout, channels
            enc.append(dict(name=f"{res}x{res}_conv", kind="conv", cin=cin, cout=cout))
        else:
            enc.append(dict(name=f"{res}x{res}_down", kind="block", mode="enc", resample="down", cin=cout, cout=cout,
                            attention=False))
        for idx in range(nb):
            cin, cout = cout, channels
            enc.append(dict(name=f"{res}x{res}_block{idx}", kind="block", mode="enc", resample="keep", cin=cin,
                            cout=cout, attention=(res in attn_res)))
    skips = [b["cout"] for b in enc]
    if not cfg.get("encode_only", False):
        for level, (channels, nb) in reversed(list(enumerate(zip(block_channels, lpb)))):
            res = image_size // 2 ** level
            if level == len(block_channels) - 1:
                dec.append(dict(name=f"{res}x{res}_in0", kind="block", mode="dec", resample="keep", cin=cout,
                                cout=cout, attention=bool(cfg.get("midblock_attention", True)), concat=False))
                dec.append(dict(name=f"{res}x{res}_in1", kind="block", mode="dec", resample="keep", cin=cout,
                                cout=cout, attention=False, concat=False))
            else:
                dec.append(dict(name=f"{res}x{res}_up", kind="block", mode="dec", resample="up", cin=cout, cout=cout,
                                attention=False, concat=False))
            for idx in range(nb + 1):
                skip = skips.pop()
                cin, cout_new = cout + skip, channels
                dec.append(dict(name=f"{res}x{res}_block{idx}", kind="block", mode="dec", resample="keep", cin=cin,
                                cout=cout_new, attention=(res in attn_res), concat=True, skip_channels=skip))
                cout = cout_new
    return enc, dec


def _attn(x, sd, prefix, num_heads):
    """UNetBlock.attn, unet_block.py:102-108 (cosine attention: per-head pixel-norm of q, k, v)."""
    y = mp_conv(x, sd[prefix + "attn_qkv.weight"])
    y = y.reshape(y.shape[0], num_heads, -1, 3, y.shape[2] * y.shape[3])
    q, k, v = normalize(y, dim=2).unbind(3)
    w = torch.einsum("nhcq,nhck->nhqk", q, k / torch.sqrt(torch.tensor(q.shape[2], dtype=q.dtype, device=q.device))).softmax(dim=3)
    y = torch.einsum("nhqk,nhck->nhcq", w, v)
    return mp_conv(y.reshape(*x.shape), sd[prefix + "attn_proj.weight"])


def unet_block(x, emb, sd, prefix, b, res_balance=0.3, attn_balance=0.3, clip_act=256.0, channels_per_head=64):
    """UNetBlock.forward, unet_block.py:116-156 (conv_type 'default', resample_type 'pooling', silu)."""
    x = resample(x, b["resample"])
    has_skip = (prefix + "conv_skip.weight") in sd
    if b["mode"] == "enc":
        if has_skip:
            x = mp_conv(x, sd[prefix + "conv_skip.weight"])
        x = normalize(x, dim=1)
    y = mp_conv(mp_silu(x), sd[prefix + "conv_res0.weight"])
    if (prefix + "emb_linear.weight") in sd:
        c = mp_conv(emb, sd[prefix + "emb_linear.weight"], gain=sd[prefix + "emb_gain"]) + 1
        c = c / torch.sqrt(torch.mean(c ** 2, dim=1, keepdim=True) + 1e-8)
        y = mp_silu(y * c.unsqueeze(2).unsqueeze(3).to(y.dtype))
    else:
        y = mp_silu(y)
    y = mp_conv(y, sd[prefix + "conv_res1.weight"])
    if b["mode"] == "dec" and has_skip:
        x = mp_conv(x, sd[prefix + "conv_skip.weight"])
    x = mp_sum([x, y], res_balance)
    if b.get("attention"):
        heads = b["cout"] // channels_per_head
        if heads:
            x = mp_sum([x, _attn(x, sd, prefix, heads)], attn_balance)
    if clip_act is not None:
        x = torch.clip(x, -clip_act, clip_act)
    return x


def compute_embeddings(sd, cfg, noise_labels, conditional_inputs):
    """EDMUnet2D.compute_embeddings, edm_unet.py:145-159."""
    mc = cfg.get("model_channels", 128)
    noise_dims = mc if cfg.get("noise_emb_dims") is None else cfg["noise_emb_dims"]
    embeds, weights = [], []
    if noise_dims > 0:
        if cfg.get("fourier_scale", 1) == "pos":
            pe = positional_embedding(noise_labels, noise_dims)
        else:
            pe = fourier_embedding(noise_labels, sd["noise_fourier.freqs"], sd["noise_fourier.phases"])
        embeds.append(mp_conv(pe, sd["noise_linear.weight"]))
        weights.append(1)
    for i, ((kind, _x, weight), inp) in enumerate(zip(cfg.get("conditional_inputs") or [], conditional_inputs or [])):
        if kind == "float":
            fe = fourier_embedding(inp, sd[f"conditional_layers.{i}.0.freqs"], sd[f"conditional_layers.{i}.0.phases"])
            embeds.append(mp_conv(fe, sd[f"conditional_layers.{i}.1.weight"]))
        elif kind == "tensor":
            embeds.append(mp_silu(mp_conv(inp, sd[f"conditional_layers.{i}.weight"])))
        elif kind == "embedding":
            # MPEmbedding.forward (mp_layers.py:233-245) looks up the UN-normalised table.
            embeds.append(F.embedding(inp, sd[f"conditional_layers.{i}.weight"]))
        weights.append(weight)
    if not embeds:
        return None
    return mp_silu(mp_sum(embeds, weights))


@torch.no_grad()
def unet_forward(sd: dict, cfg: dict, x: torch.Tensor, noise_labels: torch.Tensor, conditional_inputs=None,
                 trace: dict | None = None):
    """EDMUnet2D.forward, edm_unet.py:161-184 (return_logvar=False)."""
    enc, dec = block_plan(cfg)
    bk = cfg.get("block_kwargs") or {}
    kw = dict(res_balance=bk.get("res_balance", 0.3), attn_balance=bk.get("attn_balance", 0.3),
              clip_act=bk.get("clip_act", 256.0), channels_per_head=bk.get("channels_per_head", 64))
    emb = compute_embeddings(sd, cfg, noise_labels, conditional_inputs)
    x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
    skips = []
    for b in enc:
        if b["kind"] == "conv":
            x = mp_conv(x, sd[f"enc.{b['name']}.weight"])
        else:
            x = unet_block(x, emb, sd, f"enc.{b['name']}.", b, **kw)
        skips.append(x)
        if trace is not None:
            trace[f"enc.{b['name']}."] = x
    for b in dec:
        if b.get("concat"):
            x = mp_concat([x, skips.pop()], float(cfg.get("concat_balance", 0.3)))
        x = unet_block(x, emb, sd, f"dec.{b['name']}.", b, **kw)
        if trace is not None:
            trace[f"dec.{b['name']}."] = x
    out_gain = sd["out_gain"] if "out_gain" in sd else 1.0
    return mp_conv(x, sd["out_conv.weight"], gain=out_gain)


# ---------------------------------------------------------------------------------------------- weights
DECODER_CFG = dict(image_size=512, in_channels=5, out_channels=1, model_channels=64, model_channel_mults=[1, 2, 3, 4],
                   layers_per_block=3, emb_channels=None, noise_emb_dims=None, attn_resolutions=[],
                   midblock_attention=False, concat_balance=0.5, conditional_inputs=[], fourier_scale="pos")
"""configs/diffusion_decoder/diffusion_decoder_64-3.cfg:51-65 ([model] section)."""


def state_shapes(cfg: dict) -> dict:
    """Parameter/buffer names and shapes EDMUnet2D would register (edm_unet.py:67-143), without building modules."""
    mc = cfg.get("model_channels", 128)
    mults = cfg.get("model_channel_mults") or [1, 2, 3, 4]
    emb_ch = cfg.get("emb_channels") or mc * max(mults)
    noise_dims = mc if cfg.get("noise_emb_dims") is None else cfg["noise_emb_dims"]
    out_ch = cfg.get("out_channels") or cfg["in_channels"]
    shapes: dict = {}
    if not cfg.get("disable_out_gain", False):
        shapes["out_gain"] = ()
    if noise_dims > 0:
        if cfg.get("fourier_scale", 1) == "pos":
            shapes["noise_fourier.freqs"] = (noise_dims // 2,)
        else:
            shapes["noise_fourier.freqs"] = (noise_dims,)
            shapes["noise_fourier.phases"] = (noise_dims,)
        shapes["noise_linear.weight"] = (emb_ch, noise_dims)
    for i, (kind, xdim, _w) in enumerate(cfg.get("conditional_inputs") or []):
        if kind == "float":
            shapes[f"conditional_layers.{i}.0.freqs"] = (xdim,)
            shapes[f"conditional_layers.{i}.0.phases"] = (xdim,)
            shapes[f"conditional_layers.{i}.1.weight"] = (emb_ch, xdim)
        elif kind == "tensor":
            shapes[f"conditional_layers.{i}.weight"] = (emb_ch, xdim)
        elif kind == "embedding":
            shapes[f"conditional_layers.{i}.weight"] = (xdim, emb_ch)
    enc, dec = block_plan(cfg)
    cph = (cfg.get("block_kwargs") or {}).get("channels_per_head", 64)
    for side, blocks in (("enc", enc), ("dec", dec)):
        for b in blocks:
            p = f"{side}.{b['name']}."
            if b["kind"] == "conv":
                shapes[p + "weight"] = (b["cout"], b["cin"], 3, 3)
                continue
            shapes[p + "emb_gain"] = ()
            c0_in = b["cout"] if b["mode"] == "enc" else b["cin"]
            shapes[p + "conv_res0.weight"] = (b["cout"], c0_in, 3, 3)
            if emb_ch > 0:
                shapes[p + "emb_linear.weight"] = (b["cout"], emb_ch)
            shapes[p + "conv_res1.weight"] = (b["cout"], b["cout"], 3, 3)
            if b["cin"] != b["cout"]:
                shapes[p + "conv_skip.weight"] = (b["cout"], b["cin"], 1, 1)
            if b.get("attention") and b["cout"] // cph:
                shapes[p + "attn_qkv.weight"] = (b["cout"] * 3, b["cout"], 1, 1)
                shapes[p + "attn_proj.weight"] = (b["cout"], b["cout"], 1, 1)
    shapes["out_conv.weight"] = (out_ch, dec[-1]["cout"] if dec else enc[-1]["cout"], 3, 3)
    lv = cfg.get("logvar_channels", 128)
    shapes["logvar_fourier.freqs"] = (lv,)
    shapes["logvar_fourier.phases"] = (lv,)
    shapes["logvar_linear.weight"] = (cfg.get("n_logvar", 1), lv)
    return shapes


def procedural_state_dict(cfg: dict, seed: int = 0, emb_gain: float = 0.5, out_gain: float = 1.0) -> dict:
    """Deterministic, construction-order-independent synthetic weights: every tensor is drawn from a generator seeded
    by crc32(name) ^ seed.  emb_gain/out_gain are zero-initialised in the reference (unet_block.py:72,
    edm_unet.py:103), which would make every parity test vacuous, hence the non-zero values."""
    import zlib
    sd = {}
    for name, shape in state_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
        if name.endswith("emb_gain"):
            sd[name] = torch.tensor(emb_gain)
        elif name == "out_gain":
            sd[name] = torch.tensor(out_gain)
        elif name == "noise_fourier.freqs" and cfg.get("fourier_scale", 1) == "pos":
            half = shape[0]
            sd[name] = torch.exp(torch.arange(half) * -(math.log(10) / (half - 1)))
        elif name.endswith(".freqs"):
            sd[name] = 2 * np.pi * torch.randn(shape, generator=g)
        elif name.endswith(".phases"):
            sd[name] = 2 * np.pi * torch.rand(shape, generator=g)
        else:
            sd[name] = torch.randn(shape, generator=g)
    return sd
