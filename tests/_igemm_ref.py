"""Plain-PyTorch fp32 reference of ONE tdx_igemm_run launch (conv + fused epilogue), and a runner for the CUDA op.

Used by tests/test_igemm_gpu.py and tools/bringup_igemm.py.  Inputs are rounded to bf16 first so that the only
difference to the kernel is accumulation order and the bf16 rounding of the stored outputs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from terrain_diffusion_b200 import _lib as L
from terrain_diffusion_b200.layout import from_nc8hw8, pack_weight_segments, to_nc8hw8


def mp_silu(x):
    return F.silu(x) / 0.596


def pixelnorm(x):
    return x / (1e-4 + x.square().mean(dim=1, keepdim=True).sqrt())


@dataclass
class Case:
    name: str
    segs: list  # [(channels, taps)]
    cout: int
    n: int
    h: int
    w: int
    epi: int = 0
    resid_spatial: int = L.SP_SAME
    resid_pnorm: int = 0
    resid_scale: float = 0.7
    clip: float = 0.0
    outs: list = field(default_factory=lambda: [(L.OUT_RAW, L.SP_SAME, 1.0)])
    wscale: float = 1.0
    n_item: int = 0   # 0 = let the library choose (tdx_igemm_choose_n)
    seed: int = 0


def make_inputs(case: Case, device):
    g = torch.Generator(device="cpu").manual_seed(case.seed)
    acts, wts = [], []
    ktot = sum(c * t for c, t in case.segs)
    for (c, t) in case.segs:
        a = torch.randn(case.n, c, case.h, case.w, generator=g)
        k = 3 if t == 9 else 1
        w = torch.randn(case.cout, c, k, k, generator=g) * (case.wscale / ktot ** 0.5)
        acts.append(a.bfloat16().float().to(device))
        wts.append(w.bfloat16().float().to(device))
    cvec = (1.0 + 0.3 * torch.randn(case.n, case.cout, generator=g)).to(device)
    if case.resid_spatial == L.SP_UP2:
        rs = (case.h // 2, case.w // 2)
    elif case.resid_spatial == L.SP_DOWN2:
        rs = (case.h * 2, case.w * 2)
    else:
        rs = (case.h, case.w)
    resid = torch.randn(case.n, case.cout, *rs, generator=g).bfloat16().float().to(device)
    return acts, wts, cvec, resid


def reference(case: Case, acts, wts, cvec, resid):
    acc = None
    for a, w in zip(acts, wts):
        y = F.conv2d(a.double(), w.double(), padding=w.shape[-1] // 2)
        acc = y if acc is None else acc + y
    v = acc.float()
    if case.epi & L.EPI_EMB_SILU:
        v = mp_silu(v * cvec[:, :, None, None])
    if case.epi & L.EPI_RESID:
        r = resid
        if case.resid_spatial == L.SP_UP2:
            r = r.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        elif case.resid_spatial == L.SP_DOWN2:
            r = r[:, :, ::2, ::2]
        if case.resid_pnorm:
            r = pixelnorm(r)
        v = v + case.resid_scale * r
    if case.clip > 0:
        v = torch.clamp(v, -case.clip, case.clip)
    if case.epi & L.EPI_PNORM:
        v = pixelnorm(v)
    outs = []
    for kind, spatial, scale in case.outs:
        if kind == L.OUT_RAW:
            o = v
        elif kind == L.OUT_SILU:
            o = mp_silu(scale * v)
        else:
            o = mp_silu(v if (case.epi & L.EPI_PNORM) else pixelnorm(v))
        if spatial == L.SP_DOWN2:
            o = o[:, :, ::2, ::2]
        elif spatial == L.SP_UP2:
            o = o.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        outs.append(o)
    return outs


def run_cuda(case: Case, acts, wts, cvec, resid, rms_out=None, resid_inv=None):
    dev = acts[0].device
    a_dev = [to_nc8hw8(a) for a in acts]
    n_item = case.n_item or L.igemm_choose_n(case.cout, case.n, case.h, case.w, case.segs)
    b = pack_weight_segments(wts, n_item).to(dev)
    r_dev = to_nc8hw8(resid)
    d = L.TdxIgemmDesc()
    for i, (c, t) in enumerate(case.segs):
        d.a_ptr[i] = a_dev[i].data_ptr()
        d.a_channels[i] = c
        d.a_taps[i] = t
    d.n_seg = len(case.segs)
    d.b_packed = b.data_ptr()
    d.c_out = case.cout
    d.n_per_item = n_item
    d.n_img, d.height, d.width = case.n, case.h, case.w
    d.epi_flags = case.epi
    cvec_dev = cvec.contiguous()
    d.cvec = cvec_dev.data_ptr()
    d.resid = r_dev.data_ptr()
    d.resid_spatial = case.resid_spatial
    d.resid_pnorm = case.resid_pnorm if resid_inv is None else 0
    if rms_out is not None:        # fp32 [n, h, w]: 1 / (eps + rms) of the result, for the consumer's residual
        d.rms_out = rms_out.data_ptr()
    if resid_inv is not None:      # fp32 plane at the residual's resolution: replaces the recomputed pixel-norm
        d.resid_inv = resid_inv.data_ptr()
    d.resid_scale = case.resid_scale
    d.clip = case.clip
    bufs = []
    for i, (kind, spatial, scale) in enumerate(case.outs):
        hh, ww = case.h, case.w
        if spatial == L.SP_DOWN2:
            hh, ww = hh // 2, ww // 2
        elif spatial == L.SP_UP2:
            hh, ww = hh * 2, ww * 2
        o = torch.full((case.n, case.cout // 8, hh, ww, 8), float("nan"), dtype=torch.bfloat16, device=dev)
        bufs.append(o)
        d.out[i].ptr = o.data_ptr()
        d.out[i].kind = kind
        d.out[i].spatial = spatial
        d.out[i].scale = scale
    L.check(L.lib().tdx_igemm_run(C.byref(d), L.current_stream_ptr()))
    torch.cuda.synchronize()
    return [from_nc8hw8(o) for o in bufs]


def rel_rms(a, b):
    return float((a - b).square().mean().sqrt() / (b.square().mean().sqrt() + 1e-30))


def default_cases() -> list[Case]:
    E, R, P = L.EPI_EMB_SILU, L.EPI_RESID, L.EPI_PNORM
    return [
        Case("c64_3x3_32x32", [(64, 9)], 64, 1, 32, 32),
        Case("c64_3x3_16x8_single_tile", [(64, 9)], 64, 1, 16, 8),
        Case("c64_3x3_8x8_partial_tile", [(64, 9)], 64, 2, 8, 8),
        Case("c64_1x1_32x32", [(64, 1)], 128, 1, 32, 32),
        Case("c128_3x3_64x64_n2", [(128, 9)], 128, 2, 64, 64),
        Case("c192_3x3_32x32", [(192, 9)], 192, 1, 32, 32),
        Case("c256_3x3_32x32", [(256, 9)], 256, 1, 32, 32),
        Case("c64_3x3_256x256", [(64, 9)], 64, 1, 256, 256),
        Case("concat_256+192_to_192_64x64", [(256, 9), (192, 9)], 192, 1, 64, 64),
        Case("res1+skip_3seg", [(64, 9), (128, 1), (64, 1)], 64, 1, 64, 64),
        Case("emb_silu", [(64, 9)], 64, 2, 32, 32, epi=E),
        Case("resid_same_pnorm_3outs", [(128, 9)], 128, 1, 32, 32, epi=R, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_SAME, 0.83)]),
        Case("resid_down_pnorm_down2", [(64, 9)], 64, 1, 32, 32, epi=R, resid_spatial=L.SP_DOWN2, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_DOWN2, 1.0)]),
        Case("resid_up_up2", [(128, 9)], 128, 1, 32, 32, epi=R, resid_spatial=L.SP_UP2,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_UP2, 1.2)]),
        Case("pnorm_1x1", [(64, 1)], 128, 1, 32, 32, epi=P,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_SAME, 1.0)]),
        Case("cluster4_pnorm_persistent", [(256, 9)], 256, 2, 64, 64, epi=R, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_SAME, 0.7)]),
        Case("cluster3_pnorm_down2", [(192, 9)], 192, 1, 32, 32, epi=R, resid_spatial=L.SP_DOWN2, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_DOWN2, 1.0)]),
        Case("cluster4_k1_pnorm", [(192, 1)], 256, 1, 32, 32, epi=P,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_SAME, 1.0)]),
        Case("n128_forced_resid_pnorm", [(128, 9)], 256, 1, 32, 32, epi=R, resid_pnorm=1, n_item=128,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_SAME, 1.0)]),
        Case("n256_forced_emb", [(64, 9)], 256, 1, 32, 32, epi=E, n_item=256),
        Case("n192_forced_pnorm_2pass", [(64, 9)], 192, 1, 32, 32, epi=R, resid_pnorm=1, n_item=192,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_DOWN2, 1.0), (L.OUT_SILU, L.SP_SAME, 0.9)]),
        Case("n64_forced_c256", [(256, 9)], 256, 1, 32, 32, epi=R, n_item=64),
        Case("resident_multi_item_per_cta", [(64, 9)], 64, 3, 128, 128, epi=E),
        Case("clip_active", [(64, 9)], 64, 1, 16, 16, epi=R, clip=0.5),
        Case("clip_no_resid", [(64, 9), (64, 1)], 64, 1, 16, 16, clip=0.3),
        # ragged images: H, W multiples of 8 but not of the 16 x 8 work-item tile, several images, every spatial mode
        Case("ragged_24x40_emb", [(64, 9)], 64, 3, 24, 40, epi=E),
        Case("ragged_40x24_resid_3outs", [(128, 9)], 128, 2, 40, 24, epi=R, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_DOWN2, 1.0), (L.OUT_SILU, L.SP_UP2, 0.8)]),
        Case("ragged_56x72_resid_up", [(64, 9)], 64, 2, 56, 72, epi=R, resid_spatial=L.SP_UP2,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_SILU, L.SP_SAME, 1.1)]),
        Case("ragged_24x24_resid_down", [(64, 9)], 64, 2, 24, 24, epi=R, resid_spatial=L.SP_DOWN2, resid_pnorm=1,
             outs=[(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_SAME, 1.0)]),
    ]
