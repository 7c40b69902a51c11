"""T2/T3: every tdx_igemm_run configuration against a plain PyTorch fp32 reference of the same op (tests/_igemm_ref.py)."""
import pytest
import torch

from tests._igemm_ref import default_cases, make_inputs, reference, rel_rms, run_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", default_cases(), ids=lambda c: c.name)
def test_igemm_case(case):
    dev = torch.device("cuda:0")
    acts, wts, cvec, resid = make_inputs(case, dev)
    refs = reference(case, acts, wts, cvec, resid)
    gots = run_cuda(case, acts, wts, cvec, resid)
    for g, r in zip(gots, refs):
        assert not torch.isnan(g).any()
        # inputs are bf16-exact, accumulation is fp32: the only error is the bf16 rounding of the stored output
        assert rel_rms(g, r.bfloat16().float()) < 2e-3
        assert float((g - r).abs().max()) <= 2 ** -7 * float(r.abs().max()) + 1e-3


def test_pixel_norm_side_channel_rms_out_and_resid_inv():
    """Round 2: a producer leaves 1 / (eps + rms) per pixel (`rms_out`); the consumer that adds pixelnorm(residual)
    reads that one float (`resid_inv`) instead of every channel of the residual.  Both against the fp32 definition,
    on same-size, half-size and double-size residuals and a ragged image."""
    from terrain_diffusion_b200 import _lib as L
    from tests._igemm_ref import Case
    dev = torch.device("cuda:0")
    R = L.EPI_RESID
    outs = [(L.OUT_RAW, L.SP_SAME, 1.0), (L.OUT_PNORM_SILU, L.SP_SAME, 1.0)]
    for case in (Case("inv_same", [(64, 9)], 64, 2, 40, 24, epi=R, resid_pnorm=1, outs=outs),
                 Case("inv_up", [(128, 9)], 128, 1, 32, 32, epi=R, resid_pnorm=1, resid_spatial=L.SP_UP2, outs=outs),
                 Case("inv_down", [(64, 9)], 64, 2, 24, 24, epi=R, resid_pnorm=1, resid_spatial=L.SP_DOWN2, outs=outs)):
        acts, wts, cvec, resid = make_inputs(case, dev)
        refs = reference(case, acts, wts, cvec, resid)
        inv_plane = (1.0 / (1e-4 + resid.square().mean(dim=1).sqrt())).contiguous()      # [n, hr, wr] fp32
        rms = torch.full((case.n, case.h, case.w), float("nan"), device=dev)
        gots = run_cuda(case, acts, wts, cvec, resid, rms_out=rms, resid_inv=inv_plane)
        for g_, r in zip(gots, refs):
            assert rel_rms(g_, r.bfloat16().float()) < 2e-3
        want = 1.0 / (1e-4 + refs[0].square().mean(dim=1).sqrt())
        assert not torch.isnan(rms).any()
        assert float(((rms - want) / want).abs().max()) < 2e-3
        plain = run_cuda(case, acts, wts, cvec, resid)                                    # recomputed pixel-norm path
        for a, b in zip(plain, gots):
            assert rel_rms(a, b) < 2e-3


def test_igemm_rejects_bad_descriptors():
    import ctypes as C
    from terrain_diffusion_b200 import _lib as L
    d = L.TdxIgemmDesc()
    assert L.lib().tdx_igemm_run(C.byref(d), None) == -1
    assert b"n_seg" in L.lib().tdx_last_error()


@pytest.mark.parametrize("chans,dtypes,hw", [((1, 4), ("f32", "bf16"), (40, 24)), ((5, 0), ("f32", "f32"), (16, 8)),
                                             ((11, 0), ("f32", "f32"), (24, 16))])
def test_im2col_of_the_first_convolution_is_exact(chans, dtypes, hw):
    """tdx_im2col_run = cat([x0*s0, x1*s1, ones]) -> 3x3 neighbourhoods, zero padded (edm_unet.py:168-172 feeds a padded
    MPConv): channel k = tap*ci + c; exact up to the single bf16 rounding of each value."""
    import ctypes as C

    import torch.nn.functional as F

    from terrain_diffusion_b200 import _lib as L
    from terrain_diffusion_b200.layout import from_nc8hw8
    dev = torch.device("cuda:0")
    h, w = hw
    n = 2
    g = torch.Generator().manual_seed(5)
    srcs, scales = [], []
    for c, dt in zip(chans, dtypes):
        t = torch.randn(n, max(c, 1), h, w, generator=g)
        srcs.append(t.to(dev).to(torch.bfloat16 if dt == "bf16" else torch.float32).contiguous())
        scales.append(torch.tensor([0.37 + len(scales)], device=dev))
    ci = chans[0] + chans[1] + 1
    kpad = ((9 * ci + 63) // 64) * 64
    out = torch.full((n, kpad // 8, h, w, 8), float("nan"), dtype=torch.bfloat16, device=dev)
    d = L.TdxIm2colDesc()
    for i in range(2):
        d.src[i] = srcs[i].data_ptr() if chans[i] else None
        d.src_channels[i] = chans[i]
        d.src_dtype[i] = 1 if dtypes[i] == "bf16" else 0
        d.src_scale[i] = scales[i].data_ptr() if (chans[i] and i == 0) else None
    d.out = out.data_ptr()
    d.k_pad = kpad
    d.n_img, d.height, d.width = n, h, w
    L.check(L.lib().tdx_im2col_run(C.byref(d), L.current_stream_ptr()))
    torch.cuda.synchronize()
    parts = [srcs[0].float() * scales[0]]
    if chans[1]:
        parts.append(srcs[1].float())
    parts.append(torch.ones(n, 1, h, w, device=dev))
    x = torch.cat(parts, dim=1)                                   # [n, ci, h, w]
    cols = F.unfold(x, kernel_size=3, padding=1).view(n, ci, 9, h, w)   # unfold orders (c, tap)
    ref = cols.permute(0, 2, 1, 3, 4).reshape(n, 9 * ci, h, w).bfloat16().float()
    got = from_nc8hw8(out)
    assert torch.equal(got[:, :9 * ci], ref)
    assert torch.count_nonzero(got[:, 9 * ci:]) == 0
    d.k_pad = kpad + 64
    assert L.lib().tdx_im2col_run(C.byref(d), L.current_stream_ptr()) != 0   # wrong k_pad is rejected
