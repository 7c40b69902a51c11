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


def test_igemm_rejects_bad_descriptors():
    import ctypes as C
    from terrain_diffusion_b200 import _lib as L
    d = L.TdxIgemmDesc()
    assert L.lib().tdx_igemm_run(C.byref(d), None) == -1
    assert b"n_seg" in L.lib().tdx_last_error()
