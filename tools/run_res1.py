"""Runs the encoder res1-style igemm (64->64 @256^2, residual pixel-norm, clip, 3 outputs) a few times: ncu target."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from terrain_diffusion_b200 import _lib as L
from terrain_diffusion_b200.layout import pack_weight_segments, to_nc8hw8

dev = torch.device("cuda:0")
lib = L.lib()
cout, res = 64, 256
segs = [(64, 9)]
acts = [to_nc8hw8(torch.randn(1, c, res, res, device=dev)) for c, _ in segs]
wts = [torch.randn(cout, c, 3, 3, device=dev) * 0.02 for c, t in segs]
n_item = L.igemm_choose_n(cout, 1, res, res, segs)
b = pack_weight_segments(wts, n_item)
out = torch.empty(1, cout // 8, res, res, 8, dtype=torch.bfloat16, device=dev)
resid = to_nc8hw8(torch.randn(1, cout, res, res, device=dev))
out2 = torch.empty_like(out); out3 = torch.empty_like(out)
d = L.TdxIgemmDesc()
d.a_ptr[0] = acts[0].data_ptr(); d.a_channels[0] = 64; d.a_taps[0] = 9
d.n_seg = 1; d.b_packed = b.data_ptr(); d.c_out = cout; d.n_per_item = n_item
d.n_img, d.height, d.width = 1, res, res
d.out[0].ptr = out.data_ptr(); d.out[0].kind = L.OUT_RAW; d.out[0].scale = 1.0
if len(sys.argv) > 1 and sys.argv[1] == "res0":       # the res0-type launch: embedding scale + mp_silu, one raw output
    cvec = torch.ones(1, cout, device=dev)
    d.epi_flags = L.EPI_EMB_SILU; d.cvec = cvec.data_ptr()
else:
    d.epi_flags = L.EPI_RESID; d.resid = resid.data_ptr(); d.resid_pnorm = 1; d.resid_scale = 0.9; d.clip = 256.0
    d.out[1].ptr = out2.data_ptr(); d.out[1].kind = L.OUT_PNORM_SILU; d.out[1].scale = 1.0
    d.out[2].ptr = out3.data_ptr(); d.out[2].kind = L.OUT_SILU; d.out[2].scale = 0.8
for _ in range(4):
    L.check(lib.tdx_igemm_run(C.byref(d), L.current_stream_ptr()))
torch.cuda.synchronize()
print("done")
