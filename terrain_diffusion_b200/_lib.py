"""ctypes binding of libtdx.so (include/tdx.h).  Loading fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtdx.so"


class TdxError(RuntimeError):
    pass


class TdxOutSpec(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("kind", C.c_int32), ("spatial", C.c_int32), ("scale", C.c_float),
                ("_pad", C.c_int32)]


class TdxIgemmDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p * 3), ("a_channels", C.c_int32 * 3), ("a_taps", C.c_int32 * 3), ("n_seg", C.c_int32),
        ("b_packed", C.c_void_p), ("c_out", C.c_int32), ("n_img", C.c_int32), ("height", C.c_int32),
        ("width", C.c_int32), ("epi_flags", C.c_int32), ("cvec", C.c_void_p), ("resid", C.c_void_p),
        ("resid_spatial", C.c_int32), ("resid_pnorm", C.c_int32), ("resid_scale", C.c_float), ("clip", C.c_float),
        ("out", TdxOutSpec * 3),
    ]


OUT_NONE, OUT_RAW, OUT_SILU, OUT_PNORM_SILU = 0, 1, 2, 3
SP_SAME, SP_DOWN2, SP_UP2 = 0, 1, 2
EPI_EMB_SILU, EPI_RESID, EPI_PNORM = 1, 2, 4

_lib = None


def lib() -> C.CDLL:
    """Return the loaded library; raises TdxError if it has not been built (python -m terrain_diffusion_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TdxError(f"{LIB_PATH} not found: build it with `python -m terrain_diffusion_b200.build` "
                           "(nvcc, sm_100a). There is no fallback path.")
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    l.tdx_last_error.restype = C.c_char_p
    l.tdx_last_error.argtypes = []
    l.tdx_device_info.restype = C.c_int
    l.tdx_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.tdx_igemm_packed_weight_elems.restype = C.c_int64
    l.tdx_igemm_packed_weight_elems.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32]
    l.tdx_igemm_run.restype = C.c_int
    l.tdx_igemm_run.argtypes = [C.POINTER(TdxIgemmDesc), C.c_void_p]
    for name, (res, args) in _OPTIONAL.items():
        if hasattr(l, name):
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args


_OPTIONAL: dict = {}


def check(rc: int) -> None:
    if rc != 0:
        raise TdxError(f"libtdx error {rc}: {lib().tdx_last_error().decode()}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
