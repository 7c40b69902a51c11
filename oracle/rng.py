"""ORACLE (test infrastructure, not product): Python face of oracle/rng.c plus a pure-Python restatement for
small cases.  Reference: terrain_diffusion/inference/portable_rng.py:24-82, world_pipeline.py:58-115."""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_rng.so"
_lib = None

MASK64 = 0xFFFFFFFFFFFFFFFF
PCG_MULT = 6364136223846793005
PCG_INC = 1442695040888963407


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
        _lib.oracle_next_seed.restype = C.c_uint64
        _lib.oracle_next_seed.argtypes = [C.c_uint64]
        _lib.oracle_tile_seed.restype = C.c_uint64
        _lib.oracle_tile_seed.argtypes = [C.c_uint64, C.c_int64, C.c_int64]
        _lib.oracle_fill_standard_normal.restype = None
        _lib.oracle_fill_standard_normal.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_int]
        _lib.oracle_gaussian_noise_patch.restype = C.c_int
        _lib.oracle_gaussian_noise_patch.argtypes = [C.c_uint64] + [C.c_int64] * 7 + [C.c_void_p]
    return _lib


def next_seed(seed: int) -> int:
    return int(lib().oracle_next_seed(int(seed) & MASK64))


def tile_seed(base_seed: int, ty: int, tx: int) -> int:
    return int(lib().oracle_tile_seed(int(base_seed) & MASK64, int(ty), int(tx)))


def standard_normal(seed: int, size, dtype=np.float32) -> np.ndarray:
    out = np.empty(size, dtype=dtype)
    if out.size:
        lib().oracle_fill_standard_normal(int(seed) & MASK64, out.ctypes.data, out.size, 1 if dtype == np.float32 else 0)
    return out


def gaussian_noise_patch(base_seed, y0, x0, h, w, channels=1, tile_h=256, tile_w=256) -> np.ndarray:
    out = np.empty((channels, h, w), dtype=np.float32)
    rc = lib().oracle_gaussian_noise_patch(int(base_seed) & MASK64, y0, x0, h, w, channels, tile_h, tile_w,
                                           out.ctypes.data)
    assert rc == 0
    return out


# ---- pure-Python restatement (small n only): cross-checks the C code ----------------------------------------------
def py_pcg_next(state: int) -> tuple[int, int]:
    state = (state * PCG_MULT + PCG_INC) & MASK64
    x = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
    rot = state >> 59
    return state, ((x >> rot) | (x << ((32 - rot) & 31))) & 0xFFFFFFFF


def py_tile_seed(base_seed: int, ty: int, tx: int) -> int:
    h = (int(base_seed) & MASK64) * 0x9E3779B9
    h = (h + (int(ty) & 0xFFFFFFFF)) & MASK64
    return (h * 0x9E3779B9 + (int(tx) & 0xFFFFFFFF)) & MASK64


def py_standard_normal(seed: int, n: int) -> list[float]:
    state, out = int(seed) & MASK64, []
    while len(out) < n:
        state, u1 = py_pcg_next(state)
        state, u2 = py_pcg_next(state)
        v1 = 2.0 * (u1 + 1.0) / 4294967296.0 - 1.0
        v2 = 2.0 * (u2 + 1.0) / 4294967296.0 - 1.0
        s = v1 * v1 + v2 * v2
        if 0.0 < s < 1.0:
            f = math.sqrt(-2.0 * math.log(s) / s)
            out.append(v1 * f)
            if len(out) < n:
                out.append(v2 * f)
    return out
