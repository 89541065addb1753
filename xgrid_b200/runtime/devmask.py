"""Device-side mask compilation (csrc/kernels/xgb_mask.cu): int32 ``Grid.boundary`` ->
uint8 device mask + 128-point chunk flags + per-value histogram, in one fused pass per
uploaded piece.  Replaces several host NumPy passes over the mask."""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

_SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "kernels", "xgb_mask.cu")
PIECE = 1 << 24          # points per uploaded piece (64 MiB of int32)
_state = {}


class _Params(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("flags", C.c_void_p), ("hist", C.c_void_p),
                ("bad", C.c_void_p), ("n", C.c_int64), ("n_padded", C.c_int64)]


def _function(rt):
    fn = _state.get("fn")
    if fn is None:
        from ..config import get_config
        from . import shim
        with open(_SRC) as f:
            src = f.read()
        flags = ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"]
        root = os.path.join(".", get_config().cacheroot)
        os.makedirs(root, exist_ok=True)
        path = os.path.join(root, "mask_" + hashlib.sha256(src.encode()).hexdigest()[:24] + ".cubin")
        if os.path.exists(path):
            with open(path, "rb") as f:
                image = f.read()
        else:
            image, _ = shim.compile_cuda(src, "xgb_mask.cu", flags, {})
            tmp = path + f".{os.getpid()}.tmp"
            with open(tmp, "wb") as f:
                f.write(image)
            os.replace(tmp, path)
        module = rt.module_load(image)
        fn = _state["fn"] = rt.get_function(module, "xgb_mask_compile")
        _state["stage"] = rt.alloc(PIECE * 4)
        _state["aux"] = rt.alloc(256 * 8 + 64)
    return fn


def compile_mask(rt, boundary: np.ndarray, mask_dev: int, flags_dev: int, n_padded: int):
    """boundary: C-contiguous int32 array.  Fills mask_dev[0:n_padded] and flags_dev[0:n_padded/128].
    Returns (hist[256] as int64 array, bad flag)."""
    fn = _function(rt)
    flat = boundary.reshape(-1)
    n = flat.size
    stage, aux = _state["stage"], _state["aux"]
    rt.memset(aux, 0, 256 * 8 + 64)
    done = 0
    while done < n_padded:
        take = min(PIECE, n - done) if done < n else 0
        span = min(PIECE, n_padded - done)
        if take > 0:
            piece = flat[done:done + take]
            rt.h2d(stage, piece.ctypes.data, take * 4)
        P = _Params(stage, mask_dev + done, flags_dev + (done >> 7), aux, aux + 256 * 8, take, span)
        blocks = max(1, min(span >> 7, rt.sm_count * 16))
        rt.launch(fn, (blocks, 1, 1), (128, 1, 1), P)
        if take > 0:
            rt.sync()                 # the staging buffer is reused by the next piece
        done += span
    out = np.zeros(256 + 8, np.int64)
    rt.d2h(out.ctypes.data, aux, 256 * 8 + 64)
    rt.sync()
    return out[:256].copy(), bool(out[256] & 0xffffffff)
