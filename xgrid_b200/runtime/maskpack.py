"""Mask upload: int32 ``Grid.boundary`` -> one byte per point + 128-point chunk flags on the device,
plus the per-value histogram and a packed host copy.

The reference passes the int32 array itself to every sweep (xgrid/xgrid/__init__.py:41,66).  Here the
array is packed on the host by several threads (``xgb_mask_pack``: most 128-point chunks are all zero and
cost one OR-reduction + one memset) and only the packed bytes cross PCIe -- a quarter of the int32 array.
The packed copy doubles as the host snapshot used for change detection, index lists and the fused-pair
check, so no further host pass over the mask is needed."""
from __future__ import annotations

import ctypes as C

import numpy as np

STAGED_MIN = 8 << 20     # packed masks at least this large are uploaded through the staged path


def pack(boundary: np.ndarray, n_padded: int):
    """boundary: C-contiguous int32 array.  Returns (packed uint8 array of n_padded bytes, chunk flags,
    hist[256] as int64 array, bad flag).  Host only."""
    from . import shim
    flat = boundary.reshape(-1)
    n = flat.size
    packed = np.empty(n_padded, np.uint8)
    flags = np.empty(n_padded // 128, np.uint8)
    hist = np.zeros(256, np.uint64)
    bad = C.c_int(0)
    shim.check(shim.lib().xgb_mask_pack(flat.ctypes.data_as(C.POINTER(C.c_int32)), n, n_padded,
                                        packed.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        flags.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(bad),
                                        shim.host_threads(16)))
    return packed, flags, hist.astype(np.int64), bool(bad.value)


def upload(rt, packed: np.ndarray, flags: np.ndarray, mask_dev: int, flags_dev: int) -> None:
    """Fills mask_dev[0:n_padded] and flags_dev[0:n_padded/128] from the packed host arrays."""
    if rt.STAGED and packed.size >= STAGED_MIN:
        rt.h2d_staged(mask_dev, packed.ctypes.data, packed.size)
    else:
        rt.h2d(mask_dev, packed.ctypes.data, packed.size)
    rt.h2d(flags_dev, flags.ctypes.data, flags.size)
    rt.sync()                     # `flags` may be released after return
