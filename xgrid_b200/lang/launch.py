"""Launch side of a kernel call: marshal one by-value parameter struct per
sweep group, pick the kernel variant and launch geometry, and enqueue on the
backend's compute stream.  This is the counterpart of the reference's per-call
``serialize`` + ctypes foreign call (xgrid/util/ffi.py:30-34,
xgrid/xgrid/__init__.py:60-68) -- asynchronous, and with the descriptors
cached instead of rebuilt every call.
"""
from __future__ import annotations

import ctypes
from dataclasses import astuple

import numpy as np

from ..types import Pointer, Structure
from . import cudagen

SPARSE_FRACTION = 16      # run a mask!=0 statement over its index list when it
                          # covers less than 1/16 of the grid


def _pow2ceil(x: int) -> int:
    p = 1
    while p < x:
        p <<= 1
    return p


def dense_geometry(rows: int, cols: int, V: int):
    """blockDim=(TX,TY) threads, each owning V columns of one row; grid.x walks
    the columns fastest so consecutive CTAs share rows in L2."""
    vec_cols = (cols + V - 1) // V
    tx = min(256, max(32, _pow2ceil(vec_cols)))
    ty = max(1, min(256 // tx, _pow2ceil(rows)))
    gx = (vec_cols + tx - 1) // tx
    rb = (rows + ty - 1) // ty
    gy = min(rb, 65535)
    gz = (rb + gy - 1) // gy
    return (gx, gy, gz), (tx, ty, 1)


class Launcher:
    def __init__(self, program, grids: dict) -> None:
        from ..runtime.shim import Runtime
        self.program = program
        self.grids = grids
        self.rt = Runtime.get()
        self.launches = 0

    def _scalar_value(self, t, raw):
        if isinstance(t, Pointer):
            raw = raw.contents.value if hasattr(raw, "contents") else raw.value
            t = t.element
        if isinstance(t, Structure):
            return t.ctype(*astuple(raw))
        return raw.item() if isinstance(raw, np.generic) else raw

    def __call__(self, g: cudagen.Group, env: dict) -> None:
        lead = self.grids[g.lead]
        P = g.params_cls()
        for s in g.slots:
            grid = self.grids[s.grid]
            if grid.shape != lead.shape:
                raise Exception(f"grid '{s.grid}' has shape {grid.shape} but the sweep over '{g.lead}' "
                                f"runs on {lead.shape}; all grids of one stencil statement must agree")
            lv = grid._scratch_level() if s.level == "scratch" else grid._ring[s.level]
            setattr(P, s.field, lv.dev)
        for m in g.masks:
            grid = self.grids[m]
            setattr(P, f"m_{m}", grid._mask_dev if grid._mask_any else None)
            setattr(P, f"f_{m}", grid._flags_dev if grid._mask_any else None)
        shape = lead.shape
        cols = shape[-1]
        rows = lead.size // cols if cols else 0
        P.rows, P.cols = rows, cols
        for a, n in enumerate(shape):
            setattr(P, f"n{a}", n)
        for name in g.shapes:
            arr = getattr(P, f"shape_{name}")
            for a, n in enumerate(self.grids[name].shape):
                arr[a] = n
        for name, t in g.scalars.items():
            setattr(P, f"u_{name}", self._scalar_value(t, env[name]))

        if lead.size == 0:
            return
        variant, V = cudagen.VARIANT_DENSE, 1
        if g.sparse:
            k = g.stmts[0].sweep.mask
            count = lead._mask_count(k)
            if count == 0:
                return
            if count * SPARSE_FRACTION <= lead.size:
                variant = cudagen.VARIANT_SPARSE
                ptr, count = lead._index_list(k)
                P.list, P.count = ptr, count
        if variant == cudagen.VARIANT_DENSE:
            for cand in sorted(g.vwidths, reverse=True):
                if cols % cand == 0:
                    V = cand
                    break
            grid_dim, block_dim = dense_geometry(rows, cols, V)
        else:
            block_dim = (128, 1, 1)
            grid_dim = ((P.count + 127) // 128, 1, 1)
        fn = self.program.function(cudagen.kernel_name(g, variant, V))
        self.rt.launch(fn, grid_dim, block_dim, P)
        self.launches += 1
        if g.implicit:
            self.grids[g.stmts[0].sweep.grid.name]._swap_scratch()

    def finish(self) -> None:
        pass
