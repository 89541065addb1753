"""Loader for ``oracle/_ref/``: the reference's OWN compiled kernels -- TEST INFRASTRUCTURE ONLY.

``oracle/make_ref.py`` lets the unmodified reference JIT its generated
C99+OpenMP into ``oracle/_ref/<md5>.c.so`` (one exported symbol per kernel).
This module binds those symbols with ctypes exactly as the reference's FFI does
(xgrid/util/ffi.py:18-35: no ``argtypes``, every argument an exact ctypes
instance) and performs what ``Operator.__call__`` does around the native call
(xgrid/lang/operator.py:37-41: resize + tick every grid argument).  The grid
argument is the by-value struct of xgrid/util/typing/reference.py:39-46 /
xgrid/lang/generator.py:139-147: ``{int32 time; int32 shape[d]; T** data;
int32* boundary_mask;}``.

Used to (a) validate the C restatement ``xgrid_oracle.c`` at sizes beyond the
golden vectors (tests/test_oracle_ref.py) and (b) as the CPU baseline of
``bench.py`` (``cpu_baseline.kind == "reference"``).  Valid for 1-D and square
2-D grids below 2^31 points; never used for 3-D.
"""
from __future__ import annotations

import ctypes as C
import json
import os

from . import HostGrid

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_manifest = None
_fns: dict = {}
_structs: dict = {}


def manifest() -> dict:
    global _manifest
    if _manifest is None:
        path = os.path.join(_DIR, "manifest.json")
        if os.path.exists(path):
            with open(path) as f:
                _manifest = json.load(f)
        else:
            _manifest = {}
    return _manifest


def available(kernel: str | None = None) -> bool:
    m = manifest()
    if kernel is None:
        return bool(m)
    return kernel in m and os.path.exists(os.path.join(_DIR, m[kernel]["lib"]))


def _grid_struct(ndim: int):
    if ndim not in _structs:
        _structs[ndim] = type(f"__Grid{ndim}d_d", (C.Structure,), {
            "_fields_": [("time", C.c_int32), ("shape", C.c_int32 * ndim),
                         ("data", C.POINTER(C.POINTER(C.c_double))),
                         ("boundary_mask", C.POINTER(C.c_int32))]})
    return _structs[ndim]


class CavityConfig(C.Structure):
    """examples/cavity.py:37-44 as the by-value C struct (declaration order, natural alignment)."""
    _fields_ = [("rho", C.c_double), ("nu", C.c_double), ("dt", C.c_double), ("dx", C.c_double),
                ("dy", C.c_double)]


def _serialize(g: HostGrid):
    """xgrid/xgrid/__init__.py:60-68"""
    assert g.dtype == "float64" and g.size < 2 ** 31
    assert g.dimension == 1 or (g.dimension == 2 and g.shape[0] == g.shape[1]), \
        "the reference's addressing is only valid for 1-D and square 2-D grids (SURVEY.md F1)"
    levels = (C.POINTER(C.c_double) * len(g._data))(*[g.ptr(l, C.c_double) for l in range(len(g._data))])
    s = _grid_struct(g.dimension)(len(g._data), (C.c_int32 * g.dimension)(*g.shape), levels, g.mask_ptr())
    return s, levels


def call(kernel: str, *args) -> None:
    """``kernel(*args)`` through the reference's compiled code.  ``HostGrid`` arguments are
    ticked (depth from the manifest), floats become ``double``, a 5-field config becomes the struct."""
    m = manifest()[kernel]
    fn = _fns.get(kernel)
    if fn is None:
        fn = getattr(C.CDLL(os.path.join(_DIR, m["lib"])), m["symbol"])
        fn.restype = None
        _fns[kernel] = fn
    keep, cargs = [], []
    for a in args:
        if isinstance(a, HostGrid):
            a._op_invoke(m["depth"])
            s, lv = _serialize(a)
            keep.append(lv)
            cargs.append(s)
        elif isinstance(a, float):
            cargs.append(C.c_double(a))
        elif hasattr(a, "rho"):
            cargs.append(CavityConfig(a.rho, a.nu, a.dt, a.dx, a.dy))
        else:
            raise TypeError(f"unsupported argument for the reference arm: {a!r}")
    fn(*cargs)


# bench / test name -> reference kernel symbol
KERNEL_OF = {"ewmul": "elementwise_mul", "conv1d": "convection_1d", "conv1d_nl": "convection_1d_nonlinear",
             "diff1d": "diffusion_1d", "conv2d": "convection_2d", "diff2d": "diffusion_2d",
             "cavity": "cavity_kernel"}
