"""CPU oracle for the xgrid stencil hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``xgrid_b200/`` imports this package.  It may be used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs, and only as the checker / the timed CPU arm.

Two parts:

* ``HostGrid`` restates the reference's storage semantics
  (xgrid/xgrid/__init__.py:21-86): a list of NumPy time levels, an int32
  ``boundary`` mask, ``_extend_time`` (append zero levels, truncate to depth)
  and the tick rotation (last buffer becomes ``now``).
* ``libxgrid_oracle.so`` (``xgrid_oracle.c``) restates the generated loop
  nests of the workload kernels; the ``step_*`` wrappers below do what
  ``Operator.__call__`` does (xgrid/lang/operator.py:37-41): tick every grid
  argument, then run the statements.

Parity pin: ``tests/golden/*.npz`` were produced by the *real* reference
(``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` exists) and ``tests/test_oracle.py`` checks this oracle
against them bit-for-bit.  3-D and non-square 2-D cases cannot be produced by
the reference (its linear index is wrong there); for those the
oracle is additionally checked against independent NumPy slice restatements
(tests/test_oracle.py).  ``oracle/interp.py`` is a second, generic oracle: a
NumPy interpreter of arbitrary kernels, pinned against the same golden vectors
(tests/test_interp.py) and used by the randomized differential tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SRC = os.path.join(_HERE, "xgrid_oracle.c")
_LIB = os.path.join(_BUILD, "libxgrid_oracle.so")

_lib = None
i64 = C.c_int64
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp, no -march (the reference never passes one: no FMA)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(_BUILD, exist_ok=True)
        cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fpic", _SRC, "-o", _LIB, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"oracle build failed: {' '.join(cmd)}\n{r.stderr}")
    return _LIB


class CavityCfg(C.Structure):
    _fields_ = [("rho", C.c_double), ("nu", C.c_double), ("dt", C.c_double), ("dx", C.c_double),
                ("dy", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.xo_threads.restype = C.c_int
        l.xo_ewmul_f64.argtypes = [_dp, _dp, _dp, _ip, i64]
        l.xo_ewmul_f32.argtypes = [_fp, _fp, _fp, _ip, i64]
        l.xo_conv1d_f64.argtypes = [_dp, _dp, _ip, i64, C.c_double, C.c_double, C.c_double]
        l.xo_conv1d_nonlinear_f64.argtypes = [_dp, _dp, _ip, i64, C.c_double, C.c_double]
        l.xo_diff1d_f64.argtypes = [_dp, _dp, _ip, i64, C.c_double, C.c_double, C.c_double]
        l.xo_conv2d_f64.argtypes = [_dp, _dp, _ip, i64, i64] + [C.c_double] * 4
        l.xo_conv2d_f32.argtypes = [_fp, _fp, _ip, i64, i64] + [C.c_float] * 4
        l.xo_diff2d_f64.argtypes = [_dp, _dp, _ip, i64, i64, C.c_double]
        l.xo_heat3d_f64.argtypes = [_dp, _dp, _ip, i64, i64, i64, C.c_double]
        l.xo_cavity_f64.argtypes = [_dp] * 7 + [_ip] * 4 + [i64, i64, CavityCfg, C.c_int]
        l.xo_fill_i32.argtypes = [_ip, _ip, i64, C.c_int32]
        for name in ("xo_ewmul_f64", "xo_ewmul_f32", "xo_conv1d_f64", "xo_conv1d_nonlinear_f64",
                     "xo_diff1d_f64", "xo_conv2d_f64", "xo_conv2d_f32", "xo_diff2d_f64", "xo_heat3d_f64",
                     "xo_cavity_f64", "xo_fill_i32"):
            getattr(l, name).restype = None
        _lib = l
    return _lib


def threads() -> int:
    return int(lib().xo_threads())


def set_threads(n: int) -> int:
    """OpenMP team size of every kernel in this process -- the C port and the reference's own shared
    objects under oracle/_ref share one libgomp.  Needed under torchrun, which exports
    OMP_NUM_THREADS=1 to its workers (libgomp reads the variable once, when it is loaded)."""
    lib()
    C.CDLL("libgomp.so.1").omp_set_num_threads(int(max(1, n)))
    return threads()


# --------------------------------------------------------------------------- storage
class HostGrid:
    """xgrid/xgrid/__init__.py:21-86 on the host, with padded level buffers."""

    def __init__(self, shape, dtype=np.float64) -> None:
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64))
        stride0 = self.size // self.shape[0] if self.shape[0] else 1
        self._pad = 3 * stride0 + 64
        self._data = [self._zeros()]                      # :38
        self.boundary = np.zeros(self.shape, np.int32)    # :41

    def _zeros(self) -> np.ndarray:
        raw = np.zeros(self.size + 2 * self._pad, self.dtype)
        view = raw[self._pad:self._pad + self.size].reshape(self.shape)
        return view

    @property
    def dimension(self) -> int:
        return len(self.shape)

    def _extend_time(self, depth: int) -> None:            # :43-47
        while len(self._data) < depth:
            self._data.append(self._zeros())
        self._data = self._data[:depth]

    def _op_invoke(self, depth: int, tick: bool = True) -> None:   # :49-54
        self._extend_time(depth)
        if tick:
            self._data.insert(0, self._data.pop())

    @property
    def now(self) -> np.ndarray:                            # :70-72
        return self._data[0]

    def __getitem__(self, key):
        return self.now[key]

    def __setitem__(self, key, value) -> None:
        self.now[key] = value

    def ptr(self, level: int, ctype):
        return self._data[level].ctypes.data_as(C.POINTER(ctype))

    def mask_ptr(self):
        self.boundary = np.ascontiguousarray(self.boundary, np.int32)
        return self.boundary.ctypes.data_as(_ip)


def _ct(dtype):
    return {np.dtype(np.float64): C.c_double, np.dtype(np.float32): C.c_float,
            np.dtype(np.int32): C.c_int32}[np.dtype(dtype)]


# --------------------------------------------------------------------------- kernels = tick + statements
def step_ewmul(result: HostGrid, a: HostGrid, b: HostGrid) -> None:
    """README.md:26-28; depth 2; all three grids tick."""
    for g in (result, a, b):
        g._op_invoke(2)
    ct = _ct(result.dtype)
    fn = lib().xo_ewmul_f64 if ct is C.c_double else lib().xo_ewmul_f32
    fn(result.ptr(0, ct), a.ptr(1, ct), b.ptr(1, ct), result.mask_ptr(), result.size)


def step_conv1d(u: HostGrid, c: float, dt: float, dx: float) -> None:
    u._op_invoke(2)
    lib().xo_conv1d_f64(u.ptr(0, C.c_double), u.ptr(1, C.c_double), u.mask_ptr(), u.size, c, dt, dx)


def step_conv1d_nonlinear(u: HostGrid, dt: float, dx: float) -> None:
    u._op_invoke(2)
    lib().xo_conv1d_nonlinear_f64(u.ptr(0, C.c_double), u.ptr(1, C.c_double), u.mask_ptr(), u.size, dt, dx)


def step_diff1d(u: HostGrid, nu: float, dt: float, dx: float) -> None:
    u._op_invoke(2)
    lib().xo_diff1d_f64(u.ptr(0, C.c_double), u.ptr(1, C.c_double), u.mask_ptr(), u.size, nu, dt, dx)


def step_conv2d(u: HostGrid, c: float, dt: float, dx: float, dy: float) -> None:
    u._op_invoke(2)
    ct = _ct(u.dtype)
    fn = lib().xo_conv2d_f64 if ct is C.c_double else lib().xo_conv2d_f32
    fn(u.ptr(0, ct), u.ptr(1, ct), u.mask_ptr(), u.shape[0], u.shape[1], c, dt, dx, dy)


def step_diff2d(u: HostGrid, a: float) -> None:
    u._op_invoke(2)
    lib().xo_diff2d_f64(u.ptr(0, C.c_double), u.ptr(1, C.c_double), u.mask_ptr(), u.shape[0], u.shape[1], a)


def step_heat3d(u: HostGrid, a: float) -> None:
    u._op_invoke(2)
    lib().xo_heat3d_f64(u.ptr(0, C.c_double), u.ptr(1, C.c_double), u.mask_ptr(), *u.shape, a)


@dataclass
class Config:
    rho: float
    nu: float
    dt: float
    dx: float
    dy: float


def step_cavity(b: HostGrid, p: HostGrid, u: HostGrid, v: HostGrid, cfg: Config, nit: int = 50) -> None:
    """examples/cavity.py:74-142; depth 2; b, p, u, v all tick."""
    for g in (b, p, u, v):
        g._op_invoke(2)
    d = C.c_double
    lib().xo_cavity_f64(b.ptr(0, d), p.ptr(0, d), p.ptr(1, d), u.ptr(0, d), u.ptr(1, d), v.ptr(0, d),
                        v.ptr(1, d), b.mask_ptr(), p.mask_ptr(), u.mask_ptr(), v.mask_ptr(),
                        p.shape[0], p.shape[1], CavityCfg(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy), nit)


def step_fill_i32(a: HostGrid, value: int = 4) -> None:
    """test.py:171-172; depth 1 (no loads)."""
    a._op_invoke(1)
    lib().xo_fill_i32(a.ptr(0, C.c_int32), a.mask_ptr(), a.size, value)
