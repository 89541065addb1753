"""xgrid_b200 -- a B200-native execution backend for the xgrid DSL.

Drop-in for the reference's public API (xgrid/__init__.py:52-53):

    import xgrid_b200 as xgrid
    xgrid.init(precision="double")
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def step(u: f2, a: float) -> None:
        u[0, 0] = u[0, 0] + a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0] - 4.0 * u[0, 0])

    u = xgrid.Grid((4096, 4096), float)
    step(u, 0.1); print(u.now)

Kernels are lowered to CUDA C for sm_100a, JIT-compiled with NVRTC and launched
through the C ABI in include/xgrid_b200.h; grids live in HBM.  There is no CPU
execution path.
"""
from __future__ import annotations

import struct as _struct
from typing import Any

from .config import init, get_config
from .grid import Grid
from .lang import boundary, c
from .lang.operator import external, function, kernel
from .types import BaseType, Grid as _GridT, Integer, Void, grid, ptr

__version__ = "0.1.0"


def _dimension_typecheck(args: list) -> BaseType:
    if not isinstance(args[0], _GridT):
        raise Exception(f"Incompatible dimension to type '{args[0]}'")
    return Integer(_struct.calcsize("i"))


def _shape_typecheck(args: list) -> BaseType:
    if not isinstance(args[0], _GridT):
        raise Exception(f"Incompatible shape to type '{args[0]}'")
    if not isinstance(args[1], Integer):
        raise Exception(f"Incompatible shape dimension '{args[1]}'")
    return Integer(_struct.calcsize("i"))


def _tick_typecheck(args: list) -> BaseType:
    if not isinstance(args[0], _GridT):
        raise Exception(f"Incompatible tick to type '{args[0]}'")
    return Void()


@external(typecheck_override=_dimension_typecheck)
def dimension(grid: Any) -> int:
    ...


@external(typecheck_override=_shape_typecheck)
def shape(grid: Any, dimension: int) -> int:
    ...


@external(typecheck_override=_tick_typecheck)
def tick(grid: Any) -> None:
    ...


def flush() -> None:
    """Enqueue every deferred kernel call on the device stream (B200 extension).  Calls of
    1-D kernels are deferred so that runs of identical calls can execute several time steps
    per launch; reading ``Grid.now`` / ``.boundary`` flushes implicitly."""
    from .lang.schedule import flush_pending
    flush_pending()


def synchronize() -> None:
    """Flush deferred calls and block until every enqueued sweep has finished (B200 extension)."""
    from .runtime.shim import Runtime
    flush()
    Runtime.get().sync()


def empty_cache() -> None:
    """Return the device buffers cached by the runtime's pool to the driver (B200 extension; the pool keeps
    freed time levels for re-use because cudaFree / cudaMalloc of large buffers cost ~100 ms each)."""
    from .runtime.shim import Runtime
    if Runtime._instance is not None:
        Runtime._instance.trim_pool()


__all__ = ["kernel", "function", "init", "ptr", "grid", "boundary", "c", "external", "Grid",
           "shape", "dimension", "tick", "synchronize", "flush", "empty_cache"]
