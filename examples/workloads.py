"""The BASELINE.json workload kernels written in the xgrid DSL, verbatim from
the reference's README / test.py / examples (with the one-argument
``xgrid.boundary(k)`` form), plus the synthetic-input builders of SURVEY.md
§8d.  Shared by tests, ``bench.py`` and ``__graft_entry__.smoke()`` so that
all three exercise the same programs.

Kernels are created by factory functions because annotations such as
``xgrid.grid[float, 1]`` resolve ``float`` through ``xgrid.init(precision=...)``
at parse time (xgrid/util/typing/annotation.py:41-42).
"""

from dataclasses import dataclass

import numpy as np

import xgrid_b200 as xgrid


@dataclass
class Config:
    """examples/cavity.py:37-44"""
    rho: float
    nu: float
    dt: float
    dx: float
    dy: float


def make_kernels() -> dict:
    """Define every workload kernel under the *current* ``xgrid.init`` config."""
    f1 = xgrid.grid[float, 1]
    f2 = xgrid.grid[float, 2]
    f3 = xgrid.grid[float, 3]
    i2 = xgrid.grid[int, 2]

    @xgrid.kernel()
    def elementwise_mul(result: f1, a: f1, b: f1) -> None:      # README.md:26-28
        result[0] = a[0] * b[0]

    @xgrid.kernel()
    def convection_1d(u: f1, c: float, dt: float, dx: float) -> None:   # test.py:214-218
        u[0] = u[0] - c * dt / dx * (u[0] - u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    @xgrid.kernel()
    def convection_1d_nonlinear(u: f1, dt: float, dx: float) -> None:   # test.py:240-244
        u[0] = u[0] - u[0] * dt / dx * (u[0] - u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    @xgrid.kernel()
    def diffusion_1d(u: f1, nu: float, dt: float, dx: float) -> None:   # test.py:269-273
        u[0] = u[0] + nu * dt / dx ** 2.0 * (u[1] - 2.0 * u[0] + u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    @xgrid.kernel()
    def convection_2d(u: f2, c: float, dt: float, dx: float, dy: float) -> None:   # test.py:300-307
        cdx = c * dt / dx
        cdy = c * dt / dy
        u[0, 0] = u[0, 0] + cdx * (u[0, 0] - u[-1, 0]) - cdy * (u[0, 0] - u[0, -1])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    @xgrid.kernel()
    def diffusion_2d(u: f2, a: float) -> None:                  # 5-point, SURVEY.md §8d C3
        u[0, 0] = u[0, 0] + a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0] - 4.0 * u[0, 0])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    @xgrid.kernel()
    def diffusion_2d_open(u: f2, a: float) -> None:             # no mask: overstep-mode fixture
        u[0, 0] = u[0, 0] + a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0] - 4.0 * u[0, 0])

    @xgrid.kernel()
    def heat_3d(u: f3, a: float) -> None:                       # 7-point, SURVEY.md §8d C5
        u[0, 0, 0] = u[0, 0, 0] + a * (u[1, 0, 0] + u[-1, 0, 0] + u[0, 1, 0] + u[0, -1, 0] +
                                       u[0, 0, 1] + u[0, 0, -1] - 6.0 * u[0, 0, 0])
        with xgrid.boundary(1):
            u[0, 0, 0] = 0.0

    @xgrid.kernel()
    def fill4(a: i2) -> None:                                   # test.py:171-172
        a[0, 0] = 4

    @xgrid.kernel()
    def index_guard(a: i2) -> None:                             # test.py:186-187
        a[0, 0] = a[-1, -1][-1]

    @xgrid.kernel()
    def cavity_kernel(b: f2, p: f2, u: f2, v: f2, cfg: Config) -> None:   # examples/cavity.py:74-142
        b[0, 0] = (cfg.rho * (1.0 / cfg.dt *
                              ((u[0, 1] - u[0, -1]) /
                               (2.0 * cfg.dx) + (v[1, 0] - v[-1, 0]) / (2.0 * cfg.dy)) -
                              ((u[0, 1] - u[0, -1]) / (2.0 * cfg.dx))**2.0 -
                              2.0 * ((u[1, 0] - u[-1, 0]) / (2.0 * cfg.dy) *
                                     (v[0, 1] - v[0, -1]) / (2.0 * cfg.dx)) -
                              ((v[1, 0] - v[-1, 0]) / (2.0 * cfg.dy))**2.0))

        p[0, 0] = (((p[0, 1] + p[0, -1]) * cfg.dy**2.0 +
                    (p[1, 0] + p[-1, 0]) * cfg.dx**2.0) /
                   (2.0 * (cfg.dx**2.0 + cfg.dy**2.0)) -
                   cfg.dx**2.0 * cfg.dy**2.0 / (2.0 * (cfg.dx**2.0 + cfg.dy**2.0)) *
                   b[0, 0][0])

        with xgrid.boundary(1):
            p[0, 0] = p[0, -1][0]  # dp/dx = 0 at x = 2
        with xgrid.boundary(2):
            p[0, 0] = p[1, 0][0]   # dp/dy = 0 at y = 0
        with xgrid.boundary(3):
            p[0, 0] = p[0, 1][0]   # dp/dx = 0 at x = 0
        with xgrid.boundary(4):
            p[0, 0] = 0.0

        for _ in range(0, 50):
            p[0, 0] = (((p[0, 1][0] + p[0, -1][0]) * cfg.dy**2.0 +
                        (p[1, 0][0] + p[-1, 0][0]) * cfg.dx**2.0) /
                       (2.0 * (cfg.dx**2.0 + cfg.dy**2.0)) -
                       cfg.dx**2.0 * cfg.dy**2.0 / (2.0 * (cfg.dx**2.0 + cfg.dy**2.0)) *
                       b[0, 0][0])

            with xgrid.boundary(1):
                p[0, 0] = p[0, -1][0]
            with xgrid.boundary(2):
                p[0, 0] = p[1, 0][0]
            with xgrid.boundary(3):
                p[0, 0] = p[0, 1][0]
            with xgrid.boundary(4):
                p[0, 0] = 0.0

        u[0, 0] = (u[0, 0] -
                   u[0, 0] * cfg.dt / cfg.dx *
                   (u[0, 0] - u[0, -1]) -
                   v[0, 0] * cfg.dt / cfg.dy *
                   (u[0, 0] - u[-1, 0]) -
                   cfg.dt / (2.0 * cfg.rho * cfg.dx) * (p[0, 1][0] - p[0, -1][0]) +
                   cfg.nu * (cfg.dt / cfg.dx**2.0 *
                             (u[0, 1] - 2.0 * u[0, 0] + u[0, -1]) +
                             cfg.dt / cfg.dy**2.0 *
                             (u[1, 0] - 2.0 * u[0, 0] + u[-1, 0])))

        v[0, 0] = (v[0, 0] -
                   u[0, 0] * cfg.dt / cfg.dx *
                   (v[0, 0] - v[0, -1]) -
                   v[0, 0] * cfg.dt / cfg.dy *
                   (v[0, 0] - v[-1, 0]) -
                   cfg.dt / (2.0 * cfg.rho * cfg.dy) * (p[1, 0][0] - p[-1, 0][0]) +
                   cfg.nu * (cfg.dt / cfg.dx**2.0 *
                             (v[0, 1] - 2.0 * v[0, 0] + v[0, -1]) +
                             cfg.dt / cfg.dy**2.0 *
                             (v[1, 0] - 2.0 * v[0, 0] + v[-1, 0])))

        with xgrid.boundary(1):
            u[0, 0] = 0.0
            v[0, 0] = 0.0

        with xgrid.boundary(2):
            u[0, 0] = 1.0

    return {k: v for k, v in locals().items() if isinstance(v, xgrid.lang.operator.Operator)}


# --------------------------------------------------------------------------- synthetic inputs
def ic_1d(n: int, dtype=np.float64):
    """IC 1.0 with 2.0 on [int(.5/dx) : int(1/dx + 1)] (test.py:195-198)."""
    dx = 2.0 / (n - 1)
    u = np.ones(n, dtype)
    u[int(.5 / dx):int(1 / dx + 1)] = 2
    return u, dx


def ic_2d_box(n: int, dtype=np.float64):
    dx = 2.0 / (n - 1)
    u = np.ones((n, n), dtype)
    u[int(.5 / dx):int(1 / dx) + 1, int(.5 / dx):int(1 / dx) + 1] = 2
    return u, dx


def shell_mask(shape, value: int = 1) -> np.ndarray:
    m = np.full(shape, value, np.int32)
    m[tuple(slice(1, -1) for _ in shape)] = 0
    return m


def shell_mask_slab(global_shape, lo: int, hi: int, value: int = 1) -> np.ndarray:
    """Rows [lo, hi) of shell_mask(global_shape) without materialising the global mask."""
    local = (hi - lo,) + tuple(global_shape[1:])
    m = np.full(local, value, np.int32)
    inner = tuple(slice(1, -1) for _ in global_shape[1:])
    a = max(lo, 1) - lo
    b = min(hi, global_shape[0] - 1) - lo
    if b > a:
        m[(slice(a, b),) + inner] = 0
    return m


def cavity_masks_slab(n0: int, n1: int, lo: int, hi: int):
    """Rows [lo, hi) of cavity_masks(n0, n1) without materialising the global masks."""
    rows = hi - lo
    first, last = (lo == 0), (hi == n0)
    mu = np.zeros((rows, n1), np.int32)
    mu[:, 0] = 1
    mu[:, -1] = 1
    if first:
        mu[0, :] = 1
    if last:
        mu[-1, :] = 2
    mv = shell_mask_slab((n0, n1), lo, hi)
    mp = np.zeros((rows, n1), np.int32)
    mp[:, -1] = 1
    if first:
        mp[0, :] = 2
    mp[:, 0] = 3
    if last:
        mp[-1, :] = 4
    return mv.copy(), mp, mu, mv


def cavity_masks(n0: int, n1: int):
    """examples/cavity.py:56-68 -> (mb, mp, mu, mv)"""
    mu = np.zeros((n0, n1), np.int32)
    mu[0, :] = 1
    mu[:, 0] = 1
    mu[:, -1] = 1
    mu[-1, :] = 2
    mv = shell_mask((n0, n1))
    mp = np.zeros((n0, n1), np.int32)
    mp[:, -1] = 1
    mp[0, :] = 2
    mp[:, 0] = 3
    mp[-1, :] = 4
    mb = shell_mask((n0, n1))
    return mb, mp, mu, mv
