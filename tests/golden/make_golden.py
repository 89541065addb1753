#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container, where the reference is mounted read-only:

    cd /tmp && python /root/repo/tests/golden/make_golden.py

The reference (``/root/reference``: pure Python that generates C99+OpenMP and
compiles it with the system gcc) is imported from where it lies; it needs a
writable cwd for its JIT cache, so the script chdirs to a temp dir.  The
kernels are lifted from the reference's own tests / examples / README with the
one-argument ``xgrid.boundary(k)`` form (the two-argument form in test.py:217
is rejected by the reference's own parser):

  README.md:26-28 (elementwise_mul, fp32 default and fp64), test.py:171-192
  (int grid fill, index guard), test.py:195-221 / 228-247 / 254-276 / 283-310
  (conv1d, nonlinear, diff1d, conv2d), examples/cavity.py:40-142 (cavity),
  plus a 5-point 2-D diffusion, an unmatched-mask case, scalar /
  struct kernels (test.py:130-165) and overstep="wrap"/"limit" on a square grid.

Every fixture stores the inputs and EVERY ring level of every grid after the
run, so ring rotation and never-written cells are pinned too.  The reference
cannot travel to the GPU box; these files can.
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def levels(g):
    return {f"L{k}": np.array(a) for k, a in enumerate(g._data)}


def save(name, **arrays):
    flat = {}
    for key, val in arrays.items():
        if isinstance(val, dict):
            for k2, v2 in val.items():
                flat[f"{key}.{k2}"] = v2
        else:
            flat[key] = np.asarray(val)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print("wrote", name, sorted(flat))


def main():
    os.chdir(tempfile.mkdtemp(prefix="xgrid_golden_"))
    sys.path.insert(0, REF)
    import xgrid
    from dataclasses import dataclass
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.warn
    g = globals()

    # ------------------------------------------------------------------ fp64 block
    xgrid.init(precision="double", opt_level=3, cacheroot=".xg", parallel=True)
    f1 = xgrid.grid[float, 1]
    f2 = xgrid.grid[float, 2]
    i2 = xgrid.grid[int, 2]
    g.update(xgrid=xgrid)

    @xgrid.kernel()
    def elementwise_mul(result: f1, a: f1, b: f1) -> None:
        result[0] = a[0] * b[0]

    rng = np.random.default_rng(0)
    a, b, r = (xgrid.Grid((10000,), float) for _ in range(3))
    a_in, b_in = rng.random(10000), rng.random(10000)
    a.now[:] = a_in
    b.now[:] = b_in
    elementwise_mul(r, a, b)
    once = (levels(r), levels(a), levels(b))
    elementwise_mul(r, a, b)       # second call: ring rotates again (F4)
    save("ewmul_f64", a_in=a_in, b_in=b_in, r1=once[0], a1=once[1], b1=once[2],
         r2=levels(r), a2=levels(a), b2=levels(b))

    # conv1d / nonlinear / diffusion: test.py:195-276
    def ic1d(u, dx):
        u.now.fill(1)
        u.now[int(.5 / dx):int(1 / dx + 1)] = 2

    nx = 41
    dx = 2 / (nx - 1)

    @xgrid.kernel()
    def convection_1d(u: f1, c: float, dt: float, dx: float) -> None:
        u[0] = u[0] - c * dt / dx * (u[0] - u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    u = xgrid.Grid((nx,), float)
    ic1d(u, dx)
    u.boundary[0] = 1
    u_in = u.now.copy()
    for _ in range(25):
        convection_1d(u, 1.0, .025, dx)
    save("conv1d_f64", u_in=u_in, mask=u.boundary, params=[1.0, .025, dx], steps=25, u=levels(u))

    @xgrid.kernel()
    def convection_1d_nl(u: f1, dt: float, dx: float) -> None:
        u[0] = u[0] - u[0] * dt / dx * (u[0] - u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    u = xgrid.Grid((nx,), float)
    ic1d(u, dx)
    u.boundary[0] = 1
    u_in = u.now.copy()
    for _ in range(10):
        convection_1d_nl(u, .025, dx)
    save("conv1d_nonlinear_f64", u_in=u_in, mask=u.boundary, params=[.025, dx], steps=10, u=levels(u))

    @xgrid.kernel()
    def diffusion_1d(u: f1, nu: float, dt: float, dx: float) -> None:
        u[0] = u[0] + nu * dt / dx ** 2.0 * (u[1] - 2.0 * u[0] + u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    u = xgrid.Grid((nx,), float)
    ic1d(u, dx)
    u.boundary[0] = u.boundary[-1] = 1
    u_in = u.now.copy()
    for _ in range(20):
        diffusion_1d(u, .01, .01, dx)
    save("diff1d_f64", u_in=u_in, mask=u.boundary, params=[.01, .01, dx], steps=20, u=levels(u))

    # unmatched mask value: those cells are never written (F5)
    u = xgrid.Grid((nx,), float)
    ic1d(u, dx)
    u.boundary[0] = 1
    u.boundary[7] = 5
    u.boundary[20:23] = 9
    u_in = u.now.copy()
    for _ in range(7):
        convection_1d(u, 1.0, .025, dx)
    save("conv1d_stale_f64", u_in=u_in, mask=u.boundary, params=[1.0, .025, dx], steps=7, u=levels(u))

    # conv2d: test.py:283-310
    n2 = 101
    dx2 = 2 / (n2 - 1)
    dt2 = .5 * dx2
    nt2 = int(.7 / dt2)

    @xgrid.kernel()
    def convection_2d(u: f2, c: float, dt: float, dx: float, dy: float) -> None:
        cdx = c * dt / dx
        cdy = c * dt / dy
        u[0, 0] = u[0, 0] + cdx * (u[0, 0] - u[-1, 0]) - cdy * (u[0, 0] - u[0, -1])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    u = xgrid.Grid((n2, n2), float)
    u.now.fill(1)
    u.now[int(.5 / dx2):int(1 / dx2) + 1, int(.5 / dx2):int(1 / dx2) + 1] = 2
    u.boundary[0, :] = u.boundary[:, 0] = 1
    u_in = u.now.copy()
    for _ in range(nt2):
        convection_2d(u, 1.0, dt2, dx2, dx2)
    save("conv2d_f64", u_in=u_in, mask=u.boundary, params=[1.0, dt2, dx2, dx2], steps=nt2, u=levels(u))

    # 5-point diffusion on a square grid with a shell mask
    @xgrid.kernel()
    def diffusion_2d(u: f2, a: float) -> None:
        u[0, 0] = u[0, 0] + a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0] - 4.0 * u[0, 0])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    u = xgrid.Grid((64, 64), float)
    u.now[:] = np.random.default_rng(1).random((64, 64))
    u.boundary[0, :] = u.boundary[-1, :] = u.boundary[:, 0] = u.boundary[:, -1] = 1
    u_in = u.now.copy()
    for _ in range(30):
        diffusion_2d(u, 0.2)
    save("diff2d_f64", u_in=u_in, mask=u.boundary, params=[0.2], steps=30, u=levels(u))

    # int grid fill + index guard: test.py:168-192
    @xgrid.kernel()
    def fill4(a: i2) -> None:
        a[0, 0] = 4

    ig = xgrid.Grid((10, 10), dtype=int)
    fill4(ig)
    save("fill_i32", a=levels(ig))

    @xgrid.kernel()
    def guard(a: i2) -> None:
        a[0, 0] = a[-1, -1][-1]

    ig = xgrid.Grid((10, 10), dtype=int)
    guard(ig)
    save("indexguard_i32", a=levels(ig))

    # scalar + struct kernels: test.py:127-165
    g["TEMP"] = 10

    @xgrid.kernel()
    def add3(a: int, b: int) -> int:
        return a + b + TEMP   # noqa: F821

    @dataclass
    class Vector3i:
        x: int
        y: int
        z: int

        @xgrid.function(method=True)
        def dot(self, b: "Vector3i") -> int:
            return self.x * b.x + self.y * b.y + self.z * b.z
    g["Vector3i"] = Vector3i

    @xgrid.kernel()
    def vdot(a: Vector3i, b: Vector3i) -> int:
        return a.dot(b)

    @xgrid.kernel()
    def scal(a: float, b: float, n: int) -> float:
        acc = 0.0
        for i in range(0, n):
            if i % 2 == 0:
                acc = acc + a / b
            else:
                acc = acc - a * b ** 2.0
        return acc

    save("scalars", add3=add3(123, 456), vdot=vdot(Vector3i(1, -2, 3), Vector3i(4, 5, -6)),
         scal=scal(1.7, 0.3, 9))

    # cavity: examples/cavity.py:40-142, 101^2, 20 timesteps
    @dataclass
    class Config:
        rho: float
        nu: float
        dt: float
        dx: float
        dy: float
    g["Config"] = Config

    @xgrid.kernel()
    def cavity_kernel(b: f2, p: f2, u: f2, v: f2, cfg: Config) -> None:
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
            p[0, 0] = p[0, -1][0]
        with xgrid.boundary(2):
            p[0, 0] = p[1, 0][0]
        with xgrid.boundary(3):
            p[0, 0] = p[0, 1][0]
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

    def cavity_grids(n):
        u, v, p, b = (xgrid.Grid((n, n), float) for _ in range(4))
        u.boundary[0, :] = u.boundary[:, 0] = u.boundary[:, -1] = 1
        u.boundary[-1, :] = 2
        v.boundary.fill(1)
        v.boundary[1:-1, 1:-1] = 0
        p.boundary[:, -1] = 1
        p.boundary[0, :] = 2
        p.boundary[:, 0] = 3
        p.boundary[-1, :] = 4
        b.boundary.fill(1)
        b.boundary[1:-1, 1:-1] = 0
        return b, p, u, v

    for n, steps in ((101, 20), (41, 100)):
        b, p, u, v = cavity_grids(n)
        cfg = Config(1.0, 0.1, 0.0001, 2 / (n - 1), 2 / (n - 1))
        for _ in range(steps):
            cavity_kernel(b, p, u, v, cfg)
        save(f"cavity_{n}_f64", steps=steps, cfg=[cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy],
             mb=b.boundary, mp=p.boundary, mu=u.boundary, mv=v.boundary,
             b=levels(b), p=levels(p), u=levels(u), v=levels(v))

    # ------------------------------------------------------------------ overstep modes (square grid)
    for mode in ("wrap", "limit"):
        xgrid.init(precision="double", opt_level=3, cacheroot=".xg", parallel=True, overstep=mode)

        @xgrid.kernel(name=f"diffusion_2d_{mode}")
        def diffusion_2d_os(u: f2, a: float) -> None:
            u[0, 0] = u[0, 0] + a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0] - 4.0 * u[0, 0])

        u = xgrid.Grid((32, 32), float)
        u.now[:] = np.random.default_rng(2).random((32, 32))
        u_in = u.now.copy()
        for _ in range(10):
            diffusion_2d_os(u, 0.2)
        save(f"diff2d_{mode}_f64", u_in=u_in, params=[0.2], steps=10, u=levels(u))

    # ------------------------------------------------------------------ fp32 block (the README default)
    xgrid.init(opt_level=3, cacheroot=".xg", parallel=True)      # precision="float"
    f1s = xgrid.grid[float, 1]
    f2s = xgrid.grid[float, 2]

    @xgrid.kernel(name="elementwise_mul_f32")
    def elementwise_mul32(result: f1s, a: f1s, b: f1s) -> None:
        result[0] = a[0] * b[0]

    a, b, r = (xgrid.Grid((10000,), float) for _ in range(3))
    a.now[:] = a_in.astype(np.float32)
    b.now[:] = b_in.astype(np.float32)
    elementwise_mul32(r, a, b)
    save("ewmul_f32", a_in=a_in.astype(np.float32), b_in=b_in.astype(np.float32),
         r1=levels(r), a1=levels(a), b1=levels(b))

    @xgrid.kernel(name="diffusion_1d_f32")
    def diffusion_1d32(u: f1s, nu: float, dt: float, dx: float) -> None:
        # float grid, double literals: mixed precision by C's conversions (F6)
        u[0] = u[0] + nu * dt / dx ** 2.0 * (u[1] - 2.0 * u[0] + u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    u = xgrid.Grid((nx,), float)
    ic1d(u, dx)
    u.boundary[0] = u.boundary[-1] = 1
    u_in = u.now.copy()
    for _ in range(20):
        diffusion_1d32(u, .01, .01, dx)
    save("diff1d_f32", u_in=u_in, mask=u.boundary, params=[.01, .01, dx], steps=20, u=levels(u))

    @xgrid.kernel(name="convection_2d_f32")
    def convection_2d32(u: f2s, c: float, dt: float, dx: float, dy: float) -> None:
        cdx = c * dt / dx
        cdy = c * dt / dy
        u[0, 0] = u[0, 0] + cdx * (u[0, 0] - u[-1, 0]) - cdy * (u[0, 0] - u[0, -1])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    u = xgrid.Grid((n2, n2), float)
    u.now.fill(1)
    u.now[int(.5 / dx2):int(1 / dx2) + 1, int(.5 / dx2):int(1 / dx2) + 1] = 2
    u.boundary[0, :] = u.boundary[:, 0] = 1
    u_in = u.now.copy()
    for _ in range(nt2):
        convection_2d32(u, 1.0, dt2, dx2, dx2)
    save("conv2d_f32", u_in=u_in, mask=u.boundary, params=[1.0, dt2, dx2, dx2], steps=nt2, u=levels(u))


if __name__ == "__main__":
    main()
