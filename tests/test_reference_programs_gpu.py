"""The reference's own user programs, run unchanged except for the import line."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_readme_program(tmp_path):
    """README.md:17-58 verbatim (default init() => fp32), with the reference's
    actual post-conditions (F4): every argument was ticked, so the inputs sit one level back."""
    import xgrid_b200 as xgrid
    import numpy as np

    xgrid.init(cacheroot=str(tmp_path))

    fvec = xgrid.grid[float, 1]

    @xgrid.kernel()
    def elementwise_mul(result: fvec, a: fvec, b: fvec) -> None:
        result[0] = a[0] * b[0]

    a = xgrid.Grid((10000, ), float)
    b = xgrid.Grid((10000, ), float)

    for i in range(10000):
        a[i] = random.random()
        b[i] = random.random()

    result = xgrid.Grid((10000, ), float)
    a_in, b_in = a.now.copy(), b.now.copy()

    elementwise_mul(result, a, b)

    assert result.now.dtype == np.float32
    assert np.array_equal(result.now, a_in * b_in)
    assert not a.now.any() and np.array_equal(a._data[1], a_in)      # F4: a.now is the zero buffer now
    # the README's closing assertion compares against the ticked (zero) inputs in the reference too
    assert np.sum(result.now) != np.dot(a.now, b.now)


def test_test_py_grid_kernels(tmp_path):
    """test.py:168-192 (int grid fill, index guard) and :195-221 with its two-argument boundary form."""
    import xgrid_b200 as xgrid
    xgrid.init(comment=True, cacheroot=str(tmp_path), opt_level=3, precision="double")

    @xgrid.kernel()
    def aux(a: xgrid.grid[int, 2]) -> None:
        a[0, 0] = 4

    grid = xgrid.Grid((10, 10), dtype=int)
    aux(grid)
    for row in grid.now:
        for element in row:
            assert element == 4

    @xgrid.kernel()
    def guard(a: xgrid.grid[int, 2]) -> None:
        a[0, 0] = a[-1, -1][-1]

    grid = xgrid.Grid((10, 10), dtype=int)
    guard(grid)

    float1d = xgrid.grid[float, 1]
    nx = 41
    dx = 2 / (nx - 1)
    u = xgrid.Grid((nx,), float)
    u.now.fill(1)
    u.now[int(.5 / dx):int(1 / dx + 1)] = 2
    u.boundary[0] = 1

    @xgrid.kernel()
    def convection_1d(u: float1d, c: float, dt: float, dx: float) -> None:
        u[0] = u[0] - c * dt / dx * (u[0] - u[-1])
        with xgrid.boundary(u, 1):
            u[0] = 1.0

    un = u.now.copy()
    for _ in range(25):
        convection_1d(u, 1, .025, dx)
        new = un.copy()
        new[1:] = un[1:] - 1 * .025 / dx * (un[1:] - un[:-1])
        new[0] = 1.0
        un = new
    assert np.array_equal(u.now, un)
