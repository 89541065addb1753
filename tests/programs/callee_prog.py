"""A kernel that calls other operators WITH grid parameters (and a scalar helper): the callees' stencil
statements run on the caller's buffers, without a tick of their own (xgrid/lang/generator.py:208-212,418-419).
Executed by the reference (tests/golden/make_callee_golden.py) and -- through the `import xgrid` alias -- by the
B200 backend (tests/test_callee_gpu.py).  Load it AFTER xgrid.init(): annotations resolve `float` at import."""
import xgrid

f2 = xgrid.grid[float, 2]


@xgrid.function()
def twice(a: float) -> float:
    return a * 2.0


@xgrid.kernel()
def relax(u: f2, a: float) -> None:
    u[0, 0] = u[0, 0] * a + 0.25 * (u[0, 1][2] + u[1, 0][2])
    with xgrid.boundary(1):
        u[0, 0] = 1.5


@xgrid.kernel()
def smooth(p: f2, q: f2) -> None:
    for _ in range(0, 3):
        p[0, 0] = 0.25 * (p[0, 1][0] + p[0, -1][0] + p[1, 0][0] + p[-1, 0][0]) + q[0, 0][0]
        with xgrid.boundary(2):
            p[0, 0] = p[0, 1][0]


@xgrid.kernel()
def outer(u: f2, v: f2, a: float) -> None:
    relax(u, a)
    v[0, 0] = u[0, 0][0] + 1.0
    with xgrid.boundary(1):
        v[0, 0] = 0.0
    relax(v, twice(a))
    smooth(u, v)
