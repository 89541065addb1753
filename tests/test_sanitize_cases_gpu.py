"""Small parity cases that drive every shared-memory pipeline variant once -- bulk-copy tiled (2-D and 3-D), two
steps per pass (tiled2), fused Jacobi pairs (jacobi2), multi-step (full, tail and short) -- sized for
`compute-sanitizer` (scripts/sanitize_gpu.sh runs this file under racecheck and memcheck); also part of the normal
GPU suite."""
import numpy as np
import pytest

import oracle
import xgrid_b200 as xgrid
from examples import workloads as W
from oracle import HostGrid

pytestmark = pytest.mark.gpu


@pytest.fixture()
def k(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    return W.make_kernels()


def _pair(shape, seed, mask):
    ic = np.random.default_rng(seed).random(shape)
    u, h = xgrid.Grid(shape, float), HostGrid(shape)
    for g in (u, h):
        g.now[...] = ic
        g.boundary[...] = mask
    return u, h


def test_tiled_2d_and_two_steps_per_pass(k):
    from xgrid_b200.lang.launch import STATS
    shape = (96, 2048)
    u, h = _pair(shape, 1, W.shell_mask(shape))
    before = dict(STATS)
    for _ in range(5):                       # deferred: two passes of tiled2 + one single (tiled) step
        k["diffusion_2d"](u, 0.2)
        oracle.step_diff2d(h, 0.2)
    assert np.array_equal(u.now, h.now) and np.array_equal(u._data[1], h._data[1])
    assert STATS.get("tiled2", 0) - before.get("tiled2", 0) == 2 and STATS.get("tiled", 0) > before.get("tiled", 0)


def test_tiled_3d(k):
    shape = (20, 16, 256)
    u, h = _pair(shape, 2, W.shell_mask(shape))
    for _ in range(2):
        k["heat_3d"](u, 0.1)
        oracle.step_heat3d(h, 0.1)
        assert np.array_equal(u.now, h.now)


def test_fused_jacobi_pairs(k):
    from xgrid_b200.lang.launch import STATS
    n = 640
    mb, mp, mu, mv = W.cavity_masks(n, n)
    dx = 2.0 / (n - 1)
    cfg = W.Config(1.0, 0.1, 1e-4 * (100.0 / (n - 1)) ** 2, dx, dx)
    gs = [xgrid.Grid((n, n), float) for _ in range(4)]
    hs = [HostGrid((n, n)) for _ in range(4)]
    rng = np.random.default_rng(3)
    for g, hh, m in zip(gs, hs, (mb, mp, mu, mv)):
        ic = 1e-3 * rng.random((n, n))
        for x in (g, hh):
            x.now[...] = ic
            x.boundary[...] = m
    before = STATS.get("jacobi2", 0)
    k["cavity_kernel"](*gs, cfg)
    oracle.step_cavity(*hs, oracle.Config(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy))
    for g, hh in zip(gs, hs):
        assert np.array_equal(g.now, hh.now)
    assert STATS.get("jacobi2", 0) - before == 24             # 50 sweeps = 24 fused pairs + 2 single sweeps


@pytest.mark.parametrize("steps", [20, 40, 64])
def test_multistep_variants(k, steps):
    n = 20000
    ic, dx = W.ic_1d(n)
    mask = np.zeros(n, np.int32)
    mask[0] = mask[-1] = 1
    mask[1984] = 7
    u, h = xgrid.Grid((n,), float), HostGrid((n,))
    for g in (u, h):
        g.now[...] = ic
        g.boundary[...] = mask
    args = (0.01, 0.2 * dx * dx / 0.01, dx)
    for _ in range(steps):
        k["diffusion_1d"](u, *args)
        oracle.step_diff1d(h, *args)
    assert np.array_equal(u._data[0], h._data[0]) and np.array_equal(u._data[1], h._data[1])
