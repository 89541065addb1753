"""Randomized differential tests: random DSL programs (multiple grids, time levels 0..2,
offsets up to +-2, masked / implicit / looped statements) run on the CUDA path and on the
NumPy interpreter; every ring level of every grid must be bit-identical.  Shapes are chosen
to hit every kernel variant (dense, march, tiled, sparse, multistep, ghost re-layout)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import xgrid_b200 as xgrid
from oracle import HostGrid
from oracle.interp import Interp
from randprog import gen_inputs, gen_source, load_program

CASES = []
for seed in range(8):
    CASES.append((seed, 1, 2, [(37,), (4100,), (33000,)][seed % 3], False))
for seed in range(8, 20):
    CASES.append((seed, 2, 1 + seed % 3, [(17, 33), (40, 1030), (64, 1024), (9, 2050), (130, 66)][seed % 5], False))
for seed in range(20, 30):
    CASES.append((seed, 3, 1 + seed % 2, [(6, 7, 9), (18, 9, 130), (20, 16, 256), (5, 40, 64)][seed % 4], False))
for seed in range(30, 36):
    CASES.append((seed, 1, 1, [(20000,), (70001,), (16384,)][seed % 3], True))


@pytest.mark.parametrize("seed,ndim,ngrids,shape,single", CASES)
def test_random_program(tmp_path, seed, ndim, ngrids, shape, single):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    src = gen_source(seed, ndim, ngrids, single_1d=single)
    prog = load_program(src, str(tmp_path), f"randprog_{seed}")
    ics, masks = gen_inputs(seed, shape, ngrids)
    dev, host = [], []
    for ic, m in zip(ics, masks):
        g = xgrid.Grid(shape, float)
        g.now[...] = ic
        g.boundary[...] = m
        dev.append(g)
        h = HostGrid(shape)
        h.now[...] = ic
        h.boundary[...] = m
        host.append(h)
    ref = Interp(prog)
    a, b = 0.3, 1.7
    calls = 70 if single else 3
    for _ in range(calls):
        prog(*dev, a, b)
        ref(*host, a, b)
    for n, (g, h) in enumerate(zip(dev, host)):
        gd, hd = g._data, h._data
        assert len(gd) == len(hd), src
        for lvl, (x, y) in enumerate(zip(gd, hd)):
            if not np.array_equal(x, y, equal_nan=True):
                bad = np.argwhere(x != y)
                raise AssertionError(f"grid g{n} level {lvl}: {len(bad)} cells differ, first {bad[:4].tolist()}\n"
                                     f"shape={shape}\n{src}")


def test_every_kernel_variant_was_exercised():
    """Runs last in this file: the random programs above must have gone through every variant."""
    from xgrid_b200.lang.launch import STATS
    for variant in ("dense", "march", "tiled", "sparse", "multistep"):
        assert STATS.get(variant, 0) > 0, (variant, STATS)
