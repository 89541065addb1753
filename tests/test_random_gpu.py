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
    CASES.append((seed, 3, 1 + seed % 2, [(6, 7, 9), (18, 9, 130), (20, 16, 256), (17, 13, 300)][seed % 4], False))
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


# inserted before the coverage test at collection time? no: pytest runs tests in file order, and the
# coverage assertion above only needs the fp64 cases.  fp32 / int cases follow.
FP32_CASES = [(100 + s, nd, 1 + s % 2, shp) for s, (nd, shp) in enumerate([
    (1, (4100,)), (1, (33000,)), (2, (40, 1032)), (2, (17, 36)), (3, (18, 9, 132)), (3, (6, 7, 12))])]


@pytest.mark.parametrize("seed,ndim,ngrids,shape", FP32_CASES)
def test_random_program_fp32(tmp_path, seed, ndim, ngrids, shape):
    """precision="float": grids and `float` scalars are fp32, literals stay double, so every
    expression has C's mixed-precision typing; results must still be bit-identical."""
    xgrid.init(cacheroot=str(tmp_path / "xg"))
    src = gen_source(seed, ndim, ngrids)
    prog = load_program(src, str(tmp_path), f"randprog32_{seed}")
    ics, masks = gen_inputs(seed, shape, ngrids)
    dev, host = [], []
    for ic, m in zip(ics, masks):
        g = xgrid.Grid(shape, float)
        assert g.now.dtype == np.float32
        g.now[...] = ic.astype(np.float32)
        g.boundary[...] = m
        dev.append(g)
        h = HostGrid(shape, np.float32)
        h.now[...] = ic.astype(np.float32)
        h.boundary[...] = m
        host.append(h)
    ref = Interp(prog)
    for _ in range(3):
        prog(*dev, 0.3, 1.7)
        ref(*host, 0.3, 1.7)
    for n, (g, h) in enumerate(zip(dev, host)):
        for lvl, (x, y) in enumerate(zip(g._data, h._data)):
            assert x.dtype == np.float32
            if not np.array_equal(x, y, equal_nan=True):
                bad = np.argwhere(x != y)
                raise AssertionError(f"fp32 grid g{n} level {lvl}: {len(bad)} cells differ, first {bad[:4].tolist()}\n{src}")


def test_int_grid_stencil(tmp_path):
    """int grids: C integer arithmetic (truncating division, dividend-signed remainder) in sweeps."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    i2 = xgrid.grid[int, 2]

    @xgrid.kernel()
    def life_like(a: i2, k: int) -> None:
        a[0, 0] = (a[0, 1] + a[0, -1] + a[1, 0] + a[-1, 0] - 4 * a[0, 0]) / k + a[0, 0] % 7
        with xgrid.boundary(1):
            a[0, 0] = 0 - 3

    shape = (33, 130)
    rng = np.random.default_rng(5)
    ic = rng.integers(-1000, 1000, shape).astype(np.int32)
    m = np.zeros(shape, np.int32)
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 1
    g = xgrid.Grid(shape, int)
    g.now[...] = ic
    g.boundary[...] = m
    h = HostGrid(shape, np.int32)
    h.now[...] = ic
    h.boundary[...] = m
    ref = Interp(life_like)
    for _ in range(4):
        life_like(g, 3)
        ref(h, 3)
    assert np.array_equal(g._data[0], h._data[0]) and np.array_equal(g._data[1], h._data[1])


# ---- the same generator, but the expected values come from the UNMODIFIED reference itself
# (tests/golden/randprog.npz, written by tests/golden/make_random_golden.py): CUDA path vs reference,
# no interpreter in between.  Every third stored program, and every second one of the overstep="wrap" /
# "limit" sets (all of them are replayed through the interpreter on CPU in tests/test_interp.py).
def _reference_cases():
    import os
    out = []
    for mode, name, step in (("none", "randprog", 3), ("wrap", "randprog_wrap", 2), ("limit", "randprog_limit", 2)):
        data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        for seed in sorted({int(k.split(".")[0]) for k in data.files})[::step]:
            out.append((mode, seed, data))
    return out


_REFCASES = _reference_cases()


@pytest.mark.parametrize("mode,seed", [(m, s) for m, s, _ in _REFCASES])
def test_random_program_against_reference_outputs(tmp_path, mode, seed):
    from randprog import guard_array_ends
    data = next(d for m, s, d in _REFCASES if (m, s) == (mode, seed))
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"), overstep=mode)
    ndim, ngrids, single, *shape = (int(x) for x in data[f"{seed}.meta"])
    shape = tuple(shape)
    src = gen_source(seed, ndim, ngrids, single_1d=bool(single))
    assert src == str(data[f"{seed}.src"]), "tests/randprog.py changed: regenerate tests/golden/randprog*.npz"
    prog = load_program(src, str(tmp_path), f"randprog_ref_{seed}")
    ics, masks = gen_inputs(seed, shape, ngrids)
    if mode == "none":
        guard_array_ends(masks, shape)      # out-of-array taps are undefined in the reference
    dev = []
    for ic, m in zip(ics, masks):
        g = xgrid.Grid(shape, float)
        g.now[...] = ic
        g.boundary[...] = m
        dev.append(g)
    for _ in range(3):
        prog(*dev, 0.3, 1.7)
    for n, g in enumerate(dev):
        got = g._data
        assert len(got) == int(data[f"{seed}.g{n}.depth"])
        for lvl, x in enumerate(got):
            want = data[f"{seed}.g{n}.L{lvl}"]
            assert np.array_equal(x, want, equal_nan=True), (f"g{n} level {lvl}: "
                                                             f"{int((x != want).sum())} cells differ\n{src}")
