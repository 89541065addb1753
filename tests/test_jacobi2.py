"""Two solver iterations per pass (xgrid_b200/lang/jacobi2.py).

CPU part: the pattern matcher, the host-side chain check and a NumPy model of the in-kernel
boundary resolution, compared with two statement-at-a-time iterations (oracle/interp.py).
GPU part (-m gpu): the fused kernel against the interpreter / the cavity oracle, bit-exact."""
import numpy as np
import pytest

import xgrid_b200 as xgrid
from examples import workloads as W
from xgrid_b200.lang import jacobi2
from xgrid_b200.lang.schedule import Program
from oracle import HostGrid
from oracle.interp import Interp


def make_solvers():
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def solver5(p: f2, b: f2, w: float, n: int) -> None:
        for _ in range(0, n):
            p[0, 0] = w * (p[0, 1][0] + p[0, -1][0] + p[1, 0][0] + p[-1, 0][0]) - 0.01 * b[0, 0]
            with xgrid.boundary(1):
                p[0, 0] = p[0, -1][0]
            with xgrid.boundary(2):
                p[0, 0] = p[1, 0][0]
            with xgrid.boundary(3):
                p[0, 0] = p[0, 1][0]
            with xgrid.boundary(4):
                p[0, 0] = 0.5
            with xgrid.boundary(5):
                p[0, 0] = p[-1, 0][0]
            with xgrid.boundary(6):
                p[0, 0] = p[1, 1][0]

    @xgrid.kernel()
    def solver9(p: f2, w: float, n: int) -> None:
        for _ in range(0, n):
            p[0, 0] = w * (p[0, 1][0] + p[0, -1][0] + p[1, 0][0] + p[-1, 0][0]) + \
                0.5 * w * (p[1, 1][0] + p[-1, -1][0] + p[1, -1][0] + p[-1, 1][0]) - w * p[0, 0][0]
            with xgrid.boundary(2):
                p[0, 0] = w
            with xgrid.boundary(1):
                p[0, 0] = p[-1, -1][0]

    @xgrid.kernel()
    def not_a_pair(p: f2, w: float, n: int) -> None:
        for i in range(0, n):
            p[0, 0] = w * (p[0, 2][0] + p[0, -1][0])          # halo 2: stays step-at-a-time
            with xgrid.boundary(1):
                p[0, 0] = 1.0

    return solver5, solver9, not_a_pair


def shell_geometry(shape, rng, obstacles: int, clean: bool = True):
    """Cavity-like edge values 1..4 plus isolated interior patterns (single points and dominoes of
    two different values) on a coarse lattice, incl. value 9 that no statement handles (F5).
    clean=True keeps to patterns whose copy chains end at points no statement writes (the fused
    pass applies); clean=False also draws dominoes where a statement reads a point that a LATER
    statement writes (the launcher must fall back to step-at-a-time)."""
    n0, n1 = shape
    m = np.zeros(shape, np.int32)
    m[:, -1] = 1
    m[0, :] = 2
    m[:, 0] = 3
    m[-1, :] = 4
    vertical = [(2, 1), (2, 9), (9, 5), (1, 5), (3, 5), (9, 2)] if clean else [(2, 4), (1, 2), (2, 5), (5, 3)]
    horizontal = [(9, 1), (3, 9), (6, 9), (1, 3), (5, 2)] if clean else [(3, 1), (2, 1), (3, 4), (1, 3)]
    cells = [(r, c) for r in range(4, n0 - 4, 4) for c in range(4, n1 - 4, 4)]
    for idx in rng.choice(len(cells), size=min(obstacles, len(cells)), replace=False):
        r, c = cells[idx]
        kind = rng.integers(0, 4)
        if kind == 0:
            m[r, c] = rng.choice([1, 2, 3, 4, 5, 6, 9])
        elif kind == 1:
            m[r, c], m[r + 1, c] = vertical[rng.integers(len(vertical))]
        elif kind == 2:
            m[r, c], m[r, c + 1] = horizontal[rng.integers(len(horizontal))]
        else:
            m[r, c] = 4
    return m


# --------------------------------------------------------------------------- NumPy model of the fused pass
def model_pair(pair, mask, p0, sweep_rhs):
    """What the fused kernel computes for one pair, written per point: middle state R (sweep A,
    boundary points keep their value), boundary resolution on R, sweep B on resolved taps."""
    n0, n1 = mask.shape
    flat = mask.reshape(-1)
    total = flat.size
    rules = {r.mask: (j, r) for j, r in enumerate(pair.rules)}

    def at(arr, y):
        return arr[y] if 0 <= y < total else 0.0

    taps = [(0, 1), (0, -1), (1, 0), (-1, 0)]
    R = p0.reshape(-1).copy()
    for y in range(total):
        if flat[y] == 0:
            R[y] = sweep_rhs([at(p0.reshape(-1), y + d0 * n1 + d1) for d0, d1 in taps], y)

    def resolve(y):
        limit = len(pair.rules)
        while True:
            m = flat[y] if 0 <= y < total else 255
            hit = rules.get(int(m))
            if hit is None or hit[0] >= limit:
                return at(R, y)
            j, r = hit
            if r.const is not None:
                return float(r.const.value)
            y += r.offset[0] * n1 + r.offset[1]
            limit = j

    out = np.empty(total)
    for y in range(total):
        if flat[y] == 0:
            out[y] = sweep_rhs([resolve(y + d0 * n1 + d1) for d0, d1 in taps], y)
        else:
            out[y] = resolve(y)
    return out.reshape(mask.shape)


@pytest.fixture(scope="module")
def solvers(tmp_path_factory):
    xgrid.init(precision="double", cacheroot=str(tmp_path_factory.mktemp("xgj2")))
    return make_solvers()


def test_match_and_emit(solvers):
    s5, s9, nap = solvers
    p5, p9, pn = Program(s5), Program(s9), Program(nap)
    assert len(p5.pairs) == 1 and len(p9.pairs) == 1 and not pn.pairs
    pair = next(iter(p5.pairs.values()))
    assert [(r.mask, r.offset, r.const is not None) for r in pair.rules] == [
        (1, (0, -1), False), (2, (1, 0), False), (3, (0, 1), False), (4, (0, 0), True),
        (5, (-1, 0), False), (6, (1, 1), False)]
    assert pair.extras == [("b", 1)]
    assert "_jacobi2_v2" in p5.source and "cp.async.bulk" not in p5.source      # primitives live in the header
    assert p5.image()[:4] == b"\x7fELF" and p9.image()[:4] == b"\x7fELF"          # NVRTC sm_100a
    cav = Program(W.make_kernels()["cavity_kernel"])
    assert len(cav.pairs) == 1


def test_chains_fit(solvers):
    pair = next(iter(Program(solvers[0]).pairs.values()))
    for seed in range(6):
        assert jacobi2.chains_fit(pair, shell_geometry((40, 64), np.random.default_rng(seed), 30))
    assert jacobi2.chains_fit(pair, W.cavity_masks(64, 64)[1])
    unclean = 0
    for seed in range(6):
        unclean += not jacobi2.chains_fit(pair, shell_geometry((40, 64), np.random.default_rng(seed), 30, clean=False))
    assert unclean >= 5
    ok = np.zeros((40, 64), np.int32)
    ok[10, 10], ok[11, 10] = 2, 1                  # 2 copies from below, a point statement 1 (earlier) wrote
    ok[20, 20], ok[21, 20] = 9, 5                  # 5 copies from a point nothing writes
    assert jacobi2.chains_fit(pair, ok)
    late = np.zeros((40, 64), np.int32)
    late[10, 10], late[10, 9] = 1, 3               # 1 copies from the left, a point statement 3 (LATER) writes
    assert not jacobi2.chains_fit(pair, late)
    far = np.zeros((40, 64), np.int32)
    far[5, 5], far[6, 6], far[7, 6] = 6, 2, 1      # 6 -> (6,6) -> (7,6) -> (7,5): two rows away from the start
    assert not jacobi2.chains_fit(pair, far)
    assert not jacobi2.chains_fit(pair, np.ones((40, 64), np.int32))


def test_model_matches_two_iterations(solvers):
    s5 = solvers[0]
    pair = next(iter(Program(s5).pairs.values()))
    rng = np.random.default_rng(3)
    shape = (32, 40)
    mask = shell_geometry(shape, rng, 14)
    assert jacobi2.chains_fit(pair, mask)
    p0 = rng.random(shape)
    b0 = rng.random(shape)
    w = 0.23
    hp, hb = HostGrid(shape), HostGrid(shape)
    hp.now[...] = p0
    hb.now[...] = b0
    hp.boundary[...] = mask
    Interp(s5)(hp, hb, w, 2)
    # the call ticked both grids: p's level 0 starts as the rotated-in zero buffer
    start = np.zeros(shape)
    bflat = b0.reshape(-1)

    def rhs(t, y):
        return w * (t[0] + t[1] + t[2] + t[3]) - 0.01 * bflat[y]

    fused = model_pair(pair, mask, start, rhs)
    # boundary statements of the second iteration, statement at a time
    flat = fused.reshape(-1)
    for r in pair.rules:
        idx = np.flatnonzero(mask.reshape(-1) == r.mask)
        if r.const is not None:
            flat[idx] = float(r.const.value)
        else:
            src = idx + r.offset[0] * shape[1] + r.offset[1]
            flat[idx] = flat[src]
    assert np.array_equal(fused, hp.now)


# --------------------------------------------------------------------------- GPU
def fused_passes(trips: int) -> int:
    """Fused passes per call: pairs of iterations, kept even when the trip count is even (so the number of
    level-0 <-> scratch swaps per call stays even and the CUDA-graph arrangement repeats every 2nd call)."""
    pairs = trips // 2
    return pairs - 1 if (pairs % 2 == 1 and trips % 2 == 0) else pairs


def _run_both(kernel, shape, mask, scalars, calls, with_b, seed):
    rng = np.random.default_rng(seed)
    ics = [rng.random(shape)] + ([rng.random(shape)] if with_b else [])
    dev, host = [], []
    for n, ic in enumerate(ics):
        g = xgrid.Grid(shape, float)
        g.now[...] = ic
        h = HostGrid(shape)
        h.now[...] = ic
        if n == 0:
            g.boundary[...] = mask
            h.boundary[...] = mask
        dev.append(g)
        host.append(h)
    ref = Interp(kernel)
    for _ in range(calls):
        kernel(*dev, *scalars)
        ref(*host, *scalars)
    for n, (g, h) in enumerate(zip(dev, host)):
        for lvl, (x, y) in enumerate(zip(g._data, h._data)):
            if not np.array_equal(x, y, equal_nan=True):
                bad = np.argwhere(x != y)
                raise AssertionError(f"grid {n} level {lvl}: {len(bad)} cells differ, first {bad[:6].tolist()}, "
                                     f"mask there {[int(mask[tuple(b)]) for b in bad[:6]]}")


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n,obstacles,seed,clean", [
    ((96, 640), 6, 0, 1, True), ((130, 1024), 7, 40, 2, True), ((64, 516), 5, 25, 3, True),
    ((200, 1536), 4, 300, 4, True), ((67, 512), 3, 10, 5, True), ((140, 1100), 6, 200, 6, False),
])
def test_fused_pairs_random_geometry(solvers, shape, n, obstacles, seed, clean):
    from xgrid_b200.lang.launch import STATS
    s5 = solvers[0]
    rng = np.random.default_rng(seed)
    mask = shell_geometry(shape, rng, obstacles, clean)
    before = STATS.get("jacobi2", 0)
    _run_both(s5, shape, mask, (0.23, n), 3, True, seed)
    assert STATS.get("jacobi2", 0) - before == (3 * fused_passes(n) if clean else 0)


@pytest.mark.gpu
def test_fused_pairs_nine_point(solvers):
    from xgrid_b200.lang.launch import STATS
    s9 = solvers[1]
    shape = (150, 768)
    m = np.zeros(shape, np.int32)
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 2
    m[1:-1, 1] = 1                      # second column copies from the upper-left neighbour (a value-2 point)
    before = STATS.get("jacobi2", 0)
    _run_both(s9, shape, m, (0.11, 8), 2, False, 9)
    assert STATS.get("jacobi2", 0) - before == 8


@pytest.mark.gpu
def test_fallback_when_chains_do_not_fit(solvers):
    from xgrid_b200.lang.launch import STATS
    s5 = solvers[0]
    shape = (80, 640)
    m = shell_geometry(shape, np.random.default_rng(7), 0)
    m[5, 5], m[6, 6], m[7, 6] = 6, 2, 1            # a copy chain that leaves the on-chip window
    before = STATS.get("jacobi2", 0)
    _run_both(s5, shape, m, (0.23, 4), 2, True, 7)
    assert STATS.get("jacobi2", 0) == before


@pytest.mark.gpu
@pytest.mark.parametrize("n", [640, 1024])
def test_cavity_fused(n):
    """examples/cavity.py with the 50-sweep pressure loop running 25 fused passes per timestep."""
    import oracle
    from xgrid_b200.lang.launch import STATS
    xgrid.init(precision="double", cacheroot=".xgridtest")
    kern = W.make_kernels()["cavity_kernel"]
    masks = W.cavity_masks(n, n)
    rng = np.random.default_rng(n)
    ics = [np.zeros((n, n)), np.zeros((n, n)), 0.01 * rng.random((n, n)), 0.01 * rng.random((n, n))]
    cfg = W.Config(1.0, 0.1, 1e-4 * (100.0 / (n - 1)) ** 2, 2.0 / (n - 1), 2.0 / (n - 1))
    dev, host = [], []
    for ic, m in zip(ics, masks):
        g = xgrid.Grid((n, n), float)
        g.now[...] = ic
        g.boundary[...] = m
        h = HostGrid((n, n))
        h.now[...] = ic
        h.boundary[...] = m
        dev.append(g)
        host.append(h)
    before = STATS.get("jacobi2", 0)
    steps = 3
    for _ in range(steps):
        kern(*dev, cfg)
        oracle.step_cavity(*host, oracle.Config(cfg.rho, cfg.nu, cfg.dt, cfg.dx, cfg.dy))
    for name, g, h in zip("bpuv", dev, host):
        for lvl, (x, y) in enumerate(zip(g._data, h._data)):
            assert np.array_equal(x, y, equal_nan=True), f"{name} level {lvl}: {int((x != y).sum())} cells differ"
    # the first two calls run step-at-a-time launches (the second one records the CUDA graph)
    assert STATS.get("jacobi2", 0) - before >= fused_passes(50) * 2
