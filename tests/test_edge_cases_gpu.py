"""Edge cases of the ring / call semantics, CUDA path vs the
NumPy interpreter on HostGrid (which restates xgrid/xgrid/__init__.py:43-54 literally)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import xgrid_b200 as xgrid
from oracle import HostGrid
from oracle.interp import Interp


@pytest.fixture()
def x64(tmp_path):
    xgrid.init(precision="double", cacheroot=str(tmp_path))
    return xgrid


def pair(shape, seed=0, mask=None):
    ic = np.random.default_rng(seed).random(shape)
    g, h = xgrid.Grid(shape, float), HostGrid(shape)
    g.now[...] = ic
    h.now[...] = ic
    if mask is not None:
        g.boundary[...] = mask
        h.boundary[...] = mask
    return g, h


def same(g, h):
    gd, hd = g._data, h._data
    assert len(gd) == len(hd)
    for x, y in zip(gd, hd):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("shape", [(1,), (2,), (3,), (1, 1), (1, 7), (7, 1), (2, 2), (1, 1, 1), (1, 5, 2), (3, 1, 4)])
def test_tiny_and_unit_axis_grids(x64, shape):
    nd = len(shape)
    G = xgrid.grid[float, nd]
    zero = ", ".join("0" for _ in shape)
    left = ", ".join("-1" if a == nd - 1 else "0" for a in range(nd))
    up = ", ".join("1" if a == 0 else "0" for a in range(nd))
    src = (f"def k(u: G, a: float) -> None:\n"
           f"    u[{zero}] = u[{zero}] + a * (u[{left}] - u[{up}])\n"
           f"    with xgrid.boundary(1):\n"
           f"        u[{zero}] = 2.0\n")
    import os, importlib.util, tempfile
    d = tempfile.mkdtemp()
    path = os.path.join(d, "tiny_mod.py")
    with open(path, "w") as f:
        f.write("import xgrid_b200 as xgrid\nG = xgrid.grid[float, %d]\n@xgrid.kernel()\n" % nd + src)
    spec = importlib.util.spec_from_file_location("tiny_mod", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mask = np.zeros(shape, np.int32)
    mask.reshape(-1)[0] = 1
    g, h = pair(shape, 1, mask)
    ref = Interp(mod.k)
    for _ in range(3):
        mod.k(g, 0.25)
        ref(h, 0.25)
    same(g, h)


def test_ring_depth_changes_between_kernels(x64):
    """A depth-3 kernel extends the ring, a depth-2 kernel truncates it again
    (xgrid/xgrid/__init__.py:43-47), a depth-1 kernel leaves one level."""
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def deep(u: f1) -> None:
        u[0] = 0.5 * u[0] + 0.25 * u[1][2] + 0.25 * u[-1][-2]

    @xgrid.kernel()
    def shallow(u: f1) -> None:
        u[0] = u[0] * 0.5 + u[-1]

    @xgrid.kernel()
    def flat(u: f1) -> None:
        u[0] = 3.0

    g, h = pair((300,), 2)
    rd, rs, rf = Interp(deep), Interp(shallow), Interp(flat)
    for op, ref in ((deep, rd), (deep, rd), (shallow, rs), (deep, rd), (flat, rf), (shallow, rs), (deep, rd)):
        op(g)
        ref(h)
        same(g, h)


def test_same_grid_passed_twice_ticks_twice(x64):
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def both(a: f1, b: f1) -> None:
        a[0] = b[0] + 1.0

    g, h = pair((64,), 3)
    ref = Interp(both)
    for _ in range(3):
        both(g, g)          # operator.py:37-39 ticks once per ARGUMENT
        ref(h, h)
        same(g, h)


def test_tick_false(x64):
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel(tick=False)
    def inplace(u: f1, a: float) -> None:
        u[0] = u[0][0] * a + u[0]        # resizes the ring to depth 2 but does not rotate it

    g, h = pair((128,), 4)
    ref = Interp(inplace)
    for _ in range(3):
        inplace(g, 0.5)
        ref(h, 0.5)
        same(g, h)


def test_mismatched_shapes_raise(x64):
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def add(a: f1, b: f1) -> None:
        a[0] = b[0]

    with pytest.raises(Exception, match="shape"):
        add(xgrid.Grid((10,), float), xgrid.Grid((12,), float))
    with pytest.raises(TypeError):
        add(xgrid.Grid((10,), float), xgrid.Grid((10, 2), float))


def test_fill_and_getitem_setitem(x64):
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def shift(u: f1) -> None:
        u[0] = u[-1]

    u = xgrid.Grid((16,), float)
    u.fill(np.arange(16, dtype=np.float64), 0)
    shift(u)
    assert u[3] == 2.0 and u[0] == 0.0          # u[-1] of point 0 reads the zero ghost
    u[5] = 42.0
    shift(u)
    assert u[6] == 42.0
    assert np.array_equal(u._data[1][:5], [0, 0, 1, 2, 3])


def test_mask_change_between_calls(x64):
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def relax(u: f2) -> None:
        u[0, 0] = 0.25 * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    g, h = pair((40, 48), 5)
    ref = Interp(relax)
    for step in range(6):
        if step == 2:
            g.boundary[0, :] = 1
            h.boundary[0, :] = 1
        if step == 4:
            g.boundary[:, -1] = 1
            h.boundary[:, -1] = 1
            g.boundary[0, :] = 0
            h.boundary[0, :] = 0
        relax(g)
        ref(h)
        same(g, h)


def test_writes_through_arrays_kept_across_calls(x64):
    """`b = g.boundary` / `a = g.now` kept by the program and written BETWEEN kernel calls (plain attributes in the
    reference, xgrid/xgrid/__init__.py:38-41,70-72): the next call must see the writes; a deferred 1-D run queued
    before a mask write still runs with the mask it was called with."""
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def relax(u: f2) -> None:
        u[0, 0] = 0.25 * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0])
        with xgrid.boundary(1):
            u[0, 0] = 1.0

    g, h = pair((40, 48), 9)
    ref = Interp(relax)
    b, hb = g.boundary, h.boundary            # kept across the calls
    relax(g)
    ref(h)
    a = g.now                                 # kept: mirrors ring level 0 now, level 1 after the next call
    ha = h.now
    for step in range(5):
        if step == 1:
            b[3, :] = 1                       # through the kept arrays, never touching g.boundary again
            hb[3, :] = 1
        if step == 2:
            b += 0                            # in-place operator without a change
            np.copyto(b, np.where(hb == 1, 1, 0).astype(np.int32))
        if step == 3:
            a[5:9, 5:9] = 7.0                 # the kept .now array: in the reference it IS a ring level
            ha[5:9, 5:9] = 7.0
        relax(g)
        ref(h)
        same(g, h)
    # 1-D deferred run: 12 queued calls see the old mask, the following ones the new mask
    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def shift(u: f1, c: float) -> None:
        u[0] = u[0] - c * (u[0] - u[-1])
        with xgrid.boundary(1):
            u[0] = 1.0

    g1, h1 = pair((1 << 15,), 4)
    ref1 = Interp(shift)
    b1 = g1.boundary
    b1[0] = 1
    h1.boundary[0] = 1
    for _ in range(12):
        shift(g1, 0.5)
        ref1(h1, 0.5)
    b1[1000] = 1                              # flushes the 12 queued calls first
    h1.boundary[1000] = 1
    for _ in range(9):
        shift(g1, 0.5)
        ref1(h1, 0.5)
    same(g1, h1)


def test_element_indexing_on_device_level(x64):
    """`u[i]` / `u[i] = v` on a device-resident level move one element and keep the level on
    the device; slices still hand the level to the host."""
    f2 = xgrid.grid[float, 2]

    @xgrid.kernel()
    def bump(u: f2) -> None:
        u[0, 0] = u[0, 0] + 1.0

    u = xgrid.Grid((6, 5), float)
    u.now[...] = np.arange(30.0).reshape(6, 5)
    bump(u)
    assert u._ring[0].where == "device"
    assert u[2, 3] == 14.0 and u[-1, -1] == 30.0 and u[(0, 0)] == 1.0
    assert u._ring[0].where == "device"
    u[2, 3] = -5.0
    assert u._ring[0].where == "device"
    bump(u)
    assert u[2, 3] == -4.0                      # loads read the level the element write went to
    assert u._data[1][2, 3] == -5.0
    with pytest.raises(IndexError):
        u[6, 0]
    assert u[1].shape == (5,)                   # row slice -> NumPy view of the host mirror
