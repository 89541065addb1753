"""xgb_h2d_staged / xgb_d2h_staged (multi-threaded pageable <-> device copies, opt-in through
XGB_STAGED_COPY=1): byte-exact round trips at chunk-boundary sizes, and a Grid whose uploads / downloads
go through the staged path.  Correctness passed on a B200 with the last seconds of round 1's GPU budget;
the path has not been TIMED yet (scripts/e2e_phases.py with XGB_STAGED_COPY=1), so it stays off by default."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [1, 4 << 20, (4 << 20) + 8, 37 << 20, 129 << 20])
def test_staged_round_trip(tmp_path, nbytes):
    import xgrid_b200 as xgrid
    from xgrid_b200.runtime.shim import Runtime
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    rt = Runtime.get()
    src = np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8)
    dev = rt.alloc(nbytes)
    rt.h2d_staged(dev, src.ctypes.data, nbytes)
    src_copy = src.copy()
    src[:] = 0                                   # the call has consumed the source
    out = np.empty(nbytes, np.uint8)
    rt.d2h_staged(out.ctypes.data, dev, nbytes)
    assert np.array_equal(out, src_copy)
    ref = np.empty(nbytes, np.uint8)             # against the plain copy path
    rt.d2h(ref.ctypes.data, dev, nbytes)
    rt.sync()
    assert np.array_equal(ref, src_copy)
    rt.free(dev)


def test_grid_through_staged_path(tmp_path, monkeypatch):
    import xgrid_b200 as xgrid
    from examples import workloads as W
    from xgrid_b200.runtime.shim import Runtime
    import oracle
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    monkeypatch.setattr(Runtime, "STAGED", True)
    n = 1 << 21                                  # 16 MiB levels: above STAGED_MIN
    ic, dx = W.ic_1d(n)
    u, h = xgrid.Grid((n,), float), oracle.HostGrid((n,))
    u.now[...] = ic
    h.now[...] = ic
    u.boundary[0] = h.boundary[0] = 1
    k = W.make_kernels()["convection_1d"]
    for _ in range(70):
        k(u, 1.0, 0.5 * dx, dx)
        oracle.step_conv1d(h, 1.0, 0.5 * dx, dx)
    assert np.array_equal(u._data[0], h._data[0]) and np.array_equal(u._data[1], h._data[1])
