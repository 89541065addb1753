"""Operators called from a kernel: callees WITH grid parameters run their own sweeps on the
caller's buffers -- golden produced by the unmodified reference (tests/golden/make_callee_golden.py) -- and
`@xgrid.external` operators resolve to `__device__` functions of a CUDA header named by `includes=`."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def load_callee_program(tmp_path):
    import xgrid
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    spec = importlib.util.spec_from_file_location("callee_prog_b200", os.path.join(ROOT, "tests", "programs", "callee_prog.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return xgrid, mod


def test_callees_with_grid_parameters_match_the_reference(tmp_path, golden):
    xgrid, mod = load_callee_program(tmp_path)
    g = golden("callee_f64")
    u, v = xgrid.Grid(g["mask"].shape, float), xgrid.Grid(g["mask"].shape, float)
    u.now[...] = g["u_in"]
    v.now[...] = g["v_in"]
    u.boundary[...] = g["mask"]
    v.boundary[...] = g["mask"]
    for _ in range(int(g["steps"])):                           # direct, recorded and replayed calls
        mod.outer(u, v, float(g["a"]))
    assert mod.outer.depth == int(g["depth"]) == 3            # the callee's [2] loads set the caller's ring depth
    for name, grid in (("u", u), ("v", v)):
        levels = grid._data
        assert len(levels) == 3
        for l, arr in enumerate(levels):
            assert np.array_equal(arr, g[f"{name}.L{l}"]), f"{name}.L{l}"


def test_external_operator_is_a_device_function_from_a_cuda_header(tmp_path, monkeypatch):
    import xgrid_b200 as xgrid
    monkeypatch.chdir(tmp_path)
    (tmp_path / "userlib").mkdir()
    (tmp_path / "userlib" / "bump.h").write_text(
        "__device__ __forceinline__ double bump(double x, double k) { return x * x + k; }\n")
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))

    def _check(args):
        return args[0]

    @xgrid.external(includes=["userlib/bump.h"], typecheck_override=_check)
    def bump(x: float, k: float) -> float:
        ...

    f1 = xgrid.grid[float, 1]

    @xgrid.kernel()
    def apply(u: f1, k: float) -> None:
        u[0] = bump(u[0], k) + u[-1]

    n = 5000
    u = xgrid.Grid((n,), float)
    x = np.random.default_rng(1).random(n)
    u.now[...] = x
    apply(u, 0.5)
    want = x * x + 0.5
    want[1:] = want[1:] + x[:-1]
    assert np.array_equal(u.now, want)
    assert '#include "userlib/bump.h"' in apply.src
