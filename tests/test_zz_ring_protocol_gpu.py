"""Ring protocol on the CUDA path against the UNMODIFIED reference's outputs (tests/golden/ring_protocol.npz,
scenarios in tests/ring_cases.py): ring depth and every level after every call.  The same scenarios are
replayed through the interpreter on CPU (tests/test_interp.py); the GPU edge-case tests cover them against
the interpreter.  (Named to run last: added after round 1's GPU budget was spent.)"""
import importlib.util
import os

import numpy as np
import pytest

import xgrid_b200 as xgrid

import ring_cases as RC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scenario", sorted(RC.SCENARIOS))
def test_ring_protocol_against_reference_outputs(tmp_path, scenario):
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ring_protocol.npz"))
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    path = tmp_path / "ring_kernels.py"
    path.write_text(RC.SOURCE.replace("IMPORT_LINE", "import xgrid_b200 as xgrid"))
    spec = importlib.util.spec_from_file_location("ring_kernels_gpu", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    grids = {}
    for which in "gh":
        ic, mask = RC.initial(np, which)
        g = xgrid.Grid((RC.N,), float)
        g.now[...] = ic
        g.boundary[...] = mask
        grids[which] = g
    for step, (kernel, spec_) in enumerate(RC.SCENARIOS[scenario]):
        getattr(mod, kernel)(*RC.arguments(spec_, grids))
        for which, g in grids.items():
            levels = g._data
            assert len(levels) == int(gold[f"{scenario}.{step}.{which}.depth"]), (step, kernel, which)
            for lvl, arr in enumerate(levels):
                want = gold[f"{scenario}.{step}.{which}.L{lvl}"]
                assert np.array_equal(arr, want, equal_nan=True), (scenario, step, kernel, which, lvl)
