"""``import xgrid`` (the reference's module name) resolves to the B200 backend: same objects, the reference's
dotted module tree, and the scalar facts of the reference's test.py run as written (CPU part; the grid facts and
the cavity example run in tests/test_import_xgrid_gpu.py)."""
import os
import sys

import pytest

import test_import_xgrid_gpu as G


def test_alias_exports_the_backend_objects():
    import xgrid
    import xgrid_b200
    for name in ("kernel", "function", "init", "ptr", "grid", "boundary", "c", "external", "Grid", "shape",
                 "dimension", "tick"):                       # xgrid/__init__.py:52-53 of the reference
        assert getattr(xgrid, name) is getattr(xgrid_b200, name), name
    from xgrid.lang.operator import Operator
    from xgrid.util.console import Console
    from xgrid.util.ffi import Compiler, Library
    from xgrid.util.logging import Logger
    from xgrid.util.typing.value import Floating
    from xgrid.xgrid import Grid
    from xgrid_b200.log import Logger as L2
    assert Logger is L2 and Grid is xgrid_b200.Grid and Floating(8).width_bytes == 8
    with pytest.raises(Exception, match="gcc JIT"):
        Compiler(".x", ["gcc"])
    import io
    sink = io.StringIO()
    Console(sink).println("x")
    assert sink.getvalue() == "x\n"


def test_scalar_facts_of_test_py_run_as_written(tmp_path, monkeypatch):
    import types
    pyplot = G._Recorder("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot, mpl.cm = pyplot, types.ModuleType("matplotlib.cm")
    for k, v in (("matplotlib", mpl), ("matplotlib.pyplot", pyplot), ("matplotlib.cm", mpl.cm)):
        monkeypatch.setitem(sys.modules, k, v)
    monkeypatch.chdir(tmp_path)
    import xgrid
    piece = G._programs()["test_py_facts_3_to_10"]
    mod = G._load(tmp_path, "ref_test_facts_cpu", G.HARNESS + piece["text"])
    xgrid.init(comment=True, cacheroot=".xgridtest", opt_level=3, precision="double")
    facts = dict(mod.test.tests)
    for _ in range(5):
        facts["lang.Operator.simple"]()                      # asserts aux(a, b) == a + b + TEMP
        facts["lang.Operator.structure"]()                   # asserts aux(a, b) == a.dot(b)
    assert len(facts) == 8
