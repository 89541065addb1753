"""Front-end conformance against the UNMODIFIED reference (CPU only): every case of tests/frontend_cases.py
was run through the reference's parser + generator (tests/golden/make_frontend_golden.py ->
tests/golden/frontend_conformance.json); this backend must accept / reject the same programs, derive the same
ring depth, and return the same value from scalar-only kernels (evaluated on the host with C semantics)."""
import dataclasses
import importlib.util
import json
import os

import pytest

import xgrid_b200 as xgrid
from xgrid_b200.lang.schedule import Program

import frontend_cases as FC

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "frontend_conformance.json")) as f:
    GOLD = json.load(f)


def _load(name, directory):
    path = os.path.join(directory, f"fc_{name}.py")
    with open(path, "w") as f:
        f.write(FC.source_of(name, "import xgrid_b200 as xgrid"))
    spec = importlib.util.spec_from_file_location(f"fc_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _plain(v):
    if dataclasses.is_dataclass(v):
        return list(dataclasses.astuple(v))
    return v.item() if hasattr(v, "item") else v


def test_golden_covers_every_case():
    assert set(GOLD) == set(FC.CASES), "tests/frontend_cases.py changed: regenerate tests/golden/frontend_conformance.json"


@pytest.mark.parametrize("name", sorted(FC.CASES))
def test_case_matches_reference(tmp_path, name):
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    want = GOLD[name]
    got = {"ok": True, "depth": None, "ret": None, "error": None}
    try:
        mod = _load(name, str(tmp_path))
        prog = Program(mod.k)                 # parse + schedule + CUDA C generation (no compilation)
        got["depth"] = prog.depth
        if hasattr(mod, "CALL"):
            got["ret"] = _plain(mod.k(*mod.CALL))
    except Exception as e:                    # the reference raises plain Exceptions (Logger.dead)
        got["ok"], got["error"] = False, str(e)
    if name in FC.DEVIATIONS:
        assert got["ok"] != want["ok"], f"{name} is listed as a deliberate deviation but now matches the reference"
        return
    assert got["ok"] == want["ok"], f"reference {'accepts' if want['ok'] else 'rejects'} this program " \
                                    f"({want['error']}); here: {got['error']}"
    if want["ok"]:
        assert got["depth"] == want["depth"]
        if want["ret"] is not None and not (want["ret"] == 0 and got["ret"] is None):    # void: ctypes gives 0
            assert got["ret"] == want["ret"] and type(got["ret"]) is type(want["ret"])


def test_every_accepted_grid_kernel_compiles_unit_by_unit(tmp_path):
    """The launch path compiles each generated kernel as its own NVRTC unit (lazy JIT): every unit of every
    accepted grid kernel -- struct-element grids, dataclass methods, called operators, inline C, shape() --
    must compile for sm_100a on its own (declarations are shared, kernels never depend on each other)."""
    xgrid.init(precision="double", cacheroot=str(tmp_path / "xg"))
    units = 0
    for name in sorted(FC.CASES):
        if not GOLD[name]["ok"] or name in FC.DEVIATIONS:
            continue
        prog = Program(_load(name, str(tmp_path)).k)
        if not prog.source:
            continue                                  # scalar-only kernel: nothing runs on the device
        for names, image in prog.images():
            assert image[:4] == b"\x7fELF", (name, names)
            units += 1
    assert units >= 80
