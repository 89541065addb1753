"""Type-layout conformance (CPU only): by-value structs must have exactly the reference's C layout
(tests/golden/type_layout.json, produced by the reference's own type system through ctypes), for
precision="float" and "double".  Struct-element grids: the reference builds the NumPy element dtype as a
PACKED record (xgrid/xgrid/__init__.py:17-18) although its generated C indexes an array of naturally
ALIGNED structs, so the two only agree when no padding exists; this backend's element dtype always equals
the C layout -- identical to the reference wherever the reference is consistent with itself."""
import json
import os

import pytest

import xgrid_b200 as xgrid
from xgrid_b200.grid import parse_numpy_dtype
from xgrid_b200.types import parse_annotation

import type_cases

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "type_layout.json")) as f:
    GOLD = json.load(f)


@pytest.mark.parametrize("precision", ["float", "double"])
def test_struct_and_scalar_layout(tmp_path, precision):
    xgrid.init(precision=precision, cacheroot=str(tmp_path / "xg"))
    got = type_cases.describe(parse_annotation, parse_numpy_dtype, precision)
    assert set(got) == {k for k in GOLD if k.startswith(precision + ".")}
    consistent = 0
    for key, mine in got.items():
        want = GOLD[key]
        if "offsets" not in want:                       # int / float / bool
            assert mine == want, key
            continue
        for field in ("size", "align", "offsets"):      # the by-value C struct: always identical
            assert mine[field] == want[field], (key, field)
        # element dtype of a struct grid == the C layout ...
        assert mine["np_itemsize"] == mine["size"] and mine["np_offsets"] == mine["offsets"], key
        # ... which is the reference's dtype whenever the reference's packed record has no hidden mismatch
        if want["np_itemsize"] == want["size"]:
            consistent += 1
            assert mine["np_itemsize"] == want["np_itemsize"] and mine["np_offsets"] == want["np_offsets"], key
    assert consistent >= 1
