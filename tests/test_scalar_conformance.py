"""Scalar semantics against the UNMODIFIED reference (CPU only): 120 random scalar-only kernels
(tests/randscalar.py) were compiled with gcc and called by the reference
(tests/golden/make_scalar_golden.py -> randscalar.json); the host evaluator of this backend
(`lang/schedule.py::HostEval`, which also computes the scalar prologue of every grid kernel) must return
the same int / the same double (or, with precision="float", the same float) bit for bit."""
import importlib.util
import json
import os

import pytest

import xgrid_b200 as xgrid
from randscalar import gen_source

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = {}
for _precision, _name in (("double", "randscalar"), ("float", "randscalar_f32")):
    with open(os.path.join(HERE, "golden", _name + ".json")) as f:
        GOLD[_precision] = json.load(f)


@pytest.mark.parametrize("precision,seed", [(p, s) for p in GOLD for s in sorted(GOLD[p], key=int)])
def test_scalar_kernel_matches_reference(tmp_path, precision, seed):
    xgrid.init(precision=precision, cacheroot=str(tmp_path / "xg"))
    want = GOLD[precision][seed]
    src, args = gen_source(int(seed))
    assert src == want["src"] and list(args) == want["args"], \
        "tests/randscalar.py changed: regenerate tests/golden/randscalar.json"
    path = tmp_path / f"rs_{seed}.py"
    path.write_text(src.replace("IMPORT_LINE", "import xgrid_b200 as xgrid"))
    spec = importlib.util.spec_from_file_location(f"rs_{seed}", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.k(*args)
    if want["float"]:
        assert isinstance(got, float) and got.hex() == want["ret"], (got, float.fromhex(want["ret"]), src)
    else:
        assert isinstance(got, int) and not isinstance(got, bool) and repr(got) == want["ret"], (got, want["ret"], src)
