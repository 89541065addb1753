#!/usr/bin/env python
"""tests/golden/type_layout.json: C struct layout (ctypes) and NumPy element dtype of the reference's type
system for the dataclasses of tests/type_cases.py, with precision="float" and "double".

    cd /tmp && python /root/repo/tests/golden/make_type_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("XGRID_REFERENCE", "/root/reference"))
import type_cases      # noqa: E402
import xgrid           # noqa: E402
from xgrid.util.typing.annotation import parse_annotation      # noqa: E402
from xgrid.xgrid import parse_numpy_dtype                        # noqa: E402

out = {}
for precision in ("float", "double"):
    xgrid.init(precision=precision, cacheroot="/tmp/.xg_types")
    out.update(type_cases.describe(parse_annotation, parse_numpy_dtype, precision))
with open(os.path.join(HERE, "type_layout.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True)[:1500])
