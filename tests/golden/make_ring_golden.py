#!/usr/bin/env python
"""tests/golden/ring_protocol.npz: the scenarios of tests/ring_cases.py run by the UNMODIFIED reference; every
ring level of both grids after every call.

    cd /tmp && python /root/repo/tests/golden/make_ring_golden.py
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ring_cases as RC      # noqa: E402


def main():
    work = tempfile.mkdtemp(prefix="xgrid_ring_")
    os.chdir(work)
    sys.path.insert(0, os.environ.get("XGRID_REFERENCE", "/root/reference"))
    import xgrid
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.warn
    xgrid.init(precision="double", opt_level=2, cacheroot=".xg", parallel=True)
    path = os.path.join(work, "ring_kernels.py")
    with open(path, "w") as f:
        f.write(RC.SOURCE.replace("IMPORT_LINE", "import xgrid"))
    spec = importlib.util.spec_from_file_location("ring_kernels", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name, calls in RC.SCENARIOS.items():
        grids = {}
        for which in "gh":
            ic, mask = RC.initial(np, which)
            g = xgrid.Grid((RC.N,), float)
            g.now[...] = ic
            g.boundary[...] = mask
            grids[which] = g
        for step, (kernel, spec_) in enumerate(calls):
            getattr(mod, kernel)(*RC.arguments(spec_, grids))
            for which, g in grids.items():
                out[f"{name}.{step}.{which}.depth"] = np.array(len(g._data))
                for lvl, arr in enumerate(g._data):
                    out[f"{name}.{step}.{which}.L{lvl}"] = np.array(arr)
        print(name, [len(g._data) for g in grids.values()])
    np.savez_compressed(os.path.join(HERE, "ring_protocol.npz"), **out)
    print("wrote ring_protocol.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
