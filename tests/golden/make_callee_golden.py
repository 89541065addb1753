"""Golden for tests/programs/callee_prog.py, produced by the UNMODIFIED reference (build container only).

    python tests/golden/make_callee_golden.py
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    sys.path.insert(0, REF)
    os.chdir(tempfile.mkdtemp(prefix="xgrid_callee_"))
    import xgrid
    xgrid.init(precision="double", opt_level=3, parallel=True, cacheroot=".xg")
    spec = importlib.util.spec_from_file_location("callee_prog", os.path.join(HERE, "..", "programs", "callee_prog.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    outer = mod.outer
    n = 24                                      # square: the reference's strides are only right there (F1)
    rng = np.random.default_rng(42)
    u_in, v_in = rng.random((n, n)), rng.random((n, n))
    mask = np.zeros((n, n), np.int32)
    mask[0, :] = mask[-1, :] = mask[:, 0] = 1
    mask[:, -1] = 1
    mask[1:-1, 0] = 2                           # Neumann column of `smooth` (reads its right neighbour at level 0)
    mask[7, 9] = 1                              # value 2 has no statement in `relax` / `outer`: stale cells (F5)
    u, v = xgrid.Grid((n, n), float), xgrid.Grid((n, n), float)
    u.now[...] = u_in
    v.now[...] = v_in
    u.boundary[...] = mask
    v.boundary[...] = mask
    steps = 5
    for _ in range(steps):
        outer(u, v, 0.4)
    out = {"u_in": u_in, "v_in": v_in, "mask": mask, "steps": steps, "a": 0.4, "depth": outer.depth}
    for name, g in (("u", u), ("v", v)):
        for l, arr in enumerate(g._data):
            out[f"{name}.L{l}"] = arr
    np.savez_compressed(os.path.join(HERE, "callee_f64.npz"), **out)
    print("depth", outer.depth, "levels", len(u._data), "u.now[3,3] =", u.now[3, 3])


if __name__ == "__main__":
    main()
