#!/usr/bin/env python
"""Generate tests/golden/randprog.npz: random DSL programs run by the UNMODIFIED reference.

    cd /tmp && python /root/repo/tests/golden/make_random_golden.py

The programs are the ones ``tests/randprog.py`` generates for the randomized differential GPU tests
(several grids, time levels 0..2, offsets up to +-2, masked / implicit / looped statements, scalar
control flow).  Here their text is handed to the reference itself (``import xgrid`` instead of
``import xgrid_b200 as xgrid``), on the shapes where the reference's addressing is valid (1-D and
SQUARE 2-D), and every ring level of every grid after three calls is stored together
with the program text.  ``tests/test_interp.py::test_random_programs_match_reference`` replays them
through ``oracle/interp.py`` bit for bit -- so the interpreter the GPU tests use as their oracle is
pinned to the reference over the DSL surface, not only on the workload kernels.

Cells whose taps could leave the array get a mask value that matches no statement
(``randprog.guard_array_ends``): out-of-array reads are undefined in the reference.
Inputs are not stored: ``randprog.gen_inputs(seed, shape, ngrids)`` is deterministic
(``numpy.random.default_rng``); the stored program text guards against generator drift.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
from randprog import gen_inputs, gen_source, guard_array_ends, load_program      # noqa: E402

# (seed, ndim, ngrids, shape, single_1d)
CASES = [(s, 1, 1 + s % 3, [(37,), (301,), (1000,)][s % 3], False) for s in range(100, 112)]
CASES += [(s, 2, 1 + s % 3, [(17, 17), (24, 24), (33, 33)][s % 3], False) for s in range(112, 130)]
CASES += [(s, 1, 1, (500,), True) for s in range(130, 134)]
# overstep="wrap" / "limit": taps never leave the array, so no guard band is needed
OVERSTEP_CASES = [(s, 1, 1 + s % 2, [(301,), (1000,)][s % 2], False) for s in range(200, 202)]
OVERSTEP_CASES += [(s, 2, 1 + s % 3, [(17, 17), (24, 24), (33, 33), (40, 40)][s % 4], False) for s in range(202, 206)]
# precision="float": fp32 grids and scalars, double literals (C's mixed-precision typing)
FP32_CASES = [(s, 1, 1 + s % 2, [(301,), (1000,)][s % 2], False) for s in range(300, 303)]
FP32_CASES += [(s, 2, 1 + s % 3, [(17, 17), (24, 24), (33, 33)][s % 3], False) for s in range(303, 309)]
CALLS = 3
A, B = 0.3, 1.7


def run_cases(xgrid, cases, guard: bool, dtype=np.float64) -> dict:
    out = {}
    for seed, ndim, ngrids, shape, single in cases:
        src = gen_source(seed, ndim, ngrids, single_1d=single)
        ref_src = src.replace("import xgrid_b200 as xgrid", "import xgrid")
        assert ref_src != src
        prog = load_program(ref_src, os.getcwd(), f"refprog_{seed}_{len(os.listdir(os.getcwd()))}")
        ics, masks = gen_inputs(seed, shape, ngrids)
        if guard:
            guard_array_ends(masks, shape)      # no statement may read outside the array (undefined in the reference)
        grids = []
        for ic, m in zip(ics, masks):
            g = xgrid.Grid(shape, float)
            assert g.now.dtype == dtype
            g.now[...] = ic.astype(dtype)
            g.boundary[...] = m
            grids.append(g)
        for _ in range(CALLS):
            prog(*grids, A, B)
        out[f"{seed}.src"] = np.array(src)
        out[f"{seed}.meta"] = np.array([ndim, ngrids, int(single), *shape])
        for n, g in enumerate(grids):
            out[f"{seed}.g{n}.depth"] = np.array(len(g._data))
            for lvl, arr in enumerate(g._data):
                out[f"{seed}.g{n}.L{lvl}"] = np.array(arr)
        print("ran", seed, shape, "depth", len(grids[0]._data))
    return out


def main():
    os.chdir(tempfile.mkdtemp(prefix="xgrid_randgold_"))
    sys.path.insert(0, REF)
    import xgrid
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.warn
    for mode, cases, name, precision in (("none", CASES, "randprog", "double"),
                                         ("wrap", OVERSTEP_CASES, "randprog_wrap", "double"),
                                         ("limit", OVERSTEP_CASES, "randprog_limit", "double"),
                                         ("none", FP32_CASES, "randprog_f32", "float")):
        xgrid.init(precision=precision, opt_level=3, cacheroot=f".xg_{name}", parallel=True, overstep=mode)
        out = run_cases(xgrid, cases, guard=(mode == "none"),
                        dtype=np.float64 if precision == "double" else np.float32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
