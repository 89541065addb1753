#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference's own compiled kernels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run in the build container,
where the reference is mounted read-only at /root/reference:

    python oracle/make_ref.py

The reference is Python that emits C99+OpenMP and JIT-compiles it with the
system gcc into ``<cacheroot>/<md5>.c.so`` (xgrid/util/ffi.py:59-94).  This
script imports the reference from where it lies, points its ``cacheroot`` at
``oracle/_ref`` and calls every workload kernel once on a tiny grid, so the
reference itself writes its generated C and the shared objects there
(``init(precision="double", opt_level=3, parallel=True)`` -> ``gcc -shared
-fpic -lm -fopenmp -O3``, xgrid/util/init.py:22-33).  Nothing from the
reference's sources is copied: the kernel programs are the DSL text of
``examples/workloads.py`` (the README / test.py / examples programs),
re-imported with ``xgrid`` bound to the reference package.

``oracle/_ref/manifest.json`` records, per kernel, the shared object, the
exported symbol (generator.py:200-209: named after the kernel) and the ring
depth the reference computed (generator.py:108,428).  ``oracle/ref.py`` loads
them with ctypes on the GPU box, where /root/reference does not exist.
``oracle/_ref/`` is git-ignored (build output) but not gpurun-ignored.

Valid domain: 1-D any size, 2-D square only; heat_3d is
skipped (the reference's 3-D addressing is wrong, F1).
"""
import importlib.util
import json
import os
import shutil
import sys
import tempfile

import numpy as np

REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
WORKLOADS = os.path.join(os.path.dirname(HERE), "examples", "workloads.py")


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "xgrid")):
        print("make_ref: no reference at", REF, "- nothing to do")
        return 0
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    os.chdir(HERE)                       # Compiler.cacheroot = "./" + cacheroot (ffi.py:44)
    sys.path.insert(0, REF)
    import xgrid
    import xgrid.lang.operator  # noqa: F401  (make_kernels filters on xgrid.lang.operator.Operator)
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.warn
    xgrid.init(precision="double", opt_level=3, parallel=True, cacheroot="_ref")

    # the workload programs, with `xgrid` = the reference package (inspect.getsource needs a file)
    tmp = tempfile.mkdtemp(prefix="xgrid_ref_")
    with open(WORKLOADS) as f:
        text = f.read().replace("import xgrid_b200 as xgrid", "import xgrid\nimport xgrid.lang.operator")
    path = os.path.join(tmp, "ref_workloads.py")
    with open(path, "w") as f:
        f.write(text)
    spec = importlib.util.spec_from_file_location("ref_workloads", path)
    W = importlib.util.module_from_spec(spec)
    sys.modules["ref_workloads"] = W
    spec.loader.exec_module(W)
    kernels = W.make_kernels()

    def grid(shape):
        g = xgrid.Grid(shape, float)
        g.now[...] = np.random.default_rng(0).random(shape)
        return g

    calls = {
        "elementwise_mul": lambda k: k(grid((16,)), grid((16,)), grid((16,))),
        "convection_1d": lambda k: k(grid((16,)), 1.0, 0.01, 0.1),
        "convection_1d_nonlinear": lambda k: k(grid((16,)), 0.01, 0.1),
        "diffusion_1d": lambda k: k(grid((16,)), 0.01, 0.01, 0.1),
        "convection_2d": lambda k: k(grid((8, 8)), 1.0, 0.01, 0.1, 0.1),
        "diffusion_2d": lambda k: k(grid((8, 8)), 0.2),
        "cavity_kernel": lambda k: k(grid((8, 8)), grid((8, 8)), grid((8, 8)), grid((8, 8)),
                                     W.Config(1.0, 0.1, 1e-4, 0.1, 0.1)),
    }
    manifest = {}
    for name, call in calls.items():
        before = set(os.listdir(OUT))
        op = kernels[name]
        call(op)                                    # lazy JIT (operator.py:31-34): gcc runs here
        new = sorted(f for f in set(os.listdir(OUT)) - before if f.endswith(".so"))
        assert len(new) == 1, (name, new)
        manifest[name] = {"lib": new[0], "symbol": op.name, "depth": int(op.depth)}
        print("make_ref:", name, "->", new[0], "depth", op.depth)
    with open(os.path.join(OUT, new[0][:-3])) as f:
        cmdline = f.readline().strip()
    manifest["_meta"] = {"cc": cmdline, "reference": REF,
                         "config": "precision=double opt_level=3 parallel=True"}
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
