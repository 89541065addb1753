#!/usr/bin/env python
"""Generate tests/golden/randscalar.json: the random scalar kernels of tests/randscalar.py compiled (gcc) and
called by the UNMODIFIED reference, with precision="double" (randscalar.json) and precision="float"
(randscalar_f32.json: fp32 variables, double literals -- C's mixed-precision typing); return values are stored as repr (ints) / float.hex (floats).

    cd /tmp && python /root/repo/tests/golden/make_scalar_golden.py
"""
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
from randscalar import gen_source      # noqa: E402

SEEDS = list(range(1000, 1120))


def main():
    work = tempfile.mkdtemp(prefix="xgrid_scalar_")
    os.chdir(work)
    sys.path.insert(0, REF)
    import xgrid
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.warn
    for precision, name in (("double", "randscalar"), ("float", "randscalar_f32")):
        xgrid.init(precision=precision, opt_level=2, cacheroot=f".xg_{precision}", parallel=True)
        run(xgrid, work, precision, name)


def run(xgrid, work, precision, name):
    out = {}
    for seed in SEEDS if precision == "double" else SEEDS[:60]:
        src, args = gen_source(seed)
        path = os.path.join(work, f"rs_{precision}_{seed}.py")
        with open(path, "w") as f:
            f.write(src.replace("IMPORT_LINE", "import xgrid"))
        spec = importlib.util.spec_from_file_location(f"rs_{precision}_{seed}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ret = mod.k(*args)
        out[str(seed)] = {"src": src, "args": list(args),
                          "ret": float(ret).hex() if isinstance(ret, float) else repr(int(ret)),
                          "float": isinstance(ret, float)}
        print(seed, args, ret)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", name + ".json", "with", len(out), "kernels")


if __name__ == "__main__":
    main()
