#!/usr/bin/env python
"""Generate tests/golden/frontend_conformance.json by running tests/frontend_cases.py through the
UNMODIFIED reference front end (parser + generator; scalar-only kernels are also compiled with gcc and
called once).

    cd /tmp && python /root/repo/tests/golden/make_frontend_golden.py
"""
import dataclasses
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
import frontend_cases as FC      # noqa: E402


def load(name, text, directory):
    path = os.path.join(directory, f"fc_{name}.py")
    with open(path, "w") as f:
        f.write(text)
    spec = importlib.util.spec_from_file_location(f"fc_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def plain(v):
    if dataclasses.is_dataclass(v):
        return list(dataclasses.astuple(v))
    if hasattr(v, "item"):
        v = v.item()
    return v


def main():
    work = tempfile.mkdtemp(prefix="xgrid_fc_")
    os.chdir(work)
    sys.path.insert(0, REF)
    import xgrid
    from xgrid.lang.generator import Generator
    from xgrid.util.logging import Logger, LogLevel
    Logger.level = LogLevel.fail if hasattr(LogLevel, "fail") else LogLevel.warn
    xgrid.init(precision="double", opt_level=2, cacheroot=".xg", parallel=True)
    out = {}
    for name in FC.CASES:
        rec = {"ok": True, "error": None, "depth": None, "ret": None}
        try:
            mod = load(name, FC.source_of(name, "import xgrid"), work)
            gen = Generator(mod.k)
            gen.source
            rec["depth"] = gen.depth + 1
            if hasattr(mod, "CALL"):
                rec["ret"] = plain(mod.k(*mod.CALL))
        except BaseException as e:          # Logger.dead raises Exception((msg, ...)); gcc errors are Exceptions too
            rec["ok"] = False
            rec["error"] = str(e)[:300]
        out[name] = rec
        print(f"{name:48s} {'ok ' if rec['ok'] else 'REJ'} depth={rec['depth']} ret={rec['ret']} {rec['error'] or ''}"[:200])
    with open(os.path.join(HERE, "frontend_conformance.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote frontend_conformance.json:", sum(r["ok"] for r in out.values()), "accepted,",
          sum(not r["ok"] for r in out.values()), "rejected")


if __name__ == "__main__":
    main()
