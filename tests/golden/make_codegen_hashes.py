#!/usr/bin/env python
"""tests/golden/codegen_hashes.json: sha256 of the CUDA C generated for every workload kernel (overstep none / wrap,
precision double / float), of the template headers and of the NVRTC flags.  Regenerate ONLY together with a GPU
validation of the new device code (pytest -m gpu on a B200), and say so in the commit:

    python tests/golden/make_codegen_hashes.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def hashes(cacheroot: str) -> dict:
    import xgrid_b200 as xgrid
    from examples import workloads as W
    from xgrid_b200.lang.schedule import Program, template_headers
    out = {}
    for mode in ("none", "wrap"):
        for precision in ("double", "float"):
            xgrid.init(precision=precision, cacheroot=cacheroot, overstep=mode)
            for name, op in W.make_kernels().items():
                prog = Program(op)
                out[f"{mode}.{precision}.{name}"] = hashlib.sha256(prog.source.encode()).hexdigest()[:16]
            out[f"flags.{mode}.{precision}"] = " ".join(prog.config.nvrtc_flags)
    out["headers"] = hashlib.sha256("".join(template_headers().values()).encode()).hexdigest()[:16]
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "codegen_hashes.json"), "w") as f:
        json.dump(hashes("/tmp/.xg_hashes"), f, indent=1, sort_keys=True)
    print("wrote codegen_hashes.json")
