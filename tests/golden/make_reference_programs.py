"""Fixture generator (build container only): lifts the reference's USER PROGRAMS -- test.py facts 3-10 and the
lid-driven-cavity example up to its time loop -- as text, unmodified, into tests/golden/reference_programs.json.

They are the workload inputs SURVEY.md §8c says to lift (the reference pins no numeric results; its programs
are its specification).  tests/test_import_xgrid_gpu.py executes this text under ``import xgrid`` (the alias
package at the repo root) on the GPU box, where /root/reference does not exist.

    python tests/golden/make_reference_programs.py
"""
import hashlib
import json
import os

REF = os.environ.get("XGRID_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

PIECES = {
    # name: (file, first line, last line) -- 1-based, inclusive
    "test_py_facts_3_to_10": ("test.py", 125, 314),
    "cavity_py_setup_kernel_loop": ("examples/cavity.py", 1, 151),
}


def main() -> None:
    out = {}
    for name, (rel, lo, hi) in PIECES.items():
        path = os.path.join(REF, rel)
        with open(path) as f:
            text = f.read()
        lines = text.splitlines(keepends=True)
        out[name] = {"file": rel, "lines": [lo, hi], "sha256_of_file": hashlib.sha256(text.encode()).hexdigest(),
                     "text": "".join(lines[lo - 1:hi])}
    with open(os.path.join(HERE, "reference_programs.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, v["file"], v["lines"], len(v["text"].splitlines()), "lines")


if __name__ == "__main__":
    main()
