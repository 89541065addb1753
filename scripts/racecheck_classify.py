#!/usr/bin/env python
"""compute-sanitizer racecheck logs (--racecheck-report hazard) -> one line per class of hazard:
(severity, RAW/WAR/WAW, writer function, reader function) with counts, distinct addresses and blocks."""
import collections
import re
import sys


def classify(path: str) -> None:
    txt = open(path, errors="replace").read()
    summary = re.findall(r"RACECHECK SUMMARY: [^\n]+", txt)
    print(f"== {path}: {summary[-1] if summary else 'no summary'}")
    kinds = collections.Counter()
    addrs, blocks = collections.defaultdict(set), collections.defaultdict(set)
    for b in txt.split("========= \n"):
        m = re.search(r"(Error|Warning): (\(Warp Level Programming\) )?Potential (\w+) hazard detected at __shared__ (0x[0-9a-f]+) "
                      r"in block \((\d+),(\d+),(\d+)\)", b)
        if not m:
            continue
        fn = {}
        for role in ("Write", "Read"):
            for hit in re.finditer(role + r" Thread \((\d+),\d+,\d+\) at ([^\n]+)", b):
                name = re.sub(r"\+0x[0-9a-f]+.*", "", hit.group(2))
                name = re.sub(r"^void ", "", name).split("(")[0]
                fn.setdefault(role, name)
        key = (m.group(1), "warp-level" if m.group(2) else "cross-warp", m.group(3), fn.get("Write"), fn.get("Read"))
        kinds[key] += 1
        addrs[key].add(int(m.group(4), 16) // 8)
        blocks[key].add((m.group(5), m.group(6), m.group(7)))
    for key, n in kinds.most_common():
        print(f"   {n:7d} byte-hazards  {len(addrs[key]):6d} distinct 8-byte words  {len(blocks[key]):4d} blocks  {key}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        classify(p)
