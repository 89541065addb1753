#!/usr/bin/env python
"""gpurun_out/r2_traffic_<case>.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv, one file per bench variant; scripts/r2_profiles.sh) -> profiles/traffic.json + profiles/r2_traffic_<case>.csv.

For every case the DOMINANT kernel (largest total time) is reported: DRAM read+write bytes per launch (median over
its captured launches), launches seen, time per launch under ncu.  bench.py's roofline.traffic reads the table;
keys: <workload> (the default, temporally blocked path), <workload>:tail (short runs), <workload>:onepass."""
import csv
import json
import os
import shutil
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {   # file tag -> (traffic.json key, steps per launch of the dominant kernel)
    "heat3d": ("heat3d", 1), "conv1d": ("conv1d", 64), "conv1d_tail": ("conv1d:tail", 20),
    "conv1d_onepass": ("conv1d:onepass", 1), "conv1d_nl": ("conv1d_nl", 64), "diff1d": ("diff1d", 64),
    "conv2d": ("conv2d", 2), "conv2d_onepass": ("conv2d:onepass", 1), "diff2d": ("diff2d", 2),
    "diff2d_onepass": ("diff2d:onepass", 1), "cavity": ("cavity", 2),
}


def parse(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        d = launches.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1,
                                                                          "ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
    return list(launches.values())


def main(src_dir: str) -> None:
    table = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel of each bench "
                         "variant (ncu --metrics, --clock-control none; scripts/r2_profiles.sh, round 2). Keys: "
                         "<workload> = default path, :tail = short deferred run, :onepass = one launch per step"}
    for tag, (key, steps) in CASES.items():
        path = os.path.join(src_dir, f"r2_traffic_{tag}.csv")
        if not os.path.exists(path):
            continue
        launches = parse(path)
        if not launches:
            continue
        by = {}
        for l in launches:
            by.setdefault(l["kernel"], []).append(l)
        kernel, ls = max(by.items(), key=lambda kv: sum(x.get("gpu__time_duration.sum", 0) for x in kv[1]))
        byt = statistics.median(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in ls)
        ns = statistics.median(x.get("gpu__time_duration.sum", 0) for x in ls)
        shutil.copy(path, os.path.join(ROOT, "profiles", os.path.basename(path)))
        table[key] = {"kernel": kernel, "bytes_per_launch": int(byt), "steps_per_launch": steps,
                      "source": f"profiles/{os.path.basename(path)} ({len(ls)} launches captured, median {ns / 1e3:.1f} us per "
                                f"launch under ncu = {byt / max(ns, 1):.0f} GB/s DRAM)"}
        print(f"{key:18s} {kernel:48s} {byt / 1e6:10.1f} MB/launch {ns / 1e3:9.1f} us  x{len(ls)}")
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(table, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out"))
