#!/usr/bin/env python
"""ncu .ncu-rep -> small CSV summary (metric,value,unit) of the first captured launch, plus the
top warp-stall reasons.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max"]


def main(rep: str, out: str) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [("metric", "value", "unit"), ("Kernel Name", first[col["Kernel Name"]], "")]
    for w in WANT:
        if w in col:
            lines.append((w, first[col[w]], units[col[w]]))
    stalls = [(h, float(first[i].replace(",", ""))) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
              and first[i] not in ("", "n/a")]
    for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
        lines.append((h, f"{v:.3f}", "warps/issue"))
    with open(out, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    print(out, "\n".join(",".join(l) for l in lines[:12]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
