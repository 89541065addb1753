#!/bin/bash
# heat3d: wider column tiles (longer contiguous DRAM segments per row), one CTA per SM with 16 consumer warps
O=gpurun_out
mkdir -p $O
run() {
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --workload heat3d --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:52s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.3f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:52s} FAILED {line[:300]}")
PY
}
{
run "default (W=256, TJ=8, 2 CTAs/SM)" A=1
run "W=512: WX3=2 NSV=4 TJ=8 smem 210K" XGB_WX3=2 XGB_NSV=4 XGB_TJ=8 XGB_SMEM=215040
run "W=512: WX3=2 NSV=4 TJ=4 smem 105K" XGB_WX3=2 XGB_NSV=4 XGB_TJ=4 XGB_SMEM=107520
run "W=512: WX3=1 NSV=8 TJ=8 smem 210K" XGB_WX3=1 XGB_NSV=8 XGB_TJ=8 XGB_SMEM=215040 XGB_TILED_BATCH_MAX=200
run "W=1024: WX3=4 NSV=4 TJ=4 smem 210K" XGB_WX3=4 XGB_NSV=4 XGB_TJ=4 XGB_SMEM=215040
run "W=512: WX3=2 NSV=4 TJ=6 smem 210K" XGB_WX3=2 XGB_NSV=4 XGB_TJ=6 XGB_SMEM=215040
run "W=128: NSV=2 TJ=16 (16 warps) smem 105K" XGB_NSV=2 XGB_TJ=16 XGB_SMEM=107520
run "default (repeat)" A=1
} 2>&1 | tee $O/r2i_session7.txt
