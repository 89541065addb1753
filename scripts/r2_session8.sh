#!/bin/bash
O=gpurun_out
mkdir -p $O
run() {
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --workload "${WL:-heat3d}" --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:52s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.4f} ms  frac {d['roofline']['frac']:.3f}  launches {d['gpu_launches']} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:52s} FAILED {line[:300]}")
PY
}
{
run "default (W=256, TJ=8, 2 CTAs/SM)" A=1
run "W=512: WX3=2 NSV=4 TJ=8 smem 210K" XGB_WX3=2 XGB_NSV=4 XGB_TJ=8 XGB_SMEM=215040
run "W=512: WX3=1 NSV=8 TJ=8 smem 210K" XGB_WX3=1 XGB_NSV=8 XGB_TJ=8 XGB_SMEM=215040 XGB_TILED_BATCH_MAX=200
run "W=1024: WX3=4 NSV=4 TJ=4 smem 210K" XGB_WX3=4 XGB_NSV=4 XGB_TJ=4 XGB_SMEM=215040
WL=conv1d run "conv1d 20 steps (cached marshalling)" A=1
WL=conv1d run "conv1d 20 steps (repeat)" A=1
} 2>&1 | tee $O/r2j_session8.txt
