#!/bin/bash
# heat3d one-pass kernel: env sweeps (tile shape, stages, chunk length) + one ncu --set full capture.
O=gpurun_out
mkdir -p $O
run() {  # label, env...
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --workload heat3d --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:44s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.3f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:44s} FAILED {line[:200]}")
PY
}
{
run "default" A=1
run "default (repeat)" A=1
run "smem 72K" XGB_SMEM=73728
run "smem 150K" XGB_SMEM=153600
run "smem 200K" XGB_SMEM=204800
run "NSV=2" XGB_NSV=2
run "NSV=2 smem 72K" XGB_NSV=2 XGB_SMEM=73728
run "NSV=2 TJ=16" XGB_NSV=2 XGB_TJ=16
run "TJ=4" XGB_TJ=4
run "TJ=4 smem 72K" XGB_TJ=4 XGB_SMEM=73728
run "TJ=16 smem 200K" XGB_TJ=16 XGB_SMEM=204800
run "WX3=2 NSV=2" XGB_WX3=2 XGB_NSV=2
run "WX3=2 NSV=2 TJ=4" XGB_WX3=2 XGB_NSV=2 XGB_TJ=4
run "min_ctas 2368" XGB_MIN_CTAS=2368
run "min_ctas 4736" XGB_MIN_CTAS=4736
run "min_ctas 16384" XGB_MIN_CTAS=16384
run "march variant (no tiled)" XGB_TILED=0
run "fma build" XGB_NOP=1
} 2>&1 | tee $O/r2b_tune_heat3d.txt
# ncu: full set on the heat3d kernel (one launch)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:heat_3d.*tiled -s 3 -c 1 -o $O/r2b_heat3d_tiled \
   python bench.py --workload heat3d --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r2b_ncu.log 2>&1
tail -3 $O/r2b_ncu.log
ls -la $O/*.ncu-rep
